// cabanapic_b200 C++ host facade -- data model.
//
// Same names as the reference's src/types.h (real_t, the member enums, particle_list_t,
// interpolator_array_t, accumulator_array_t, accumulator_array_sa_t, field_array_t,
// particle_mover_t, VOXEL, RANK_TO_INDEX; reference src/types.h:4-8,31-58,64-108,116-134,138-195),
// so decks/*.cxx and a driver written for CabanaPIC compile unchanged -- but the arrays are
// DEVICE-BACKED: each one is a host mirror (what deck initialisers and diagnostics see through
// Cabana::slice<>() / access()) plus residency flags; the hot-path calls in push.h / fields.h /
// accumulator.h / interpolator.h run on the B200 through the C ABI (include/cabanapic_b200.h)
// and move data only when a side that is out of date is actually touched.
#ifndef CABANAPIC_B200_TYPES_H
#define CABANAPIC_B200_TYPES_H

#ifndef REAL_TYPE
#define real_t float
#else
#define real_t REAL_TYPE
#endif

#include <Kokkos_Core.hpp>   // include/compat: host containers + loop dispatch with the Kokkos/Cabana names
#include <Cabana_Core.hpp>

#include <memory>

#ifndef CELL_BLOCK_FACTOR
#define CELL_BLOCK_FACTOR 32
#endif
const size_t cell_blocking = CELL_BLOCK_FACTOR;

using MemorySpace = Kokkos::DefaultExecutionSpace::memory_space;
using ExecutionSpace = Kokkos::DefaultExecutionSpace;

enum UserParticleFields { PositionX = 0, PositionY, PositionZ, VelocityX, VelocityY, VelocityZ, Weight, Cell_Index };
enum InterpolatorFields {
    EX = 0, DEXDY, DEXDZ, D2EXDYDZ, EY, DEYDZ, DEYDX, D2EYDZDX, EZ, DEZDX, DEZDY, D2EZDXDY,
    CBX, DCBXDX, CBY, DCBYDY, CBZ, DCBZDZ
};
enum FieldFields { FIELD_EX = 0, FIELD_EY, FIELD_EZ, FIELD_CBX, FIELD_CBY, FIELD_CBZ, FIELD_JFX, FIELD_JFY, FIELD_JFZ };
namespace accumulator_var { enum a_v { jx = 0, jy = 1, jz = 2 }; }
#define ACCUMULATOR_VAR_COUNT 3
#define ACCUMULATOR_ARRAY_LENGTH 4

namespace cabanapic {

// Which copy of a device-backed array is current.  Shared by all handles to the same array
// (the reference passes these containers by value as ref-counted handles).
struct Residency {
    bool host_valid = true;     // the host mirror holds the current values
    bool device_valid = false;  // the GPU context holds the current values
};

enum class ArrayKind { particles, fields, interpolators, accumulators };

// A Cabana-style AoSoA whose authoritative copy may live on the GPU.  Host access goes through
// Cabana::slice<M>() (overloaded below): it first brings the mirror up to date, then assumes the
// caller may write and marks the device copy stale.
template <ArrayKind K, class Members, int VL>
class DeviceBacked : public Cabana::AoSoA<Members, Kokkos::HostSpace, VL> {
   public:
    using base = Cabana::AoSoA<Members, Kokkos::HostSpace, VL>;
    static constexpr ArrayKind kind = K;
    DeviceBacked() : base(), res_(std::make_shared<Residency>()) {}
    DeviceBacked(const std::string& label, std::size_t n) : base(label, n), res_(std::make_shared<Residency>()) {}
    Residency& residency() const { return *res_; }
    const base& host() const { return *this; }
    void host_access() const;     // defined in device.h (needs the runtime)

   private:
    std::shared_ptr<Residency> res_;
};

}  // namespace cabanapic

using ParticleDataTypes = Cabana::MemberTypes<real_t, real_t, real_t, real_t, real_t, real_t, real_t, int>;
using InterpolatorDataTypes = Cabana::MemberTypes<real_t, real_t, real_t, real_t, real_t, real_t, real_t, real_t, real_t,
                                                  real_t, real_t, real_t, real_t, real_t, real_t, real_t, real_t, real_t>;
using FieldDataTypes = Cabana::MemberTypes<real_t, real_t, real_t, real_t, real_t, real_t, real_t, real_t, real_t>;

using particle_list_t = cabanapic::DeviceBacked<cabanapic::ArrayKind::particles, ParticleDataTypes, 32>;
using interpolator_array_t = cabanapic::DeviceBacked<cabanapic::ArrayKind::interpolators, InterpolatorDataTypes, cell_blocking>;
using field_array_t = cabanapic::DeviceBacked<cabanapic::ArrayKind::fields, FieldDataTypes, cell_blocking>;

namespace cabanapic {
// accumulator_array_t: the reference's Kokkos::View<real_t*[3][4]> (src/types.h:120), device-backed.
class Accumulators : public Kokkos::View<real_t* [ACCUMULATOR_VAR_COUNT][ACCUMULATOR_ARRAY_LENGTH]> {
   public:
    using base = Kokkos::View<real_t* [ACCUMULATOR_VAR_COUNT][ACCUMULATOR_ARRAY_LENGTH]>;
    Accumulators() : base(), res_(std::make_shared<Residency>()) {}
    Accumulators(const std::string& label, std::size_t n) : base(label, n), res_(std::make_shared<Residency>()) {}
    Residency& residency() const { return *res_; }
    void host_access() const;
    // element access refreshes the mirror first, like a slice
    real_t& operator()(std::size_t i, std::size_t j, std::size_t k) const { host_access(); return base::operator()(i, j, k); }

   private:
    std::shared_ptr<Residency> res_;
};
// The scatter view of the reference (Kokkos::Experimental::ScatterView, src/types.h:122-123) has no
// work left to do: the deposit's privatisation lives inside the push kernel.  It survives as a
// handle so that example.cpp's create_scatter_view / contribute / reset_except lines still compile.
struct ScatterHandle {
    Accumulators target;
    template <class T> void reset_except(const T&) {}
    void reset() {}
};
}  // namespace cabanapic
using accumulator_array_t = cabanapic::Accumulators;
using accumulator_array_sa_t = cabanapic::ScatterHandle;

#include "grid.h"

class particle_mover_t {
   public:
    real_t dispx, dispy, dispz;
    int32_t i;
};

// ix + gx*(iy + gy*iz)  <->  (ix, iy, iz), gx = _x, gy = _y   (reference src/types.h:184-195)
#define RANK_TO_INDEX(rank, ix, iy, iz, _x, _y)            \
    {                                                      \
        const int r_ = (rank);                             \
        const int row_ = r_ / int(_x);                     \
        (ix) = r_ - row_ * int(_x);                        \
        (iz) = row_ / int(_y);                             \
        (iy) = row_ - (iz) * int(_y);                      \
    }
#define VOXEL(x, y, z, nx, ny, nz, NG) ((x) + ((nx) + (NG * 2)) * ((y) + ((ny) + (NG * 2)) * (z)))

#endif  // CABANAPIC_B200_TYPES_H
