// Facade of the reference's src/fields.h: Field_Solver<Solver_Type> with the EM / ES_1D policies
// (:274-379, :467-545, :550-720) and dump_energies (:722-763), executed on the GPU.
#ifndef CABANAPIC_B200_FIELDS_H
#define CABANAPIC_B200_FIELDS_H

#include <fstream>
#include "Cabana_Parallel.hpp"
#include "Cabana_DeepCopy.hpp"
#include "input/deck.h"

// Policy tags: the arithmetic of both solvers lives in cabanapic_b200/csrc/cpic_fields.cuh; which one
// runs is fixed when the context is created (-DES_FIELD_SOLVER, as in example/example.cpp:24-35).
class EM_Field_Solver {
   public:
    static constexpr int kind = CPIC_SOLVER_EM;
};
class ES_Field_Solver_1D {
   public:
    static constexpr int kind = CPIC_SOLVER_ES_1D;
};
class ES_Field_Solver {   // declared by the reference, never instantiated (example/example.cpp:27-33)
   public:
    static constexpr int kind = CPIC_SOLVER_ES_1D;
};

template <typename Solver_Type>
class Field_Solver : public Solver_Type {
   public:
    // the reference's constructor zeroes all nine field arrays (src/fields.h:279-315)
    Field_Solver(field_array_t& fields) {
#ifdef ES_FIELD_SOLVER
        static_assert(Solver_Type::kind == CPIC_SOLVER_ES_1D, "built with -DES_FIELD_SOLVER: use ES_Field_Solver_1D");
#else
        static_assert(Solver_Type::kind == CPIC_SOLVER_EM, "EM build: use EM_Field_Solver (or compile with -DES_FIELD_SOLVER)");
#endif
        auto ex = Cabana::slice<FIELD_EX>(fields);   auto ey = Cabana::slice<FIELD_EY>(fields);   auto ez = Cabana::slice<FIELD_EZ>(fields);
        auto bx = Cabana::slice<FIELD_CBX>(fields);  auto by = Cabana::slice<FIELD_CBY>(fields);  auto bz = Cabana::slice<FIELD_CBZ>(fields);
        auto jx = Cabana::slice<FIELD_JFX>(fields);  auto jy = Cabana::slice<FIELD_JFY>(fields);  auto jz = Cabana::slice<FIELD_JFZ>(fields);
        for (std::size_t i = 0; i < fields.size(); ++i) {
            ex(i) = 0; ey(i) = 0; ez(i) = 0; bx(i) = 0; by(i) = 0; bz(i) = 0; jx(i) = 0; jy(i) = 0; jz(i) = 0;
        }
    }

    // reference :352-364 -> EM :668-719 (stencil + periodic ghost copy of cB); ES_1D: no-op (:470-482)
    void advance_b(field_array_t& fields, real_t px, real_t py, real_t pz, size_t, size_t, size_t, size_t) {
        cabanapic::Runtime& rt = cabanapic::Runtime::get();
        rt.need_on_device(fields);
        rt.check(cpic_advance_b(rt.ctx(), px, py, pz), "cpic_advance_b");
        rt.device_wrote(fields);
    }
    // reference :365-378 -> EM :618-665 (J fold, J ghost copy, stencil); ES_1D :511-544
    void advance_e(field_array_t& fields, real_t px, real_t py, real_t pz, size_t, size_t, size_t, size_t, real_t dt_eps0) {
        cabanapic::Runtime& rt = cabanapic::Runtime::get();
        rt.need_on_device(fields);
        rt.check(cpic_advance_e(rt.ctx(), px, py, pz, dt_eps0), "cpic_advance_e");
        rt.device_wrote(fields);
    }
    // 0.5 * sum(E^2) over the interior (EM :556-587) or over every cell (ES_1D :484-509); summed in
    // double on the device, returned in real_t like the reference
    real_t e_energy(field_array_t& fields, real_t, real_t, real_t, size_t, size_t, size_t, size_t) {
        double e = 0, b = 0;
        energies(fields, e, b);
        return (real_t)e;
    }
    real_t b_energy(field_array_t& fields, real_t, real_t, real_t, size_t, size_t, size_t, size_t) {
        double e = 0, b = 0;
        energies(fields, e, b);
        return (real_t)b;
    }
    // ex along the x line through (y,z) = (1,1), the reference's ex1d dump format (:318-350)
    void dump_fields(FILE* fp, field_array_t& fields, real_t xmin, real_t, real_t, real_t dx, real_t, real_t, size_t nx, size_t ny,
                     size_t nz, size_t ng) {
        const bool device_was_current = fields.residency().device_valid;
        auto ex = Cabana::slice<FIELD_EX>(fields);
        fields.residency().device_valid = device_was_current;     // read-only use of the mirror
        for (size_t i = 1; i < nx + 1; i++) {
            const real_t x = xmin + (i - 0.5) * dx;
            const size_t ii = VOXEL(i, 1, 1, nx, ny, nz, ng);
            fprintf(fp, "%e %e\n", x, ex(ii));
        }
        fprintf(fp, "\n\n");
    }

   private:
    void energies(field_array_t& fields, double& e, double& b) {
        cabanapic::Runtime& rt = cabanapic::Runtime::get();
        rt.need_on_device(fields);
        rt.check(cpic_energies(rt.ctx(), &e, &b), "cpic_energies");
    }
};

// `step time e_energy [b_energy]` appended to energies.txt, one line per call (reference :722-763;
// the file is truncated only for step 0, which the reference's loop never produces -- delete a
// stale energies.txt before a run, as with the reference).
template <typename field_solver_t>
void dump_energies(field_solver_t& field_solver, field_array_t& fields, int step, real_t time, real_t px, real_t py, real_t pz,
                   size_t nx, size_t ny, size_t nz, size_t ng) {
    const real_t e_en = field_solver.e_energy(fields, px, py, pz, nx, ny, nz, ng);
    std::ofstream energy_file;
    if (step == 0) energy_file.open("energies.txt", std::ofstream::out | std::ofstream::trunc);
    else energy_file.open("energies.txt", std::ios::app);
    energy_file << step << " " << time << " " << e_en;
#ifndef ES_FIELD_SOLVER
    const real_t b_en = field_solver.b_energy(fields, px, py, pz, nx, ny, nz, ng);
    energy_file << " " << b_en;
    printf("%d %f %e %e\n", step, time, e_en, b_en);
#else
    printf("%d %f %e\n", step, time, e_en);
#endif
    energy_file << std::endl;
    energy_file.close();
}

#endif  // CABANAPIC_B200_FIELDS_H
