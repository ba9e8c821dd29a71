// Shared device/host helpers for the cabanapic_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <type_traits>

namespace cpic {

// Grid geometry incl. ghosts.  Voxel index = x + gx*(y + gy*z)  (reference: VOXEL(), src/types.h:195).
struct Grid {
    int nx, ny, nz, ng;
    int gx, gy, gz;      // n + 2*ng
    int sy, sz;          // strides: sy = gx, sz = gx*gy
    long long nc;        // gx*gy*gz
    int per;             // bit a set: axis a (0 x, 1 y, 2 z) is periodic inside this context; a cleared bit
                         // means the ghost layer along that axis belongs to a neighbouring slab (multi-GPU)
};

inline Grid make_grid(int nx, int ny, int nz, int ng) {
    Grid g;
    g.nx = nx; g.ny = ny; g.nz = nz; g.ng = ng;
    g.gx = nx + 2 * ng; g.gy = ny + 2 * ng; g.gz = nz + 2 * ng;
    g.sy = g.gx; g.sz = g.gx * g.gy;
    g.nc = (long long)g.gx * g.gy * g.gz;
    g.per = 7;
    return g;
}

// Interpolator record stride in reals.  18 coefficients per cell (src/types.h:64-84);
// the float record is padded to 20 (80 B) so a record is five aligned 128-bit words.
template <class R> struct IpStride;
template <> struct IpStride<float>  { static constexpr int value = 20; };
template <> struct IpStride<double> { static constexpr int value = 18; };

enum { F_EX = 0, F_EY, F_EZ, F_CBX, F_CBY, F_CBZ, F_JFX, F_JFY, F_JFZ, F_N };
enum { I_EX = 0, I_DEXDY, I_DEXDZ, I_D2EXDYDZ, I_EY, I_DEYDZ, I_DEYDX, I_D2EYDZDX, I_EZ, I_DEZDX, I_DEZDY,
       I_D2EZDXDY, I_CBX, I_DCBXDX, I_CBY, I_DCBYDY, I_CBZ, I_DCBZDZ, I_N };

// Device view of the nine field arrays (struct-of-arrays, each nc long).
template <class R>
struct Fields {
    R* c[F_N];
};

// ---- particle store: an array of records ------------------------------------------------------------
// One particle is one record of the reference's eight AoSoA members (src/types.h:31-58) in two 4-wide
// halves: pos = (dx, dy, dz, cell) and mom = (ux, uy, uz, w).  In float that is 32 bytes = exactly one
// DRAM/L2 sector, read with one 256-bit load (LDG.E.ENL2.256 on sm_100a) and -- the reason for this
// layout -- written as ONE FULL sector wherever the particle lands, which is what lets the push and the
// sort scatter particles into cell order at copy speed; the struct-of-arrays store this replaced turned
// every moved particle into eight partial-sector writes (158 B/particle of DRAM traffic for 64 useful,
// profiles/r01_push2_reorder_v1_256x256x32_ncu.md; tools/ubench/scatter_layout.cu).
// The cell index travels in the bits of pos.w (int32 in float records, int64 in double records) and is
// only ever moved, never used in arithmetic.
template <class R> struct PHalf;
template <> struct __align__(16) PHalf<float>  { float x, y, z, w; };
template <> struct __align__(32) PHalf<double> { double x, y, z, w; };
template <class R> struct PRec;
template <> struct __align__(32) PRec<float>  { PHalf<float> pos, mom; };
template <> struct __align__(64) PRec<double> { PHalf<double> pos, mom; };

__device__ __forceinline__ float  cell_to_real(int c, float)  { return __int_as_float(c); }
__device__ __forceinline__ double cell_to_real(int c, double) { return __longlong_as_double((long long)c); }
__device__ __forceinline__ int real_to_cell(float v)  { return __float_as_int(v); }
__device__ __forceinline__ int real_to_cell(double v) { return (int)__double_as_longlong(v); }

template <class R>
struct Particles {
    PRec<R>* rec;
    __device__ __forceinline__ int cell(long long n) const { return real_to_cell(rec[n].pos.w); }
    __device__ __forceinline__ void store_pos(long long n, R x, R y, R z, int c) const {
        PHalf<R> h; h.x = x; h.y = y; h.z = z; h.w = cell_to_real(c, R(0));
        rec[n].pos = h;
    }
    __device__ __forceinline__ void store_mom(long long n, R ux, R uy, R uz, R w) const {
        PHalf<R> h; h.x = ux; h.y = uy; h.z = uz; h.w = w;
        rec[n].mom = h;
    }
};

// Multiply-add under the context's floating-point policy.  This translation unit is
// compiled with -fmad=false, so `a * b + c` below really is two roundings; the fused
// form is only used when the caller asked for CPIC_FP_CONTRACT.
template <bool FMA>
__device__ __forceinline__ float madd(float a, float b, float c) {
    if constexpr (FMA) return __fmaf_rn(a, b, c); else return a * b + c;
}
template <bool FMA>
__device__ __forceinline__ double madd(double a, double b, double c) {
    if constexpr (FMA) return __fma_rn(a, b, c); else return a * b + c;
}
// a*b - c*d
template <bool FMA, class R>
__device__ __forceinline__ R mdiff(R a, R b, R c, R d) {
    if constexpr (FMA) return madd<true>(a, b, -(c * d)); else return a * b - c * d;
}

}  // namespace cpic
