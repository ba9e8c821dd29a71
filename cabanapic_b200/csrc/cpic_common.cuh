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
