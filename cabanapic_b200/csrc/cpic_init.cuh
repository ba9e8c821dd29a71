// Device-side particle initialiser for the synthetic uniform thermal plasma.
// The reference runs its default initialisers in the execution space as well
// (Cabana::simd_parallel_for in src/input/deck.h:155-157); decks that need host code
// (rand(), std::cout) keep using the host path + cpic_upload_particles.
//
// Counter-based RNG: Philox-4x32-10 keyed by the seed, counter = global particle index, so
// every GPU (and the numpy mirror in cabanapic_b200/decks.py, which consumes the same
// words in the same roles) can generate any slice of the global particle list independently.
#pragma once
#include "cpic_common.cuh"
#include "cpic_particles.cuh"

namespace cpic {

__device__ __forceinline__ void philox4x32_10(unsigned long long ctr, unsigned stream, unsigned long long key,
                                              unsigned (&out)[4]) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    unsigned x0 = (unsigned)ctr, x1 = (unsigned)(ctr >> 32), x2 = stream, x3 = 0u;
    unsigned k0 = (unsigned)key, k1 = (unsigned)(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, x0), lo0 = M0 * x0;
        const unsigned hi1 = __umulhi(M1, x2), lo1 = M1 * x2;
        const unsigned n0 = hi1 ^ x1 ^ k0, n2 = hi0 ^ x3 ^ k1;
        x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = x0; out[1] = x1; out[2] = x2; out[3] = x3;
}

__device__ __forceinline__ double u01(unsigned v) { return ((double)v + 0.5) * (1.0 / 4294967296.0); }

struct UniformPlasmaArgs {
    long long first, count;   // global particle indices [first, first+count)
    int gnx, gny, gnz;        // global interior grid
    int nppc;
    int z0;                   // first global interior z-plane owned by this context (slab mode), else 0
    int lnx, lny;             // local interior extents in x, y (== global)
    unsigned long long seed;
    double vth[3];
    double weight;
};

// Particle k sits in global interior cell k / nppc (x fastest): the store starts cell-sorted.
// Offsets ~ U(-1,1); momenta ~ N(0, vth) by Box-Muller, evaluated in double then narrowed.
template <class R>
__global__ void __launch_bounds__(256) k_init_uniform_plasma(Particles<R> p, UniformPlasmaArgs a) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    if (n >= a.count) return;
    const unsigned long long k = (unsigned long long)(a.first + n);
    unsigned ra[4], rb[4];
    philox4x32_10(k, 0u, a.seed, ra);
    philox4x32_10(k, 1u, a.seed, rb);
    PRec<R> r;
    r.pos.x = (R)(2.0 * u01(ra[0]) - 1.0);
    r.pos.y = (R)(2.0 * u01(ra[1]) - 1.0);
    r.pos.z = (R)(2.0 * u01(ra[2]) - 1.0);
    const double r1 = sqrt(-2.0 * log(u01(ra[3])));
    const double r2 = sqrt(-2.0 * log(u01(rb[0])));
    const double t1 = 6.283185307179586 * u01(rb[1]);
    const double t2 = 6.283185307179586 * u01(rb[2]);
    r.mom.x = (R)(a.vth[0] * r1 * cos(t1));
    r.mom.y = (R)(a.vth[1] * r1 * sin(t1));
    r.mom.z = (R)(a.vth[2] * r2 * cos(t2));
    r.mom.w = (R)a.weight;
    const long long c = (long long)(k / (unsigned long long)a.nppc);
    const int ix = (int)(c % a.gnx);
    const int iy = (int)((c / a.gnx) % a.gny);
    const int iz = (int)(c / ((long long)a.gnx * a.gny));
    r.pos.w = cell_to_real((ix + 1) + (a.lnx + 2) * ((iy + 1) + (a.lny + 2) * (iz - a.z0 + 1)), R(0));
    p.rec[n] = r;
}

// Kinetic energy diagnostic (absent upstream): out[0] += sum_p w_p (gamma_p - 1), gamma = sqrt(1 + u.u),
// evaluated as u.u / (gamma + 1) in double (no cancellation for cold particles); block tree + one atomic per block.
template <class R>
__global__ void __launch_bounds__(256) k_kinetic_energy(Particles<R> p, long long np, double* __restrict__ out) {
    double e = 0.0;
    for (long long n = blockIdx.x * 256LL + threadIdx.x; n < np; n += (long long)gridDim.x * 256) {
        const PHalf<R> m = p.rec[n].mom;
        const double u2 = (double)m.x * m.x + (double)m.y * m.y + (double)m.z * m.z;
        e += (double)m.w * (u2 / (sqrt(1.0 + u2) + 1.0));
    }
    __shared__ double se[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0) se[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += se[k];
        atomicAdd(out, t);
    }
}

// State digest (bench / multi-GPU parity block): out[0] particles, [1] sum of weights, [2] particles whose cell is not
// an interior voxel, [3] particles with an offset outside [-1,1], [4] kinetic energy -- one pass, double accumulation.
// np_dev (optional): the particle count lives on the device (slab mode).
template <class R>
__global__ void __launch_bounds__(256) k_state_digest(Particles<R> p, long long np, const long long* __restrict__ np_dev,
                                                      Grid g, double* __restrict__ out) {
    if (np_dev) np = *np_dev;
    double v[5] = {0, 0, 0, 0, 0};
    for (long long n = blockIdx.x * 256LL + threadIdx.x; n < np; n += (long long)gridDim.x * 256) {
        const PRec<R> r = p.rec[n];
        const int c = real_to_cell(r.pos.w);
        const int ix = c % g.gx, iy = (c / g.gx) % g.gy, iz = c / (g.gx * g.gy);
        const bool interior = c >= 0 && ix >= g.ng && ix < g.nx + g.ng && iy >= g.ng && iy < g.ny + g.ng && iz >= g.ng && iz < g.nz + g.ng;
        const bool inside = fabs((double)r.pos.x) <= 1.0 && fabs((double)r.pos.y) <= 1.0 && fabs((double)r.pos.z) <= 1.0;
        const double u2 = (double)r.mom.x * r.mom.x + (double)r.mom.y * r.mom.y + (double)r.mom.z * r.mom.z;
        v[0] += 1.0; v[1] += (double)r.mom.w; v[2] += interior ? 0.0 : 1.0; v[3] += inside ? 0.0 : 1.0;
        v[4] += (double)r.mom.w * (u2 / (sqrt(1.0 + u2) + 1.0));
    }
    __shared__ double se[8][5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) se[threadIdx.x >> 5][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += se[w][threadIdx.x];
        atomicAdd(out + threadIdx.x, t);
    }
}

}  // namespace cpic
