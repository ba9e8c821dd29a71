// Counting sort of the particle store by cell index -- the step the reference left
// commented out (Cabana::sortByKey by Cell_Index, example/example.cpp:224-228).
// Three phases: histogram of cells, exclusive scan over cells, scatter of all 8 members
// into the second particle buffer.  Cell-sorted order is what makes the push kernel's
// interpolator loads warp broadcasts and its deposit a single warp-aggregated row update.
#pragma once
#include "cpic_common.cuh"
#include "cpic_particles.cuh"

namespace cpic {

// Also validates cell indices (the reference has no check: decks/2stream-short.cxx at HEAD
// overruns the grid silently).  bad[0] counts out-of-range cells.
__global__ void __launch_bounds__(256) k_cell_histogram(const int* __restrict__ cell, long long np, long long nc,
                                                        unsigned* __restrict__ count, unsigned* __restrict__ bad) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    if (n >= np) return;
    const int c = cell[n];
    if (c < 0 || c >= nc) { atomicAdd(bad, 1u); return; }
    atomicAdd(count + c, 1u);
}

__global__ void __launch_bounds__(256) k_check_cells(const int* __restrict__ cell, long long np, long long nc,
                                                     unsigned* __restrict__ bad) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    if (n >= np) return;
    const int c = cell[n];
    if (c < 0 || c >= nc) atomicAdd(bad, 1u);
}

// Block-wise exclusive scan of 2048 elements per block (256 threads x 8); block totals go to
// bsum for the next level.
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = 256 * SCAN_ITEMS;
__global__ void __launch_bounds__(256) k_scan_tile(const unsigned* __restrict__ in, unsigned* __restrict__ out,
                                                   long long n, unsigned* __restrict__ bsum) {
    __shared__ unsigned warp_tot[8];
    const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS], run = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        run += v[k];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    unsigned woff = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) woff += (k < w) ? warp_tot[k] : 0u;
    unsigned ex = woff + inc - run;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 255 && bsum) bsum[blockIdx.x] = woff + inc;
}
__global__ void __launch_bounds__(256) k_scan_add(unsigned* __restrict__ out, long long n,
                                                  const unsigned* __restrict__ boff) {
    const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    const unsigned add = boff[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

// Scatter: slot = cursor[cell]++ (cursor starts at the scanned offsets).  Lanes of a warp
// that share a cell claim a contiguous block with one atomic, keeping their relative order.
template <class R>
__global__ void __launch_bounds__(256) k_sort_scatter(Particles<R> src, Particles<R> dst, long long np,
                                                      unsigned* __restrict__ cursor) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = n < np;
    const int c = valid ? src.cell[n] : -1 - lane;
    const unsigned peers = __match_any_sync(0xffffffffu, c);
    const int leader = __ffs(peers) - 1;
    const int rank = __popc(peers & ((1u << lane) - 1u));
    unsigned base = 0;
    if (valid && lane == leader) base = atomicAdd(cursor + c, (unsigned)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!valid) return;
    const long long d = (long long)base + rank;
    dst.dx[d] = src.dx[n]; dst.dy[d] = src.dy[n]; dst.dz[d] = src.dz[n];
    dst.ux[d] = src.ux[n]; dst.uy[d] = src.uy[n]; dst.uz[d] = src.uz[n];
    dst.w[d] = src.w[n]; dst.cell[d] = c;
}

}  // namespace cpic
