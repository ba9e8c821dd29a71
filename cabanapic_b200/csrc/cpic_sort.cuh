// Counting sort of the particle store by cell index -- the step the reference left
// commented out (Cabana::sortByKey by Cell_Index, example/example.cpp:224-228).
// Three phases: histogram of cells, exclusive scan over cells, scatter of the particle records
// into the second particle buffer.  Cell-sorted order is what makes the push kernel's
// interpolator loads warp broadcasts and its deposit a single warp-aggregated row update.
#pragma once
#include "cpic_common.cuh"
#include "cpic_particles.cuh"

namespace cpic {

// Also validates cell indices (the reference has no check: decks/2stream-short.cxx at HEAD
// overruns the grid silently).  bad[0] counts out-of-range cells.
// Length of the run of equal keys that starts at this lane (0 for lanes that do not start a
// run).  Particles are nearly cell-sorted, so a warp usually holds 1-3 runs: one atomic per
// run instead of one per particle.  Keys of invalid (tail) lanes must be unique negatives.
__device__ __forceinline__ int run_length_at_head(int key, int lane, int& rank_in_run, int& head_lane) {
    const unsigned full = 0xffffffffu;
    const int prev = __shfl_up_sync(full, key, 1);
    const bool head = (lane == 0) || (key != prev);
    const unsigned heads = __ballot_sync(full, head);
    const unsigned below = heads & ((2u << lane) - 1u);          // heads at or below this lane
    head_lane = 31 - __clz(below);
    rank_in_run = lane - head_lane;
    const unsigned above = (lane == 31) ? 0u : (heads >> (lane + 1));
    const int next = above ? (lane + __ffs(above)) : 32;
    return head ? (next - lane) : 0;
}

template <class R>
__global__ void __launch_bounds__(256) k_cell_histogram(Particles<R> p, long long np, long long nc,
                                                        unsigned* __restrict__ count, unsigned* __restrict__ bad) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int c = (n < np) ? p.cell(n) : -1 - lane;
    if (n < np && (c < 0 || c >= nc)) { atomicAdd(bad, 1u); c = -1 - lane; }
    int rank, head_lane;
    const int len = run_length_at_head(c, lane, rank, head_lane);
    if (len > 0 && c >= 0) atomicAdd(count + c, (unsigned)len);
}

template <class R>
__global__ void __launch_bounds__(256) k_check_cells(Particles<R> p, long long np, long long nc,
                                                     unsigned* __restrict__ bad) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    if (n >= np) return;
    const int c = p.cell(n);
    if (c < 0 || c >= nc) atomicAdd(bad, 1u);
}

// count[cell[j]] += 1 for a short list of cells (particles appended after a migration)
__global__ void __launch_bounds__(256) k_hist_add(const int* __restrict__ cell, long long n, long long nc,
                                                  unsigned* __restrict__ count) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= n) return;
    const int c = cell[j];
    if (c >= 0 && c < nc) atomicAdd(count + c, 1u);
}

// Block-wise exclusive scan of 2048 elements per block (256 threads x 8); block totals go to
// bsum for the next level.
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = 256 * SCAN_ITEMS;
// (in and out may alias -- every launch is in place -- so neither is __restrict__)
__global__ void __launch_bounds__(256) k_scan_tile(const unsigned* in, unsigned* out, long long n, unsigned* bsum) {
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
__global__ void __launch_bounds__(256) k_scan_add(unsigned* out, long long n, const unsigned* boff) {
    const long long base = (long long)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    const unsigned add = boff[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

// Scatter: slot = cursor[cell]++ (cursor starts at the scanned offsets).  Each run of equal cells
// inside a warp claims a contiguous block with one atomic, keeping its relative order.  A record is one
// 256-bit load and one 256-bit store (float): the whole payload is in registers before the atomic's
// round trip is needed, and every particle lands as one full 32-byte sector wherever it goes.
template <class R>
__global__ void __launch_bounds__(256) k_sort_scatter(Particles<R> src, Particles<R> dst, long long np,
                                                      unsigned* __restrict__ cursor) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = n < np;
    int c = -1 - lane;
    PRec<R> r;
    if (valid) { r = src.rec[n]; c = real_to_cell(r.pos.w); }
    int rank, head_lane;
    const int len = run_length_at_head(c, lane, rank, head_lane);
    unsigned base = 0;
    if (len > 0 && valid) base = atomicAdd(cursor + c, (unsigned)len);
    base = __shfl_sync(0xffffffffu, base, head_lane);
    if (!valid) return;
    dst.rec[(long long)base + rank] = r;
}

}  // namespace cpic
