// Particle migration between z-slabs (multi-GPU slab mode; new work -- the reference is single
// process, its only traces of this are the commented neighbour logic in src/move_p.h:327-346).
//
// With the z axis not periodic inside a context, the mover leaves a particle that crossed the
// slab's low/high z face in the ghost plane z = 0 / z = nz+1 (what the reference does for any
// non-periodic boundary, src/move_p.h:257-352).  These kernels pull such particles out of the
// store into two struct-of-arrays send buffers (cell index already re-based to the receiving
// slab's numbering) and close the holes they leave, so the store stays dense.  k_pack_records /
// k_unpack_records convert between such member arrays and the record store; the host transfers
// (cpic_upload_particles / cpic_download_particles) go through them chunk by chunk as well.
//
// Send buffer layout (both buffers, capacity `cap` particles): member m (dx dy dz ux uy uz w) at
// byte offset m*cap*sizeof(R); cell (int32) at byte offset 7*cap*sizeof(R).
#pragma once
#include "cpic_common.cuh"
#include "cpic_particles.cuh"

namespace cpic {

template <class R>
struct SendBuf {
    R* m[7];
    int* cell;
};
template <class R>
inline SendBuf<R> carve_sendbuf(void* base, long long cap) {
    SendBuf<R> b;
    char* p = static_cast<char*>(base);
    for (int k = 0; k < 7; ++k) b.m[k] = reinterpret_cast<R*>(p + (size_t)k * cap * sizeof(R));
    b.cell = reinterpret_cast<int*>(p + (size_t)7 * cap * sizeof(R));
    return b;
}

__device__ __forceinline__ int z_side(int cell, int plane, int nz) {   // 0: stays, 1: low ghost, 2: high ghost
    const int iz = cell / plane;
    return iz == 0 ? 1 : (iz == nz + 1 ? 2 : 0);
}

// counters: [0] n_lo, [1] n_hi, [2] overflow flag
template <class R>
__global__ void __launch_bounds__(256) k_extract_mark(Particles<R> p, long long np, int plane, int nz, SendBuf<R> lo,
                                                      SendBuf<R> hi, long long cap, int rebase_lo, int rebase_hi,
                                                      unsigned* __restrict__ counters) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    if (n >= np) return;
    const int c = p.cell(n);
    const int side = z_side(c, plane, nz);
    if (!side) return;
    const unsigned slot = atomicAdd(counters + (side - 1), 1u);
    if (slot >= cap) { counters[2] = 1u; return; }
    SendBuf<R>& b = side == 1 ? lo : hi;
    const PRec<R> r = p.rec[n];
    b.m[0][slot] = r.pos.x; b.m[1][slot] = r.pos.y; b.m[2][slot] = r.pos.z;
    b.m[3][slot] = r.mom.x; b.m[4][slot] = r.mom.y; b.m[5][slot] = r.mom.z; b.m[6][slot] = r.mom.w;
    b.cell[slot] = c + (side == 1 ? rebase_lo : rebase_hi);
}

// After the counts are known: np_new = np - n_out.  Holes at index < np_new are filled with the
// staying particles found at index >= np_new (there are exactly as many).  lists: [0..cap) hole
// indices, [cap..2cap) donor indices; counters [3] n_holes, [4] n_donors.
template <class R>
__global__ void __launch_bounds__(256) k_extract_lists(Particles<R> p, long long np, long long np_new, int plane, int nz,
                                                       unsigned* __restrict__ lists, long long cap,
                                                       unsigned* __restrict__ counters) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    if (n >= np) return;
    const bool leaving = z_side(p.cell(n), plane, nz) != 0;
    if (n < np_new) {
        if (leaving) lists[atomicAdd(counters + 3, 1u)] = (unsigned)n;
    } else if (!leaving) {
        lists[cap + atomicAdd(counters + 4, 1u)] = (unsigned)n;
    }
}
template <class R>
__global__ void __launch_bounds__(256) k_extract_fill(Particles<R> p, const unsigned* __restrict__ lists, long long cap,
                                                      const unsigned* __restrict__ counters) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= counters[3]) return;
    const unsigned h = lists[j], d = lists[cap + j];
    p.rec[h] = p.rec[d];
}

// List-driven variants: the push left the store indices of the ghost-plane particles in `list` (PushArgs::
// leave_list), so the extraction costs O(leavers) instead of two passes over the store.
template <class R>
__global__ void __launch_bounds__(256) k_extract_mark_list(Particles<R> p, const unsigned* __restrict__ list, long long nl,
                                                           int plane, int nz, SendBuf<R> lo, SendBuf<R> hi, long long cap,
                                                           int rebase_lo, int rebase_hi, unsigned* __restrict__ counters) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= nl) return;
    const long long n = list[j];
    const PRec<R> r = p.rec[n];
    const int c = real_to_cell(r.pos.w);
    const int side = z_side(c, plane, nz);
    if (!side) { counters[5] = 1u; return; }        // cannot happen: the list holds ghost-plane particles only
    const unsigned slot = atomicAdd(counters + (side - 1), 1u);
    if (slot >= cap) { counters[2] = 1u; return; }
    SendBuf<R>& b = side == 1 ? lo : hi;
    b.m[0][slot] = r.pos.x; b.m[1][slot] = r.pos.y; b.m[2][slot] = r.pos.z;
    b.m[3][slot] = r.mom.x; b.m[4][slot] = r.mom.y; b.m[5][slot] = r.mom.z; b.m[6][slot] = r.mom.w;
    b.cell[slot] = c + (side == 1 ? rebase_lo : rebase_hi);
}
// threads [0, nl): listed particles below np_new are holes; threads [nl, nl + (np - np_new)): the staying
// particles of the tail [np_new, np) are the donors
template <class R>
__global__ void __launch_bounds__(256) k_extract_lists_list(Particles<R> p, const unsigned* __restrict__ list, long long nl,
                                                            long long np, long long np_new, int plane, int nz,
                                                            unsigned* __restrict__ lists, long long cap,
                                                            unsigned* __restrict__ counters) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j < nl) {
        const unsigned n = list[j];
        if (n < np_new) lists[atomicAdd(counters + 3, 1u)] = n;
    } else {
        const long long n = np_new + (j - nl);
        if (n < np && z_side(p.cell(n), plane, nz) == 0) lists[cap + atomicAdd(counters + 4, 1u)] = (unsigned)n;
    }
}

// Struct-of-arrays exchange buffer (SendBuf, or a staging chunk of a host transfer) <-> records.
template <class R>
__global__ void __launch_bounds__(256) k_pack_records(Particles<R> p, long long first, SendBuf<R> b, long long n) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= n) return;
    PRec<R> r;
    r.pos.x = b.m[0][j]; r.pos.y = b.m[1][j]; r.pos.z = b.m[2][j]; r.pos.w = cell_to_real(b.cell[j], R(0));
    r.mom.x = b.m[3][j]; r.mom.y = b.m[4][j]; r.mom.z = b.m[5][j]; r.mom.w = b.m[6][j];
    p.rec[first + j] = r;
}
// k_pack_records for data arriving from the host: a cell index outside [0, nc) is counted in *bad and stored
// as cell 0 (a ghost cell), so that a push of the chunk before the host has seen the count stays in bounds.
template <class R>
__global__ void __launch_bounds__(256) k_pack_records_checked(Particles<R> p, long long first, SendBuf<R> b, long long n,
                                                              long long nc, unsigned* __restrict__ bad) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= n) return;
    int c = b.cell[j];
    if (c < 0 || c >= nc) { atomicAdd(bad, 1u); c = 0; }
    PRec<R> r;
    r.pos.x = b.m[0][j]; r.pos.y = b.m[1][j]; r.pos.z = b.m[2][j]; r.pos.w = cell_to_real(c, R(0));
    r.mom.x = b.m[3][j]; r.mom.y = b.m[4][j]; r.mom.z = b.m[5][j]; r.mom.w = b.m[6][j];
    p.rec[first + j] = r;
}
template <class R>
__global__ void __launch_bounds__(256) k_unpack_records(Particles<R> p, long long first, SendBuf<R> b, long long n) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= n) return;
    const PRec<R> r = p.rec[first + j];
    b.m[0][j] = r.pos.x; b.m[1][j] = r.pos.y; b.m[2][j] = r.pos.z;
    b.m[3][j] = r.mom.x; b.m[4][j] = r.mom.y; b.m[5][j] = r.mom.z; b.m[6][j] = r.mom.w;
    b.cell[j] = real_to_cell(r.pos.w);
}

}  // namespace cpic
