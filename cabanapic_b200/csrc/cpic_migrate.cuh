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

// ---- the same extraction with every count on the device (cpic_slab_extract_async): no host round trip, so
// the host can enqueue whole steps ahead of the device.  Grids are sized for the send capacity and loop.
// dc (device counts): [0] np  [1] error flags  [2] leavers of the last extraction (lo)  [3] (hi).
// Error bits: 1 send buffer overflow, 2 leaver list overflow / inconsistent, 4 store capacity exceeded.
template <class R>
__global__ void __launch_bounds__(256) k_extract_mark_dev(Particles<R> p, const unsigned* __restrict__ list,
                                                          const unsigned* __restrict__ nl_ptr, unsigned list_cap,
                                                          int plane, int nz, SendBuf<R> lo, SendBuf<R> hi, long long cap,
                                                          int rebase_lo, int rebase_hi, unsigned* __restrict__ counters,
                                                          long long* __restrict__ dc) {
    const unsigned nl = *nl_ptr;
    if (nl > list_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) dc[1] |= 2; return; }
    // whole warps iterate together; one counter atomic per warp and side instead of one per leaver
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    for (long long base = blockIdx.x * 256LL + (threadIdx.x & ~31); base < nl; base += gridDim.x * 256LL) {
        const long long j = base + lane;
        PRec<R> r;
        int c = 0, side = 0;
        bool stray = false;
        if (j < nl) {
            r = p.rec[list[j]];
            c = real_to_cell(r.pos.w);
            side = z_side(c, plane, nz);
            stray = side == 0;
        }
        const unsigned m1 = __ballot_sync(0xffffffffu, side == 1), m2 = __ballot_sync(0xffffffffu, side == 2);
        if (__any_sync(0xffffffffu, stray) && lane == 0) counters[5] = 1u;      // cannot happen: the list holds ghost-plane particles only
        unsigned b1 = 0, b2 = 0;
        if (lane == 0) {
            if (m1) b1 = atomicAdd(counters + 0, (unsigned)__popc(m1));
            if (m2) b2 = atomicAdd(counters + 1, (unsigned)__popc(m2));
        }
        b1 = __shfl_sync(0xffffffffu, b1, 0); b2 = __shfl_sync(0xffffffffu, b2, 0);
        if (!side) continue;
        const unsigned slot = side == 1 ? b1 + __popc(m1 & lt) : b2 + __popc(m2 & lt);
        if (slot >= cap) { counters[2] = 1u; continue; }
        SendBuf<R>& b = side == 1 ? lo : hi;
        b.m[0][slot] = r.pos.x; b.m[1][slot] = r.pos.y; b.m[2][slot] = r.pos.z;
        b.m[3][slot] = r.mom.x; b.m[4][slot] = r.mom.y; b.m[5][slot] = r.mom.z; b.m[6][slot] = r.mom.w;
        b.cell[slot] = c + (side == 1 ? rebase_lo : rebase_hi);
    }
}
// holes (listed particles below np_new) and donors (stayers of the tail [np_new, np)), as k_extract_lists_list
template <class R>
__global__ void __launch_bounds__(256) k_extract_lists_dev(Particles<R> p, const unsigned* __restrict__ list,
                                                           const unsigned* __restrict__ nl_ptr, int plane, int nz,
                                                           unsigned* __restrict__ lists, long long cap,
                                                           unsigned* __restrict__ counters, const long long* __restrict__ dc) {
    if (dc[1] || counters[2] || counters[5]) return;
    const long long nl = *nl_ptr, np = dc[0], n_out = (long long)counters[0] + counters[1], np_new = np - n_out;
    for (long long j = blockIdx.x * 256LL + threadIdx.x; j < nl + n_out; j += gridDim.x * 256LL) {
        if (j < nl) {
            const unsigned n = list[j];
            if (n < np_new) lists[atomicAdd(counters + 3, 1u)] = n;
        } else {
            const long long n = np_new + (j - nl);
            if (n < np && z_side(p.cell(n), plane, nz) == 0) lists[cap + atomicAdd(counters + 4, 1u)] = (unsigned)n;
        }
    }
}
template <class R>
__global__ void __launch_bounds__(256) k_extract_fill_dev(Particles<R> p, const unsigned* __restrict__ lists, long long cap,
                                                          const unsigned* __restrict__ counters, const long long* __restrict__ dc) {
    if (dc[1] || counters[2] || counters[5]) return;
    const long long nh = counters[3];
    for (long long j = blockIdx.x * 256LL + threadIdx.x; j < nh; j += gridDim.x * 256LL) p.rec[lists[j]] = p.rec[lists[cap + j]];
}
// one thread: new particle count, the counts the neighbours need (int64 n_lo, n_hi), error flags
__global__ void k_extract_finish_dev(const unsigned* __restrict__ counters, const unsigned* __restrict__ nl_ptr,
                                     long long* __restrict__ dc, long long* __restrict__ counts_out) {
    const long long n_lo = counters[0], n_hi = counters[1];
    if (counters[2]) dc[1] |= 1;
    if (counters[5] || n_lo + n_hi != (long long)*nl_ptr || counters[3] != counters[4]) dc[1] |= 2;
    if (dc[1]) { counts_out[0] = counts_out[1] = 0; dc[2] = dc[3] = 0; return; }
    dc[0] -= n_lo + n_hi;
    dc[2] = n_lo; dc[3] = n_hi;
    counts_out[0] = n_lo; counts_out[1] = n_hi;
}
// arrivals: *m_ptr particles of a capacity-`cap` exchange buffer go to the end of the store; the cell histogram
// of the last push is kept current.  (k_append_finish_dev then advances the count.)
template <class R>
__global__ void __launch_bounds__(256) k_append_dev(Particles<R> p, SendBuf<R> b, long long cap, const long long* __restrict__ m_ptr,
                                                    long long store_cap, long long nc, unsigned* __restrict__ hist,
                                                    long long* __restrict__ dc) {
    if (dc[1]) return;      // an earlier step of this migration flagged an error: leave the store alone (reported by sync_np)
    const long long m = *m_ptr, np = dc[0];
    if (m < 0 || m > cap || np + m > store_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) dc[1] |= 4; return; }
    for (long long j = blockIdx.x * 256LL + threadIdx.x; j < m; j += gridDim.x * 256LL) {
        const int c = b.cell[j];
        PRec<R> r;
        r.pos.x = b.m[0][j]; r.pos.y = b.m[1][j]; r.pos.z = b.m[2][j]; r.pos.w = cell_to_real(c, R(0));
        r.mom.x = b.m[3][j]; r.mom.y = b.m[4][j]; r.mom.z = b.m[5][j]; r.mom.w = b.m[6][j];
        p.rec[np + j] = r;
        if (hist && c >= 0 && c < nc) atomicAdd(hist + c, 1u);
    }
}
__global__ void k_append_finish_dev(const long long* __restrict__ m_ptr, long long* __restrict__ dc) {
    if (!dc[1]) dc[0] += *m_ptr;
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
// k_pack_records for data arriving from the host: a cell index outside [lo, hi) -- the whole grid, or a slab's interior
// planes -- is counted in *bad and stored as cell lo (a ghost / edge cell inside the grid), so that a push of the chunk
// before the host has seen the count stays in bounds.
template <class R>
__global__ void __launch_bounds__(256) k_pack_records_checked(Particles<R> p, long long first, SendBuf<R> b, long long n,
                                                              long long lo, long long hi, unsigned* __restrict__ bad) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= n) return;
    int c = b.cell[j];
    if (c < lo || c >= hi) { atomicAdd(bad, 1u); c = (int)lo; }
    PRec<R> r;
    r.pos.x = b.m[0][j]; r.pos.y = b.m[1][j]; r.pos.z = b.m[2][j]; r.pos.w = cell_to_real(c, R(0));
    r.mom.x = b.m[3][j]; r.mom.y = b.m[4][j]; r.mom.z = b.m[5][j]; r.mom.w = b.m[6][j];
    p.rec[first + j] = r;
}
// records at listed store positions -> member arrays (the patches of cpic_mgpu_step_host)
template <class R>
__global__ void __launch_bounds__(256) k_gather_records(Particles<R> p, const unsigned* __restrict__ list, SendBuf<R> b, long long n) {
    const long long j = blockIdx.x * 256LL + threadIdx.x;
    if (j >= n) return;
    const PRec<R> r = p.rec[list[j]];
    b.m[0][j] = r.pos.x; b.m[1][j] = r.pos.y; b.m[2][j] = r.pos.z;
    b.m[3][j] = r.mom.x; b.m[4][j] = r.mom.y; b.m[5][j] = r.mom.z; b.m[6][j] = r.mom.w;
    b.cell[j] = real_to_cell(r.pos.w);
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
