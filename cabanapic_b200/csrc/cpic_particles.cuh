// Particle kernels: Boris push + VPIC-style cell-crossing mover + current deposit, and
// the one-off uncenter rotation.
//
// Reference: push<>, src/push.h:65-295; move_p<>, src/move_p.h:59-374 (with
// detect_leaving_domain :9-54); uncenter_particles, src/uncenter_p.h:27-98.
//
// Layout: particles are 32-byte (float) / 64-byte (double) records (PRec, cpic_common.cuh), one
// thread per particle here, so a warp reads 32 consecutive records with one 256-bit load per lane
// (two in double) and writes the two halves it changes.  The interpolator record of the particle's
// cell is fetched with 128-bit loads through the read-only path (cell-sorted particles make these
// warp-wide broadcasts of one record).
//
// Deposit: every streak adds 12 values (jx[4] jy[4] jz[4]) to its cell's accumulator row.
// The first streak of every particle -- the only one for the ~87 % that stay in their
// cell -- goes to the particle's *current* cell, so for cell-sorted particles a whole
// warp targets one row: the warp transposes-and-reduces the 12x32 values with 16 shuffles
// and issues ONE 12-lane atomic row update instead of 384 atomics.  Later streaks of cell
// crossers (rare, scattered over the 6 neighbours) use 128-bit vector atomics directly.
#pragma once
#include "cpic_common.cuh"

namespace cpic {

template <class R>
struct PushArgs {
    Particles<R> p;
    long long np;
    const R* ip;   // [nc][IpStride]
    R* acc;        // [nc][12]
    R qdt_2mc, cdt_dx, cdt_dy, cdt_dz, qsp;
    int nx, ny, nz, ng, gx, gy;
    unsigned magic_gx, magic_gy;   // ceil(2^32 / gx), ceil(2^32 / gy)
    int periodic;     // bit a: wrap along axis a (0 x, 1 y, 2 z); cleared: the particle stays in the ghost cell
    int reflect;      // bit a: the domain faces of axis a reflect particles (Boundary::Reflect, src/move_p.h:298-324)
    int dep_thresh;   // mixed warps: runs at least this long are warp-reduced, shorter ones use direct atomics
    int dep_rounds;   // mixed warps: at most this many peel rounds before falling back to direct atomics
    unsigned long long* stats;  // optional: [0] movers [1] crossings [2..7] wraps per face
    unsigned* hist;             // optional: per-cell count of the particles' NEW cells (feeds the next sort)
    // reordering push (k_push2<..., REORD>): the advanced particles are written to `dst`, each to a slot of
    // the segment of the cell it occupied when the step began (cursor = exclusive scan of that histogram)
    Particles<R> dst;
    unsigned* cursor;
    // slab mode (z not periodic inside this context): the mover appends the store index of every particle
    // it leaves in a z ghost plane (cell < leave_lo or cell >= leave_hi) to leave_list, so that the migration
    // that follows touches only those instead of scanning the whole store
    unsigned* leave_list;     // optional
    unsigned* leave_count;
    unsigned leave_cap;
    unsigned leave_off;        // added to the listed indices (a push over a sub-range of the store: cpic_step_host)
    int leave_lo, leave_hi;
    // CPIC_DEPOSIT_ORDERED (k_push<DEPOSIT = 4>): nothing is added during the push; streak j of particle n writes its cell
    // and its 12 currents to event n * ev_k + j, and k_ordered_accumulate adds the events of every cell in (particle,
    // streak) order afterwards -- the order of the reference's serial loop
    int* ev_cell;              // [np * ev_k], -1 = unused
    R* ev_row;                 // [np * ev_k][12]
    int ev_k;
    unsigned* ev_overflow;     // set when a particle has more than ev_k streaks
    int priv_nc;               // > 0: k_push2<PRIV> keeps a block-private accumulator (+ histogram) of this many cells in shared memory
    const long long* np_dev;   // optional: the particle count lives on the device (overrides np; k_push2 only)
};

// ---------------------------------------------------------------------------------------
// The 12 quadrant currents of one streak.  Reference: CALC_J, src/push.h:218-232 and
// accumulate_j, src/move_p.h:156-170 (identical arithmetic).  d = streak midpoint,
// u = half displacement in cell units, v5 = the q*ux*uy*uz/3 correction.
template <bool FMA, class R>
__device__ __forceinline__ void streak_currents(R q, R ux, R uy, R uz, R dx, R dy, R dz, R v5, R (&a)[12]) {
    const R one = R(1);
#define CPIC_QUAD(U, DA, DB, O)                                   \
    {                                                             \
        R v0, v1, v2, v3, v4;                                     \
        v4 = q * (U);                                             \
        if constexpr (FMA) {                                      \
            v0 = madd<true>(-v4, (DA), v4);                       \
            v1 = madd<true>(v4, (DA), v4);                        \
            const R hi = one + (DB), lo = one - (DB);             \
            v2 = madd<true>(v0, hi, -v5);                         \
            v3 = madd<true>(v1, hi, v5);                          \
            v0 = madd<true>(v0, lo, v5);                          \
            v1 = madd<true>(v1, lo, -v5);                         \
        } else {                                                  \
            v1 = v4 * (DA);                                       \
            v0 = v4 - v1;                                         \
            v1 += v4;                                             \
            v4 = one + (DB);                                      \
            v2 = v0 * v4;                                         \
            v3 = v1 * v4;                                         \
            v4 = one - (DB);                                      \
            v0 *= v4;                                             \
            v1 *= v4;                                             \
            v0 += v5;                                             \
            v1 -= v5;                                             \
            v2 -= v5;                                             \
            v3 += v5;                                             \
        }                                                         \
        a[(O) + 0] = v0; a[(O) + 1] = v1; a[(O) + 2] = v2; a[(O) + 3] = v3; \
    }
    CPIC_QUAD(ux, dy, dz, 0)
    CPIC_QUAD(uy, dz, dx, 4)
    CPIC_QUAD(uz, dx, dy, 8)
#undef CPIC_QUAD
}

// ---------------------------------------------------------------------------------------
// Accumulator row updates.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}
template <class R>
__device__ __forceinline__ void row_add_scalar(R* row, const R (&a)[12]) {
#pragma unroll
    for (int j = 0; j < 12; ++j) atomicAdd(row + j, a[j]);
}
__device__ __forceinline__ void row_add_vec(float* row, const float (&a)[12]) {
    red_add_v4(row + 0, a[0], a[1], a[2], a[3]);
    red_add_v4(row + 4, a[4], a[5], a[6], a[7]);
    red_add_v4(row + 8, a[8], a[9], a[10], a[11]);
}
__device__ __forceinline__ void row_add_vec(double* row, const double (&a)[12]) { row_add_scalar(row, a); }
// Deposit target of k_push2: the global accumulator (one 128-bit reduction), or -- PRIV, grids of a few hundred
// cells where every reduction of the machine would hit the same handful of L2 lines -- a block-private copy in
// shared memory that is added to the global one once, when the block retires.
template <bool PRIV>
__device__ __forceinline__ void acc_add4(float* __restrict__ gacc, float* sacc, int c, int eg, float x, float y, float z, float w) {
    if constexpr (PRIV) {
        float* p = sacc + c * 12 + eg * 4;
        atomicAdd(p + 0, x); atomicAdd(p + 1, y); atomicAdd(p + 2, z); atomicAdd(p + 3, w);
    } else {
        red_add_v4(gacc + (long long)c * 12 + eg * 4, x, y, z, w);
    }
}
template <bool PRIV>
__device__ __forceinline__ void hist_add(unsigned* __restrict__ ghist, unsigned* shist, int c, unsigned v) {
    if constexpr (PRIV) atomicAdd(shist + c, v); else atomicAdd(ghist + c, v);
}

// Sum a[0..11] over the lanes in `mask`-selected contributions (others pass zeros) with a
// transpose-reduce: after the 5 exchange stages lane L holds the warp total of entry
// e(L) = 8*b4 + 4*b3 + 2*b2 + b1 (b_k = bit k of L); lanes with bit 0 set hold duplicates.
template <class R>
__device__ __forceinline__ R warp_transpose_sum12(const R (&a)[12], int lane) {
    const unsigned full = 0xffffffffu;
    R r8[8], r4[4], r2[2], r1;
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const R lo = a[j], hi = (j + 8 < 12) ? a[j + 8] : R(0);
        const R keep = h4 ? hi : lo, send = h4 ? lo : hi;
        r8[j] = keep + __shfl_xor_sync(full, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const R keep = h3 ? r8[j + 4] : r8[j], send = h3 ? r8[j] : r8[j + 4];
        r4[j] = keep + __shfl_xor_sync(full, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const R keep = h2 ? r4[j + 2] : r4[j], send = h2 ? r4[j] : r4[j + 2];
        r2[j] = keep + __shfl_xor_sync(full, send, 4);
    }
    {
        const R keep = h1 ? r2[1] : r2[0], send = h1 ? r2[0] : r2[1];
        r1 = keep + __shfl_xor_sync(full, send, 2);
    }
    r1 += __shfl_xor_sync(full, r1, 1);
    return r1;
}

// CPIC_DEPOSIT_ORDERED: record streak j of particle n
template <class R>
__device__ __forceinline__ void record_event(const PushArgs<R>& a, long long n, int j, int cell, const R (&cur)[12]) {
    if (j >= a.ev_k) { *a.ev_overflow = 1u; return; }
    const long long e = n * a.ev_k + j;
    a.ev_cell[e] = cell;
    R* row = a.ev_row + e * 12;
#pragma unroll
    for (int k = 0; k < 12; ++k) row[k] = cur[k];
}
// one thread per (cell, entry): the events of the cell are contiguous in `order` (stable sort of the event indices by
// cell), in (particle, streak) order; lo / hi by binary search in the sorted keys.  acc[c][k] = (..((acc + e0) + e1) + ..)
template <class R>
__global__ void __launch_bounds__(256) k_ordered_accumulate(const unsigned* __restrict__ keys, const unsigned* __restrict__ order,
                                                            long long m, const R* __restrict__ ev_row, R* __restrict__ acc, long long nc) {
    const long long t = blockIdx.x * 256LL + threadIdx.x;
    if (t >= nc * 12) return;
    const unsigned c = (unsigned)(t / 12);
    const int k = (int)(t % 12);
    long long lo = 0, hi = m;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (keys[mid] < c) lo = mid + 1; else hi = mid; }
    R s = acc[t];
    for (long long i = lo; i < m && keys[i] == c; ++i) s += ev_row[(long long)order[i] * 12 + k];
    acc[t] = s;
}
__global__ void __launch_bounds__(256) k_iota(unsigned* v, long long n) {
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i < n) v[i] = (unsigned)i;
}

// First-streak deposit; must be called by all 32 lanes of the warp (converged).
template <class R, int DEPOSIT>
__device__ __forceinline__ void deposit_first(R* __restrict__ acc, bool valid, int ii, const R (&a)[12], int lane,
                                              int thresh, int rounds) {
    if constexpr (DEPOSIT == 1) {
        if (valid) row_add_scalar(acc + (long long)ii * 12, a);
    } else if constexpr (DEPOSIT == 2) {
        if (valid) row_add_vec(acc + (long long)ii * 12, a);
    } else if constexpr (DEPOSIT == 4) {
        // (ordered mode: the caller records the event instead)
    } else {
        const unsigned full = 0xffffffffu;
        const unsigned vmask = __ballot_sync(full, valid);
        if (vmask == 0) return;
        const int leader = __ffs(vmask) - 1;
        const int c0 = __shfl_sync(full, ii, leader);
        unsigned same = __ballot_sync(full, valid && ii == c0);
        const int e = ((lane >> 1) & 7) | ((lane & 16) >> 1);  // entry this lane ends up holding
        if (same == vmask) {  // the common case for cell-sorted particles: one row per warp
            // (lanes that do not contribute pass all-zero currents, so no masking is needed)
            const R tot = warp_transpose_sum12(a, lane);
            if (!(lane & 1) && e < 12) atomicAdd(acc + (long long)c0 * 12 + e, tot);
            return;
        }
        // Mixed warp: peel off one cell at a time (sorted particles give 2-3 runs per warp).  A
        // run that is long enough is summed with the warp transpose; short runs and whatever is
        // left after four rounds (unsorted input) update their rows directly.
        unsigned todo = vmask;
        bool mine_done = !valid;
        for (int round = 0; todo && round < rounds; ++round) {
            const int ld = __ffs(todo) - 1;
            const int c = __shfl_sync(full, ii, ld);
            const bool mine = !mine_done && ii == c;
            const unsigned grp = __ballot_sync(full, mine);
            if (__popc(grp) >= thresh) {
                R z[12];
#pragma unroll
                for (int j = 0; j < 12; ++j) z[j] = mine ? a[j] : R(0);
                const R tot = warp_transpose_sum12(z, lane);
                if (!(lane & 1) && e < 12) atomicAdd(acc + (long long)c * 12 + e, tot);
            } else if (mine) {
                row_add_vec(acc + (long long)ii * 12, a);
            }
            if (mine) mine_done = true;
            todo &= ~grp;
        }
        if (!mine_done) row_add_vec(acc + (long long)ii * 12, a);
    }
}

// ---------------------------------------------------------------------------------------
// One streak of the mover.  Reference: src/move_p.h:101-137,154,195-203.  On entry (x,y,z)
// is the stored position and (rx,ry,rz) the remaining half displacement; on exit the
// position has advanced by twice the streak, the remainder has shrunk, (sx..,mx..,v5) describe
// the streak, and the return value is the axis whose face ended it (3 = end of track).
// The bare literals 3.4e38, 2, 0.5 and (1./3.) of the reference are doubles/ints mixed into
// real_t arithmetic; only the last one changes a rounding (final multiply in double).
template <class R>
__device__ __forceinline__ int mover_streak(R& x, R& y, R& z, R& rx, R& ry, R& rz, R q, R& sx, R& sy, R& sz,
                                            R& mx, R& my, R& mz, R& v5, R& dirv) {
    sx = rx; sy = ry; sz = rz;
    const R d0 = (sx > 0) ? R(1) : R(-1);
    const R d1 = (sy > 0) ? R(1) : R(-1);
    const R d2 = (sz > 0) ? R(1) : R(-1);
    const R big = R(3.4e38);
    const R v0 = (sx == 0) ? big : (d0 - x) / sx;
    const R v1 = (sy == 0) ? big : (d1 - y) / sy;
    const R v2 = (sz == 0) ? big : (d2 - z) / sz;
    R v3 = R(2);
    int axis = 3;
    if (v0 < v3) { v3 = v0; axis = 0; }
    if (v1 < v3) { v3 = v1; axis = 1; }
    if (v2 < v3) { v3 = v2; axis = 2; }
    v3 *= R(0.5);
    sx *= v3; sy *= v3; sz *= v3;
    mx = x + sx; my = y + sy; mz = z + sz;
    v5 = (R)((double)(q * sx * sy * sz) * (1. / 3.));
    rx -= sx; ry -= sy; rz -= sz;
    x += sx + sx; y += sy + sy; z += sz + sz;
    dirv = (axis == 0) ? d0 : (axis == 1 ? d1 : d2);
    return axis;
}

// Cell update when a streak ended on a face.  Reference: src/move_p.h:218-244 (neighbour),
// :9-54 (ghost-layer test, later tests override earlier, one ghost layer assumed),
// :257-288 (periodic wrap), :351-352 (new voxel).  Returns the face code 0..5 (-x -y -z +x +y +z)
// in the low bits and, if the neighbour was a ghost layer, 8 + that layer's code above them.
// Boundary::Reflect (a.reflect; the reference keeps VPIC's block in comments, :298-324): when the face is a domain
// face of a reflecting axis the particle keeps its cell and the CROSS_REFLECTED bit is set -- the caller then leaves the
// position on the face and reverses the momentum component and the remaining displacement along that axis.
constexpr int CROSS_REFLECTED = 8;
template <class R>
__device__ __forceinline__ int cross_face(int& ii, int axis, R dirv, const PushArgs<R>& a) {
    int face = axis;
    if (dirv > 0) face += 3;
    // RANK_TO_INDEX (src/types.h:184-193) with the two divisions done by multiply-high with the
    // host-computed ceil(2^32/d) and one correction step (exact for 0 <= ii < 2^31, d < 2^16)
    int iy = (int)__umulhi((unsigned)ii, a.magic_gx);
    if (iy * a.gx > ii) --iy;
    int ix = ii - iy * a.gx;
    int iz = (int)__umulhi((unsigned)iy, a.magic_gy);
    if (iz * a.gy > iy) --iz;
    iy -= iz * a.gy;
    if (face == 0) ix--;
    if (face == 1) iy--;
    if (face == 2) iz--;
    if (face == 3) ix++;
    if (face == 4) iy++;
    if (face == 5) iz++;
    if (a.reflect & (1 << axis)) {
        const int coord = axis == 0 ? ix : (axis == 1 ? iy : iz), top = axis == 0 ? a.nx : (axis == 1 ? a.ny : a.nz);
        if (coord == 0 || coord == top + 1) return face | CROSS_REFLECTED;
    }
    // detect_leaving_domain: later tests override earlier ones.  a.periodic is a per-axis mask
    // (7 = the reference); an axis whose bit is clear (slab mode: that ghost layer belongs to a
    // neighbour) neither wraps nor hides the wrap of another axis, so its tests are skipped.
    const int per = a.periodic;
    int leaving = -1;
    if ((per & 1) && ix == 0) leaving = 0;
    if ((per & 2) && iy == 0) leaving = 1;
    if ((per & 4) && iz == 0) leaving = 2;
    if ((per & 1) && ix == a.nx + 1) leaving = 3;
    if ((per & 2) && iy == a.ny + 1) leaving = 4;
    if ((per & 4) && iz == a.nz + 1) leaving = 5;
    if (leaving >= 0) {
        if (leaving == 0) ix = (a.nx - 1) + a.ng;
        else if (leaving == 1) iy = (a.ny - 1) + a.ng;
        else if (leaving == 2) iz = (a.nz - 1) + a.ng;
        else if (leaving == 3) ix = a.ng;
        else if (leaving == 4) iy = a.ng;
        else iz = a.ng;
    }
    ii = ix + a.gx * (iy + a.gy * iz);
    return face | (leaving >= 0 ? ((8 + leaving) << 4) : 0);
}

// Load the (padded) interpolator record of cell ii with 128-bit read-only loads.
template <class R>
__device__ __forceinline__ void load_record(const R* __restrict__ ip, int ii, R (&f)[IpStride<R>::value]) {
    constexpr int S = IpStride<R>::value;
    using V = typename std::conditional<sizeof(R) == 4, float4, double2>::type;
    constexpr int PER = 16 / sizeof(R);
    const V* src = reinterpret_cast<const V*>(ip + (long long)ii * S);
#pragma unroll
    for (int k = 0; k < S / PER; ++k) {
        const V v = __ldg(src + k);
        *reinterpret_cast<V*>(&f[k * PER]) = v;
    }
}

// ---------------------------------------------------------------------------------------
// The push kernel: persistent and warp-autonomous.
//
// The grid is sized to fill the machine once (SMs x resident blocks); every warp walks the
// particle store in 32-particle tiles with a grid stride, so consecutive warps touch
// consecutive 128-byte lines of each member array and no block barrier or block launch is on
// the per-tile path.  Per tile:
//   1. each lane advances one particle (gather, Boris, displacement);
//   2. particles that stay in their cell (~87 %) store their new position and put their 12
//      currents into the warp-aggregated row update (deposit_first);
//   3. particles that leave their cell are NOT moved yet: the lane appends a compact mover
//      record to the warp's private shared-memory list -- the in-kernel form of VPIC's
//      particle_mover_t list that the reference kept only in comments (src/push.h:271-291,
//      src/types.h:175-179).  Whenever the list holds a full warp's worth it is drained
//      densely, one mover per lane, so the divergent cell-crossing loop (IEEE divides, index
//      arithmetic, 1-4 streaks) runs with 32 active lanes instead of 3-4.
// The next tile's eight member loads are issued before the current tile's arithmetic
// (PREFETCH) so DRAM latency overlaps compute within the warp, not just across warps.
#ifndef PUSH_MIN_BLOCKS
#define PUSH_MIN_BLOCKS 4
#endif
constexpr int PUSH_WARPS = 8;          // warps per block
constexpr int MOVER_CAP = 64;          // per-warp list capacity (drained at >= 32)

template <class R, int CAP = MOVER_CAP>
struct WarpMoverList {
    R x[CAP], y[CAP], z[CAP], rx[CAP], ry[CAP], rz[CAP], q[CAP];
    int cell[CAP];
    unsigned idx[CAP];   // global particle index (< 2^31)
};

// the reordering push of k_push3 writes a mover's whole record from the drain, so its list also carries the new
// momentum and the weight
template <class R, int CAP = MOVER_CAP>
struct WarpMoverListP : WarpMoverList<R, CAP> {
    R ux[CAP], uy[CAP], uz[CAP], w[CAP];
};

// Drain list entries [first, first+32) (lanes beyond `count` idle).  Reference: move_p,
// src/move_p.h:93-371 -- streak, deposit into the current cell, then either stop (end of
// track) or cross the face into the neighbour and continue.
// OUTOFPLACE: the list's idx are slots of a.dst (reordering push) and the cell is always written.
template <class R, bool FMA, int DEPOSIT, bool STATS, class List, bool OUTOFPLACE = false, bool PRIV = false>
__device__ __forceinline__ void drain_movers(const PushArgs<R>& a, List& ml, int first, int count,
                                             int lane, unsigned long long& n_cross, unsigned long long (&n_wrap)[6],
                                             float* sacc = nullptr, unsigned* shist = nullptr) {
    const int m = first + lane;
    bool leaves = false;
    unsigned leaver = 0;
    if (lane < count) {
        R px = ml.x[m], py = ml.y[m], pz = ml.z[m];
        R dx = ml.rx[m], dy = ml.ry[m], dz = ml.rz[m];
        const R qq = ml.q[m];
        int c = ml.cell[m];
        unsigned flip = 0;
        int streak = 0;
        for (;;) {
            R sx, sy, sz, mx, my, mz, v5, dirv;
            const int axis = mover_streak(px, py, pz, dx, dy, dz, qq, sx, sy, sz, mx, my, mz, v5, dirv);
            R jc[12];
            streak_currents<FMA>(qq, sx, sy, sz, mx, my, mz, v5, jc);
            if constexpr (DEPOSIT == 4) {
                record_event(a, (long long)ml.idx[m], streak++, c, jc);
            } else if constexpr (PRIV) {
                acc_add4<true>(nullptr, sacc, c, 0, (float)jc[0], (float)jc[1], (float)jc[2], (float)jc[3]);
                acc_add4<true>(nullptr, sacc, c, 1, (float)jc[4], (float)jc[5], (float)jc[6], (float)jc[7]);
                acc_add4<true>(nullptr, sacc, c, 2, (float)jc[8], (float)jc[9], (float)jc[10], (float)jc[11]);
            } else if constexpr (DEPOSIT == 1) row_add_scalar(a.acc + (long long)c * 12, jc);
            else row_add_vec(a.acc + (long long)c * 12, jc);
            if (axis == 3) break;
            // snap onto the face, move to the neighbour, re-enter from its other side
            const int code = cross_face(c, axis, dirv, a);
            if (code & CROSS_REFLECTED) {      // reflecting wall: stay on the face, turn around
                if (axis == 0) { px = dirv; dx = -dx; }
                if (axis == 1) { py = dirv; dy = -dy; }
                if (axis == 2) { pz = dirv; dz = -dz; }
                flip ^= 1u << axis;
            } else {
                if (axis == 0) px = -dirv;
                if (axis == 1) py = -dirv;
                if (axis == 2) pz = -dirv;
            }
            if (STATS) {
                ++n_cross;
                if (code >> 4) ++n_wrap[(code >> 4) - 8];
            }
        }
        const long long pn = ml.idx[m];
        if (flip) {      // the momentum half was stored by the main path: reverse the reflected components in place
            PRec<R>* rec = OUTOFPLACE ? a.dst.rec : a.p.rec;
            PHalf<R> mo = rec[pn].mom;
            if (flip & 1u) mo.x = -mo.x;
            if (flip & 2u) mo.y = -mo.y;
            if (flip & 4u) mo.z = -mo.z;
            rec[pn].mom = mo;
        }
        leaves = a.leave_list && (c < a.leave_lo || c >= a.leave_hi);
        leaver = (unsigned)pn;
        if constexpr (OUTOFPLACE) {
            a.dst.store_pos(pn, px, py, pz, c);
            hist_add<PRIV>(a.hist, shist, c, 1u);
        } else {
            a.p.store_pos(pn, px, py, pz, c);
            if (a.hist) hist_add<PRIV>(a.hist, shist, c, 1u);
        }
    }
    __syncwarp();
    if (a.leave_list) {      // slab mode: list the particles left in a z ghost plane, one counter atomic per warp
        const unsigned lm = __ballot_sync(0xffffffffu, leaves);
        if (lm) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(a.leave_count, (unsigned)__popc(lm));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned j = base + __popc(lm & ((1u << lane) - 1u));
            if (leaves && j < a.leave_cap) a.leave_list[j] = leaver + a.leave_off;
        }
    }
}

template <class R, bool FMA, int DEPOSIT, bool STATS, bool PREFETCH>
__global__ void __launch_bounds__(PUSH_WARPS * 32, (sizeof(R) == 4 ? PUSH_MIN_BLOCKS : (PUSH_MIN_BLOCKS + 1) / 2))
k_push(PushArgs<R> a) {
    __shared__ WarpMoverList<R> lists[PUSH_WARPS];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpMoverList<R>& ml = lists[warp];
    const long long ntiles = (a.np + 31) / 32;
    const long long stride = (long long)gridDim.x * PUSH_WARPS;
    const R one = R(1.), one_third = R(1. / 3.), two_fifteenths = R(2. / 15.);
    int nlist = 0;                                    // movers waiting in this warp's list
    unsigned long long n_mov = 0, n_cross = 0, n_wrap[6] = {0, 0, 0, 0, 0, 0};

    long long tile = (long long)blockIdx.x * PUSH_WARPS + warp;
    // registers for the tile being processed (+ the next one when prefetching)
    int ii = 0;
    R x = 0, y = 0, z = 0, ux = 0, uy = 0, uz = 0, w = 0;
    if (tile < ntiles) {
        const long long n = tile * 32 + lane;
        if (n < a.np) {
            const PRec<R> r = a.p.rec[n];
            ii = real_to_cell(r.pos.w); x = r.pos.x; y = r.pos.y; z = r.pos.z;
            ux = r.mom.x; uy = r.mom.y; uz = r.mom.z; w = r.mom.w;
        }
    }
    for (; tile < ntiles; tile += stride) {
        const long long n = tile * 32 + lane;
        const bool valid = n < a.np;
        // next tile's members: issue the loads now, consume them next iteration
        int ii_n = 0;
        R x_n = 0, y_n = 0, z_n = 0, ux_n = 0, uy_n = 0, uz_n = 0, w_n = 0;
        if (PREFETCH) {
            const long long nn = (tile + stride) * 32 + lane;
            if (nn < a.np) {
                const PRec<R> r = a.p.rec[nn];
                ii_n = real_to_cell(r.pos.w); x_n = r.pos.x; y_n = r.pos.y; z_n = r.pos.z;
                ux_n = r.mom.x; uy_n = r.mom.y; uz_n = r.mom.z; w_n = r.mom.w;
            }
        }

        bool mover = false, stay = false;
        R rx = 0, ry = 0, rz = 0, q = 0;
        R cur[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) cur[j] = R(0);

        if (valid) {
            R f[IpStride<R>::value];
            load_record(a.ip, ii, f);
            q = w * a.qsp;
            // src/push.h:124-138 -- trilinear E, linear B at the particle
            const R hax = a.qdt_2mc * madd<FMA>(z, madd<FMA>(y, f[I_D2EXDYDZ], f[I_DEXDZ]), madd<FMA>(y, f[I_DEXDY], f[I_EX]));
            const R hay = a.qdt_2mc * madd<FMA>(x, madd<FMA>(z, f[I_D2EYDZDX], f[I_DEYDX]), madd<FMA>(z, f[I_DEYDZ], f[I_EY]));
            const R haz = a.qdt_2mc * madd<FMA>(y, madd<FMA>(x, f[I_D2EZDXDY], f[I_DEZDY]), madd<FMA>(x, f[I_DEZDX], f[I_EZ]));
            const R cbx = madd<FMA>(x, f[I_DCBXDX], f[I_CBX]);
            const R cby = madd<FMA>(y, f[I_DCBYDY], f[I_CBY]);
            const R cbz = madd<FMA>(z, f[I_DCBZDZ], f[I_CBZ]);

            ux += hax; uy += hay; uz += haz;                                  // half E kick, :144-146
            // :148 -- sqrtf even when real_t is double (argument rounds to float first)
            R v0 = a.qdt_2mc / (R)sqrtf((float)(one + madd<FMA>(ux, ux, madd<FMA>(uy, uy, uz * uz))));
            R v1 = madd<FMA>(cbx, cbx, madd<FMA>(cby, cby, cbz * cbz));
            R v2 = (v0 * v0) * v1;
            R v3 = v0 * madd<FMA>(v2, madd<FMA>(v2, two_fifteenths, one_third), one);
            R v4 = v3 / madd<FMA>(v1, v3 * v3, one);
            v4 += v4;
            v0 = madd<FMA>(v3, mdiff<FMA>(uy, cbz, uz, cby), ux);              // Boris u', :155-157
            v1 = madd<FMA>(v3, mdiff<FMA>(uz, cbx, ux, cbz), uy);
            v2 = madd<FMA>(v3, mdiff<FMA>(ux, cby, uy, cbx), uz);
            ux = madd<FMA>(v4, mdiff<FMA>(v1, cbz, v2, cby), ux);              // rotation, :158-160
            uy = madd<FMA>(v4, mdiff<FMA>(v2, cbx, v0, cbz), uy);
            uz = madd<FMA>(v4, mdiff<FMA>(v0, cby, v1, cbx), uz);
            ux += hax; uy += hay; uz += haz;                                  // second half kick
            a.p.store_mom(n, ux, uy, uz, w);                                  // :165-167

            v0 = one / (R)sqrtf((float)(one + madd<FMA>(ux, ux, madd<FMA>(uy, uy, uz * uz))));  // :169
            ux *= a.cdt_dx; uy *= a.cdt_dy; uz *= a.cdt_dz;                   // this order, :171-176
            ux *= v0; uy *= v0; uz *= v0;
            v0 = x + ux; v1 = y + uy; v2 = z + uz;                            // streak midpoint
            v3 = v0 + ux; v4 = v1 + uy;                                       // new position
            const R v5n = v2 + uz;

            if (v3 <= one && v4 <= one && v5n <= one && -v3 <= one && -v4 <= one && -v5n <= one) {  // :187
                stay = true;
                a.p.store_pos(n, v3, v4, v5n, ii);
                const R v5 = q * ux * uy * uz * one_third;                    // :203
                streak_currents<FMA>(q, ux, uy, uz, v0, v1, v2, v5, cur);
            } else {
                mover = true;
                rx = ux; ry = uy; rz = uz;                                    // local_pm, :261-263
            }
        }

        // in-cell particles: one streak into the particle's own cell
        if constexpr (DEPOSIT == 4) { if (stay) record_event(a, n, 0, ii, cur); }
        else deposit_first<R, DEPOSIT>(a.acc, stay, ii, cur, lane, a.dep_thresh, a.dep_rounds);

        // movers: append to the warp's list (warp-synchronous, no atomics)
        const unsigned mm = __ballot_sync(0xffffffffu, mover);
        if (mm) {
            if (mover) {
                const int m = nlist + __popc(mm & ((1u << lane) - 1u));
                ml.x[m] = x; ml.y[m] = y; ml.z[m] = z;
                ml.rx[m] = rx; ml.ry[m] = ry; ml.rz[m] = rz;
                ml.q[m] = q; ml.cell[m] = ii; ml.idx[m] = (unsigned)n;
            }
            nlist += __popc(mm);
            if (STATS) n_mov += mover ? 1 : 0;
            __syncwarp();
            if (nlist >= 32) {       // drain the newest 32 (keeps the rest at the front)
                nlist -= 32;
                drain_movers<R, FMA, DEPOSIT, STATS>(a, ml, nlist, 32, lane, n_cross, n_wrap);
            }
        }

        if (PREFETCH) {
            ii = ii_n; x = x_n; y = y_n; z = z_n; ux = ux_n; uy = uy_n; uz = uz_n; w = w_n;
        } else {
            const long long nn = (tile + stride) * 32 + lane;
            if (nn < a.np) {
                const PRec<R> r = a.p.rec[nn];
                ii = real_to_cell(r.pos.w); x = r.pos.x; y = r.pos.y; z = r.pos.z;
                ux = r.mom.x; uy = r.mom.y; uz = r.mom.z; w = r.mom.w;
            }
        }
    }
    if (nlist > 0) drain_movers<R, FMA, DEPOSIT, STATS>(a, ml, 0, nlist, lane, n_cross, n_wrap);

    if (STATS) {
        __syncwarp();
        const unsigned full = 0xffffffffu;
        unsigned long long v[8];
        v[0] = n_mov; v[1] = n_cross;
#pragma unroll
        for (int k = 0; k < 6; ++k) v[2 + k] = n_wrap[k];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(full, v[k], o);
            if (lane == 0 && v[k]) atomicAdd(a.stats + k, v[k]);
        }
    }
}

// Reference: uncenter_particles, src/uncenter_p.h:27-98 -- backward half Boris rotation
// (qdt_4mc = -0.5*qdt_2mc) followed by one +half E kick; positions untouched.  Uses sqrt in
// real_t here (the reference casts a double sqrt), not sqrtf.
template <class R, bool FMA>
__global__ void __launch_bounds__(256) k_uncenter(Particles<R> p, long long np, const R* __restrict__ ip, R qdt_2mc) {
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    if (n >= np) return;
    const R one = R(1.), one_third = R(1. / 3.), two_fifteenths = R(2. / 15.);
    const R qdt_4mc = (R)(-0.5 * (double)qdt_2mc);
    R f[IpStride<R>::value];
    const PRec<R> r = p.rec[n];
    load_record(ip, real_to_cell(r.pos.w), f);
    const R x = r.pos.x, y = r.pos.y, z = r.pos.z;
    const R hax = qdt_2mc * madd<FMA>(z, madd<FMA>(y, f[I_D2EXDYDZ], f[I_DEXDZ]), madd<FMA>(y, f[I_DEXDY], f[I_EX]));
    const R hay = qdt_2mc * madd<FMA>(x, madd<FMA>(z, f[I_D2EYDZDX], f[I_DEYDX]), madd<FMA>(z, f[I_DEYDZ], f[I_EY]));
    const R haz = qdt_2mc * madd<FMA>(y, madd<FMA>(x, f[I_D2EZDXDY], f[I_DEZDY]), madd<FMA>(x, f[I_DEZDX], f[I_EZ]));
    const R cbx = madd<FMA>(x, f[I_DCBXDX], f[I_CBX]);
    const R cby = madd<FMA>(y, f[I_DCBYDY], f[I_CBY]);
    const R cbz = madd<FMA>(z, f[I_DCBZDZ], f[I_CBZ]);
    R ux = r.mom.x, uy = r.mom.y, uz = r.mom.z;
    R v0 = qdt_4mc / (R)sqrt((double)(one + madd<FMA>(ux, ux, madd<FMA>(uy, uy, uz * uz))));
    R v1 = madd<FMA>(cbx, cbx, madd<FMA>(cby, cby, cbz * cbz));
    R v2 = (v0 * v0) * v1;
    R v3 = v0 * madd<FMA>(v2, madd<FMA>(v2, two_fifteenths, one_third), one);
    R v4 = v3 / madd<FMA>(v1, v3 * v3, one);
    v4 += v4;
    v0 = madd<FMA>(v3, mdiff<FMA>(uy, cbz, uz, cby), ux);
    v1 = madd<FMA>(v3, mdiff<FMA>(uz, cbx, ux, cbz), uy);
    v2 = madd<FMA>(v3, mdiff<FMA>(ux, cby, uy, cbx), uz);
    ux = madd<FMA>(v4, mdiff<FMA>(v1, cbz, v2, cby), ux);
    uy = madd<FMA>(v4, mdiff<FMA>(v2, cbx, v0, cbz), uy);
    uz = madd<FMA>(v4, mdiff<FMA>(v0, cby, v1, cbx), uz);
    ux += hax; uy += hay; uz += haz;
    p.store_mom(n, ux, uy, uz, r.mom.w);
}

}  // namespace cpic
