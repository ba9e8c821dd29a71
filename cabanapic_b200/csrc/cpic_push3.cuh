// k_push3 -- the float push + mover + deposit + reordering kernel, third generation (sm_100a only):
// BLOCK-OWNED CELL RANGES with TMA-staged interpolators.
//
// Same arithmetic as k_push / k_push2 (reference: push<>, src/push.h:65-295; move_p<>, src/move_p.h:59-374); what
// changes is who processes what.  k_push2 is warp-autonomous over particle tiles: every lane gathers its interpolator
// records from L1/L2 (5-10 LDG.128 per lane and tile, ~7 distinct lines each), ~13 % "foreigners" break every run of
// equal cells, so a tile pays ~18 slot-claim atomics and ~60 scattered reduction lines
// (profiles/r03_push2_reorder_steady_256x256x64_ncu.md: L1TEX tag stage 70 % busy, issue 46 %).  Here:
//
//   * the store is exactly segmented: src = concatenation over cells c of segment(c) = [sin[c], sin[c+1]) (the
//     exclusive scan under which the previous reordering push / counting sort wrote it).  A CTA owns a CHUNK of
//     consecutive cells [c0, c0+ch) -- normally whole x-rows -- and therefore a contiguous range of the store;
//   * ONE elected thread brings the chunk's interpolator records (ch x 80 B, contiguous) and its slice of the segment
//     bounds into shared memory with two TMA bulk copies (cp.async.bulk.shared::cluster.global + mbarrier
//     complete_tx); every gather of a particle whose cell is inside the chunk is a shared-memory load;
//   * a particle is NATIVE when its index lies inside the segment of its own cell (it did not change cell in the
//     previous step: ~87 %), otherwise a FOREIGNER.  Natives along a tile are sorted by cell by construction, so their
//     destination slots need one claim atomic per (tile, cell) instead of one per broken run, their records leave as
//     contiguous 256-bit stores, and their first-streak currents go through the per-warp deposit rows and a
//     BRANCH-FREE segmented sum (predicated red.v4 on a cell change).  Foreigners claim their slot individually and
//     deposit with three direct red.v4 -- rare lanes, no effect on the native path;
//   * a warp owns a contiguous sub-range of the chunk's tiles; movers go to the per-warp list and are drained densely
//     as before, but the drain now writes the mover's WHOLE record once (momentum travels in the list).
#pragma once
#include "cpic_push2.cuh"

namespace cpic {

// Block shape: 5 warps x 4 CTAs per SM = 20 warps at 96 registers.  Measured (profiles/r2_push3_block_shapes.log, 256x256x64 x
// 64 ppc, ms): 5x4 5.71 | 10x2 5.83 | 12x2 5.96 | 4x4 5.98 | 4x5 6.00 | 8x2 6.05 | 6x3 6.15 | 8x3 6.16 (round-2 start) |
// 20x1 6.19 | 11x2 6.20 | 3x6 6.22 | 9x2 6.25 | 16x1 6.38 | 7x3 6.51.  What helps: 96 instead of 80 registers (no loop
// counter / pointer re-materialisation, 20 instead of 48 bytes of spill), a warp count per SM divisible by the four
// schedulers, and small CTAs (the chunk barriers and the TMA wait stall fewer warps).
#ifndef PUSH3_NWARPS
#define PUSH3_NWARPS 5
#endif
#ifndef PUSH3_MIN_BLOCKS
#define PUSH3_MIN_BLOCKS 4
#endif
constexpr int PUSH3_WARPS = PUSH3_NWARPS;
constexpr int PUSH3_CH_MAX = 264;      // cells per chunk (one x-row of the 256^3 deck incl. ghosts = 258)
constexpr int PUSH3_MOVER_CAP = 64;

struct Push3Args {
    PushArgs<float> a;       // p = src, dst, ip, acc, hist (new cells), cursor (= scan of the current cells' histogram, mutable)
    const unsigned* sin;     // [nc + 1 (+ pad)]: immutable segment bounds of src
    // chunk geometry: a z-plane (plane = gx*gy cells) is cut into cpp chunks of ch cells (the last one shorter), so a
    // chunk never straddles two planes; nchunks = cpp * gz.  Work order: blocks of yblock chunks, all planes of a block
    // before the next block -- the z-neighbours of a row are then visited within ~yblock chunks of each other and their
    // accumulator / interpolator / cursor lines are still in L2 (in plain voxel order a plane = 272 MB of particle
    // traffic lies between them at 256^3 x 64)
    int ch, cpp, gz, plane, yblock, nchunks, nc;
    unsigned* work;          // dynamic work counter (zeroed by the host before the launch)
};

// A warp's list of cell-crossers waiting for the drain (the in-kernel form of VPIC's particle_mover_t list the reference
// kept in comments, src/push.h:271-291): 48-byte entries {x y z cell | rx ry rz slot | ux uy uz w}, three 128-bit words --
// an append is 3 STS.128 and a drain read 3 LDS.128 (stride 12 words: conflict-free), where thirteen separate arrays
// cost 13 shared-memory wavefronts per append with only the few mover lanes active.  The charge q = w * qsp is
// recomputed by the drain (the same single multiplication).
struct MoverList3 {
    float4 e[PUSH3_MOVER_CAP * 3];
};

struct Push3Smem {
    float ip[PUSH3_CH_MAX * 20];                 // the chunk's interpolator records (TMA destination, 16-byte aligned)
    unsigned sin[PUSH3_CH_MAX + 8];              // sin[c0 & ~3 ...]: the chunk's slice of the segment bounds (TMA destination)
    unsigned long long mbar;                     // transaction barrier of the two bulk copies
    int chunk;                                   // the chunk the CTA works on (dynamic scheduling)
    int pad_;
    MoverList3 lists[PUSH3_WARPS];
    float rows[PUSH3_WARPS][32 * PUSH2_ROW];
    int rcell[PUSH3_WARPS][32];
    int rcnt[PUSH3_WARPS][32];
};

// ---- TMA / mbarrier (PTX ISA 8.x; SASS: UBLKCP, SYNCS) -----------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CPIC_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CPIC_MBAR_DONE;\n"
        "bra CPIC_MBAR_WAIT;\n"
        "CPIC_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// predicated reductions: no branch around a rare flush (the compiler turned the `if` of segsum_rows into BSSY/BRA/BSYNC
// triples -- 175 control-flow instructions per tile in profiles/r03_push2_instruction_mix.md)
__device__ __forceinline__ void red_add_v4_if(bool p, float* addr, float x, float y, float z, float w) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n}\n" ::"l"(addr),
        "f"(x), "f"(y), "f"(z), "f"(w), "r"((unsigned)p)
        : "memory");
}
__device__ __forceinline__ void red_add_u32_if(bool p, unsigned* addr, unsigned v) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p red.relaxed.gpu.global.add.u32 [%0], %1;\n}\n" ::"l"(addr), "r"(v),
                 "r"((unsigned)p)
                 : "memory");
}

// Branch-free segmented sum of the warp's 32 deposit rows (see segsum_rows): lane -> (row group rg, entry group eg).
// Rows of a tile's native stayers are sorted by cell; a row with cell -1 holds exact zeros and never breaks a run.
// Lanes 24..31 count the native stayers per cell the same way (histogram of the new cells).
__device__ __forceinline__ void segsum_rows3(const float* rows, const int* rcell, const int* rcnt, float* __restrict__ gacc,
                                             unsigned* __restrict__ ghist, int lane) {
    const int rg = lane < 24 ? lane / 3 : lane - 24;
    const int4 c4 = reinterpret_cast<const int4*>(rcell)[rg];
    if (lane < 24) {
        const int eg = lane - 3 * rg;
        const float* src = rows + (rg * 4) * PUSH2_ROW + eg * 4;
        float4 s = *reinterpret_cast<const float4*>(src);
        int c = c4.x;
#define CPIC_SEG3(CN, K)                                                                                   \
        {                                                                                                  \
            const float4 v = *reinterpret_cast<const float4*>(src + (K) * PUSH2_ROW);                      \
            const bool chg = (CN) >= 0 && c >= 0 && (CN) != c;                                             \
            red_add_v4_if(chg, gacc + (long long)c * 12 + eg * 4, s.x, s.y, s.z, s.w);                     \
            const float keep = chg ? 0.f : 1.f;                                                            \
            s.x = fmaf(s.x, keep, v.x); s.y = fmaf(s.y, keep, v.y); s.z = fmaf(s.z, keep, v.z); s.w = fmaf(s.w, keep, v.w); \
            c = (CN) >= 0 ? (CN) : c;                                                                      \
        }
        CPIC_SEG3(c4.y, 1)
        CPIC_SEG3(c4.z, 2)
        CPIC_SEG3(c4.w, 3)
#undef CPIC_SEG3
        red_add_v4_if(c >= 0, gacc + (long long)c * 12 + eg * 4, s.x, s.y, s.z, s.w);
    } else {
        const int4 n4 = reinterpret_cast<const int4*>(rcnt)[rg];
        int c = c4.x;
        unsigned cnt = (unsigned)n4.x;
#define CPIC_SEGC3(CN, NN)                                                                                 \
        {                                                                                                  \
            const bool chg = (CN) >= 0 && c >= 0 && (CN) != c;                                             \
            red_add_u32_if(chg && cnt != 0u, ghist + c, cnt);                                              \
            cnt = (chg ? 0u : cnt) + (unsigned)(NN);                                                       \
            c = (CN) >= 0 ? (CN) : c;                                                                      \
        }
        CPIC_SEGC3(c4.y, n4.y)
        CPIC_SEGC3(c4.z, n4.z)
        CPIC_SEGC3(c4.w, n4.w)
#undef CPIC_SEGC3
        red_add_u32_if(c >= 0 && cnt != 0u, ghist + c, cnt);
    }
}

// Drain list entries [first, first+32): the move_p loop (src/move_p.h:93-371) per lane, then ONE 256-bit store of the
// mover's whole record at the slot the main path claimed for it.
template <bool FMA, bool STATS>
__device__ __forceinline__ void drain_movers3(const PushArgs<float>& a, const MoverList3& ml, int first, int count, int lane,
                                              unsigned long long& n_cross, unsigned long long (&n_wrap)[6]) {
    const int m = first + lane;
    bool leaves = false;
    unsigned leaver = 0;
    if (lane < count) {
        const float4 e0 = ml.e[3 * m], e1 = ml.e[3 * m + 1], e2 = ml.e[3 * m + 2];
        float px = e0.x, py = e0.y, pz = e0.z;
        float dx = e1.x, dy = e1.y, dz = e1.z;
        const float qq = __fmul_rn(e2.w, a.qsp);
        int c = __float_as_int(e0.w);
        unsigned flip = 0;
        for (;;) {
            float sx, sy, sz, mx, my, mz, v5, dirv;
            const int axis = mover_streak(px, py, pz, dx, dy, dz, qq, sx, sy, sz, mx, my, mz, v5, dirv);
            float jc[12];
            streak_currents<FMA>(qq, sx, sy, sz, mx, my, mz, v5, jc);
            row_add_vec(a.acc + (long long)c * 12, jc);
            if (axis == 3) break;
            const int code = cross_face(c, axis, dirv, a);
            if (code & CROSS_REFLECTED) {      // reflecting wall (Boundary::Reflect): stay on the face, turn around
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
        const unsigned pn = __float_as_uint(e1.w);
        leaves = a.leave_list && (c < a.leave_lo || c >= a.leave_hi);
        leaver = pn;
        PRec<float> o;
        o.pos.x = px; o.pos.y = py; o.pos.z = pz; o.pos.w = cell_to_real(c, 0.f);
        o.mom.x = e2.x; o.mom.y = e2.y; o.mom.z = e2.z; o.mom.w = e2.w;
        if (flip & 1u) o.mom.x = -o.mom.x;
        if (flip & 2u) o.mom.y = -o.mom.y;
        if (flip & 4u) o.mom.z = -o.mom.z;
        a.dst.rec[pn] = o;
        atomicAdd(a.hist + c, 1u);
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

template <bool FMA, bool STATS, bool FASTDS>
__global__ void __launch_bounds__(PUSH3_WARPS * 32, PUSH3_MIN_BLOCKS) k_push3(const __grid_constant__ Push3Args q, float one_rt) {
    extern __shared__ __align__(128) unsigned char push3_smem_raw[];
    Push3Smem& sm = *reinterpret_cast<Push3Smem*>(push3_smem_raw);
    const PushArgs<float>& a = q.a;
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    MoverList3& ml = sm.lists[warp];
    float* rows = sm.rows[warp];
    int* rcell = sm.rcell[warp];
    int* rcnt = sm.rcnt[warp];
    P2 P{one_rt};
    const float one = 1.f, one_third = (float)(1. / 3.), two_fifteenths = (float)(2. / 15.);
    int nlist = 0;
    const unsigned sm_ip = smem_u32(sm.ip);
    unsigned long long n_mov = 0, n_cross = 0, n_wrap[6] = {0, 0, 0, 0, 0, 0};
    const PRec<float>* __restrict__ grec = a.p.rec;
    PRec<float> rzero;
    rzero.pos = PHalf<float>{0.f, 0.f, 0.f, 0.f}; rzero.mom = rzero.pos;

    if (threadIdx.x == 0) mbar_init(&sm.mbar, 1);
    __syncthreads();
    unsigned phase = 0;
    // work items: chunks [0, nchunks) of q.ch cells each, then tail items of 64*PUSH3_WARPS*8 particles behind the segments
    if (a.np_dev && a.np_dev[1]) return;      // a migration flagged an overflow / inconsistency (reported by sync_np)
    const unsigned np_ = (unsigned)(a.np_dev ? *a.np_dev : a.np);
    const unsigned tail_lo = q.sin[q.nc];      // particles [tail_lo, np) lie behind the segments (slab mode: arrivals of a migration)
    const unsigned tail_n = np_ > tail_lo ? np_ - tail_lo : 0u;
    constexpr unsigned TAIL_ITEM = 64u * PUSH3_WARPS * 8u;
    const int nitems = q.nchunks + (int)((tail_n + TAIL_ITEM - 1u) / TAIL_ITEM);

    for (;;) {
        __syncthreads();                      // the previous item's readers are done with sm.ip / sm.sin / sm.chunk
        if (threadIdx.x == 0) sm.chunk = (int)atomicAdd(q.work, 1u);
        __syncthreads();
        const int item = sm.chunk;
        if (item >= nitems) break;
        int c0 = 0, chn = 0;
        unsigned P0, P1;
        const unsigned* sS = sm.sin;
        if (item < q.nchunks) {
            const int per = q.yblock * q.gz, jb_last = (q.cpp - 1) / q.yblock;
            const int jb = min(item / per, jb_last), rem = item - jb * per;
            const int bc = jb == jb_last ? q.cpp - jb * q.yblock : q.yblock;
            const int zz = rem / bc, j = jb * q.yblock + rem % bc;
            c0 = zz * q.plane + j * q.ch;
            const int c1 = min(c0 + q.ch, (zz + 1) * q.plane);
            chn = c1 - c0;
            P0 = q.sin[c0]; P1 = q.sin[c1];
            if (P1 > np_) P1 = np_;           // (slab mode: the store shrank below the segments)
            if (P0 >= P1) continue;           // no particles in this chunk (uniform over the CTA)
            if (threadIdx.x == 0) {
                const int c0a = c0 & ~3;
                const unsigned ns = (unsigned)(((c1 + 1 - c0a) + 3) & ~3);
                mbar_expect_tx(&sm.mbar, (unsigned)chn * 80u + ns * 4u);
                bulk_g2s(sm.ip, a.ip + (long long)c0 * 20, (unsigned)chn * 80u, &sm.mbar);
                bulk_g2s(sm.sin, q.sin + c0a, ns * 4u, &sm.mbar);
            }
            sS = sm.sin + (c0 & 3);           // sS[k] = sin[c0 + k], k = 0 .. chn
            mbar_wait(&sm.mbar, phase);
            phase ^= 1u;
        } else {                              // a tail item: every particle is a foreigner, nothing to stage
            P0 = tail_lo + (unsigned)(item - q.nchunks) * TAIL_ITEM;
            P1 = min(P0 + TAIL_ITEM, np_);
        }
        // this warp's contiguous range of the item's 64-particle tiles
        const unsigned ntile = (P1 - P0 + 63u) >> 6;
        const unsigned r0 = P0 + 64u * ((unsigned)warp * ntile / PUSH3_WARPS);
        const unsigned r1 = min(P0 + 64u * ((unsigned)(warp + 1) * ntile / PUSH3_WARPS), P1);

        PRec<float> rA = rzero, rB = rzero;
        {
            const unsigned iA = r0 + 2u * lane;
            if (iA < r1) rA = grec[iA];
            if (iA + 1u < r1) rB = grec[iA + 1u];
        }
        for (unsigned i0 = r0; i0 < r1; i0 += 64u) {
            const unsigned iA = i0 + 2u * lane, iB = iA + 1u;
            const bool validA = iA < r1, validB = iB < r1;
            // unpack this tile's records into the packed operands, then request the next tile's records INTO THE SAME
            // REGISTERS (consumed next iteration; nothing may touch them before).  No second record buffer, no copy at the
            // end of the iteration; a lane without a next particle keeps its old, finite record (it is masked by valid*).
            const int cA = validA ? real_to_cell(rA.pos.w) : c0;
            const int cB = validB ? real_to_cell(rB.pos.w) : cA;
            float2 x = make_float2(rA.pos.x, rB.pos.x), y = make_float2(rA.pos.y, rB.pos.y), z = make_float2(rA.pos.z, rB.pos.z);
            float2 ux = make_float2(rA.mom.x, rB.mom.x), uy = make_float2(rA.mom.y, rB.mom.y), uz = make_float2(rA.mom.z, rB.mom.z);
            const float2 w = make_float2(rA.mom.w, rB.mom.w);
            {
                const unsigned nA = iA + 64u;
                if (nA < r1) rA = grec[nA];
                if (nA + 1u < r1) rB = grec[nA + 1u];
            }
            const unsigned oA = (unsigned)(cA - c0), oB = (unsigned)(cB - c0);
            const bool inA = oA < (unsigned)chn, inB = oB < (unsigned)chn;
            // native: the particle's index lies inside the segment of its own cell
            bool natA = false, natB = false;
            if (validA && inA) { const unsigned lo = sS[oA]; natA = (iA - lo) < (sS[oA + 1] - lo); }
            if (validB && inB) { const unsigned lo = sS[oB]; natB = (iB - lo) < (sS[oB + 1] - lo); }
            // the natives of a tile are sorted by cell: cells cf .. cl, lane k claims the slots of cell cf + k
            const int cf = __reduce_min_sync(full, natA ? cA : (natB ? cB : 0x7fffffff));
            const int cl = __reduce_max_sync(full, natB ? cB : (natA ? cA : -1));
            int ncell = cl - cf + 1;
            if (ncell > 32) { natA = false; natB = false; ncell = 0; }      // very sparse cells: everybody takes the general path
            const bool forA = validA && !natA, forB = validB && !natB;
            unsigned mycnt = 0;
            int rkA = 0, rkB = 0;
            for (int k = 0; k < ncell; ++k) {
                const int c = cf + k;
                const bool hA = natA && cA == c, hB = natB && cB == c;
                const unsigned mA = __ballot_sync(full, hA), mB = __ballot_sync(full, hB);
                if (lane == k) mycnt = __popc(mA) + __popc(mB);
                if (hA) rkA = __popc(mA & lt);
                if (hB) rkB = __popc(mA) + __popc(mB & lt);
            }
            // destination slots (in the segment of the cell the particle is in now); the atomics' round trip overlaps
            // the gather and the Boris rotation
            unsigned base = 0, fsA = 0, fsB = 0;
            if (mycnt) base = atomicAdd(a.cursor + cf + lane, mycnt);
            if (forA) fsA = atomicAdd(a.cursor + cA, 1u);
            if (forB) fsB = atomicAdd(a.cursor + cB, 1u);

            // ---- field gather (src/push.h:74-138): from the TMA-staged chunk, or from global memory for a foreigner
            // whose cell lies outside the chunk
            float2 hax, hay, haz, cbx, cby, cbz;
            {
                const float4* pa = inA ? reinterpret_cast<const float4*>(sm.ip + oA * 20u)
                                        : reinterpret_cast<const float4*>(a.ip + (long long)cA * 20);
                const float4* pb = inB ? reinterpret_cast<const float4*>(sm.ip + oB * 20u)
                                        : reinterpret_cast<const float4*>(a.ip + (long long)cB * 20);
                float fA[20], fB[20];
#pragma unroll
                for (int k = 0; k < 5; ++k) *reinterpret_cast<float4*>(&fA[4 * k]) = pa[k];
#pragma unroll
                for (int k = 0; k < 5; ++k) *reinterpret_cast<float4*>(&fB[4 * k]) = pb[k];
#define F2(k) make_float2(fA[k], fB[k])
                hax = P.mul(P.madd<FMA>(z, P.madd<FMA>(y, F2(I_D2EXDYDZ), F2(I_DEXDZ)), P.madd<FMA>(y, F2(I_DEXDY), F2(I_EX))), a.qdt_2mc);
                hay = P.mul(P.madd<FMA>(x, P.madd<FMA>(z, F2(I_D2EYDZDX), F2(I_DEYDX)), P.madd<FMA>(z, F2(I_DEYDZ), F2(I_EY))), a.qdt_2mc);
                haz = P.mul(P.madd<FMA>(y, P.madd<FMA>(x, F2(I_D2EZDXDY), F2(I_DEZDY)), P.madd<FMA>(x, F2(I_DEZDX), F2(I_EZ))), a.qdt_2mc);
                cbx = P.madd<FMA>(x, F2(I_DCBXDX), F2(I_CBX));
                cby = P.madd<FMA>(y, F2(I_DCBYDY), F2(I_CBY));
                cbz = P.madd<FMA>(z, F2(I_DCBZDZ), F2(I_CBZ));
#undef F2
            }
            const float2 qq = P.mul(w, a.qsp);

            // ---- Boris push (src/push.h:144-167)
            ux = P.add(ux, hax); uy = P.add(uy, hay); uz = P.add(uz, haz);
            float2 v0, v1, v2, v3, v4;
            {
                const float2 g2 = P.add(P.madd<FMA>(ux, ux, P.madd<FMA>(uy, uy, P.mul(uz, uz))), one);
                if (FASTDS && __all_sync(full, safe_below(g2))) v0 = div2_fast(P.bc(a.qdt_2mc), sqrt2_fast(g2));
                else v0 = make_float2(__fdiv_rn(a.qdt_2mc, __fsqrt_rn(g2.x)), __fdiv_rn(a.qdt_2mc, __fsqrt_rn(g2.y)));   // :148
            }
            v1 = P.madd<FMA>(cbx, cbx, P.madd<FMA>(cby, cby, P.mul(cbz, cbz)));
            v2 = P.mul(P.mul(v0, v0), v1);
            v3 = P.mul(v0, P.madd<FMA>(v2, P.madd<FMA>(v2, two_fifteenths, one_third), one));
            {
                const float2 den = P.madd<FMA>(v1, P.mul(v3, v3), one);
                if (FASTDS && __all_sync(full, safe_below(den))) v4 = div2_fast(v3, den);
                else v4 = make_float2(__fdiv_rn(v3.x, den.x), __fdiv_rn(v3.y, den.y));
            }
            v4 = P.add(v4, v4);
            v0 = P.madd<FMA>(v3, P.mdiff<FMA>(uy, cbz, uz, cby), ux);
            v1 = P.madd<FMA>(v3, P.mdiff<FMA>(uz, cbx, ux, cbz), uy);
            v2 = P.madd<FMA>(v3, P.mdiff<FMA>(ux, cby, uy, cbx), uz);
            ux = P.madd<FMA>(v4, P.mdiff<FMA>(v1, cbz, v2, cby), ux);
            uy = P.madd<FMA>(v4, P.mdiff<FMA>(v2, cbx, v0, cbz), uy);
            uz = P.madd<FMA>(v4, P.mdiff<FMA>(v0, cby, v1, cbx), uz);
            ux = P.add(ux, hax); uy = P.add(uy, hay); uz = P.add(uz, haz);
            const float2 pux = ux, puy = uy, puz = uz;      // the new momentum (:165-167)

            // ---- displacement (src/push.h:169-182)
            {
                const float2 g2 = P.add(P.madd<FMA>(ux, ux, P.madd<FMA>(uy, uy, P.mul(uz, uz))), one);
                if (FASTDS && __all_sync(full, safe_below(g2))) v0 = div2_fast(P.bc(one), sqrt2_fast(g2));
                else v0 = make_float2(__fdiv_rn(one, __fsqrt_rn(g2.x)), __fdiv_rn(one, __fsqrt_rn(g2.y)));
            }
            ux = P.mul(ux, a.cdt_dx); uy = P.mul(uy, a.cdt_dy); uz = P.mul(uz, a.cdt_dz);
            ux = P.mul(ux, v0); uy = P.mul(uy, v0); uz = P.mul(uz, v0);
            const float2 mx = P.add(x, ux), my = P.add(y, uy), mz = P.add(z, uz);        // streak midpoint
            const float2 nx_ = P.add(mx, ux), ny_ = P.add(my, uy), nz_ = P.add(mz, uz);  // new position

            const bool inpA = fabsf(nx_.x) <= one && fabsf(ny_.x) <= one && fabsf(nz_.x) <= one;   // :187
            const bool inpB = fabsf(nx_.y) <= one && fabsf(ny_.y) <= one && fabsf(nz_.y) <= one;
            const bool stayA = validA && inpA, stayB = validB && inpB;
            const bool movA = validA && !inpA, movB = validB && !inpB;

            // the claimed slots: natives take consecutive slots of their cell's claim, foreigners their own
            const unsigned nbA = __shfl_sync(full, base, (cA - cf) & 31), nbB = __shfl_sync(full, base, (cB - cf) & 31);
            const unsigned dA = natA ? nbA + (unsigned)rkA : fsA;
            const unsigned dB = natB ? nbB + (unsigned)rkB : fsB;
            // a stayer's whole record in one full-sector store (a mover's is written by the drain)
            {
                PRec<float> o;
                if (stayA) {
                    o.pos.x = nx_.x; o.pos.y = ny_.x; o.pos.z = nz_.x; o.pos.w = cell_to_real(cA, 0.f);
                    o.mom.x = pux.x; o.mom.y = puy.x; o.mom.z = puz.x; o.mom.w = w.x;
                    a.dst.rec[dA] = o;
                }
                if (stayB) {
                    o.pos.x = nx_.y; o.pos.y = ny_.y; o.pos.z = nz_.y; o.pos.w = cell_to_real(cB, 0.f);
                    o.mom.x = pux.y; o.mom.y = puy.y; o.mom.z = puz.y; o.mom.w = w.y;
                    a.dst.rec[dB] = o;
                }
            }

            // ---- first-streak currents of the pair (src/push.h:203-254), packed.  Movers and invalid lanes get
            // charge 0 (every current is a product with q: exact zeros, no selects).
            {
                const bool nsA = natA && stayA, nsB = natB && stayB;            // native stayers -> deposit rows
                const bool pairB = nsB && (!nsA || cB == cA);                     // B shares the row (or has it alone)
                const float2 qd = make_float2(stayA ? qq.x : 0.f, stayB ? qq.y : 0.f);
                float2 cur[12];
                const float2 v5 = P.mul(P.mul(P.mul(P.mul(qd, ux), uy), uz), one_third);   // :203
                streak_currents2<FMA>(P, qd, ux, uy, uz, mx, my, mz, v5, cur);
                __syncwarp();
                const float wA = nsA ? 1.f : 0.f, wB = pairB ? 1.f : 0.f;      // exact 0/1 weights instead of 24 selects
                float4* r4 = reinterpret_cast<float4*>(rows + lane * PUSH2_ROW);
#define RW(j) fmaf(cur[j].y, wB, cur[j].x * wA)
                r4[0] = make_float4(RW(0), RW(1), RW(2), RW(3));
                r4[1] = make_float4(RW(4), RW(5), RW(6), RW(7));
                r4[2] = make_float4(RW(8), RW(9), RW(10), RW(11));
#undef RW
                rcell[lane] = nsA ? cA : (pairB ? cB : -1);
                rcnt[lane] = (nsA ? 1 : 0) + (pairB ? 1 : 0);
                // everything else that stays deposits directly: foreigners, and a native B whose pair straddles two cells
                const bool dirA = stayA && !nsA, dirB = stayB && !pairB;
                float* const ga = a.acc + (long long)cA * 12;
                float* const gb = a.acc + (long long)cB * 12;
                red_add_v4_if(dirA, ga + 0, cur[0].x, cur[1].x, cur[2].x, cur[3].x);
                red_add_v4_if(dirA, ga + 4, cur[4].x, cur[5].x, cur[6].x, cur[7].x);
                red_add_v4_if(dirA, ga + 8, cur[8].x, cur[9].x, cur[10].x, cur[11].x);
                red_add_u32_if(dirA, a.hist + cA, 1u);
                red_add_v4_if(dirB, gb + 0, cur[0].y, cur[1].y, cur[2].y, cur[3].y);
                red_add_v4_if(dirB, gb + 4, cur[4].y, cur[5].y, cur[6].y, cur[7].y);
                red_add_v4_if(dirB, gb + 8, cur[8].y, cur[9].y, cur[10].y, cur[11].y);
                red_add_u32_if(dirB, a.hist + cB, 1u);
                __syncwarp();
                segsum_rows3(rows, rcell, rcnt, a.acc, a.hist, lane);
            }

            // ---- movers: append to the warp's list, drain densely (src/push.h:261-269 -> move_p)
            const unsigned mA = __ballot_sync(full, movA), mB = __ballot_sync(full, movB);
            if (mA | mB) {
                if (STATS) n_mov += (movA ? 1 : 0) + (movB ? 1 : 0);
                if (mA) {
                    if (movA) {
                        const int m = nlist + __popc(mA & lt);
                        ml.e[3 * m] = make_float4(x.x, y.x, z.x, __int_as_float(cA));
                        ml.e[3 * m + 1] = make_float4(ux.x, uy.x, uz.x, __uint_as_float(dA));
                        ml.e[3 * m + 2] = make_float4(pux.x, puy.x, puz.x, w.x);
                    }
                    nlist += __popc(mA);
                    __syncwarp();
                    if (nlist >= 32) {
                        nlist -= 32;
                        drain_movers3<FMA, STATS>(a, ml, nlist, 32, lane, n_cross, n_wrap);
                    }
                }
                if (mB) {
                    if (movB) {
                        const int m = nlist + __popc(mB & lt);
                        ml.e[3 * m] = make_float4(x.y, y.y, z.y, __int_as_float(cB));
                        ml.e[3 * m + 1] = make_float4(ux.y, uy.y, uz.y, __uint_as_float(dB));
                        ml.e[3 * m + 2] = make_float4(pux.y, puy.y, puz.y, w.y);
                    }
                    nlist += __popc(mB);
                    __syncwarp();
                    if (nlist >= 32) {
                        nlist -= 32;
                        drain_movers3<FMA, STATS>(a, ml, nlist, 32, lane, n_cross, n_wrap);
                    }
                }
            }
        }
    }
    if (nlist > 0) drain_movers3<FMA, STATS>(a, ml, 0, nlist, lane, n_cross, n_wrap);

    if (STATS) {
        __syncwarp();
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

}  // namespace cpic
