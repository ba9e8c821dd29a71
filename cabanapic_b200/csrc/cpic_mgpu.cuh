// Multi-GPU layer of the C ABI (include/cabanapic_b200_mgpu.h): one process per GPU, NCCL called from C++.
// Included at the end of cpic_capi.cu (it drives a context through CtxBase).  The exchange choreography is the one
// cabanapic_b200/dist.py prototyped in Python + torch in round 1; here it is host C++ and three small kernels:
//   * copy planes (J ghost copy, cB ghost copy) are sent from and received INTO the field arrays themselves -- a
//     z-plane of a member is one contiguous run of gx*gy reals -- so they need no pack / unpack at all;
//   * add planes (ghost accumulator planes, the z sweeps of the J fold) are received into scratch and added by
//     k_plane_add / k_rect_add;
//   * the particle migration uses the device-counted extraction / append of the context (cpic_slab_*_async).
// Transport of the slab exchanges: PEER MEMORY over NVLink / NVSwitch.  At create time every rank exports its mailbox
// (receive buffers + two arrival flags) through CUDA IPC and maps its two z neighbours'; an exchange
// is then ONE kernel (k_p2p_put) that stores this rank's planes / leaver records straight into the neighbours' mailboxes
// -- the particle payload cut to the device-side leaver count -- and raises the neighbours' arrival flags with a system-scope release once its last block is done, followed
// by a one-warp kernel (k_p2p_wait) that acquires this rank's own flags.  No NCCL kernel, no rendezvous, no host
// involvement; the sequence numbers live in device memory, so the whole step still replays as a CUDA graph.  Why no
// credits are needed: between two uses of the same landing buffer the sender has always waited for a later message of
// the same neighbour, which that neighbour only sends after consuming the earlier one (stream order) -- every landing
// buffer is reused two or more phases later (accumulator planes / J fold planes share scratch, the cB and J copy planes
// have their own slots).  Nothing lands in
// the field arrays themselves: a neighbour may be a whole phase ahead, and this rank's own x/y ghost passes still write
// the z ghost planes then (measured: in-place landing changed 13 of 3.4e7 migrations at 8 ranks) -- ghost planes are
// unpacked from the mailbox by k_planes_unpack after the wait.  NCCL send/recv
// (five groups per step, ~110 us each: profiles/r03_slab_timeline_2gpu_256x256x64.log) stays as the fallback when IPC
// mapping is unavailable (CPIC_MGPU_P2P=0 forces it) and carries the one-time handle exchange.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <chrono>
#include <thread>

#include "../../include/cabanapic_b200_mgpu.h"

namespace {

// ---- NCCL, resolved at run time ------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
    bool load() {
        if (lib) return true;
        const char* names[] = {getenv("CPIC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { why = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return false; }
#define CPIC_NCCL_SYM(field, name) \
        field = reinterpret_cast<decltype(field)>(dlsym(lib, name)); \
        if (!field) { why = std::string("NCCL symbol missing: ") + name; lib = nullptr; return false; }
        CPIC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        CPIC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        CPIC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        CPIC_NCCL_SYM(GroupStart, "ncclGroupStart")
        CPIC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        CPIC_NCCL_SYM(Send, "ncclSend")
        CPIC_NCCL_SYM(Recv, "ncclRecv")
        CPIC_NCCL_SYM(AllReduce, "ncclAllReduce")
        CPIC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CPIC_NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;
thread_local std::string g_mgpu_error;
static_assert(sizeof(ncclUniqueId) == CPIC_MGPU_ID_BYTES, "ncclUniqueId is 128 bytes");

// ---- kernels of the exchange ---------------------------------------------------------------------------------------------
template <class R>
__global__ void __launch_bounds__(256) k_plane_add(R* __restrict__ dst, const R* __restrict__ src, long long n) {
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
// dst[y][x] += src[y][x] on the rectangle [y0,y1) x [x0,x1) of a gy x gx plane (the z sweeps of the J fold, src/fields.h:126-151)
template <class R>
__global__ void __launch_bounds__(256) k_rect_add(R* __restrict__ dst, const R* __restrict__ src, int gx, int y0, int y1, int x0, int x1) {
    const int w = x1 - x0;
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i >= (long long)w * (y1 - y0)) return;
    const int y = y0 + (int)(i / w), x = x0 + (int)(i % w);
    dst[(long long)y * gx + x] += src[(long long)y * gx + x];
}
// ---- peer-memory exchange ------------------------------------------------------------------------------------------------
struct PutSeg {
    const char* src;
    char* dst;                 // in a neighbour's memory (IPC mapping)
    long long bytes;           // multiple of 4
    const long long* n_ptr;    // if set: copy min(*n_ptr * elem, bytes) bytes (device-side leaver count)
    long long elem;
};
constexpr int P2P_MAXSEG = 24;
struct PutArgs {
    PutSeg seg[P2P_MAXSEG];
    unsigned* done;                    // block counter (local)
    unsigned long long* sent;          // [2] messages sent so far: [0] downwards, [1] upwards (local)
    unsigned long long* flag[2];       // [0] the lower neighbour's "from above" flag, [1] the upper neighbour's "from below" flag
};
struct WaitArgs {
    const unsigned long long* flag[2]; // this rank's arrival flags: [0] from below, [1] from above (null: nothing expected)
    unsigned long long* expect;        // [2] messages consumed so far (local)
    long long* err;                    // err[1] |= 8: a neighbour did not arrive in time (this layer's own word)
    long long* dc;                     // the context's device counts (may be null): dc[1] |= 8 stops its kernels too
    unsigned long long timeout_ns;
};
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// grid (blocks per segment, segments): 128-bit stores over NVLink; the last block to finish publishes the arrival
__global__ void __launch_bounds__(256) k_p2p_put(const __grid_constant__ PutArgs a) {
    const PutSeg& s = a.seg[blockIdx.y];
    long long bytes = s.bytes;
    if (s.n_ptr) {
        long long n = *s.n_ptr;
        if (n < 0) n = 0;
        bytes = n * s.elem < bytes ? n * s.elem : bytes;
    }
    const long long tid = blockIdx.x * 256LL + threadIdx.x, nthr = gridDim.x * 256LL;
    long long w0 = 0;                  // 4-byte words already copied by the 128-bit loop
    if ((((unsigned long long)s.src | (unsigned long long)s.dst) & 15ull) == 0) {
        const long long n16 = bytes >> 4;
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(s.src);
        uint4* __restrict__ dst = reinterpret_cast<uint4*>(s.dst);
        for (long long i = tid; i < n16; i += nthr) dst[i] = src[i];
        w0 = n16 * 4;
    }
    {
        const unsigned* __restrict__ src = reinterpret_cast<const unsigned*>(s.src);
        unsigned* __restrict__ dst = reinterpret_cast<unsigned*>(s.dst);
        for (long long i = w0 + tid; i < (bytes >> 2); i += nthr) dst[i] = src[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(a.done, 1u) == gridDim.x * gridDim.y - 1u) {
            *a.done = 0u;
            __threadfence_system();
            for (int d = 0; d < 2; ++d)
                if (a.flag[d]) st_release_sys(a.flag[d], ++a.sent[d]);
        }
    }
}
// test hook (CPIC_P2P_SKEW_US): holds a rank back before it sends, so that the tests see neighbours a phase apart
__global__ void k_p2p_skew(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { __nanosleep(1000); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}
__global__ void k_p2p_wait(const WaitArgs a) {
    const int d = threadIdx.x;
    if (d >= 2 || !a.flag[d]) return;
    const unsigned long long want = ++a.expect[d];
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(a.flag[d]) < want) {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > a.timeout_ns) {
            a.err[1] |= 8;
            if (a.dc) a.dc[1] |= 8;
            break;
        }
    }
}

// copy planes out of the mailbox into the z ghost planes: blockIdx.y = (direction, member)
struct UnpackArgs { const char* src[6]; char* dst[6]; long long bytes; };
__global__ void __launch_bounds__(256) k_planes_unpack(const UnpackArgs a) {
    const uint4* __restrict__ s = reinterpret_cast<const uint4*>(a.src[blockIdx.y]);
    uint4* __restrict__ d = reinterpret_cast<uint4*>(a.dst[blockIdx.y]);
    const long long n16 = a.bytes >> 4;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n16; i += gridDim.x * 256LL) d[i] = s[i];
    const unsigned* __restrict__ s4 = reinterpret_cast<const unsigned*>(a.src[blockIdx.y]);
    unsigned* __restrict__ d4 = reinterpret_cast<unsigned*>(a.dst[blockIdx.y]);
    for (long long i = n16 * 4 + blockIdx.x * 256LL + threadIdx.x; i < (a.bytes >> 2); i += gridDim.x * 256LL) d4[i] = s4[i];
}

__global__ void k_count_accumulate(const long long* __restrict__ sent, long long* __restrict__ total) {
    total[0] += sent[0]; total[1] += sent[1];
}

struct Mgpu {
    CtxBase* c = nullptr;
    cpic_params gprm{};      // the global box
    int rank = 0, world = 1, up = 0, down = 0, mode = CPIC_MGPU_REPLICATED;
    int z0 = 0, nzl = 0, nzl_down = 0;
    ncclComm_t comm = nullptr;
    std::string err;
    // slab exchange buffers (device)
    long long send_cap = 0;
    char *send_lo = nullptr, *send_hi = nullptr, *recv_dn = nullptr, *recv_up = nullptr;
    long long* cnt = nullptr;       // [0,1] sent (lo, hi) in the last step, [2,3] received (from below, from above), [4,5] totals sent
    char* scratch = nullptr;        // 2 accumulator planes (or 2 + 2 field planes)
    double* diag = nullptr;         // 8 doubles
    // CUDA graph of a pair of fused slab steps
    cudaGraphExec_t gexec = nullptr;
    bool graph_failed = false, used_graph = false;
    cpic_consts graph_k{};
    unsigned graph_sig = 0;      // CtxBase::state_signature() at capture time
    long long steps_done = 0;

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int ctx(int rc) { if (rc) err = c->err; return rc; }
    int nccl(ncclResult_t r, const char* what) {
        if (r == ncclSuccess) return CPIC_OK;
        return fail(CPIC_E_CUDA, "%s: NCCL: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    }
    size_t rb() const { return (size_t)c->prm.real_bytes; }
    ncclDataType_t rtype() const { return c->prm.real_bytes == 4 ? ncclFloat : ncclDouble; }
    long long plane() const { return (long long)c->g.gx * c->g.gy; }
    char* field(int m, long long z) {      // member m, plane z
        void* p = nullptr; int64_t cnt_ = 0, stride = 0;
        c->device_ptr(16, &p, &cnt_, &stride);
        return (char*)p + ((size_t)m * stride + (size_t)z * plane()) * rb();
    }
    char* accp(long long z) {
        void* p = nullptr; int64_t cnt_ = 0, stride = 0;
        c->device_ptr(18, &p, &cnt_, &stride);
        return (char*)p + (size_t)z * plane() * 12 * rb();
    }

    // ring exchange of `n` pairs: send up[i] to the upper and dn[i] to the lower neighbour, receive the lower neighbour's
    // "up" message into from_dn[i] and the upper neighbour's "down" message into from_up[i] -- ONE NCCL group.  With two
    // ranks both neighbours are the same peer; the op order (sends: up, down; receives: from below, from above) keeps the
    // pairs matched.  With one rank it is two device copies.
    // p_up / p_dn: where s_up / s_dn land in the upper / lower neighbour's memory (peer-memory transport); n_up / n_dn:
    // device-side element counts of a struct-of-arrays particle payload (8 member arrays of soa_cap elements each)
    struct Xfer {
        const void* s_up; const void* s_dn; void* r_dn; void* r_up; size_t bytes;
        void* p_up = nullptr; void* p_dn = nullptr;
        const long long* n_up = nullptr; const long long* n_dn = nullptr; long long soa_cap = 0;
    };
    // ---- peer-memory transport state
    bool p2p = false;
    char* mail = nullptr;                 // this rank's mailbox (exported): flags, scratch planes, counts, payload buffers
    size_t mail_bytes = 0;
    struct MailLayout { long long flags, scratch, cnt, recv_dn, recv_up, send_cap, nz, nc_pad, plane, bytes, copy; } lay{}, peer_lay[2]{};
    char* peer_mail[2] = {nullptr, nullptr};      // [0] lower, [1] upper neighbour
    void* p2p_state = nullptr;            // done (u32 @0), sent[2] (@8), expect[2] (@24), error words (long long[2] @40)
    unsigned long long p2p_timeout_ns = 120ull * 1000000000ull;
    unsigned long long p2p_skew_ns = 0;   // test hook: ranks with rank % 3 == 1 idle this long before every send
    // landing area of the copy planes in a mailbox: per kind (0: cB, 1: J) [from below: 3 planes][from above: 3 planes],
    // plane stride 256-aligned.  J and cB planes must NOT share slots: their exchanges are consecutive phases, and a
    // neighbour that has my J planes already may send its cB planes before I have unpacked its J planes (a landing
    // buffer is safe to reuse two phases later: the neighbour's next message to me is sent after its unpack).
    size_t copy_stride() const { return ((size_t)plane() * rb() + 255) & ~(size_t)255; }
    char* copy_slot(char* mailbox, const MailLayout& L, int kind, int from, int i) const {
        return mailbox + L.copy + (size_t)((kind * 2 + from) * 3 + i) * copy_stride();
    }
    char* peer_scratch(int dir, size_t off) const { return peer_mail[dir] + peer_lay[dir].scratch + off; }
    int ring_p2p(const Xfer* x, int n, const char* what) {
        PutArgs a{};
        int ns = 0;
        bool to_dn = false, to_up = false, from_dn = false, from_up = false;
        auto add = [&](const void* src, void* dst, size_t bytes, const long long* np_, long long elem) {
            if (ns < P2P_MAXSEG) a.seg[ns] = PutSeg{(const char*)src, (char*)dst, (long long)bytes, np_, elem};
            ++ns;
        };
        for (int i = 0; i < n; ++i) {
            const Xfer& t = x[i];
            if (t.soa_cap > 0) {          // particle payload: the first *n elements of each of the 8 member arrays
                for (int k = 0; k < 8; ++k) {      // (cell: int32 = 4 bytes = rb() in float; slab mode is float only)
                    const size_t off = (size_t)k * t.soa_cap * rb();
                    const long long cu = peer_lay[1].send_cap, cd = peer_lay[0].send_cap;      // the receivers carve by THEIR capacity
                    if (t.s_up) add((const char*)t.s_up + off, (char*)t.p_up + (size_t)k * cu * rb(), (size_t)std::min(t.soa_cap, cu) * rb(), t.n_up, (long long)rb());
                    if (t.s_dn) add((const char*)t.s_dn + off, (char*)t.p_dn + (size_t)k * cd * rb(), (size_t)std::min(t.soa_cap, cd) * rb(), t.n_dn, (long long)rb());
                }
            } else {
                if (t.s_up) add(t.s_up, t.p_up, t.bytes, nullptr, 0);
                if (t.s_dn) add(t.s_dn, t.p_dn, t.bytes, nullptr, 0);
            }
            to_up |= t.s_up != nullptr; to_dn |= t.s_dn != nullptr;
            from_dn |= t.r_dn != nullptr; from_up |= t.r_up != nullptr;
        }
        if (ns > P2P_MAXSEG) return fail(CPIC_E_INVALID, "%s: too many segments for one peer-memory exchange", what);
        for (int i = 0; i < ns; ++i) if (!a.seg[i].dst) return fail(CPIC_E_INVALID, "%s: peer destination missing", what);
        char* st = (char*)p2p_state;
        a.done = (unsigned*)st;
        a.sent = (unsigned long long*)(st + 8);
        a.flag[0] = to_dn ? (unsigned long long*)(peer_mail[0] + peer_lay[0].flags + 128) : nullptr;   // their "from above"
        a.flag[1] = to_up ? (unsigned long long*)(peer_mail[1] + peer_lay[1].flags) : nullptr;         // their "from below"
        if (p2p_skew_ns && rank % 3 == 1) k_p2p_skew<<<1, 1, 0, c->stream>>>(p2p_skew_ns);
        if (ns > 0) {
            k_p2p_put<<<dim3(32, (unsigned)ns), 256, 0, c->stream>>>(a);
            ++c->launches;
            int rc = c->check_launch("k_p2p_put");
            if (rc) return ctx(rc);
        }
        WaitArgs w{};
        w.flag[0] = from_dn ? (const unsigned long long*)(mail + lay.flags) : nullptr;
        w.flag[1] = from_up ? (const unsigned long long*)(mail + lay.flags + 128) : nullptr;
        w.expect = (unsigned long long*)(st + 24);
        w.err = (long long*)(st + 40);
        w.dc = c->device_counts();
        w.timeout_ns = p2p_timeout_ns;
        k_p2p_wait<<<1, 32, 0, c->stream>>>(w);
        ++c->launches;
        return ctx(c->check_launch("k_p2p_wait"));
    }
    int ring(const Xfer* x, int n, const char* what) {
        if (p2p) return ring_p2p(x, n, what);
        if (world == 1) {
            for (int i = 0; i < n; ++i) {
                if (x[i].s_up && x[i].r_dn) cudaMemcpyAsync(x[i].r_dn, x[i].s_up, x[i].bytes, cudaMemcpyDeviceToDevice, c->stream);
                if (x[i].s_dn && x[i].r_up) cudaMemcpyAsync(x[i].r_up, x[i].s_dn, x[i].bytes, cudaMemcpyDeviceToDevice, c->stream);
            }
            return c->cuda(cudaGetLastError(), what);
        }
        int rc;
        if ((rc = nccl(g_nccl.GroupStart(), what))) return rc;
        for (int i = 0; i < n; ++i) {
            if (x[i].s_up) g_nccl.Send(x[i].s_up, x[i].bytes, ncclChar, up, comm, c->stream);
            if (x[i].s_dn) g_nccl.Send(x[i].s_dn, x[i].bytes, ncclChar, down, comm, c->stream);
        }
        for (int i = 0; i < n; ++i) {
            if (x[i].r_dn) g_nccl.Recv(x[i].r_dn, x[i].bytes, ncclChar, down, comm, c->stream);
            if (x[i].r_up) g_nccl.Recv(x[i].r_up, x[i].bytes, ncclChar, up, comm, c->stream);
        }
        ++c->launches;
        return nccl(g_nccl.GroupEnd(), what);
    }
    template <class R>
    int plane_add(char* dst, const char* src, long long n) {
        k_plane_add<R><<<blocks_for(n), 256, 0, c->stream>>>((R*)dst, (const R*)src, n);
        return c->check_launch("k_plane_add");
    }
    int plane_add_any(char* dst, const char* src, long long n) {
        return rb() == 4 ? plane_add<float>(dst, src, n) : plane_add<double>(dst, src, n);
    }
    int rect_add(char* dst, const char* src, int y0, int y1, int x0, int x1) {
        const long long n = (long long)(x1 - x0) * (y1 - y0);
        if (rb() == 4) k_rect_add<float><<<blocks_for(n), 256, 0, c->stream>>>((float*)dst, (const float*)src, c->g.gx, y0, y1, x0, x1);
        else k_rect_add<double><<<blocks_for(n), 256, 0, c->stream>>>((double*)dst, (const double*)src, c->g.gx, y0, y1, x0, x1);
        return c->check_launch("k_rect_add");
    }

    // Map the two z neighbours' mailboxes (CUDA IPC; the 64-byte handles and the layouts travel once over
    // NCCL).  On any failure the NCCL transport stays in place -- but every rank must take the same decision, so the
    // outcome is agreed on with one more ring exchange.
    struct Blob { cudaIpcMemHandle_t mail; MailLayout lay; long long ok; long long pad[5]; };
    int setup_p2p() {
        Blob mine{}, *dev = nullptr;
        Blob got[2]{};
        cudaDeviceSynchronize();      // the mailbox's flags are zero before anybody can learn its address
        bool ok = cudaIpcGetMemHandle(&mine.mail, mail) == cudaSuccess;
        cudaGetLastError();
        mine.lay = lay; mine.ok = ok ? 1 : 0;
        if (cudaMalloc(&dev, 3 * sizeof(Blob)) != cudaSuccess) return fail(CPIC_E_NOMEM, "cudaMalloc(IPC handles)");
        auto exchange = [&]() -> int {
            cudaMemcpyAsync(dev, &mine, sizeof(Blob), cudaMemcpyHostToDevice, c->stream);
            Xfer x{dev, dev, dev + 1, dev + 2, sizeof(Blob)};
            int rc = ring(&x, 1, "IPC handle exchange");
            if (rc) return rc;
            cudaMemcpyAsync(got, dev + 1, 2 * sizeof(Blob), cudaMemcpyDeviceToHost, c->stream);
            return c->cuda(cudaStreamSynchronize(c->stream), "IPC handle exchange");
        };
        int rc = exchange();
        if (rc) { cudaFree(dev); return rc; }
        ok = ok && got[0].ok && got[1].ok;
        std::string why = ok ? "" : "cudaIpcGetMemHandle failed on a rank";
        if (ok) {
            for (int d = 0; d < 2 && ok; ++d) {
                if (d == 1 && up == down) { peer_mail[1] = peer_mail[0]; peer_lay[1] = peer_lay[0]; break; }
                void* pm = nullptr;
                const cudaError_t e1 = cudaIpcOpenMemHandle(&pm, got[d].mail, cudaIpcMemLazyEnablePeerAccess);
                if (e1 != cudaSuccess) {
                    why = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e1);
                    cudaGetLastError();
                    ok = false;
                    break;
                }
                peer_mail[d] = (char*)pm; peer_lay[d] = got[d].lay;
            }
        }
        if (ok && cudaMalloc(&p2p_state, 64) != cudaSuccess) { ok = false; why = "cudaMalloc"; }
        if (ok) cudaMemset(p2p_state, 0, 64);
        // second round: did everybody succeed?
        mine.ok = ok ? 1 : 0;
        rc = exchange();
        cudaFree(dev);
        if (rc) return rc;
        if (!(ok && got[0].ok && got[1].ok)) {
            teardown_p2p();
            return fail(CPIC_E_UNSUPPORTED, "%s", ok ? "a neighbour could not map peer memory" : why.c_str());
        }
        if (const char* e = getenv("CPIC_P2P_TIMEOUT_S")) p2p_timeout_ns = (unsigned long long)(atof(e) * 1e9);
        if (const char* e = getenv("CPIC_P2P_SKEW_US")) p2p_skew_ns = (unsigned long long)(atof(e) * 1e3);
        p2p = true;
        return CPIC_OK;
    }
    void teardown_p2p() {
        if (p2p && comm) {      // nobody may still be storing into my mailbox: one NCCL round trip with both neighbours
            p2p = false;
            long long* t = nullptr;
            if (cudaMalloc(&t, 4 * sizeof(long long)) == cudaSuccess) {
                Xfer x{t, t + 1, t + 2, t + 3, sizeof(long long)};
                ring(&x, 1, "teardown");
                cudaStreamSynchronize(c->stream);
                cudaFree(t);
            }
        }
        p2p = false;
        for (int d = 0; d < 2; ++d) {
            if (d == 1 && up == down) break;
            if (peer_mail[d]) cudaIpcCloseMemHandle(peer_mail[d]);
        }
        peer_mail[0] = peer_mail[1] = nullptr;
        if (p2p_state) { cudaFree(p2p_state); p2p_state = nullptr; }
        cudaGetLastError();
    }
    // did a peer-memory wait time out?  (host-synchronising callers only)
    int check_p2p() {
        if (!p2p_state) return CPIC_OK;
        long long h[2] = {0, 0};
        cudaMemcpyAsync(h, (char*)p2p_state + 40, sizeof h, cudaMemcpyDeviceToHost, c->stream);
        int rc = c->cuda(cudaStreamSynchronize(c->stream), "check_p2p");
        if (rc) return ctx(rc);
        if (h[1] & 8) return fail(CPIC_E_CUDA, "peer-memory exchange: a neighbour rank did not arrive within %.0f s", p2p_timeout_ns * 1e-9);
        return CPIC_OK;
    }

    // ---- slab pieces (example/example.cpp:248-266 with the z neighbours woven in)
    int exchange_accumulators(bool with_particles) {
        const long long nz = c->g.nz;
        const size_t ab = (size_t)plane() * 12 * rb();
        char* a_dn = scratch;
        char* a_up = scratch + ab;
        int rc;
        Xfer x[3];
        int n = 0;
        x[n++] = Xfer{accp(nz + 1), accp(0), a_dn, a_up, ab};
        if (p2p) { x[0].p_up = peer_scratch(1, 0); x[0].p_dn = peer_scratch(0, ab); }      // their a_dn / a_up
        if (with_particles) {
            const int rebase_hi = (int)(-nz * plane()), rebase_lo = (int)(nzl_down * plane());
            if ((rc = ctx(c->slab_extract_async(send_lo, send_hi, send_cap, cnt, rebase_lo, rebase_hi)))) return rc;
            x[n++] = Xfer{cnt + 1, cnt + 0, cnt + 2, cnt + 3, sizeof(long long)};
            const size_t pb = (size_t)send_cap * (7 * rb() + 4);
            x[n++] = Xfer{send_hi, send_lo, recv_dn, recv_up, pb};
            if (p2p) {
                x[1].p_up = peer_mail[1] + peer_lay[1].cnt + 2 * sizeof(long long);      // their cnt[2]: received from below
                x[1].p_dn = peer_mail[0] + peer_lay[0].cnt + 3 * sizeof(long long);      // their cnt[3]: received from above
                x[2].p_up = peer_mail[1] + peer_lay[1].recv_dn;
                x[2].p_dn = peer_mail[0] + peer_lay[0].recv_up;
                x[2].n_up = cnt + 1; x[2].n_dn = cnt + 0; x[2].soa_cap = send_cap;
            }
        }
        if ((rc = ring(x, n, "accumulator / particle exchange"))) return rc;
        if ((rc = plane_add_any(accp(1), a_dn, plane() * 12))) return rc;       // the lower neighbour's high ghost plane is my plane 1
        if ((rc = plane_add_any(accp(nz), a_up, plane() * 12))) return rc;      // the upper neighbour's low ghost plane is my plane nz
        cudaMemsetAsync(accp(0), 0, ab, c->stream);
        cudaMemsetAsync(accp(nz + 1), 0, ab, c->stream);
        if (with_particles) {
            if ((rc = ctx(c->slab_append_async(recv_dn, send_cap, cnt + 2)))) return rc;
            if ((rc = ctx(c->slab_append_async(recv_up, send_cap, cnt + 3)))) return rc;
            k_count_accumulate<<<1, 1, 0, c->stream>>>(cnt, cnt + 4);
            if ((rc = c->check_launch("k_count_accumulate"))) return rc;
        }
        return CPIC_OK;
    }
    // ghost copy along z of three consecutive members m0..m0+2: my top plane is the upper neighbour's ghost plane 0, my
    // plane 1 the lower neighbour's ghost plane nz+1 (src/fields.h:80-98); received in place
    int exchange_copy_planes(int m0) {
        const long long nz = c->g.nz;
        const size_t pb = (size_t)plane() * rb();
        const int kind = m0 == F_JFX ? 1 : 0;
        Xfer x[3];
        for (int i = 0; i < 3; ++i) {
            x[i] = Xfer{field(m0 + i, nz), field(m0 + i, 1), field(m0 + i, 0), field(m0 + i, nz + 1), pb};
            if (p2p) {      // into the neighbours' mailboxes: my top plane is "from below" up there, my plane 1 "from above" down there
                x[i].p_up = copy_slot(peer_mail[1], peer_lay[1], kind, 0, i);
                x[i].p_dn = copy_slot(peer_mail[0], peer_lay[0], kind, 1, i);
            }
        }
        int rc = ring(x, 3, "ghost-plane copy");
        if (rc || !p2p) return rc;
        UnpackArgs u{};
        for (int i = 0; i < 3; ++i) {
            u.src[i] = copy_slot(mail, lay, kind, 0, i); u.dst[i] = field(m0 + i, 0);
            u.src[3 + i] = copy_slot(mail, lay, kind, 1, i); u.dst[3 + i] = field(m0 + i, nz + 1);
        }
        u.bytes = (long long)pb;
        k_planes_unpack<<<dim3(8, 6), 256, 0, c->stream>>>(u);
        ++c->launches;
        return ctx(c->check_launch("k_planes_unpack"));
    }
    int slab_advance_b(double hx, double hy, double hz) {
        int rc;
        if ((rc = ctx(c->stencil_only(0, hx, hy, hz, 0.0)))) return rc;      // src/fields.h:692-717
        if ((rc = ctx(c->update_ghosts(2)))) return rc;                      // :718, x and y faces
        return exchange_copy_planes(F_CBX);                                  // z faces
    }
    int slab_advance_e(const cpic_consts& k) {
        const long long nz = c->g.nz;
        const int nx = c->g.nx, ny = c->g.ny;
        const size_t pb = (size_t)plane() * rb();
        int rc;
        // periodic fold of J (src/fields.h:126-183); the z sweeps are the plane exchange: jfy's comes first, jfx's second
        if ((rc = ctx(c->fold_phase(0)))) return rc;
        char* rx = scratch;
        char* ry = scratch + pb;
        Xfer x[2] = {Xfer{field(F_JFX, nz + 1), nullptr, rx, nullptr, pb}, Xfer{field(F_JFY, nz + 1), nullptr, ry, nullptr, pb}};
        if (p2p) { x[0].p_up = peer_scratch(1, 0); x[1].p_up = peer_scratch(1, pb); }      // their rx / ry
        if ((rc = ring(x, 2, "J fold planes"))) return rc;
        if ((rc = rect_add(field(F_JFY, 1), ry, 1, ny + 1, 1, nx + 2))) return rc;      // :146-151
        if ((rc = ctx(c->fold_phase(1)))) return rc;
        if ((rc = rect_add(field(F_JFX, 1), rx, 1, ny + 2, 1, nx + 1))) return rc;      // :136-141
        // ghost copy of J (:643): x, y locally, z planes from the neighbours; then the E stencil (:646-664)
        if ((rc = ctx(c->update_ghosts(1)))) return rc;
        if ((rc = exchange_copy_planes(F_JFX))) return rc;
        return ctx(c->stencil_only(1, k.px, k.py, k.pz, k.dt_eps0));
    }
    void halves(const cpic_consts& k, double& hx, double& hy, double& hz) const {
        if (c->prm.real_bytes == 4) { hx = (double)(0.5f * (float)k.px); hy = (double)(0.5f * (float)k.py); hz = (double)(0.5f * (float)k.pz); }
        else { hx = 0.5 * k.px; hy = 0.5 * k.py; hz = 0.5 * k.pz; }
    }
    int slab_step(const cpic_consts& k) {
        int rc;
        double hx, hy, hz;
        halves(k, hx, hy, hz);
        if ((rc = ctx(c->load_interpolator()))) return rc;
        if ((rc = ctx(c->clear_accumulator()))) return rc;
        if ((rc = ctx(c->push_reorder(k)))) return rc;
        if ((rc = exchange_accumulators(true))) return rc;
        if ((rc = ctx(c->unload_accumulator(k)))) return rc;
        if ((rc = slab_advance_b(hx, hy, hz))) return rc;
        if ((rc = slab_advance_e(k))) return rc;
        return slab_advance_b(hx, hy, hz);
    }
    // One step for a caller whose slab lives in HOST memory: the particles stream through the device chunk by chunk
    // (H2D / in-place push / D2H overlapped, Ctx::host_push_phase), then the usual exchanges and the field side, then the
    // host copy is patched where the migration changed the store (holes filled from the tail, arrivals appended).
    int slab_step_host(const cpic_consts& k, const void* const in[8], void* const out[8], long long n, long long out_cap,
                       long long* n_out, const void* const fin[9], void* const fout[9]) {
        int rc;
        double hx, hy, hz;
        halves(k, hx, hy, hz);
        if ((rc = ctx(c->host_push_phase(k, in, out, n, fin, true)))) return rc;
        if ((rc = exchange_accumulators(true))) return rc;
        if ((rc = ctx(c->unload_accumulator(k)))) return rc;
        if ((rc = slab_advance_b(hx, hy, hz))) return rc;
        if ((rc = slab_advance_e(k))) return rc;
        if ((rc = slab_advance_b(hx, hy, hz))) return rc;
        if ((rc = ctx(c->host_fields_out(fout)))) return rc;
        if ((rc = ctx(c->host_patch_phase(out, n, out_cap, n_out)))) return rc;
        return check_p2p();
    }
    int replicated_step(const cpic_consts& k, bool sort, bool fused) {
        int rc;
        double hx, hy, hz;
        halves(k, hx, hy, hz);
        if (sort && (rc = ctx(c->sort()))) return rc;
        if ((rc = ctx(c->load_interpolator()))) return rc;
        if ((rc = ctx(c->clear_accumulator()))) return rc;
        if ((rc = ctx(fused ? c->push_reorder(k) : c->push(k)))) return rc;
        if ((rc = reduce_accumulator())) return rc;
        if ((rc = ctx(c->unload_accumulator(k)))) return rc;
        if ((rc = ctx(c->advance_b(hx, hy, hz)))) return rc;
        if ((rc = ctx(c->advance_e(k.px, k.py, k.pz, k.dt_eps0)))) return rc;
        return ctx(c->advance_b(hx, hy, hz));
    }
    int reduce_accumulator() {
        if (mode == CPIC_MGPU_SLAB) return exchange_accumulators(false);
        if (world == 1) return CPIC_OK;
        ++c->launches;
        return nccl(g_nccl.AllReduce(accp(0), accp(0), (size_t)c->g.nc * 12, rtype(), ncclSum, comm, c->stream), "ncclAllReduce(accumulator)");
    }

    // a pair of fused slab steps as one CUDA graph (the particle double buffer and the cell-count ping-pong are back
    // where they started after two steps)
    int capture_pair(const cpic_consts& k) {
        cudaGraph_t g = nullptr;
        const long long l0 = c->launches;
        c->capturing = true;
        cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
        int rc = e == cudaSuccess ? CPIC_OK : c->cuda(e, "cudaStreamBeginCapture");
        if (!rc) rc = slab_step(k);
        if (!rc) rc = slab_step(k);
        cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
        c->capturing = false;
        graph_launches = c->launches - l0;
        c->launches = l0;                    // captured, not executed
        if (rc || e2 != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            return rc ? rc : fail(CPIC_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e2));
        }
        e = cudaGraphInstantiate(&gexec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { gexec = nullptr; cudaGetLastError(); return fail(CPIC_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
        graph_k = k;
        graph_sig = c->state_signature();      // (a pair of steps returns to this state)
        return CPIC_OK;
    }
    long long graph_launches = 0;
};

bool same_consts(const cpic_consts& a, const cpic_consts& b) { return memcmp(&a, &b, sizeof a) == 0; }

#define MGPU_OR_FAIL(m_) \
    if (!(m_)) return CPIC_E_INVALID; \
    Mgpu* m = reinterpret_cast<Mgpu*>(m_); \
    { cudaError_t e_ = cudaSetDevice(m->c->prm.device); if (e_ != cudaSuccess) return m->fail(CPIC_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e_)); }

}  // namespace

extern "C" {

const char* cpic_mgpu_last_error(const cpic_mgpu* m) {
    if (!m) return g_mgpu_error.c_str();
    return reinterpret_cast<const Mgpu*>(m)->err.c_str();
}

int cpic_mgpu_unique_id(void* id_out) {
    if (!id_out) { g_mgpu_error = "unique_id: null"; return CPIC_E_INVALID; }
    if (!g_nccl.load()) { g_mgpu_error = g_nccl.why; return CPIC_E_UNSUPPORTED; }
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) { g_mgpu_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return CPIC_E_CUDA; }
    memcpy(id_out, &id, sizeof id);
    return CPIC_OK;
}

int cpic_mgpu_bootstrap_file(const char* path, int32_t rank, int32_t world, double timeout_s, void* id_out) {
    if (!path || !id_out || rank < 0 || rank >= world) { g_mgpu_error = "bootstrap_file: bad arguments"; return CPIC_E_INVALID; }
    if (rank == 0) {
        int rc = cpic_mgpu_unique_id(id_out);
        if (rc) return rc;
        const std::string tmp = std::string(path) + ".tmp";
        FILE* f = fopen(tmp.c_str(), "wb");
        if (!f || fwrite(id_out, 1, CPIC_MGPU_ID_BYTES, f) != CPIC_MGPU_ID_BYTES) { if (f) fclose(f); g_mgpu_error = "bootstrap_file: cannot write " + tmp; return CPIC_E_INVALID; }
        fclose(f);
        if (rename(tmp.c_str(), path) != 0) { g_mgpu_error = std::string("bootstrap_file: cannot rename to ") + path; return CPIC_E_INVALID; }
        return CPIC_OK;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        FILE* f = fopen(path, "rb");
        if (f) {
            const size_t n = fread(id_out, 1, CPIC_MGPU_ID_BYTES, f);
            fclose(f);
            if (n == CPIC_MGPU_ID_BYTES) return CPIC_OK;
        }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
            g_mgpu_error = std::string("bootstrap_file: timed out waiting for ") + path;
            return CPIC_E_INVALID;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
}

int cpic_mgpu_create(const cpic_params* global, int32_t rank, int32_t world, const void* unique_id, int32_t mode,
                     int64_t send_capacity, cpic_mgpu** out) {
    if (!global || !out || world < 1 || rank < 0 || rank >= world || (world > 1 && !unique_id)) { g_mgpu_error = "mgpu_create: bad arguments"; return CPIC_E_INVALID; }
    *out = nullptr;
    if (mode == CPIC_MGPU_AUTO) {
        const long long nc = (long long)(global->nx + 2) * (global->ny + 2) * (global->nz + 2);
        mode = (world > 1 && global->nz >= 2 * world && nc >= (1ll << 18)) ? CPIC_MGPU_SLAB : CPIC_MGPU_REPLICATED;
    }
    if (mode != CPIC_MGPU_SLAB && mode != CPIC_MGPU_REPLICATED) { g_mgpu_error = "mgpu_create: unknown mode"; return CPIC_E_INVALID; }
    if (mode == CPIC_MGPU_SLAB && global->boundary != CPIC_BOUNDARY_PERIODIC) { g_mgpu_error = "mgpu_create: slab mode is periodic only"; return CPIC_E_UNSUPPORTED; }
    if (mode == CPIC_MGPU_SLAB) {
        if (global->real_bytes != 4 || global->solver != CPIC_SOLVER_EM || !global->enable_sort) { g_mgpu_error = "mgpu_create: slab mode needs float, the EM solver and enable_sort"; return CPIC_E_UNSUPPORTED; }
        if (global->nz < world) { g_mgpu_error = "mgpu_create: fewer z-planes than ranks"; return CPIC_E_INVALID; }
    }
    if (world > 1 && !g_nccl.load()) { g_mgpu_error = g_nccl.why; return CPIC_E_UNSUPPORTED; }
    Mgpu* m = new Mgpu();
    m->gprm = *global; m->rank = rank; m->world = world; m->mode = mode;
    m->up = (rank + 1) % world; m->down = (rank + world - 1) % world;
    cpic_params lp = *global;
    m->z0 = 0; m->nzl = global->nz; m->nzl_down = global->nz;
    if (mode == CPIC_MGPU_SLAB) {      // balanced contiguous z ranges
        const int base = global->nz / world, rem = global->nz % world;
        auto nz_of = [&](int r) { return base + (r < rem ? 1 : 0); };
        m->nzl = nz_of(rank);
        m->z0 = rank * base + std::min(rank, rem);
        m->nzl_down = nz_of(m->down);
        lp.nz = m->nzl;
    }
    cpic_ctx* ctx = nullptr;
    int rc = cpic_create(&lp, &ctx);
    if (rc) { g_mgpu_error = cpic_last_error(nullptr); delete m; return rc; }
    m->c = reinterpret_cast<CtxBase*>(ctx);
    CtxBase* c = m->c;
    auto bail = [&](int code, const std::string& why) { g_mgpu_error = why; cpic_mgpu_destroy(reinterpret_cast<cpic_mgpu*>(m)); return code; };
    if (mode == CPIC_MGPU_SLAB && (world > 1 || getenv("CPIC_MGPU_OPEN_Z"))) c->g.per = 3;      // z belongs to the neighbours
    const bool open_z = mode == CPIC_MGPU_SLAB && !(c->g.per & 4);
    if (world > 1) {
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof id);
        const ncclResult_t r = g_nccl.CommInitRank(&m->comm, world, id, rank);
        if (r != ncclSuccess) return bail(CPIC_E_CUDA, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
    }
    if (cudaMalloc(&m->diag, 8 * sizeof(double)) != cudaSuccess) return bail(CPIC_E_NOMEM, "cudaMalloc");
    if (open_z) {
        const long long per_plane_cap = global->max_particles / std::max(1, m->nzl);
        m->send_cap = send_capacity > 0 ? send_capacity : std::max<long long>(4096, per_plane_cap / 20);
        m->send_cap = (m->send_cap + 63) & ~63ll;      // member arrays of the payload stay 16-byte aligned
        const size_t pb = (size_t)m->send_cap * (7 * m->rb() + 4);
        for (char** b : {&m->send_lo, &m->send_hi})
            if (cudaMalloc(b, pb) != cudaSuccess) return bail(CPIC_E_NOMEM, "cudaMalloc(migration buffers)");
        // everything a neighbour writes into lives in ONE allocation, the mailbox (exported through CUDA IPC below)
        auto up256 = [](size_t v) { return (v + 255) & ~(size_t)255; };
        Mgpu::MailLayout& L = m->lay;
        size_t off = 0;
        L.flags = (long long)off; off += 256;
        L.cnt = (long long)off; off += 256;
        L.scratch = (long long)off; off = up256(off + (size_t)m->plane() * 12 * m->rb() * 2);
        L.recv_dn = (long long)off; off = up256(off + pb);
        L.recv_up = (long long)off; off = up256(off + pb);
        L.copy = (long long)off; off = up256(off + 12 * m->copy_stride());
        L.bytes = (long long)off; L.send_cap = m->send_cap; L.nz = c->g.nz; L.plane = m->plane();
        { void* fp = nullptr; int64_t n_ = 0, stride = 0; c->device_ptr(16, &fp, &n_, &stride); L.nc_pad = stride; }
        if (cudaMalloc(&m->mail, off) != cudaSuccess) return bail(CPIC_E_NOMEM, "cudaMalloc(mailbox)");
        m->mail_bytes = off;
        cudaMemset(m->mail, 0, 512);
        m->cnt = reinterpret_cast<long long*>(m->mail + L.cnt);
        m->scratch = m->mail + L.scratch;
        m->recv_dn = m->mail + L.recv_dn;
        m->recv_up = m->mail + L.recv_up;
        if (world > 1) {
            const char* e = getenv("CPIC_MGPU_P2P");
            if (!e || atoi(e) != 0) {
                const int prc = m->setup_p2p();
                if (prc) fprintf(stderr, "[cabanapic_b200 rank %d] peer-memory exchange unavailable (%s): using NCCL send/recv\n", rank, m->err.c_str());
            }
        }
    }
    *out = reinterpret_cast<cpic_mgpu*>(m);
    return CPIC_OK;
}

void cpic_mgpu_destroy(cpic_mgpu* mm) {
    if (!mm) return;
    Mgpu* m = reinterpret_cast<Mgpu*>(mm);
    if (m->c) { cudaSetDevice(m->c->prm.device); cudaStreamSynchronize(m->c->stream); }
    if (m->gexec) cudaGraphExecDestroy(m->gexec);      // the graph references NCCL work: release it before the communicator
    m->teardown_p2p();
    if (m->comm) g_nccl.CommDestroy(m->comm);
    cudaFree(m->send_lo); cudaFree(m->send_hi); cudaFree(m->mail); cudaFree(m->diag);
    if (m->c) cpic_destroy(reinterpret_cast<cpic_ctx*>(m->c));
    delete m;
}

int cpic_mgpu_transport(const cpic_mgpu* mm) {
    if (!mm) return CPIC_MGPU_TRANSPORT_NONE;
    const Mgpu* m = reinterpret_cast<const Mgpu*>(mm);
    if (m->world == 1) return CPIC_MGPU_TRANSPORT_NONE;
    return m->p2p ? CPIC_MGPU_TRANSPORT_PEER_MEMORY : CPIC_MGPU_TRANSPORT_NCCL;
}

cpic_ctx* cpic_mgpu_context(cpic_mgpu* m) { return m ? reinterpret_cast<cpic_ctx*>(reinterpret_cast<Mgpu*>(m)->c) : nullptr; }

int cpic_mgpu_layout(const cpic_mgpu* mm, int32_t* mode, int32_t* z0, int32_t* nzl) {
    if (!mm) return CPIC_E_INVALID;
    const Mgpu* m = reinterpret_cast<const Mgpu*>(mm);
    if (mode) *mode = m->mode;
    if (z0) *z0 = m->z0;
    if (nzl) *nzl = m->nzl;
    return CPIC_OK;
}

int cpic_mgpu_init_uniform_plasma(cpic_mgpu* mm, int32_t nppc, uint64_t seed, double vthx, double vthy, double vthz, double weight) {
    MGPU_OR_FAIL(mm);
    const cpic_params& g = m->gprm;
    const long long per_plane = (long long)g.nx * g.ny * nppc, total = per_plane * g.nz;
    long long first, count;
    if (m->mode == CPIC_MGPU_SLAB) { first = m->z0 * per_plane; count = m->nzl * per_plane; }
    else { first = total * m->rank / m->world; count = total * (m->rank + 1) / m->world - first; }
    return m->ctx(cpic_init_uniform_plasma(reinterpret_cast<cpic_ctx*>(m->c), first, count, g.nx, g.ny, g.nz, nppc, m->z0, seed, vthx, vthy, vthz, weight));
}

int cpic_mgpu_reduce_accumulator(cpic_mgpu* mm) {
    MGPU_OR_FAIL(mm);
    if (m->mode == CPIC_MGPU_SLAB && (m->c->g.per & 4)) return CPIC_OK;      // one periodic slab: nothing to exchange
    return m->reduce_accumulator();
}

int cpic_mgpu_step(cpic_mgpu* mm, const cpic_consts* k, int64_t nsteps, int32_t sort_interval, int32_t use_graph) {
    MGPU_OR_FAIL(mm);
    if (!k || nsteps < 0 || sort_interval < CPIC_SORT_FUSED) return m->fail(CPIC_E_INVALID, "mgpu_step: bad arguments");
    CtxBase* c = m->c;
    int rc = CPIC_OK;
    m->used_graph = false;
    c->rec(6);      // cpic_last_ms(ctx, 3): device time of this call
    struct Stamp { CtxBase* c; ~Stamp() { c->rec(7); c->ev_valid[3] = true; } } stamp{c};
    if (m->mode == CPIC_MGPU_SLAB) {
        if (c->g.per & 4) {      // a single periodic slab is the ordinary fused loop
            return m->ctx(cpic_step(reinterpret_cast<cpic_ctx*>(c), k, nsteps, sort_interval, nullptr));
        }
        if (sort_interval != CPIC_SORT_FUSED) return m->fail(CPIC_E_UNSUPPORTED, "mgpu_step: slab mode runs the reordering push (sort_interval = CPIC_SORT_FUSED) only");
        int64_t s = 0;
        if (use_graph && !m->graph_failed && nsteps >= 2 && m->steps_done >= 2 && c->dev_count && c->seg_valid) {
            // an odd number of eager steps (or a host-resident step) since the capture leaves the other halves of the
            // double buffers current: the graph would read stale ones -- capture again
            if (m->gexec && (!same_consts(*k, m->graph_k) || m->graph_sig != c->state_signature())) { cudaGraphExecDestroy(m->gexec); m->gexec = nullptr; }
            if (!m->gexec && (rc = m->capture_pair(*k))) {
                fprintf(stderr, "[cabanapic_b200 rank %d] CUDA-graph capture of the slab step failed, running eagerly: %s\n", m->rank, m->err.c_str());
                m->graph_failed = true;
                rc = CPIC_OK;
            }
            if (m->gexec) {
                for (; s + 2 <= nsteps; s += 2) {
                    if (cudaGraphLaunch(m->gexec, c->stream) != cudaSuccess) return m->fail(CPIC_E_CUDA, "cudaGraphLaunch: %s", cudaGetErrorString(cudaGetLastError()));
                    c->launches += m->graph_launches;
                }
                m->used_graph = s > 0;
            }
        }
        for (; s < nsteps && !rc; ++s) rc = m->slab_step(*k);
        m->steps_done += nsteps;
        return rc;
    }
    for (int64_t s = 0; s < nsteps && !rc; ++s) {
        const bool fused = sort_interval == CPIC_SORT_FUSED && !c->few_cells();
        const bool sort = (sort_interval > 0 && s % sort_interval == 0) || (sort_interval == CPIC_SORT_FUSED && c->few_cells() && c->fb_steps++ % 8 == 0);
        rc = m->replicated_step(*k, sort, fused);
    }
    m->steps_done += nsteps;
    return rc;
}

int cpic_mgpu_step_host(cpic_mgpu* mm, const cpic_consts* k, const void* const in[8], void* const out[8], int64_t n,
                        int64_t out_capacity, int64_t* n_out, const void* const fields_in[9], void* const fields_out[9]) {
    MGPU_OR_FAIL(mm);
    if (!k || !in || !fields_in || n < 0) return m->fail(CPIC_E_INVALID, "mgpu_step_host: bad arguments");
    for (int f = 0; f < 9; ++f) if (!fields_in[f]) return m->fail(CPIC_E_INVALID, "mgpu_step_host: null field member %d", f);
    if (n > 0) for (int f = 0; f < 8; ++f) if (!in[f]) return m->fail(CPIC_E_INVALID, "mgpu_step_host: null particle member %d", f);
    CtxBase* c = m->c;
    if (m->mode != CPIC_MGPU_SLAB) return m->fail(CPIC_E_UNSUPPORTED, "mgpu_step_host: slab mode only");
    m->used_graph = false;
    long long no = n;
    int rc;
    if (c->g.per & 4) rc = m->ctx(c->step_host(*k, in, out, n, fields_in, fields_out, nullptr));      // one periodic slab: the single-GPU path
    else rc = m->slab_step_host(*k, in, out, n, out_capacity, &no, fields_in, fields_out);
    if (!rc && n_out) *n_out = no;
    if (!rc) m->steps_done += 1;
    return rc;
}

int cpic_mgpu_prepare_graph(cpic_mgpu* mm, const cpic_consts* k) {
    MGPU_OR_FAIL(mm);
    if (!k) return m->fail(CPIC_E_INVALID, "prepare_graph: null consts");
    CtxBase* c = m->c;
    if (m->mode != CPIC_MGPU_SLAB || (c->g.per & 4) || m->graph_failed || m->steps_done < 2 || !c->dev_count || !c->seg_valid)
        return m->fail(CPIC_E_UNSUPPORTED, "prepare_graph: needs slab mode after at least two eager fused steps");
    if (m->gexec && same_consts(*k, m->graph_k) && m->graph_sig == c->state_signature()) return CPIC_OK;
    if (m->gexec) { cudaGraphExecDestroy(m->gexec); m->gexec = nullptr; }
    const int rc = m->capture_pair(*k);
    if (rc) m->graph_failed = true;
    return rc;
}

static int mgpu_counts(Mgpu* m, int off, int64_t out[2]) {
    out[0] = out[1] = 0;
    if (!m->cnt) return CPIC_OK;
    long long h[2];
    int rc = m->c->cuda(cudaMemcpyAsync(h, m->cnt + off, sizeof h, cudaMemcpyDeviceToHost, m->c->stream), "D2H counts");
    if (!rc) rc = m->c->cuda(cudaStreamSynchronize(m->c->stream), "migration_counts");
    if (rc) return m->ctx(rc);
    out[0] = h[0]; out[1] = h[1];
    return m->check_p2p();
}
int cpic_mgpu_migration_counts(cpic_mgpu* mm, int64_t out[2]) { MGPU_OR_FAIL(mm); if (!out) return CPIC_E_INVALID; return mgpu_counts(m, 4, out); }
int cpic_mgpu_last_migration(cpic_mgpu* mm, int64_t out[2]) { MGPU_OR_FAIL(mm); if (!out) return CPIC_E_INVALID; return mgpu_counts(m, 0, out); }

int cpic_mgpu_state_digest(cpic_mgpu* mm, double out[8]) {
    MGPU_OR_FAIL(mm);
    if (!out) return CPIC_E_INVALID;
    CtxBase* c = m->c;
    int rc = m->ctx(c->digest_async(m->diag));
    if (rc) return rc;
    double h[8];
    if ((rc = m->ctx(c->cuda(cudaMemcpyAsync(h, m->diag, sizeof h, cudaMemcpyDeviceToHost, c->stream), "D2H digest")))) return rc;
    if ((rc = m->ctx(c->cuda(cudaStreamSynchronize(c->stream), "state_digest")))) return rc;
    h[5] *= 0.5; h[6] = c->prm.solver == CPIC_SOLVER_EM ? 0.5 * h[6] : 0.0;
    h[7] = 0.0;
    if (m->cnt) { int64_t t[2]; if ((rc = mgpu_counts(m, 4, t))) return rc; h[7] = (double)(t[0] + t[1]); }
    if (m->world > 1) {
        // fields (and their energies) are replicated in REPLICATED mode: every rank already holds the total
        const double e5 = h[5], e6 = h[6];
        if ((rc = m->ctx(c->cuda(cudaMemcpyAsync(m->diag, h, sizeof h, cudaMemcpyHostToDevice, c->stream), "H2D digest")))) return rc;
        if ((rc = m->nccl(g_nccl.AllReduce(m->diag, m->diag, 8, ncclDouble, ncclSum, m->comm, c->stream), "ncclAllReduce(digest)"))) return rc;
        if ((rc = m->ctx(c->cuda(cudaMemcpyAsync(h, m->diag, sizeof h, cudaMemcpyDeviceToHost, c->stream), "D2H digest")))) return rc;
        if ((rc = m->ctx(c->cuda(cudaStreamSynchronize(c->stream), "state_digest")))) return rc;
        if (m->mode == CPIC_MGPU_REPLICATED) { h[5] = e5; h[6] = e6; }
    }
    memcpy(out, h, sizeof h);
    return CPIC_OK;
}

int cpic_mgpu_energies(cpic_mgpu* mm, double* e_energy, double* b_energy) {
    double d[8];
    int rc = cpic_mgpu_state_digest(mm, d);
    if (rc) return rc;
    if (e_energy) *e_energy = d[5];
    if (b_energy) *b_energy = d[6];
    return CPIC_OK;
}

int cpic_mgpu_sync(cpic_mgpu* mm) { MGPU_OR_FAIL(mm); return m->ctx(m->c->cuda(cudaStreamSynchronize(m->c->stream), "sync")); }
int cpic_mgpu_used_graph(const cpic_mgpu* mm) { return mm && reinterpret_cast<const Mgpu*>(mm)->used_graph ? 1 : 0; }

/* single-context form of the digest (cabanapic_b200.h) */
int cpic_state_digest(cpic_ctx* ctx, double out[8]) {
    CTX_OR_FAIL(ctx);
    if (!out) return c->fail(CPIC_E_INVALID, "state_digest: null");
    double* d = c->energy_scratch();
    int rc = c->digest_async(d);
    if (rc) return rc;
    double h[8];
    if ((rc = c->cuda(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream), "D2H digest"))) return rc;
    if ((rc = c->cuda(cudaStreamSynchronize(c->stream), "state_digest"))) return rc;
    h[5] *= 0.5; h[6] = c->prm.solver == CPIC_SOLVER_EM ? 0.5 * h[6] : 0.0; h[7] = 0.0;
    memcpy(out, h, sizeof h);
    return CPIC_OK;
}

}  // extern "C"
