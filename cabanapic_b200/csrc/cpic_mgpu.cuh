// Multi-GPU layer of the C ABI (include/cabanapic_b200_mgpu.h): one process per GPU, NCCL called from C++.
// Included at the end of cpic_capi.cu (it drives a context through CtxBase).  The exchange choreography is the one
// cabanapic_b200/dist.py prototyped in Python + torch in round 1; here it is host C++ and three small kernels:
//   * copy planes (J ghost copy, cB ghost copy) are sent from and received INTO the field arrays themselves -- a
//     z-plane of a member is one contiguous run of gx*gy reals -- so they need no pack / unpack at all;
//   * add planes (ghost accumulator planes, the z sweeps of the J fold) are received into scratch and added by
//     k_plane_add / k_rect_add;
//   * the particle migration uses the device-counted extraction / append of the context (cpic_slab_*_async).
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
    struct Xfer { const void* s_up; const void* s_dn; void* r_dn; void* r_up; size_t bytes; };
    int ring(const Xfer* x, int n, const char* what) {
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
        if (with_particles) {
            const int rebase_hi = (int)(-nz * plane()), rebase_lo = (int)(nzl_down * plane());
            if ((rc = ctx(c->slab_extract_async(send_lo, send_hi, send_cap, cnt, rebase_lo, rebase_hi)))) return rc;
            x[n++] = Xfer{cnt + 1, cnt + 0, cnt + 2, cnt + 3, sizeof(long long)};
            const size_t pb = (size_t)send_cap * (7 * rb() + 4);
            x[n++] = Xfer{send_hi, send_lo, recv_dn, recv_up, pb};
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
        Xfer x[3];
        for (int i = 0; i < 3; ++i) x[i] = Xfer{field(m0 + i, nz), field(m0 + i, 1), field(m0 + i, 0), field(m0 + i, nz + 1), pb};
        return ring(x, 3, "ghost-plane copy");
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
        const size_t pb = (size_t)m->send_cap * (7 * m->rb() + 4);
        for (char** b : {&m->send_lo, &m->send_hi, &m->recv_dn, &m->recv_up})
            if (cudaMalloc(b, pb) != cudaSuccess) return bail(CPIC_E_NOMEM, "cudaMalloc(migration buffers)");
        if (cudaMalloc(&m->cnt, 6 * sizeof(long long)) != cudaSuccess) return bail(CPIC_E_NOMEM, "cudaMalloc");
        cudaMemset(m->cnt, 0, 6 * sizeof(long long));
        if (cudaMalloc(&m->scratch, (size_t)m->plane() * 12 * m->rb() * 2) != cudaSuccess) return bail(CPIC_E_NOMEM, "cudaMalloc(plane scratch)");
    }
    *out = reinterpret_cast<cpic_mgpu*>(m);
    return CPIC_OK;
}

void cpic_mgpu_destroy(cpic_mgpu* mm) {
    if (!mm) return;
    Mgpu* m = reinterpret_cast<Mgpu*>(mm);
    if (m->c) { cudaSetDevice(m->c->prm.device); cudaStreamSynchronize(m->c->stream); }
    if (m->gexec) cudaGraphExecDestroy(m->gexec);      // the graph references NCCL work: release it before the communicator
    if (m->comm) g_nccl.CommDestroy(m->comm);
    cudaFree(m->send_lo); cudaFree(m->send_hi); cudaFree(m->recv_dn); cudaFree(m->recv_up);
    cudaFree(m->cnt); cudaFree(m->scratch); cudaFree(m->diag);
    if (m->c) cpic_destroy(reinterpret_cast<cpic_ctx*>(m->c));
    delete m;
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
            if (m->gexec && !same_consts(*k, m->graph_k)) { cudaGraphExecDestroy(m->gexec); m->gexec = nullptr; }
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

int cpic_mgpu_prepare_graph(cpic_mgpu* mm, const cpic_consts* k) {
    MGPU_OR_FAIL(mm);
    if (!k) return m->fail(CPIC_E_INVALID, "prepare_graph: null consts");
    CtxBase* c = m->c;
    if (m->mode != CPIC_MGPU_SLAB || (c->g.per & 4) || m->graph_failed || m->steps_done < 2 || !c->dev_count || !c->seg_valid)
        return m->fail(CPIC_E_UNSUPPORTED, "prepare_graph: needs slab mode after at least two eager fused steps");
    if (m->gexec && same_consts(*k, m->graph_k)) return CPIC_OK;
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
    return CPIC_OK;
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
