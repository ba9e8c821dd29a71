// C ABI of cabanapic_b200 (include/cabanapic_b200.h): context, device memory, launch logic.
// Host side is plain C++; all compute is in the hand-written sm_100a kernels included below.
// There is deliberately no CPU fallback anywhere in this file.
#include "../../include/cabanapic_b200.h"

#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "cpic_common.cuh"
#include "cpic_fields.cuh"
#include "cpic_particles.cuh"
#include "cpic_push2.cuh"
#include "cpic_push3.cuh"
#include "cpic_sort.cuh"
#include "cpic_init.cuh"
#include "cpic_migrate.cuh"

using namespace cpic;

namespace {

thread_local std::string g_create_error;

struct CtxBase {
    cpic_params prm{};
    Grid g{};
    std::string err;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    long long np = 0;
    long long launches = 0;
    cudaEvent_t ev[8]{};   // 0/1 push, 2/3 sort, 4/5 field side, 6/7 step
    bool ev_valid[4] = {false, false, false, false};
    bool capturing = false;    // the stream is being captured into a CUDA graph (cpic_mgpu_step): timing events become external nodes
    void rec(int i) { cudaEventRecordWithFlags(ev[i], stream, capturing ? cudaEventRecordExternal : cudaEventRecordDefault); }
    bool want_stats = false;
    bool want_hist = false;    // set by cpic_step before a push that is followed by a sort
    bool hist_valid = false;   // cell_count holds the histogram of the current cells (from the last push)
    bool cursor_valid = false; // cell_count holds the exclusive scan of that histogram (ready for a reordering push)
    bool seg_valid = false;    // the store is the concatenation of the per-cell segments seg[seg_cur] (k_push3), up to the
                               // hole filling / appended tail of a slab migration
    bool leavers_valid = false; // slab mode: the last push listed the particles it left in the z ghost planes
    bool ghost_clean = false;   // no particle sits in a z ghost plane (true after init_uniform_plasma or an extraction;
                                // unknown after an upload): only then does the mover's list find every leaver
    // opt-in per-phase profile of cpic_step: 5 events per step (start, after sort, before push,
    // after push, end) on the context's stream
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    long long prof_steps = 0;
    virtual ~CtxBase() { for (auto e : prof_ev) cudaEventDestroy(e); }

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int cuda(cudaError_t e, const char* what) {
        if (e == cudaSuccess) return CPIC_OK;
        return fail(e == cudaErrorMemoryAllocation ? CPIC_E_NOMEM : CPIC_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    int check_launch(const char* what) {
        ++launches;
        return cuda(cudaGetLastError(), what);
    }

    virtual int upload_particles(const void* const m[7], const int32_t* cell, long long n) = 0;
    virtual int download_particles(void* const m[7], int32_t* cell, long long cap, long long* n_out) = 0;
    virtual int upload_fields(const void* const f[9]) = 0;
    virtual int download_fields(void* const f[9]) = 0;
    virtual int xfer_interp(void* host, bool up) = 0;
    virtual int xfer_acc(void* host, bool up) = 0;
    virtual int load_interpolator() = 0;
    virtual int initialize_interpolator() = 0;
    virtual int clear_accumulator() = 0;
    virtual int push(const cpic_consts& k) = 0;
    virtual int unload_accumulator(const cpic_consts& k) = 0;
    virtual int advance_b(double px, double py, double pz) = 0;
    virtual int advance_e(double px, double py, double pz, double dt_eps0) = 0;
    virtual int uncenter(double qdt_2mc) = 0;
    virtual int energies_async(double* dev_out2) = 0;
    virtual int kinetic_async(double* dev_out) = 0;
    virtual int digest_async(double* dev_out8) = 0;     // [0..4] from k_state_digest, [5] [6] field energy sums (not halved)
    virtual int update_ghosts(int which) = 0;
    virtual int sort() = 0;
    virtual int push_reorder(const cpic_consts& k) = 0;
    virtual int prepare_reorder() = 0;
    virtual int init_uniform(const UniformPlasmaArgs& a) = 0;
    virtual int device_ptr(int which, void** ptr, int64_t* count, int64_t* stride) = 0;
    virtual int fold_phase(int phase) = 0;
    virtual int stencil_only(int which, double px, double py, double pz, double dt_eps0) = 0;
    virtual int extract_z(void* lo, void* hi, long long cap, long long* n_lo, long long* n_hi, int rebase_lo, int rebase_hi) = 0;
    virtual int append_device(const void* buf, long long cap, long long n) = 0;
    virtual int slab_extract_async(void* lo, void* hi, long long cap, long long* counts_dev, int rebase_lo, int rebase_hi) = 0;
    virtual int slab_append_async(const void* buf, long long cap, const long long* count_dev) = 0;
    virtual bool few_cells() const = 0;   // grid small enough for the block-private accumulator (k_push2<PRIV>)
    virtual bool is_species_of(const CtxBase* p) const = 0;
    virtual int sync_np() = 0;      // device-count mode -> host-count mode (synchronises); no-op otherwise
    virtual long long* device_counts() = 0;      // dc (may be null before the first device-counted migration)
    // which of the double buffers (particle store, segment bounds, cell counts) are current: a captured pair of steps
    // bakes these in, so a graph may only be replayed from the state it was captured in
    virtual unsigned state_signature() const = 0;
    long long fb_steps = 0;         // steps taken in the few-cells fallback of CPIC_SORT_FUSED (sort when % 8 == 0)
    bool dev_count = false;         // slab mode: np lives in dc[0] on the device, the host's np is stale
    virtual int step_host(const cpic_consts& k, const void* const in[8], void* const out[8], long long n,
                          const void* const fin[9], void* const fout[9], double* energies) = 0;
    // the two halves of a host-resident step that the multi-GPU layer weaves its exchanges between (cpic_mgpu_step_host)
    virtual int host_push_phase(const cpic_consts& k, const void* const in[8], void* const out[8], long long n,
                                const void* const fin[9], bool slab) = 0;
    virtual int host_patch_phase(void* const out[8], long long n_in, long long out_capacity, long long* n_out) = 0;
    virtual int host_fields_out(void* const fout[9]) = 0;
    virtual double* energy_scratch() = 0;
    virtual unsigned long long* stats_dev() = 0;
};

inline unsigned blocks_for(long long n, int per = 256) { return (unsigned)((n + per - 1) / per); }

template <class R>
struct Ctx final : CtxBase {
    static constexpr int S = IpStride<R>::value;
    long long nc_pad = 0;
    // particle store: two buffers of records (second only with enable_sort), capacity a multiple of 64
    long long cap = 0;
    char* pbuf[2] = {nullptr, nullptr};
    Particles<R> P[2];
    int cur = 0;
    // struct-of-arrays staging chunk for host transfers (the C ABI exchanges the eight members as separate
    // arrays, like Cabana slices): 8 copies per chunk + one pack/unpack kernel, all on the context's stream
    static constexpr long long XFER_CHUNK = 1ll << 23;
    char* xfer = nullptr;
    long long xfer_cap = 0;
    R* fields = nullptr;       // 9 * nc_pad
    R* interp = nullptr;       // nc * S
    R* acc = nullptr;          // nc * 12
    unsigned* cell_count = nullptr;   // nc (+ scan scratch)
    unsigned* cell_count2 = nullptr;  // nc: the histogram the reordering push writes while it consumes cell_count
    unsigned* scan_l1 = nullptr;
    unsigned* scan_l2 = nullptr;
    long long n_l1 = 0, n_l2 = 0;
    unsigned* bad = nullptr;
    // A SPECIES context (cpic_create_species): its own particle store, sort state and step constants, but the field,
    // interpolator and accumulator arrays and the stream of its parent -- every push deposits into the same J.
    Ctx<R>* parent = nullptr;
    double* en_dev = nullptr;          // 8 doubles scratch
    unsigned long long* stats = nullptr;  // 8 counters

    ~Ctx() override {
        if (stream || true) {
            cudaSetDevice(prm.device);
            for (auto& e : ev) if (e) cudaEventDestroy(e);
            cudaFree(pbuf[0]); cudaFree(pbuf[1]); cudaFree(xfer);
            if (!parent) { cudaFree(fields); cudaFree(interp); cudaFree(acc); }
            cudaFree(seg[0]); cudaFree(seg[1]); cudaFree(cursor3); cudaFree(work3); cudaFree(cell_count); cudaFree(cell_count2); cudaFree(scan_l1); cudaFree(scan_l2); cudaFree(bad); cudaFree(en_dev); cudaFree(stats); cudaFree(mig_counters); cudaFree(mig_lists); cudaFree(leave_list); cudaFree(leave_count); cudaFree(dc);
            cudaFree(ev_cell); cudaFree(ev_row); cudaFree(ev_keys); cudaFree(ev_order[0]); cudaFree(ev_order[1]); cudaFree(ev_tmp); cudaFree(ev_overflow);
            for (auto b : hs_buf) cudaFree(b);
            for (auto e : hs_ev) if (e) cudaEventDestroy(e);
            if (hs_up) cudaStreamDestroy(hs_up);
            if (hs_dn) cudaStreamDestroy(hs_dn);
            if (own_stream && stream) cudaStreamDestroy(stream);
        }
    }

    Fields<R> F() const {
        Fields<R> f;
        for (int m = 0; m < F_N; ++m) f.c[m] = fields + (long long)m * nc_pad;
        return f;
    }

    int ensure_xfer() {
        if (xfer) return CPIC_OK;
        xfer_cap = std::min<long long>(cap, XFER_CHUNK);
        return cuda(cudaMalloc(&xfer, (size_t)xfer_cap * (7 * sizeof(R) + sizeof(int))), "cudaMalloc(transfer staging)");
    }

    int init() {
        int rc;
        if ((rc = cuda(cudaSetDevice(prm.device), "cudaSetDevice"))) return rc;
        if (parent) { stream = parent->stream; own_stream = false; }
        else {
            if ((rc = cuda(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate"))) return rc;
            own_stream = true;
        }
        for (auto& e : ev)
            if ((rc = cuda(cudaEventCreate(&e), "cudaEventCreate"))) return rc;
        // developer tuning knobs (not part of the ABI)
        if (const char* e = getenv("CPIC_PUSH_PREFETCH")) push_prefetch = atoi(e) != 0;
        if (const char* e = getenv("CPIC_PUSH_GRID")) push_grid = atoi(e);
        if (const char* e = getenv("CPIC_PUSH_V1")) use_push2 = atoi(e) == 0;
        if (const char* e = getenv("CPIC_PUSH2_FASTDS")) push2_fastds = atoi(e) != 0;
        if (const char* e = getenv("CPIC_PUSH2_PRIV")) push2_priv = atoi(e) != 0;
        if (const char* e = getenv("CPIC_DEP_THRESH")) dep_thresh = atoi(e);
        if (const char* e = getenv("CPIC_DEP_ROUNDS")) dep_rounds = atoi(e);
        nc_pad = (g.nc + 63) / 64 * 64;
        cap = (prm.max_particles + 63) / 64 * 64;
        if (cap < 64) cap = 64;
        const size_t pbytes = (size_t)cap * sizeof(PRec<R>);
        for (int b = 0; b < (prm.enable_sort ? 2 : 1); ++b) {
            if ((rc = cuda(cudaMalloc(&pbuf[b], pbytes), "cudaMalloc(particles)"))) return rc;
            P[b].rec = reinterpret_cast<PRec<R>*>(pbuf[b]);
        }
        if (parent) { fields = parent->fields; interp = parent->interp; acc = parent->acc; }
        else {
            if ((rc = cuda(cudaMalloc(&fields, (size_t)nc_pad * F_N * sizeof(R)), "cudaMalloc(fields)"))) return rc;
            if ((rc = cuda(cudaMalloc(&interp, (size_t)g.nc * S * sizeof(R)), "cudaMalloc(interpolators)"))) return rc;
            if ((rc = cuda(cudaMalloc(&acc, (size_t)g.nc * 12 * sizeof(R)), "cudaMalloc(accumulators)"))) return rc;
        }
        if ((rc = cuda(cudaMalloc(&bad, sizeof(unsigned)), "cudaMalloc"))) return rc;
        if ((rc = cuda(cudaMalloc(&en_dev, 8 * sizeof(double)), "cudaMalloc"))) return rc;
        if ((rc = cuda(cudaMalloc(&stats, 8 * sizeof(unsigned long long)), "cudaMalloc"))) return rc;
        if (prm.enable_sort) {
            n_l1 = (g.nc + 1 + SCAN_TILE - 1) / SCAN_TILE;       // (+1: the placing push scans a sentinel entry as well)
            n_l2 = (n_l1 + SCAN_TILE - 1) / SCAN_TILE;
            if (n_l2 > SCAN_TILE) return fail(CPIC_E_INVALID, "grid too large for the 3-level cell scan");
            if ((rc = cuda(cudaMalloc(&cell_count, (size_t)(g.nc + 1) * sizeof(unsigned)), "cudaMalloc(cell_count)"))) return rc;
            if ((rc = cuda(cudaMalloc(&cell_count2, (size_t)(g.nc + 1) * sizeof(unsigned)), "cudaMalloc(cell_count2)"))) return rc;
            if ((rc = cuda(cudaMalloc(&scan_l1, (size_t)n_l1 * sizeof(unsigned)), "cudaMalloc"))) return rc;
            if ((rc = cuda(cudaMalloc(&scan_l2, (size_t)n_l2 * sizeof(unsigned)), "cudaMalloc"))) return rc;
        }
        // Field_Solver ctor zeroes the fields (src/fields.h:279-315); interpolators are zeroed by
        // initialize_interpolator (src/interpolator.cpp:125-172); Kokkos::View zero-initialises.
        if ((rc = init_push3())) return rc;
        if (!parent) {
            cudaMemsetAsync(fields, 0, (size_t)nc_pad * F_N * sizeof(R), stream);
            cudaMemsetAsync(interp, 0, (size_t)g.nc * S * sizeof(R), stream);
            cudaMemsetAsync(acc, 0, (size_t)g.nc * 12 * sizeof(R), stream);
        }
        cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), stream);
        return cuda(cudaStreamSynchronize(stream), "init");
    }

    // ------------------------------------------------------------------ transfers
    int upload_particles(const void* const m[7], const int32_t* cell, long long n) override {
        if (n < 0 || n > cap) return fail(CPIC_E_CAPACITY, "upload_particles: %lld particles exceed capacity %lld", n, cap);
        int rc;
        if ((rc = ensure_xfer())) return rc;
        cudaMemsetAsync(bad, 0, sizeof(unsigned), stream);
        for (long long first = 0; first < n; first += xfer_cap) {
            const long long cn = std::min(xfer_cap, n - first);
            SendBuf<R> b = carve_sendbuf<R>(xfer, xfer_cap);
            for (int k = 0; k < 7; ++k)
                if ((rc = cuda(cudaMemcpyAsync(b.m[k], (const R*)m[k] + first, (size_t)cn * sizeof(R), cudaMemcpyHostToDevice, stream), "H2D particles"))) return rc;
            if ((rc = cuda(cudaMemcpyAsync(b.cell, cell + first, (size_t)cn * sizeof(int), cudaMemcpyHostToDevice, stream), "H2D cell"))) return rc;
            k_pack_records<R><<<blocks_for(cn), 256, 0, stream>>>(P[cur], first, b, cn);
            if ((rc = check_launch("k_pack_records"))) return rc;
        }
        np = n;
        hist_valid = false; cursor_valid = false; seg_valid = false; leavers_valid = false; ghost_clean = false;
        // bounds-check the cell indices once on upload (would have caught decks/2stream-short.cxx)
        if (n > 0) {
            k_check_cells<R><<<blocks_for(n), 256, 0, stream>>>(P[cur], n, g.nc, bad);
            if ((rc = check_launch("k_check_cells"))) return rc;
        }
        unsigned nbad = 0;
        if ((rc = cuda(cudaMemcpyAsync(&nbad, bad, sizeof(unsigned), cudaMemcpyDeviceToHost, stream), "D2H"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(stream), "upload_particles"))) return rc;
        if (nbad) { np = 0; return fail(CPIC_E_BAD_CELL, "upload_particles: %u particles have a cell index outside [0,%lld)", nbad, g.nc); }
        return CPIC_OK;
    }
    int download_particles(void* const m[7], int32_t* cell, long long capacity, long long* n_out) override {
        if (capacity < np) return fail(CPIC_E_CAPACITY, "download_particles: buffer holds %lld, need %lld", capacity, np);
        int rc;
        if ((rc = ensure_xfer())) return rc;
        for (long long first = 0; first < np; first += xfer_cap) {
            const long long cn = std::min(xfer_cap, np - first);
            SendBuf<R> b = carve_sendbuf<R>(xfer, xfer_cap);
            k_unpack_records<R><<<blocks_for(cn), 256, 0, stream>>>(P[cur], first, b, cn);
            if ((rc = check_launch("k_unpack_records"))) return rc;
            for (int k = 0; k < 7; ++k)
                if (m[k] && (rc = cuda(cudaMemcpyAsync((R*)m[k] + first, b.m[k], (size_t)cn * sizeof(R), cudaMemcpyDeviceToHost, stream), "D2H particles"))) return rc;
            if (cell && (rc = cuda(cudaMemcpyAsync(cell + first, b.cell, (size_t)cn * sizeof(int), cudaMemcpyDeviceToHost, stream), "D2H cell"))) return rc;
        }
        if (n_out) *n_out = np;
        return cuda(cudaStreamSynchronize(stream), "download_particles");
    }
    int upload_fields(const void* const f[9]) override {
        int rc;
        for (int m = 0; m < F_N; ++m)
            if ((rc = cuda(cudaMemcpyAsync(fields + (long long)m * nc_pad, f[m], (size_t)g.nc * sizeof(R), cudaMemcpyHostToDevice, stream), "H2D fields"))) return rc;
        return cuda(cudaStreamSynchronize(stream), "upload_fields");
    }
    int download_fields(void* const f[9]) override {
        int rc;
        for (int m = 0; m < F_N; ++m)
            if (f[m] && (rc = cuda(cudaMemcpyAsync(f[m], fields + (long long)m * nc_pad, (size_t)g.nc * sizeof(R), cudaMemcpyDeviceToHost, stream), "D2H fields"))) return rc;
        return cuda(cudaStreamSynchronize(stream), "download_fields");
    }
    int xfer_interp(void* host, bool up) override {
        // host layout [nc][18]; device records are padded to S reals
        int rc;
        if (up)
            rc = cuda(cudaMemcpy2DAsync(interp, S * sizeof(R), host, 18 * sizeof(R), 18 * sizeof(R), (size_t)g.nc, cudaMemcpyHostToDevice, stream), "H2D interpolators");
        else
            rc = cuda(cudaMemcpy2DAsync(host, 18 * sizeof(R), interp, S * sizeof(R), 18 * sizeof(R), (size_t)g.nc, cudaMemcpyDeviceToHost, stream), "D2H interpolators");
        if (rc) return rc;
        return cuda(cudaStreamSynchronize(stream), "xfer_interp");
    }
    int xfer_acc(void* host, bool up) override {
        int rc = up ? cuda(cudaMemcpyAsync(acc, host, (size_t)g.nc * 12 * sizeof(R), cudaMemcpyHostToDevice, stream), "H2D accumulators")
                    : cuda(cudaMemcpyAsync(host, acc, (size_t)g.nc * 12 * sizeof(R), cudaMemcpyDeviceToHost, stream), "D2H accumulators");
        if (rc) return rc;
        return cuda(cudaStreamSynchronize(stream), "xfer_acc");
    }

    // ------------------------------------------------------------------ field side
    int load_interpolator() override {
        Box b{g.ng, g.ng, g.ng, g.nx, g.ny, g.nz};
        k_load_interpolator<R><<<blocks_for(b.count()), 256, 0, stream>>>(F(), interp, g, b);
        return check_launch("k_load_interpolator");
    }
    int initialize_interpolator() override {
        return cuda(cudaMemsetAsync(interp, 0, (size_t)g.nc * S * sizeof(R), stream), "initialize_interpolator");
    }
    int clear_accumulator() override {
        return cuda(cudaMemsetAsync(acc, 0, (size_t)g.nc * 12 * sizeof(R), stream), "clear_accumulator");
    }
    int unload_accumulator(const cpic_consts& k) override {
        // src/accumulator.cpp:66-68: products in real_t, 0.25/x in double, narrowed to real_t
        const R dx = (R)k.dx, dy = (R)k.dy, dz = (R)k.dz, dt = (R)k.dt;
        const R cx = (R)(0.25 / (double)(R)(dy * dz * dt));
        const R cy = (R)(0.25 / (double)(R)(dz * dx * dt));
        const R cz = (R)(0.25 / (double)(R)(dx * dy * dt));
        Box b{g.ng, g.ng, g.ng, g.nx + 1, g.ny + 1, g.nz + 1};
        k_unload_accumulator<R><<<blocks_for(b.count()), 256, 0, stream>>>(F(), acc, g, b, cx, cy, cz);
        return check_launch("k_unload_accumulator");
    }
    int ghost_copy(int m0) {
        Fields<R> f = F();
        k_ghost_copy3<R><<<blocks_for(ghost_cell_count(g)), 256, 0, stream>>>(f.c[m0], f.c[m0 + 1], f.c[m0 + 2], g);
        return check_launch("k_ghost_copy3");
    }
    int ghost_fold() {
        Fields<R> f = F();
        const long long m0 = std::max({(long long)g.nx * (g.nz + 1), (long long)g.ny * (g.nx + 1), (long long)g.nz * (g.ny + 1)});
        const long long m1 = std::max({(long long)g.nx * (g.ny + 1), (long long)g.ny * (g.nz + 1), (long long)g.nz * (g.nx + 1)});
        k_ghost_fold<R, 0><<<dim3(blocks_for(m0), 3), 256, 0, stream>>>(f.c[F_JFX], f.c[F_JFY], f.c[F_JFZ], g);
        int rc = check_launch("k_ghost_fold<0>");
        if (rc) return rc;
        k_ghost_fold<R, 1><<<dim3(blocks_for(m1), 3), 256, 0, stream>>>(f.c[F_JFX], f.c[F_JFY], f.c[F_JFZ], g);
        return check_launch("k_ghost_fold<1>");
    }
    int fold_phase(int phase) override {
        Fields<R> f = F();
        const long long m0 = std::max({(long long)g.nx * (g.nz + 1), (long long)g.ny * (g.nx + 1), (long long)g.nz * (g.ny + 1)});
        const long long m1 = std::max({(long long)g.nx * (g.ny + 1), (long long)g.ny * (g.nz + 1), (long long)g.nz * (g.nx + 1)});
        if (phase == 0) k_ghost_fold<R, 0><<<dim3(blocks_for(m0), 3), 256, 0, stream>>>(f.c[F_JFX], f.c[F_JFY], f.c[F_JFZ], g);
        else k_ghost_fold<R, 1><<<dim3(blocks_for(m1), 3), 256, 0, stream>>>(f.c[F_JFX], f.c[F_JFY], f.c[F_JFZ], g);
        return check_launch("k_ghost_fold");
    }
    // ------------------------------------------------------------------ slab migration
    unsigned* leave_list = nullptr;     // store indices of the particles the last push left in a z ghost plane
    unsigned* leave_count = nullptr;
    long long leave_cap = 0;
    // (re)arm the leaver list for a push in slab mode; no-op while z is periodic inside this context
    int arm_leave_list(PushArgs<R>& a) {
        a.leave_list = nullptr; a.leave_count = nullptr; a.leave_cap = 0; a.leave_lo = 0; a.leave_hi = 0; a.leave_off = 0;
        leavers_valid = false;
        if ((g.per & 4) || prm.boundary != CPIC_BOUNDARY_PERIODIC || !ghost_clean) return CPIC_OK;
        int rc;
        if (!leave_list) {
            leave_cap = cap / 4 + 65536;
            if ((rc = cuda(cudaMalloc(&leave_list, (size_t)leave_cap * sizeof(unsigned)), "cudaMalloc(leaver list)"))) return rc;
            if ((rc = cuda(cudaMalloc(&leave_count, sizeof(unsigned)), "cudaMalloc"))) return rc;
        }
        cudaMemsetAsync(leave_count, 0, sizeof(unsigned), stream);
        a.leave_list = leave_list; a.leave_count = leave_count; a.leave_cap = (unsigned)leave_cap;
        a.leave_lo = g.gx * g.gy; a.leave_hi = (g.nz + 1) * g.gx * g.gy;
        leavers_valid = true;
        return CPIC_OK;
    }
    unsigned* mig_counters = nullptr;   // 8 counters
    unsigned* mig_lists = nullptr;      // 2 * mig_cap indices: [0,mig_cap) holes, [mig_cap,2*mig_cap) donors
    long long mig_cap = 0;
    int extract_z(void* lo, void* hi, long long cap_send, long long* n_lo, long long* n_hi, int rebase_lo, int rebase_hi) override {
        int rc;
        if (cap_send < 1) return fail(CPIC_E_INVALID, "extract_z_leavers: capacity must be positive");
        if (!mig_counters && (rc = cuda(cudaMalloc(&mig_counters, 8 * sizeof(unsigned)), "cudaMalloc"))) return rc;
        if (mig_cap < 2 * cap_send) {      // up to n_lo + n_hi <= 2*cap_send holes, and as many donors
            cudaFree(mig_lists); mig_lists = nullptr;
            if ((rc = cuda(cudaMalloc(&mig_lists, (size_t)4 * cap_send * sizeof(unsigned)), "cudaMalloc(migration lists)"))) return rc;
            mig_cap = 2 * cap_send;
        }
        *n_lo = *n_hi = 0;
        cursor_valid = false;
        if (np == 0) return CPIC_OK;
        cudaMemsetAsync(mig_counters, 0, 8 * sizeof(unsigned), stream);
        const int plane = g.gx * g.gy;
        // the last push listed the particles it left in the ghost planes: O(leavers) instead of two passes
        // over the store (unless the list overflowed)
        long long nl = -1;
        if (leavers_valid) {
            unsigned hl = 0;
            if ((rc = cuda(cudaMemcpyAsync(&hl, leave_count, sizeof hl, cudaMemcpyDeviceToHost, stream), "D2H"))) return rc;
            if ((rc = cuda(cudaStreamSynchronize(stream), "extract_z_leavers"))) return rc;
            if ((long long)hl <= leave_cap) nl = hl;
        }
        if (nl == 0) { ghost_clean = true; return CPIC_OK; }
        if (nl > 0)
            k_extract_mark_list<R><<<blocks_for(nl), 256, 0, stream>>>(P[cur], leave_list, nl, plane, g.nz, carve_sendbuf<R>(lo, cap_send),
                                                                      carve_sendbuf<R>(hi, cap_send), cap_send, rebase_lo, rebase_hi, mig_counters);
        else
            k_extract_mark<R><<<blocks_for(np), 256, 0, stream>>>(P[cur], np, plane, g.nz, carve_sendbuf<R>(lo, cap_send),
                                                                 carve_sendbuf<R>(hi, cap_send), cap_send, rebase_lo, rebase_hi, mig_counters);
        if ((rc = check_launch("k_extract_mark"))) return rc;
        unsigned h[6];
        if ((rc = cuda(cudaMemcpyAsync(h, mig_counters, sizeof h, cudaMemcpyDeviceToHost, stream), "D2H"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(stream), "extract_z_leavers"))) return rc;
        if (h[2] || h[0] > cap_send || h[1] > cap_send)
            return fail(CPIC_E_CAPACITY, "extract_z_leavers: %u/%u leavers exceed the send-buffer capacity %lld", h[0], h[1], cap_send);
        if (nl > 0 && (h[5] || (long long)h[0] + h[1] != nl))
            return fail(CPIC_E_CUDA, "extract_z_leavers: the leaver list of the last push is inconsistent (%lld listed, %u + %u found)", nl, h[0], h[1]);
        *n_lo = h[0]; *n_hi = h[1];
        const long long n_out = (long long)h[0] + h[1];
        ghost_clean = true;
        if (n_out == 0) return CPIC_OK;
        const long long np_new = np - n_out;
        if (nl > 0)
            k_extract_lists_list<R><<<blocks_for(nl + n_out), 256, 0, stream>>>(P[cur], leave_list, nl, np, np_new, plane, g.nz, mig_lists, mig_cap, mig_counters);
        else
            k_extract_lists<R><<<blocks_for(np), 256, 0, stream>>>(P[cur], np, np_new, plane, g.nz, mig_lists, mig_cap, mig_counters);
        if ((rc = check_launch("k_extract_lists"))) return rc;
        leavers_valid = false;
        k_extract_fill<R><<<blocks_for(n_out), 256, 0, stream>>>(P[cur], mig_lists, mig_cap, mig_counters);
        if ((rc = check_launch("k_extract_fill"))) return rc;
        np = np_new;
        // the cell histogram of the last push stays usable: everything that sat in the two ghost planes is gone
        // (the hole filling only permutes the rest)
        if (hist_valid) {
            cudaMemsetAsync(cell_count, 0, (size_t)plane * sizeof(unsigned), stream);
            cudaMemsetAsync(cell_count + (size_t)(g.nz + 1) * plane, 0, (size_t)plane * sizeof(unsigned), stream);
        }
        return CPIC_OK;
    }
    int append_device(const void* buf, long long cap_buf, long long n) override {
        if (n < 0 || n > cap_buf) return fail(CPIC_E_INVALID, "append_particles_device: bad count");
        if (np + n > cap) return fail(CPIC_E_CAPACITY, "append_particles_device: %lld + %lld particles exceed capacity %lld", np, n, cap);
        if (n == 0) return CPIC_OK;
        cursor_valid = false;
        SendBuf<R> b = carve_sendbuf<R>(const_cast<void*>(buf), cap_buf);
        k_pack_records<R><<<blocks_for(n), 256, 0, stream>>>(P[cur], np, b, n);
        int rc;
        if ((rc = check_launch("k_pack_records"))) return rc;
        if (hist_valid) {       // keep the histogram of the last push current: count the arrivals
            k_hist_add<<<blocks_for(n), 256, 0, stream>>>(b.cell, n, g.nc, cell_count);
            if ((rc = check_launch("k_hist_add"))) return rc;
        }
        np += n;
        return CPIC_OK;
    }
    // ------------------------------------------------------------------ slab migration, counts on the device
    long long* dc = nullptr;            // [0] np [1] error flags [2] [3] leavers of the last extraction
    long long* device_counts() override { return dc; }
    unsigned state_signature() const override {
        return (unsigned)cur | ((unsigned)seg_cur << 1) | (cell_count < cell_count2 ? 4u : 0u) | (dev_count ? 8u : 0u) | (seg_valid ? 16u : 0u);
    }
    int enter_dev_count() {
        if (dev_count) return CPIC_OK;
        int rc;
        if (!dc && (rc = cuda(cudaMalloc(&dc, 4 * sizeof(long long)), "cudaMalloc"))) return rc;
        const long long h[4] = {np, 0, 0, 0};
        if ((rc = cuda(cudaMemcpyAsync(dc, h, sizeof h, cudaMemcpyHostToDevice, stream), "H2D counts"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(stream), "enter_dev_count"))) return rc;
        dev_count = true;
        return CPIC_OK;
    }
    int sync_np() override {
        if (!dev_count) return CPIC_OK;
        long long h[4] = {0, 0, 0, 0};
        int rc;
        if ((rc = cuda(cudaMemcpyAsync(h, dc, sizeof h, cudaMemcpyDeviceToHost, stream), "D2H counts"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(stream), "sync_np"))) return rc;
        dev_count = false;
        np = h[0];
        if (h[1] & 1) return fail(CPIC_E_CAPACITY, "slab_extract_async: leavers exceeded the send-buffer capacity");
        if (h[1] & 4) return fail(CPIC_E_CAPACITY, "slab_append_async: arrivals exceeded the store capacity %lld", cap);
        if (h[1] & 8) return fail(CPIC_E_CUDA, "multi-GPU exchange: a neighbour rank did not arrive in time");
        if (h[1]) return fail(CPIC_E_CUDA, "slab_extract_async: the leaver list of a push overflowed or was inconsistent");
        return CPIC_OK;
    }
    int slab_extract_async(void* lo, void* hi, long long cap_send, long long* counts_dev, int rebase_lo, int rebase_hi) override {
        int rc;
        if (!can_reorder()) return fail(CPIC_E_UNSUPPORTED, "slab_extract_async: needs the float reordering push (enable_sort, warp deposit)");
        if (cap_send < 1) return fail(CPIC_E_INVALID, "slab_extract_async: capacity must be positive");
        if (!leavers_valid) {
            // The last push did not list its leavers (first step after an upload: the ghost planes were not known
            // to be empty): one synchronising scan-based extraction, after which every push arms its list.
            if ((rc = sync_np())) return rc;
            long long a = 0, b = 0;
            if ((rc = extract_z(lo, hi, cap_send, &a, &b, rebase_lo, rebase_hi))) return rc;
            if ((rc = enter_dev_count())) return rc;
            const long long h[2] = {a, b};
            if ((rc = cuda(cudaMemcpyAsync(counts_dev, h, sizeof h, cudaMemcpyHostToDevice, stream), "H2D counts"))) return rc;
            return cuda(cudaStreamSynchronize(stream), "slab_extract_async");
        }
        if ((rc = enter_dev_count())) return rc;
        if (!mig_counters && (rc = cuda(cudaMalloc(&mig_counters, 8 * sizeof(unsigned)), "cudaMalloc"))) return rc;
        if (mig_cap < 2 * cap_send) {
            cudaFree(mig_lists); mig_lists = nullptr;
            if ((rc = cuda(cudaMalloc(&mig_lists, (size_t)4 * cap_send * sizeof(unsigned)), "cudaMalloc(migration lists)"))) return rc;
            mig_cap = 2 * cap_send;
        }
        cursor_valid = false;
        cudaMemsetAsync(mig_counters, 0, 8 * sizeof(unsigned), stream);
        const int plane = g.gx * g.gy;
        const unsigned nb = blocks_for(2 * cap_send);
        k_extract_mark_dev<R><<<nb, 256, 0, stream>>>(P[cur], leave_list, leave_count, (unsigned)leave_cap, plane, g.nz, carve_sendbuf<R>(lo, cap_send),
                                                    carve_sendbuf<R>(hi, cap_send), cap_send, rebase_lo, rebase_hi, mig_counters, dc);
        if ((rc = check_launch("k_extract_mark_dev"))) return rc;
        k_extract_lists_dev<R><<<2 * nb, 256, 0, stream>>>(P[cur], leave_list, leave_count, plane, g.nz, mig_lists, mig_cap, mig_counters, dc);
        if ((rc = check_launch("k_extract_lists_dev"))) return rc;
        k_extract_fill_dev<R><<<nb, 256, 0, stream>>>(P[cur], mig_lists, mig_cap, mig_counters, dc);
        if ((rc = check_launch("k_extract_fill_dev"))) return rc;
        k_extract_finish_dev<<<1, 1, 0, stream>>>(mig_counters, leave_count, dc, counts_dev);
        if ((rc = check_launch("k_extract_finish_dev"))) return rc;
        leavers_valid = false;
        ghost_clean = true;
        if (hist_valid) {      // everything that sat in the two ghost planes is gone (the hole filling only permutes the rest)
            cudaMemsetAsync(cell_count, 0, (size_t)plane * sizeof(unsigned), stream);
            cudaMemsetAsync(cell_count + (size_t)(g.nz + 1) * plane, 0, (size_t)plane * sizeof(unsigned), stream);
        }
        return CPIC_OK;
    }
    int slab_append_async(const void* buf, long long cap_buf, const long long* count_dev) override {
        int rc;
        if (!dev_count) return fail(CPIC_E_INVALID, "slab_append_async: call slab_extract_async first");
        cursor_valid = false;
        SendBuf<R> b = carve_sendbuf<R>(const_cast<void*>(buf), cap_buf);
        k_append_dev<R><<<blocks_for(cap_buf), 256, 0, stream>>>(P[cur], b, cap_buf, count_dev, cap, g.nc, hist_valid ? cell_count : nullptr, dc);
        if ((rc = check_launch("k_append_dev"))) return rc;
        k_append_finish_dev<<<1, 1, 0, stream>>>(count_dev, dc);
        return check_launch("k_append_finish_dev");
    }
    int update_ghosts(int which) override {
        if (reflecting()) return CPIC_OK;      // no periodic images
        if (which == 0) return ghost_fold();
        if (which == 3) return fold_phase(0);
        if (which == 4) return fold_phase(1);
        if (which == 1) return ghost_copy(F_JFX);
        if (which == 2) return ghost_copy(F_CBX);
        return fail(CPIC_E_INVALID, "update_ghosts: which=%d", which);
    }
    int stencil_only(int which, double px, double py, double pz, double dt_eps0) override {
        if (which == 0) {
            if (prm.solver == CPIC_SOLVER_ES_1D) return CPIC_OK;
            Box b{1, 1, 1, g.nx, g.ny, g.nz};
            k_advance_b<R><<<blocks_for(b.count()), 256, 0, stream>>>(F(), g, b, (R)px, (R)py, (R)pz);
            return check_launch("k_advance_b");
        }
        if (prm.solver == CPIC_SOLVER_ES_1D) {
            k_advance_e_es<R><<<blocks_for(g.nc), 256, 0, stream>>>(F(), g.nc, (R)dt_eps0);
            return check_launch("k_advance_e_es");
        }
        Box b{1, 1, 1, g.nx + 1, g.ny + 1, g.nz + 1};
        k_advance_e_em<R><<<blocks_for(b.count()), 256, 0, stream>>>(F(), g, b, (R)px, (R)py, (R)pz, (R)dt_eps0);
        return check_launch("k_advance_e_em");
    }
    int advance_b(double px, double py, double pz) override {
        if (prm.solver == CPIC_SOLVER_ES_1D) return CPIC_OK;  // src/fields.h:470-482: no-op
        Box b{1, 1, 1, g.nx, g.ny, g.nz};
        k_advance_b<R><<<blocks_for(b.count()), 256, 0, stream>>>(F(), g, b, (R)px, (R)py, (R)pz);
        int rc = check_launch("k_advance_b");
        if (rc || reflecting()) return rc;      // (conducting walls: the normal cB on a wall never changes, nothing to copy)
        return ghost_copy(F_CBX);   // src/fields.h:718
    }
    bool reflecting() const { return prm.boundary == CPIC_BOUNDARY_REFLECT; }
    int advance_e(double px, double py, double pz, double dt_eps0) override {
        int rc;
        if (reflecting()) {      // conducting walls: no periodic fold / copy; stencil, then E_tang = 0 on the walls
            if ((rc = stencil_only(1, px, py, pz, dt_eps0))) return rc;
            Fields<R> f = F();
            k_pec_walls<R><<<blocks_for(g.nc), 256, 0, stream>>>(f.c[F_EX], f.c[F_EY], f.c[F_EZ], g);
            return check_launch("k_pec_walls");
        }
        rc = ghost_fold();                                       // src/fields.h:642 / :530
        if (rc) return rc;
        if (prm.solver == CPIC_SOLVER_ES_1D) {
            k_advance_e_es<R><<<blocks_for(g.nc), 256, 0, stream>>>(F(), g.nc, (R)dt_eps0);
            return check_launch("k_advance_e_es");
        }
        if ((rc = ghost_copy(F_JFX))) return rc;                 // src/fields.h:643
        Box b{1, 1, 1, g.nx + 1, g.ny + 1, g.nz + 1};
        k_advance_e_em<R><<<blocks_for(b.count()), 256, 0, stream>>>(F(), g, b, (R)px, (R)py, (R)pz, (R)dt_eps0);
        return check_launch("k_advance_e_em");
    }
    int energies_async(double* dev_out2) override {
        cudaMemsetAsync(dev_out2, 0, 2 * sizeof(double), stream);
        const bool em = prm.solver == CPIC_SOLVER_EM;
        Box b = em ? Box{1, 1, 1, g.nx, g.ny, g.nz} : Box{0, 0, 0, g.gx, g.gy, g.gz};
        unsigned nb = blocks_for(b.count());
        if (nb > 148 * 8) nb = 148 * 8;
        k_energy<R><<<nb, 256, 0, stream>>>(F(), g, b, em ? 1 : 0, dev_out2);
        return check_launch("k_energy");
    }
    int kinetic_async(double* dev_out) override {
        cudaMemsetAsync(dev_out, 0, sizeof(double), stream);
        if (np == 0) return CPIC_OK;
        unsigned nb = blocks_for(np);
        if (nb > 148 * 8) nb = 148 * 8;
        k_kinetic_energy<R><<<nb, 256, 0, stream>>>(P[cur], np, dev_out);
        return check_launch("k_kinetic_energy");
    }
    int digest_async(double* dev_out8) override {
        cudaMemsetAsync(dev_out8, 0, 8 * sizeof(double), stream);
        unsigned nb = blocks_for(dev_count ? cap : std::max<long long>(np, 1));
        if (nb > 148 * 8) nb = 148 * 8;
        k_state_digest<R><<<nb, 256, 0, stream>>>(P[cur], np, dev_count ? dc : nullptr, g, dev_out8);
        int rc = check_launch("k_state_digest");
        if (rc) return rc;
        return energies_async(dev_out8 + 5);
    }
    double* energy_scratch() override { return en_dev; }
    unsigned long long* stats_dev() override { return stats; }

    // ------------------------------------------------------------------ particles
    // Persistent launch: one wave of blocks that fills the machine (SM count x resident blocks
    // per SM), each warp grid-striding over 32-particle tiles.
    int push_grid = 0;
    int dep_thresh = 16, dep_rounds = 4;
    bool push_prefetch = true;
    template <bool FMA, int DEP, bool ST, bool PF>
    int launch_push_k(const PushArgs<R>& a) {
        auto kern = k_push<R, FMA, DEP, ST, PF>;
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PUSH_WARPS * 32, 0);
        if (per_sm < 1) per_sm = 1;
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, prm.device);
        long long blocks = (long long)sms * per_sm;
        const long long need = (a.np + PUSH_WARPS * 32 - 1) / (PUSH_WARPS * 32);
        if (blocks > need) blocks = need;
        if (push_grid > 0) blocks = std::min<long long>(push_grid, need);
        kern<<<(unsigned)blocks, PUSH_WARPS * 32, 0, stream>>>(a);
        return check_launch("k_push");
    }
    template <bool FMA, int DEP>
    int launch_push(const PushArgs<R>& a) {
        if (want_stats) return push_prefetch ? launch_push_k<FMA, DEP, true, true>(a) : launch_push_k<FMA, DEP, true, false>(a);
        return push_prefetch ? launch_push_k<FMA, DEP, false, true>(a) : launch_push_k<FMA, DEP, false, false>(a);
    }
    // second-generation float kernel (packed FP32x2, two particles per thread); deposit mode WARP only
    bool use_push2 = true, push2_fastds = true;
    // Grids of up to PRIV_MAX_CELLS cells (the reference's 1-D decks: 34 x 3 x 3 = 306): every block keeps a private
    // accumulator + histogram in shared memory -- a global reduction per streak would serialise the whole machine
    // on a few hundred L2 lines (BASELINE configs[1], 1e8 particles on 32 cells: 86 -> ms/step, profiles/r03_*c2*)
    static constexpr long long PRIV_MAX_CELLS = 1024;
    bool push2_priv = true;
    bool use_priv() const { return push2_priv && g.nc <= PRIV_MAX_CELLS; }
    bool few_cells() const override { return use_priv() && can_reorder(); }
    bool is_species_of(const CtxBase* p) const override { return parent != nullptr && static_cast<const CtxBase*>(parent) == p; }
    template <bool FMA, bool ST, bool FD>
    int launch_push2(const PushArgs<float>& a) {
        if (!ST && a.priv_nc > 0) return a.hist ? launch_push2h<FMA, ST, FD, true, false, true>(a) : launch_push2h<FMA, ST, FD, false, false, true>(a);
        return a.hist ? launch_push2h<FMA, ST, FD, true>(a) : launch_push2h<FMA, ST, FD, false>(a);
    }
    template <bool FMA, bool ST, bool FD>
    int launch_push2r(const PushArgs<float>& a) {
        if (!ST && a.priv_nc > 0) return launch_push2h<FMA, ST, FD, true, true, true>(a);
        return launch_push2h<FMA, ST, FD, true, true>(a);
    }
    template <bool FMA, bool ST, bool FD, bool H, bool RE = false, bool PV = false>
    int launch_push2h(const PushArgs<float>& a) {
        auto kern = k_push2<FMA, ST, FD, H, RE, PV>;
        const size_t smem = sizeof(Push2Smem) + (PV ? (size_t)a.priv_nc * 13 * sizeof(float) : 0);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PUSH2_WARPS * 32, smem);
        if (per_sm < 1) per_sm = 1;
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, prm.device);
        long long blocks = (long long)sms * per_sm;
        const long long need = (a.np + PUSH2_WARPS * 64 - 1) / (PUSH2_WARPS * 64);
        if (blocks > need) blocks = need;
        if (push_grid > 0) blocks = std::min<long long>(push_grid, need);
        volatile float one = 1.0f;     // a runtime value as far as the compiler is concerned (cpic_push2.cuh)
        kern<<<(unsigned)blocks, PUSH2_WARPS * 32, smem, stream>>>(a, one);
        return check_launch("k_push2");
    }
    // The reordering push (cpic_push2.cuh, REORD): cpic_push + the cell ordering of the particle store in one
    // pass.  prepare_reorder() makes cell_count the exclusive scan of the current cells' histogram (from the
    // previous reordering push when there was one); push_reorder() consumes it.
    // ---- block-owned reordering push (cpic_push3.cuh): chunks of cells per CTA, TMA-staged interpolators
    unsigned* seg[2] = {nullptr, nullptr};   // exclusive scans (nc + 1 entries + pad): segment bounds of the store, ping-pong
    unsigned* cursor3 = nullptr;             // the mutable copy of the destination bounds the push claims slots from
    unsigned* work3 = nullptr;               // dynamic work counter
    int seg_cur = 0;                         // seg[seg_cur] describes P[cur] while seg_valid
    bool use_push3 = true;
    int p3_ch = 0, p3_cpp = 0, p3_yblock = 16;
    bool push3_ok() const { return use_push3 && can_reorder() && !use_priv() && seg[0] != nullptr; }
    int init_push3() {
        int rc;
        if (const char* e = getenv("CPIC_PUSH3")) use_push3 = atoi(e) != 0;
        if (const char* e = getenv("CPIC_PUSH3_YBLOCK")) p3_yblock = std::max(1, atoi(e));
        if (!std::is_same<R, float>::value || !prm.enable_sort) return CPIC_OK;
        for (auto& b : seg)
            if ((rc = cuda(cudaMalloc(&b, (size_t)(g.nc + 16) * sizeof(unsigned)), "cudaMalloc(segment bounds)"))) return rc;
        if ((rc = cuda(cudaMalloc(&cursor3, (size_t)(g.nc + 16) * sizeof(unsigned)), "cudaMalloc(cursor)"))) return rc;
        if ((rc = cuda(cudaMalloc(&work3, sizeof(unsigned)), "cudaMalloc"))) return rc;
        // chunk = whole x-rows while they fit (and while there are enough chunks to balance the CTAs), else a piece of a row
        const int plane = g.gx * g.gy;
        int ch = PUSH3_CH_MAX;
        if (g.gx <= PUSH3_CH_MAX) {
            int rows = PUSH3_CH_MAX / g.gx;
            while (rows > 1 && (long long)((g.gy + rows - 1) / rows) * g.gz < 16ll * 148 * PUSH3_MIN_BLOCKS) --rows;
            ch = rows * g.gx;
        }
        if (const char* e = getenv("CPIC_PUSH3_CH")) ch = std::min(PUSH3_CH_MAX, std::max(4, atoi(e)));
        p3_ch = std::min(ch, plane);
        p3_cpp = (plane + p3_ch - 1) / p3_ch;
        return CPIC_OK;
    }
    // exclusive scan of the current cells' histogram (cell_count, left intact) into dstbuf[0 .. nc]; dstbuf[nc] = total
    int scan_hist_into(unsigned* dstbuf) {
        int rc;
        if ((rc = cuda(cudaMemcpyAsync(dstbuf, cell_count, (size_t)g.nc * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream), "D2D histogram"))) return rc;
        cudaMemsetAsync(dstbuf + g.nc, 0, 16 * sizeof(unsigned), stream);
        return scan_buf(dstbuf, g.nc + 1);
    }
    // bring the store into segment form: counting sort by cell, keeping the bounds (seg) and the histogram (cell_count)
    int establish_segments() {
        int rc;
        if (!hist_valid) {
            cudaMemsetAsync(cell_count, 0, (size_t)g.nc * sizeof(unsigned), stream);
            cudaMemsetAsync(bad, 0, sizeof(unsigned), stream);
            k_cell_histogram<R><<<blocks_for(np), 256, 0, stream>>>(P[cur], np, g.nc, cell_count, bad);
            if ((rc = check_launch("k_cell_histogram"))) return rc;
            if ((rc = check_bad_cells("push_reorder"))) return rc;
            hist_valid = true;
        }
        rec(2);
        if ((rc = scan_hist_into(seg[seg_cur]))) return rc;
        if ((rc = cuda(cudaMemcpyAsync(cursor3, seg[seg_cur], (size_t)g.nc * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream), "D2D cursor"))) return rc;
        k_sort_scatter<R><<<blocks_for(np), 256, 0, stream>>>(P[cur], P[cur ^ 1], np, cursor3);
        if ((rc = check_launch("k_sort_scatter"))) return rc;
        cur ^= 1;
        rec(3);
        ev_valid[1] = true;
        cursor_valid = false; leavers_valid = false;
        seg_valid = true;
        return CPIC_OK;
    }
    template <bool FMA, bool ST, bool FD>
    int launch_push3(const Push3Args& q) {
        auto kern = k_push3<FMA, ST, FD>;
        const size_t smem = sizeof(Push3Smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, PUSH3_WARPS * 32, smem);
        if (per_sm < 1) per_sm = 1;
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, prm.device);
        long long blocks = (long long)sms * per_sm;
        if (push_grid > 0) blocks = push_grid;
        volatile float one = 1.0f;     // a runtime value as far as the compiler is concerned (cpic_push2.cuh)
        kern<<<(unsigned)blocks, PUSH3_WARPS * 32, smem, stream>>>(q, one);
        return check_launch("k_push3");
    }
    int push_reorder3(const cpic_consts& k) {
        int rc;
        if (!seg_valid) {
            if (dev_count && (rc = sync_np())) return rc;
            if (np == 0) return CPIC_OK;
            if ((rc = establish_segments())) return rc;
        }
        if constexpr (std::is_same<R, float>::value) {
            const int so = seg_cur ^ 1;
            // destination bounds = scan of the histogram of the cells the particles are in now
            if ((rc = scan_hist_into(seg[so]))) return rc;
            if ((rc = cuda(cudaMemcpyAsync(cursor3, seg[so], (size_t)g.nc * sizeof(unsigned), cudaMemcpyDeviceToDevice, stream), "D2D cursor"))) return rc;
            cudaMemsetAsync(cell_count2, 0, (size_t)g.nc * sizeof(unsigned), stream);
            cudaMemsetAsync(work3, 0, sizeof(unsigned), stream);
            Push3Args q;
            q.a = push_args(k);
            if ((rc = arm_leave_list(q.a))) return rc;
            q.a.dst = P[cur ^ 1];
            q.a.cursor = cursor3;
            q.a.hist = cell_count2;
            q.a.priv_nc = 0;
            if (dev_count) { q.a.np_dev = dc; q.a.np = cap; }
            q.sin = seg[seg_cur];
            q.ch = p3_ch; q.cpp = p3_cpp; q.gz = g.gz; q.plane = g.gx * g.gy; q.yblock = p3_yblock;
            q.nchunks = p3_cpp * g.gz; q.nc = (int)g.nc;
            q.work = work3;
            if (want_stats) cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), stream);
            rec(0);
            const bool fma = prm.fp_mode == CPIC_FP_CONTRACT;
            const float aq = fabsf((float)q.a.qdt_2mc);
            const bool fd = push2_fastds && (aq == 0.f || (aq > 1e-12f && aq < 1e12f));
            if (want_stats) rc = fma ? launch_push3<true, true, false>(q) : launch_push3<false, true, false>(q);
            else if (fd) rc = fma ? launch_push3<true, false, true>(q) : launch_push3<false, false, true>(q);
            else rc = fma ? launch_push3<true, false, false>(q) : launch_push3<false, false, false>(q);
            rec(1);
            ev_valid[0] = true;
            if (rc) return rc;
            cur ^= 1;
            seg_cur = so;
            std::swap(cell_count, cell_count2);      // cell_count: histogram of the cells the particles are in now
            hist_valid = true;
            cursor_valid = false;
            want_hist = false;
        }
        return CPIC_OK;
    }
    bool can_reorder() const {
        const int dep = prm.deposit_mode == CPIC_DEPOSIT_AUTO ? CPIC_DEPOSIT_WARP : prm.deposit_mode;
        return std::is_same<R, float>::value && use_push2 && prm.enable_sort && dep == CPIC_DEPOSIT_WARP;
    }
    int prepare_reorder() override {
        if (!prm.enable_sort) return fail(CPIC_E_INVALID, "push_reorder: context was created with enable_sort=0");
        if (push3_ok()) return CPIC_OK;      // (k_push3 keeps its own bounds, see push_reorder3)
        if ((np == 0 && !dev_count) || cursor_valid) return CPIC_OK;
        int rc;
        if (!hist_valid) {
            cudaMemsetAsync(cell_count, 0, (size_t)g.nc * sizeof(unsigned), stream);
            cudaMemsetAsync(bad, 0, sizeof(unsigned), stream);
            k_cell_histogram<R><<<blocks_for(np), 256, 0, stream>>>(P[cur], np, g.nc, cell_count, bad);
            if ((rc = check_launch("k_cell_histogram"))) return rc;
            if ((rc = check_bad_cells("push_reorder"))) return rc;
        }
        hist_valid = false;
        if ((rc = scan_cells())) return rc;
        cursor_valid = true;
        return CPIC_OK;
    }
    int push_reorder(const cpic_consts& k) override {
        if (!can_reorder()) {          // double / other deposit modes: a counting sort, then the in-place push
            int rc = sort();
            return rc ? rc : push(k);
        }
        if (push3_ok()) return push_reorder3(k);
        if (np == 0 && !dev_count) return CPIC_OK;      // (device-count mode: the host's np is stale)
        int rc;
        if ((rc = prepare_reorder())) return rc;
        if constexpr (std::is_same<R, float>::value) {
            PushArgs<float> a = push_args(k);
            if ((rc = arm_leave_list(a))) return rc;
            a.dst = P[cur ^ 1];
            a.cursor = cell_count;
            a.hist = cell_count2;
            if (dev_count) { a.np_dev = dc; a.np = cap; }      // (a.np then only sizes the grid)
            cudaMemsetAsync(cell_count2, 0, (size_t)g.nc * sizeof(unsigned), stream);
            if (want_stats) cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), stream);
            rec(0);
            const bool fma = prm.fp_mode == CPIC_FP_CONTRACT;
            const float aq = fabsf((float)a.qdt_2mc);
            const bool fd = push2_fastds && (aq == 0.f || (aq > 1e-12f && aq < 1e12f));
            if (want_stats) rc = fma ? launch_push2r<true, true, false>(a) : launch_push2r<false, true, false>(a);
            else if (fd) rc = fma ? launch_push2r<true, false, true>(a) : launch_push2r<false, false, true>(a);
            else rc = fma ? launch_push2r<true, false, false>(a) : launch_push2r<false, false, false>(a);
            rec(1);
            ev_valid[0] = true;
            if (rc) return rc;
            cur ^= 1;
            std::swap(cell_count, cell_count2);      // cell_count: histogram of the cells the particles are in now
            hist_valid = true;
            cursor_valid = false;
            seg_valid = false;
            want_hist = false;
        }
        return CPIC_OK;
    }
    PushArgs<R> push_args(const cpic_consts& k) {
        PushArgs<R> a;
        a.p = P[cur]; a.np = np; a.ip = interp; a.acc = acc;
        a.qdt_2mc = (R)k.qdt_2mc; a.cdt_dx = (R)k.cdt_dx; a.cdt_dy = (R)k.cdt_dy; a.cdt_dz = (R)k.cdt_dz; a.qsp = (R)k.qsp;
        a.nx = g.nx; a.ny = g.ny; a.nz = g.nz; a.ng = g.ng; a.gx = g.gx; a.gy = g.gy;
        a.magic_gx = (unsigned)(((1ull << 32) + g.gx - 1) / g.gx); a.magic_gy = (unsigned)(((1ull << 32) + g.gy - 1) / g.gy);
        a.periodic = prm.boundary == CPIC_BOUNDARY_PERIODIC ? g.per : 0;
        a.reflect = reflecting() ? 7 : 0;
        a.stats = stats;
        a.hist = nullptr;
        a.dst = P[cur]; a.cursor = nullptr;
        a.leave_list = nullptr; a.leave_count = nullptr; a.leave_cap = 0; a.leave_lo = 0; a.leave_hi = 0; a.leave_off = 0;
        a.dep_thresh = dep_thresh; a.dep_rounds = dep_rounds;
        a.np_dev = nullptr;
        a.ev_cell = nullptr; a.ev_row = nullptr; a.ev_k = 0; a.ev_overflow = nullptr;
        a.priv_nc = use_priv() ? (int)g.nc : 0;
        return a;
    }
    int push(const cpic_consts& k) override {
        if (np == 0) return CPIC_OK;
        PushArgs<R> a;
        a.np_dev = nullptr; a.priv_nc = use_priv() ? (int)g.nc : 0;
        a.ev_cell = nullptr; a.ev_row = nullptr; a.ev_k = 0; a.ev_overflow = nullptr;
        a.dst = P[cur]; a.cursor = nullptr;
        a.p = P[cur]; a.np = np; a.ip = interp; a.acc = acc;
        a.qdt_2mc = (R)k.qdt_2mc; a.cdt_dx = (R)k.cdt_dx; a.cdt_dy = (R)k.cdt_dy; a.cdt_dz = (R)k.cdt_dz; a.qsp = (R)k.qsp;
        a.nx = g.nx; a.ny = g.ny; a.nz = g.nz; a.ng = g.ng; a.gx = g.gx; a.gy = g.gy;
        a.magic_gx = (unsigned)(((1ull << 32) + g.gx - 1) / g.gx); a.magic_gy = (unsigned)(((1ull << 32) + g.gy - 1) / g.gy);
        a.periodic = prm.boundary == CPIC_BOUNDARY_PERIODIC ? g.per : 0;
        a.reflect = reflecting() ? 7 : 0;
        a.stats = stats;
        a.hist = nullptr;
        hist_valid = false; cursor_valid = false; seg_valid = false;
        { int rc0 = arm_leave_list(a); if (rc0) return rc0; }
        const int dep_mode = prm.deposit_mode == CPIC_DEPOSIT_AUTO ? CPIC_DEPOSIT_WARP : prm.deposit_mode;
        const bool p2 = std::is_same<R, float>::value && use_push2 && dep_mode == CPIC_DEPOSIT_WARP;
        if (want_hist && prm.enable_sort && p2) {      // the next step sorts: let k_push2 count the new cells
            cudaMemsetAsync(cell_count, 0, (size_t)g.nc * sizeof(unsigned), stream);
            a.hist = cell_count;
            hist_valid = true;
        }
        want_hist = false;
        a.dep_thresh = dep_thresh; a.dep_rounds = dep_rounds;
        if (want_stats) cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), stream);
        rec(0);
        const int rc = launch_inplace(a);
        rec(1);
        ev_valid[0] = true;
        return rc;
    }
    // the in-place push kernel for this context's real type, deposit mode and floating-point policy
    int launch_inplace(const PushArgs<R>& a) {
        const int dep = prm.deposit_mode == CPIC_DEPOSIT_AUTO ? CPIC_DEPOSIT_WARP : prm.deposit_mode;
        const bool fma = prm.fp_mode == CPIC_FP_CONTRACT;
        if constexpr (std::is_same<R, float>::value) {
            if (use_push2 && dep == CPIC_DEPOSIT_WARP) {
                // packed sqrt/div fast path only when qdt_2mc is a well-scaled normal number (or zero)
                const float aq = fabsf((float)a.qdt_2mc);
                const bool fd = push2_fastds && (aq == 0.f || (aq > 1e-12f && aq < 1e12f));
                if (want_stats) return fma ? launch_push2<true, true, false>(a) : launch_push2<false, true, false>(a);
                if (fd) return fma ? launch_push2<true, false, true>(a) : launch_push2<false, false, true>(a);
                return fma ? launch_push2<true, false, false>(a) : launch_push2<false, false, false>(a);
            }
        }
        if (dep == CPIC_DEPOSIT_ATOMIC) return fma ? launch_push<true, 1>(a) : launch_push<false, 1>(a);
        if (dep == CPIC_DEPOSIT_ATOMIC_V4) return fma ? launch_push<true, 2>(a) : launch_push<false, 2>(a);
        if (dep == CPIC_DEPOSIT_ORDERED) return ordered_push(a, fma);
        return fma ? launch_push<true, 3>(a) : launch_push<false, 3>(a);
    }
    // CPIC_DEPOSIT_ORDERED (parity runs): the push records every streak (cell + 12 currents) as event (particle, streak);
    // a stable radix sort of the event indices by cell (cub) makes each cell's events contiguous in (particle, streak)
    // order, and k_ordered_accumulate adds them to the accumulator in that order -- the reference's serial summation
    // order (src/push.h:218-254, src/move_p.h:154-190), hence bit-identical accumulators in strict FP mode.
    static constexpr int EV_K = 8;                  // streaks per particle (1 + face crossings of one step)
    static constexpr long long EV_MAX_NP = 1ll << 22;
    int* ev_cell = nullptr;
    R* ev_row = nullptr;
    unsigned *ev_keys = nullptr, *ev_order[2] = {nullptr, nullptr}, *ev_overflow = nullptr;
    void* ev_tmp = nullptr;
    size_t ev_tmp_bytes = 0;
    long long ev_cap = 0;
    int ordered_push(PushArgs<R> a, bool fma) {
        int rc;
        if (a.np > EV_MAX_NP) return fail(CPIC_E_CAPACITY, "CPIC_DEPOSIT_ORDERED: %lld particles in one push exceed the mode's limit %lld", (long long)a.np, EV_MAX_NP);
        const long long m = (long long)a.np * EV_K;
        if (m > ev_cap) {
            cudaFree(ev_cell); cudaFree(ev_row); cudaFree(ev_keys); cudaFree(ev_order[0]); cudaFree(ev_order[1]); cudaFree(ev_tmp);
            ev_cell = nullptr; ev_row = nullptr; ev_keys = nullptr; ev_order[0] = ev_order[1] = nullptr; ev_tmp = nullptr; ev_cap = 0;
            if ((rc = cuda(cudaMalloc(&ev_cell, (size_t)m * sizeof(int)), "cudaMalloc(event cells)"))) return rc;
            if ((rc = cuda(cudaMalloc(&ev_row, (size_t)m * 12 * sizeof(R)), "cudaMalloc(event rows)"))) return rc;
            if ((rc = cuda(cudaMalloc(&ev_keys, (size_t)m * sizeof(unsigned)), "cudaMalloc(event keys)"))) return rc;
            for (auto& o : ev_order)
                if ((rc = cuda(cudaMalloc(&o, (size_t)m * sizeof(unsigned)), "cudaMalloc(event order)"))) return rc;
            ev_tmp_bytes = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, ev_tmp_bytes, (const unsigned*)nullptr, (unsigned*)nullptr, (const unsigned*)nullptr,
                                            (unsigned*)nullptr, (int)m, 0, 32, stream);
            if ((rc = cuda(cudaMalloc(&ev_tmp, ev_tmp_bytes + 16), "cudaMalloc(sort scratch)"))) return rc;
            ev_cap = m;
        }
        if (!ev_overflow && (rc = cuda(cudaMalloc(&ev_overflow, sizeof(unsigned)), "cudaMalloc"))) return rc;
        cudaMemsetAsync(ev_cell, 0xff, (size_t)m * sizeof(int), stream);      // -1: unused (sorts behind every cell)
        cudaMemsetAsync(ev_overflow, 0, sizeof(unsigned), stream);
        a.ev_cell = ev_cell; a.ev_row = ev_row; a.ev_k = EV_K; a.ev_overflow = ev_overflow;
        if ((rc = fma ? launch_push<true, 4>(a) : launch_push<false, 4>(a))) return rc;
        k_iota<<<blocks_for(m), 256, 0, stream>>>(ev_order[0], m);
        if ((rc = check_launch("k_iota"))) return rc;
        size_t tb = ev_tmp_bytes;
        if ((rc = cuda(cub::DeviceRadixSort::SortPairs(ev_tmp, tb, reinterpret_cast<const unsigned*>(ev_cell), ev_keys, ev_order[0], ev_order[1],
                                                       (int)m, 0, 32, stream), "cub::DeviceRadixSort"))) return rc;
        k_ordered_accumulate<R><<<blocks_for(g.nc * 12), 256, 0, stream>>>(ev_keys, ev_order[1], m, ev_row, acc, g.nc);
        if ((rc = check_launch("k_ordered_accumulate"))) return rc;
        unsigned over = 0;
        if ((rc = cuda(cudaMemcpyAsync(&over, ev_overflow, sizeof over, cudaMemcpyDeviceToHost, stream), "D2H"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(stream), "ordered deposit"))) return rc;
        if (over) return fail(CPIC_E_CAPACITY, "CPIC_DEPOSIT_ORDERED: a particle crossed more than %d cell faces in one step", EV_K - 1);
        return CPIC_OK;
    }
    // ------------------------------------------------------------------ host-resident step (cpic_step_host)
    // Three streams: `hs_up` carries the H2D copies of chunk i+1, the context's stream packs / pushes / unpacks
    // chunk i, `hs_dn` carries the D2H copies of chunk i-1.  Two staging chunks per direction; events order
    // the hand-overs.  The push of a chunk is the ordinary in-place kernel on a sub-range of the store.
    static constexpr long long HS_CHUNK = 1ll << 22;
    cudaStream_t hs_up = nullptr, hs_dn = nullptr;
    char* hs_buf[4] = {nullptr, nullptr, nullptr, nullptr};   // up0 up1 dn0 dn1
    long long hs_cap = 0;
    cudaEvent_t hs_ev[8]{};   // up_done[2] up_free[2] dn_ready[2] dn_free[2]
    int ensure_host_stream() {
        if (hs_up) return CPIC_OK;
        int rc;
        hs_cap = HS_CHUNK;
        if (const char* e = getenv("CPIC_HOST_CHUNK")) hs_cap = std::max(64ll, atoll(e) / 64 * 64);   // developer / test knob
        hs_cap = std::min<long long>(cap, hs_cap);
        if ((rc = cuda(cudaStreamCreateWithFlags(&hs_up, cudaStreamNonBlocking), "cudaStreamCreate"))) return rc;
        if ((rc = cuda(cudaStreamCreateWithFlags(&hs_dn, cudaStreamNonBlocking), "cudaStreamCreate"))) return rc;
        for (auto& e : hs_ev)
            if ((rc = cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate"))) return rc;
        for (auto& b : hs_buf)
            if ((rc = cuda(cudaMalloc(&b, (size_t)hs_cap * (7 * sizeof(R) + sizeof(int))), "cudaMalloc(host-step staging)"))) return rc;
        return CPIC_OK;
    }
    // First half of a host-resident step: fields up, interpolators, then the particles streamed through the device chunk
    // by chunk (H2D / pack / in-place push / unpack / D2H on three streams).  slab: z is open -- the leavers stay in the
    // store (and in `out`, to be patched by host_patch_phase) with their store indices listed for the migration, and a
    // particle arriving from the host in a z ghost plane counts as a bad cell.
    int host_push_phase(const cpic_consts& k, const void* const in[8], void* const out[8], long long n,
                        const void* const fin[9], bool slab) override {
        if (n < 0 || n > cap) return fail(CPIC_E_CAPACITY, "step_host: %lld particles exceed capacity %lld", n, cap);
        int rc;
        if (dev_count && (rc = sync_np())) return rc;
        if ((rc = ensure_host_stream())) return rc;
        rec(6);
        // fields first: the interpolators every chunk's push gathers from (example/example.cpp:233-236)
        for (int m = 0; m < F_N; ++m)
            if ((rc = cuda(cudaMemcpyAsync(fields + (long long)m * nc_pad, fin[m], (size_t)g.nc * sizeof(R), cudaMemcpyHostToDevice, stream), "H2D fields"))) return rc;
        if ((rc = load_interpolator())) return rc;
        if ((rc = clear_accumulator())) return rc;
        cudaMemsetAsync(bad, 0, sizeof(unsigned), stream);
        if (want_stats) cudaMemsetAsync(stats, 0, 8 * sizeof(unsigned long long), stream);
        np = n;
        hist_valid = false; cursor_valid = false; seg_valid = false; leavers_valid = false; ghost_clean = false; want_hist = false;
        PushArgs<R> a0 = push_args(k);
        const long long plane = (long long)g.gx * g.gy;
        const long long c_lo = slab ? plane : 0, c_hi = slab ? (long long)(g.nz + 1) * plane : g.nc;
        if (slab) {      // every particle is checked to lie in an interior plane below: the ghost planes are empty
            ghost_clean = true;
            if ((rc = arm_leave_list(a0))) return rc;
        }
        const int32_t* cin = static_cast<const int32_t*>(in[7]);
        long long chunk = 0;
        for (long long first = 0; first < n; first += hs_cap, ++chunk) {
            const long long cn = std::min(hs_cap, n - first);
            const int b = (int)(chunk & 1);
            SendBuf<R> up = carve_sendbuf<R>(hs_buf[b], hs_cap), dn = carve_sendbuf<R>(hs_buf[2 + b], hs_cap);
            // H2D of this chunk, once the pack of chunk-2 has drained the staging buffer
            if (chunk >= 2) cudaStreamWaitEvent(hs_up, hs_ev[2 + b], 0);
            for (int m = 0; m < 7; ++m)
                if ((rc = cuda(cudaMemcpyAsync(up.m[m], (const R*)in[m] + first, (size_t)cn * sizeof(R), cudaMemcpyHostToDevice, hs_up), "H2D particles"))) return rc;
            if ((rc = cuda(cudaMemcpyAsync(up.cell, cin + first, (size_t)cn * sizeof(int), cudaMemcpyHostToDevice, hs_up), "H2D cell"))) return rc;
            cudaEventRecord(hs_ev[0 + b], hs_up);
            // records, push (src/push.h + src/move_p.h on this chunk), members
            cudaStreamWaitEvent(stream, hs_ev[0 + b], 0);
            k_pack_records_checked<R><<<blocks_for(cn), 256, 0, stream>>>(P[cur], first, up, cn, c_lo, c_hi, bad);
            if ((rc = check_launch("k_pack_records_checked"))) return rc;
            cudaEventRecord(hs_ev[2 + b], stream);
            PushArgs<R> a = a0;
            a.p.rec = P[cur].rec + first; a.dst = a.p; a.np = cn;
            a.leave_off = (unsigned)first;      // the leaver list holds indices of the whole store
            if ((rc = launch_inplace(a))) return rc;
            if (out) {
                if (chunk >= 2) cudaStreamWaitEvent(stream, hs_ev[6 + b], 0);
                k_unpack_records<R><<<blocks_for(cn), 256, 0, stream>>>(P[cur], first, dn, cn);
                if ((rc = check_launch("k_unpack_records"))) return rc;
                cudaEventRecord(hs_ev[4 + b], stream);
                cudaStreamWaitEvent(hs_dn, hs_ev[4 + b], 0);
                for (int m = 0; m < 7; ++m)
                    if (out[m] && (rc = cuda(cudaMemcpyAsync((R*)out[m] + first, dn.m[m], (size_t)cn * sizeof(R), cudaMemcpyDeviceToHost, hs_dn), "D2H particles"))) return rc;
                if (out[7] && (rc = cuda(cudaMemcpyAsync((int32_t*)out[7] + first, dn.cell, (size_t)cn * sizeof(int), cudaMemcpyDeviceToHost, hs_dn), "D2H cell"))) return rc;
                cudaEventRecord(hs_ev[6 + b], hs_dn);
            }
        }
        return CPIC_OK;
    }
    int host_fields_out(void* const fout[9]) override {
        int rc;
        if (fout)
            for (int m = 0; m < F_N; ++m)
                if (fout[m] && (rc = cuda(cudaMemcpyAsync(fout[m], fields + (long long)m * nc_pad, (size_t)g.nc * sizeof(R), cudaMemcpyDeviceToHost, stream), "D2H fields"))) return rc;
        return CPIC_OK;
    }
    // Second half in slab mode, after the migration: the device store is final; `out` holds what the chunks downloaded
    // before it.  Patch it: the holes the leavers left (filled from the store's tail by the extraction) and the range the
    // arrivals were appended to.  Synchronises.
    int host_patch_phase(void* const out[8], long long n_in, long long out_capacity, long long* n_out) override {
        int rc;
        unsigned hc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long hd[4] = {0, 0, 0, 0};
        if (!dev_count || !mig_counters) return fail(CPIC_E_INVALID, "step_host: no device-counted migration preceded the patch");
        if ((rc = cuda(cudaMemcpyAsync(hc, mig_counters, sizeof hc, cudaMemcpyDeviceToHost, stream), "D2H migration counters"))) return rc;
        if ((rc = cuda(cudaMemcpyAsync(hd, dc, sizeof hd, cudaMemcpyDeviceToHost, stream), "D2H counts"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(hs_dn), "step_host (D2H)"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(hs_up), "step_host (H2D)"))) return rc;
        unsigned nbad = 0;
        if ((rc = cuda(cudaMemcpyAsync(&nbad, bad, sizeof(unsigned), cudaMemcpyDeviceToHost, stream), "D2H"))) return rc;
        if ((rc = sync_np())) return rc;      // (synchronises; reports overflows of the migration)
        if (nbad) { np = 0; return fail(CPIC_E_BAD_CELL, "step_host: %u particles lie outside the slab's interior planes", nbad); }
        const long long n_after = n_in - hd[2] - hd[3], n_final = np, nh = hc[3];
        if (n_out) *n_out = n_final;
        rec(7);
        ev_valid[3] = true;
        if (!out) return CPIC_OK;
        if (n_final > out_capacity) return fail(CPIC_E_CAPACITY, "step_host: %lld particles after the migration exceed the output capacity %lld", n_final, out_capacity);
        // (a) arrivals: the store range [n_after, n_final)
        for (long long first = n_after; first < n_final; first += hs_cap) {
            const long long cn = std::min(hs_cap, n_final - first);
            SendBuf<R> dn = carve_sendbuf<R>(hs_buf[2], hs_cap);
            k_unpack_records<R><<<blocks_for(cn), 256, 0, stream>>>(P[cur], first, dn, cn);
            if ((rc = check_launch("k_unpack_records"))) return rc;
            for (int m = 0; m < 7; ++m)
                if (out[m] && (rc = cuda(cudaMemcpyAsync((R*)out[m] + first, dn.m[m], (size_t)cn * sizeof(R), cudaMemcpyDeviceToHost, stream), "D2H arrivals"))) return rc;
            if (out[7] && (rc = cuda(cudaMemcpyAsync((int32_t*)out[7] + first, dn.cell, (size_t)cn * sizeof(int), cudaMemcpyDeviceToHost, stream), "D2H arrivals"))) return rc;
            if ((rc = cuda(cudaStreamSynchronize(stream), "step_host (arrivals)"))) return rc;
        }
        // (b) holes: positions mig_lists[0 .. nh) now hold particles moved in from the tail
        std::vector<unsigned> idx;
        std::vector<char> rec_host;
        for (long long j0 = 0; j0 < nh; j0 += hs_cap) {
            const long long cn = std::min(hs_cap, nh - j0);
            SendBuf<R> dn = carve_sendbuf<R>(hs_buf[2], hs_cap);
            k_gather_records<R><<<blocks_for(cn), 256, 0, stream>>>(P[cur], mig_lists + j0, dn, cn);
            if ((rc = check_launch("k_gather_records"))) return rc;
            idx.resize((size_t)cn);
            rec_host.resize((size_t)cn * (7 * sizeof(R) + sizeof(int)));
            if ((rc = cuda(cudaMemcpyAsync(idx.data(), mig_lists + j0, (size_t)cn * sizeof(unsigned), cudaMemcpyDeviceToHost, stream), "D2H hole list"))) return rc;
            for (int m = 0; m < 7; ++m)
                if ((rc = cuda(cudaMemcpyAsync(rec_host.data() + (size_t)m * cn * sizeof(R), dn.m[m], (size_t)cn * sizeof(R), cudaMemcpyDeviceToHost, stream), "D2H patches"))) return rc;
            if ((rc = cuda(cudaMemcpyAsync(rec_host.data() + (size_t)7 * cn * sizeof(R), dn.cell, (size_t)cn * sizeof(int), cudaMemcpyDeviceToHost, stream), "D2H patches"))) return rc;
            if ((rc = cuda(cudaStreamSynchronize(stream), "step_host (patches)"))) return rc;
            for (int m = 0; m < 7; ++m) {
                if (!out[m]) continue;
                R* o = static_cast<R*>(out[m]);
                const R* v = reinterpret_cast<const R*>(rec_host.data() + (size_t)m * cn * sizeof(R));
                for (long long j = 0; j < cn; ++j) o[idx[(size_t)j]] = v[j];
            }
            if (out[7]) {
                int32_t* o = static_cast<int32_t*>(out[7]);
                const int32_t* v = reinterpret_cast<const int32_t*>(rec_host.data() + (size_t)7 * cn * sizeof(R));
                for (long long j = 0; j < cn; ++j) o[idx[(size_t)j]] = v[j];
            }
        }
        return CPIC_OK;
    }
    int step_host(const cpic_consts& k, const void* const in[8], void* const out[8], long long n,
                  const void* const fin[9], void* const fout[9], double* energies) override {
        int rc;
        if ((rc = host_push_phase(k, in, out, n, fin, false))) return rc;
        // field side of the step (example/example.cpp:248-266); the last chunks' D2H overlaps it
        const double hx = sizeof(R) == 4 ? (double)(0.5f * (float)k.px) : 0.5 * k.px;
        const double hy = sizeof(R) == 4 ? (double)(0.5f * (float)k.py) : 0.5 * k.py;
        const double hz = sizeof(R) == 4 ? (double)(0.5f * (float)k.pz) : 0.5 * k.pz;
        if ((rc = unload_accumulator(k))) return rc;
        if ((rc = advance_b(hx, hy, hz))) return rc;
        if ((rc = advance_e(k.px, k.py, k.pz, k.dt_eps0))) return rc;
        if ((rc = advance_b(hx, hy, hz))) return rc;
        double h[2] = {0, 0};
        if (energies) {
            if ((rc = energies_async(en_dev))) return rc;
            if ((rc = cuda(cudaMemcpyAsync(h, en_dev, sizeof h, cudaMemcpyDeviceToHost, stream), "D2H energies"))) return rc;
        }
        if ((rc = host_fields_out(fout))) return rc;
        unsigned nbad = 0;
        if ((rc = cuda(cudaMemcpyAsync(&nbad, bad, sizeof(unsigned), cudaMemcpyDeviceToHost, stream), "D2H"))) return rc;
        rec(7);
        ev_valid[3] = true;
        if ((rc = cuda(cudaStreamSynchronize(stream), "step_host"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(hs_dn), "step_host (D2H)"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(hs_up), "step_host (H2D)"))) return rc;
        if (nbad) { np = 0; return fail(CPIC_E_BAD_CELL, "step_host: %u particles have a cell index outside [0,%lld)", nbad, g.nc); }
        if (energies) {
            energies[0] = 0.5 * h[0];
            energies[1] = prm.solver == CPIC_SOLVER_EM ? 0.5 * h[1] : 0.0;
        }
        return CPIC_OK;
    }
    int uncenter(double qdt_2mc) override {
        if (np == 0) return CPIC_OK;
        if (prm.fp_mode == CPIC_FP_CONTRACT) k_uncenter<R, true><<<blocks_for(np), 256, 0, stream>>>(P[cur], np, interp, (R)qdt_2mc);
        else k_uncenter<R, false><<<blocks_for(np), 256, 0, stream>>>(P[cur], np, interp, (R)qdt_2mc);
        return check_launch("k_uncenter");
    }

    int scan_cells(long long n = -1) { return scan_buf(cell_count, n < 0 ? g.nc : n); }
    int scan_buf(unsigned* buf, long long n) {
        // exclusive scan of buf[0, n) in place (3 levels of 2048-wide tiles); n = nc or nc + 1
        unsigned* const cell_count = buf;
        const long long n_l1 = (n + SCAN_TILE - 1) / SCAN_TILE, n_l2 = (n_l1 + SCAN_TILE - 1) / SCAN_TILE;
        k_scan_tile<<<(unsigned)n_l1, 256, 0, stream>>>(cell_count, cell_count, n, scan_l1);
        int rc = check_launch("k_scan_tile");
        if (rc) return rc;
        if (n_l1 > 1) {
            k_scan_tile<<<(unsigned)n_l2, 256, 0, stream>>>(scan_l1, scan_l1, n_l1, scan_l2);
            if ((rc = check_launch("k_scan_tile"))) return rc;
            if (n_l2 > 1) {
                k_scan_tile<<<1, 256, 0, stream>>>(scan_l2, scan_l2, n_l2, nullptr);
                if ((rc = check_launch("k_scan_tile"))) return rc;
                k_scan_add<<<(unsigned)n_l2, 256, 0, stream>>>(scan_l1, n_l1, scan_l2);
                if ((rc = check_launch("k_scan_add"))) return rc;
            }
            k_scan_add<<<(unsigned)n_l1, 256, 0, stream>>>(cell_count, n, scan_l1);
            if ((rc = check_launch("k_scan_add"))) return rc;
        }
        return CPIC_OK;
    }
    // the histogram kernel counts out-of-range cell indices (set_num_particles over uninitialised records, writes through
    // device_ptr): report them instead of scattering with them
    int check_bad_cells(const char* what) {
        unsigned nbad = 0;
        int rc;
        if ((rc = cuda(cudaMemcpyAsync(&nbad, bad, sizeof(unsigned), cudaMemcpyDeviceToHost, stream), "D2H"))) return rc;
        if ((rc = cuda(cudaStreamSynchronize(stream), what))) return rc;
        if (nbad) return fail(CPIC_E_BAD_CELL, "%s: %u particles have a cell index outside [0,%lld)", what, nbad, g.nc);
        return CPIC_OK;
    }
    int sort() override {
        if (!prm.enable_sort) return fail(CPIC_E_INVALID, "sort_particles: context was created with enable_sort=0");
        if (np == 0) return CPIC_OK;
        int rc;
        rec(2);
        if (!hist_valid) {
            cudaMemsetAsync(cell_count, 0, (size_t)g.nc * sizeof(unsigned), stream);
            cudaMemsetAsync(bad, 0, sizeof(unsigned), stream);
            k_cell_histogram<R><<<blocks_for(np), 256, 0, stream>>>(P[cur], np, g.nc, cell_count, bad);
            if ((rc = check_launch("k_cell_histogram"))) return rc;
            if ((rc = check_bad_cells("sort_particles"))) return rc;
        }
        hist_valid = false; cursor_valid = false; seg_valid = false; leavers_valid = false;
        if ((rc = scan_cells())) return rc;
        k_sort_scatter<R><<<blocks_for(np), 256, 0, stream>>>(P[cur], P[cur ^ 1], np, cell_count);
        if ((rc = check_launch("k_sort_scatter"))) return rc;
        cur ^= 1;
        rec(3);
        ev_valid[1] = true;
        return CPIC_OK;
    }

    int init_uniform(const UniformPlasmaArgs& a) override {
        if (a.count < 0 || a.count > cap) return fail(CPIC_E_CAPACITY, "init_uniform_plasma: %lld particles exceed capacity %lld", a.count, cap);
        if (a.gnx != g.nx || a.gny != g.ny) return fail(CPIC_E_INVALID, "init_uniform_plasma: x/y extents must equal the context's");
        np = a.count;
        hist_valid = false; cursor_valid = false; seg_valid = false; leavers_valid = false; ghost_clean = true;
        if (np == 0) return CPIC_OK;
        k_init_uniform_plasma<R><<<blocks_for(np), 256, 0, stream>>>(P[cur], a);
        return check_launch("k_init_uniform_plasma");
    }

    int device_ptr(int which, void** ptr, int64_t* count, int64_t* stride) override {
        // particle members live inside the records: member `which` of particle n is at ptr + n*stride reals
        // (record order dx dy dz cell ux uy uz w; the cell is an int32/int64 in a real-sized slot)
        if (which >= 0 && which < 8) {
            static const int slot[8] = {0, 1, 2, 4, 5, 6, 7, 3};
            *ptr = reinterpret_cast<R*>(P[cur].rec) + slot[which];
            if (count) *count = cap;
            if (stride) *stride = 8;
            return CPIC_OK;
        }
        if (which == 16) { *ptr = fields; if (count) *count = g.nc; if (stride) *stride = nc_pad; return CPIC_OK; }
        if (which == 17) { *ptr = interp; if (count) *count = g.nc; if (stride) *stride = S; return CPIC_OK; }
        if (which == 18) { *ptr = acc; if (count) *count = g.nc; if (stride) *stride = 12; return CPIC_OK; }
        return fail(CPIC_E_INVALID, "device_ptr: which=%d", which);
    }
};

int validate(const cpic_params& p, std::string& why) {
    char buf[256];
    if (p.nx < 1 || p.ny < 1 || p.nz < 1) { why = "nx, ny, nz must be >= 1"; return CPIC_E_INVALID; }
    if (p.ng != 1) { why = "ng must be 1 (the reference's mover hard-wires one ghost layer, src/move_p.h:19-47)"; return CPIC_E_INVALID; }
    if (p.real_bytes != 4 && p.real_bytes != 8) { why = "real_bytes must be 4 or 8"; return CPIC_E_INVALID; }
    if (p.solver != CPIC_SOLVER_EM && p.solver != CPIC_SOLVER_ES_1D) { why = "unknown solver"; return CPIC_E_INVALID; }
    if (p.solver == CPIC_SOLVER_ES_1D && (p.ny > 1 || p.nz > 1)) { why = "ES field solver supports 1D only (example/example.cpp:69-74)"; return CPIC_E_INVALID; }
    if (p.boundary != CPIC_BOUNDARY_PERIODIC && p.boundary != CPIC_BOUNDARY_REFLECT) { why = "unknown boundary"; return CPIC_E_INVALID; }
    if (p.boundary == CPIC_BOUNDARY_REFLECT && p.solver != CPIC_SOLVER_EM) { why = "Boundary::Reflect (reflecting particles + conducting walls) is implemented for the EM solver only"; return CPIC_E_UNSUPPORTED; }
    if (p.max_particles < 0 || p.max_particles > (1ll << 31) - 64) { why = "max_particles must be in [0, 2^31)"; return CPIC_E_INVALID; }
    const long long nc = (long long)(p.nx + 2) * (p.ny + 2) * (p.nz + 2);
    if (nc > (1ll << 31) - 1) { snprintf(buf, sizeof buf, "%lld cells overflow the int cell index", nc); why = buf; return CPIC_E_INVALID; }
    if (p.fp_mode != CPIC_FP_STRICT && p.fp_mode != CPIC_FP_CONTRACT) { why = "unknown fp_mode"; return CPIC_E_INVALID; }
    if (p.deposit_mode < 0 || p.deposit_mode > CPIC_DEPOSIT_ORDERED) { why = "unknown deposit_mode"; return CPIC_E_INVALID; }
    return CPIC_OK;
}

#define CTX_OR_FAIL(ctx) \
    if (!(ctx)) return CPIC_E_INVALID; \
    CtxBase* c = reinterpret_cast<CtxBase*>(ctx); \
    { cudaError_t e_ = cudaSetDevice(c->prm.device); if (e_ != cudaSuccess) return c->cuda(e_, "cudaSetDevice"); }

// entry points that read the host's particle count: leave the device-count mode of the slab exchange first
#define CTX_HOSTNP(ctx) \
    CTX_OR_FAIL(ctx); \
    { int rc_np_ = c->sync_np(); if (rc_np_) return rc_np_; }

}  // namespace

extern "C" {

int cpic_abi_version(void) { return CPIC_ABI_VERSION; }

const char* cpic_last_error(const cpic_ctx* ctx) {
    if (!ctx) return g_create_error.c_str();
    return reinterpret_cast<const CtxBase*>(ctx)->err.c_str();
}

int cpic_create(const cpic_params* params, cpic_ctx** out) {
    if (!params || !out) { g_create_error = "cpic_create: null argument"; return CPIC_E_INVALID; }
    *out = nullptr;
    int rc = validate(*params, g_create_error);
    if (rc) return rc;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return CPIC_E_CUDA;
    }
    if (params->device < 0 || params->device >= ndev) { g_create_error = "device ordinal out of range"; return CPIC_E_INVALID; }
    CtxBase* c = (params->real_bytes == 4) ? static_cast<CtxBase*>(new Ctx<float>()) : static_cast<CtxBase*>(new Ctx<double>());
    c->prm = *params;
    c->g = make_grid(params->nx, params->ny, params->nz, params->ng);
    rc = (params->real_bytes == 4) ? static_cast<Ctx<float>*>(c)->init() : static_cast<Ctx<double>*>(c)->init();
    if (rc) { g_create_error = c->err; delete c; return rc; }
    *out = reinterpret_cast<cpic_ctx*>(c);
    return CPIC_OK;
}

int cpic_create_species(cpic_ctx* parent, int64_t max_particles, cpic_ctx** out) {
    CTX_OR_FAIL(parent);
    if (!out || max_particles < 0 || max_particles > (1ll << 31) - 64) return c->fail(CPIC_E_INVALID, "create_species: bad arguments");
    *out = nullptr;
    int rc;
    CtxBase* sp;
    if (c->prm.real_bytes == 4) {
        auto* ch = new Ctx<float>();
        ch->parent = static_cast<Ctx<float>*>(c);
        if (ch->parent->parent) { delete ch; return c->fail(CPIC_E_INVALID, "create_species: the parent must be a context of cpic_create"); }
        ch->prm = c->prm; ch->prm.max_particles = max_particles; ch->g = c->g;
        rc = ch->init();
        sp = ch;
    } else {
        auto* ch = new Ctx<double>();
        ch->parent = static_cast<Ctx<double>*>(c);
        if (ch->parent->parent) { delete ch; return c->fail(CPIC_E_INVALID, "create_species: the parent must be a context of cpic_create"); }
        ch->prm = c->prm; ch->prm.max_particles = max_particles; ch->g = c->g;
        rc = ch->init();
        sp = ch;
    }
    if (rc) { c->err = sp->err; delete sp; return rc; }
    *out = reinterpret_cast<cpic_ctx*>(sp);
    return CPIC_OK;
}

int cpic_step_species(cpic_ctx* ctx, cpic_ctx* const* species, const cpic_consts* consts, int32_t nspecies, int64_t nsteps,
                      int32_t sort_interval, double* energies) {
    CTX_OR_FAIL(ctx);
    if (!species || !consts || nspecies < 1 || nsteps < 0 || sort_interval < CPIC_SORT_FUSED) return c->fail(CPIC_E_INVALID, "step_species: bad arguments");
    for (int s = 0; s < nspecies; ++s) {
        if (!species[s]) return c->fail(CPIC_E_INVALID, "step_species: null species");
        CtxBase* sp = reinterpret_cast<CtxBase*>(species[s]);
        if (sp != c && !sp->is_species_of(c)) return c->fail(CPIC_E_INVALID, "step_species: species %d does not belong to this context", s);
        int rc = sp->sync_np();
        if (rc) { c->err = sp->err; return rc; }
    }
    int rc = CPIC_OK;
    double* en = nullptr;
    if (energies && nsteps > 0)
        if ((rc = c->cuda(cudaMalloc(&en, (size_t)nsteps * 2 * sizeof(double)), "cudaMalloc(energies)"))) return rc;
    const cpic_consts* k = &consts[0];      // the field constants (dt, dx, px, dt_eps0 ...) are common to all species
    const double hx = (c->prm.real_bytes == 4) ? (double)(0.5f * (float)k->px) : 0.5 * k->px;
    const double hy = (c->prm.real_bytes == 4) ? (double)(0.5f * (float)k->py) : 0.5 * k->py;
    const double hz = (c->prm.real_bytes == 4) ? (double)(0.5f * (float)k->pz) : 0.5 * k->pz;
    c->rec(6);
    for (int64_t st = 0; st < nsteps && !rc; ++st) {
        // example/example.cpp:221-266 with the push repeated per species between clear and unload (the two-species form
        // of the VPIC decks, decks/vpic/2stream-em0.cxx:207-208: every species deposits into the same accumulator)
        rc = c->load_interpolator();
        if (!rc) rc = c->clear_accumulator();
        for (int s = 0; s < nspecies && !rc; ++s) {
            CtxBase* sp = reinterpret_cast<CtxBase*>(species[s]);
            const bool fused = sort_interval == CPIC_SORT_FUSED && !sp->few_cells();
            if (!fused && sort_interval > 0 && st % sort_interval == 0) rc = sp->sort();
            if (!rc) rc = fused ? sp->push_reorder(consts[s]) : sp->push(consts[s]);
            if (rc) c->err = sp->err;
        }
        if (!rc) rc = c->unload_accumulator(*k);
        if (!rc) rc = c->advance_b(hx, hy, hz);
        if (!rc) rc = c->advance_e(k->px, k->py, k->pz, k->dt_eps0);
        if (!rc) rc = c->advance_b(hx, hy, hz);
        if (!rc && en) rc = c->energies_async(en + 2 * st);
    }
    c->rec(7);
    c->ev_valid[3] = true;
    if (!rc && en) {
        rc = c->cuda(cudaMemcpyAsync(energies, en, (size_t)nsteps * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream), "D2H energies");
        if (!rc) rc = c->cuda(cudaStreamSynchronize(c->stream), "step_species");
        for (int64_t st = 0; st < nsteps && !rc; ++st) {
            energies[2 * st] *= 0.5;
            energies[2 * st + 1] = (c->prm.solver == CPIC_SOLVER_EM) ? 0.5 * energies[2 * st + 1] : 0.0;
        }
    }
    if (en) { cudaStreamSynchronize(c->stream); cudaFree(en); }
    return rc;
}

void cpic_destroy(cpic_ctx* ctx) {
    if (!ctx) return;
    CtxBase* c = reinterpret_cast<CtxBase*>(ctx);
    cudaSetDevice(c->prm.device);
    cudaStreamSynchronize(c->stream);
    delete c;
}

int cpic_sync(cpic_ctx* ctx) { CTX_OR_FAIL(ctx); return c->cuda(cudaStreamSynchronize(c->stream), "sync"); }
int cpic_num_cells(const cpic_ctx* ctx, int64_t* out) { if (!ctx || !out) return CPIC_E_INVALID; *out = reinterpret_cast<const CtxBase*>(ctx)->g.nc; return CPIC_OK; }
int cpic_num_particles(const cpic_ctx* ctx, int64_t* out) {
    if (!out) return CPIC_E_INVALID;
    CTX_HOSTNP(const_cast<cpic_ctx*>(ctx));
    *out = c->np;
    return CPIC_OK;
}

int cpic_upload_particles(cpic_ctx* ctx, const void* dx, const void* dy, const void* dz, const void* ux, const void* uy,
                          const void* uz, const void* w, const int32_t* cell, int64_t n) {
    CTX_HOSTNP(ctx);
    if (n > 0 && (!dx || !dy || !dz || !ux || !uy || !uz || !w || !cell)) return c->fail(CPIC_E_INVALID, "upload_particles: null member array");
    const void* m[7] = {dx, dy, dz, ux, uy, uz, w};
    return c->upload_particles(m, cell, n);
}
int cpic_download_particles(cpic_ctx* ctx, void* dx, void* dy, void* dz, void* ux, void* uy, void* uz, void* w,
                            int32_t* cell, int64_t capacity, int64_t* n_out) {
    CTX_HOSTNP(ctx);
    void* m[7] = {dx, dy, dz, ux, uy, uz, w};
    long long n = 0;
    int rc = c->download_particles(m, cell, capacity, &n);
    if (n_out) *n_out = n;
    return rc;
}
int cpic_upload_fields(cpic_ctx* ctx, const void* const fields[9]) {
    CTX_OR_FAIL(ctx);
    for (int m = 0; m < 9; ++m) if (!fields || !fields[m]) return c->fail(CPIC_E_INVALID, "upload_fields: null member %d", m);
    return c->upload_fields(fields);
}
int cpic_download_fields(cpic_ctx* ctx, void* const fields[9]) { CTX_OR_FAIL(ctx); if (!fields) return c->fail(CPIC_E_INVALID, "download_fields: null"); return c->download_fields(fields); }
int cpic_upload_interpolators(cpic_ctx* ctx, const void* interp) { CTX_OR_FAIL(ctx); if (!interp) return c->fail(CPIC_E_INVALID, "null"); return c->xfer_interp(const_cast<void*>(interp), true); }
int cpic_download_interpolators(cpic_ctx* ctx, void* interp) { CTX_OR_FAIL(ctx); if (!interp) return c->fail(CPIC_E_INVALID, "null"); return c->xfer_interp(interp, false); }
int cpic_upload_accumulators(cpic_ctx* ctx, const void* acc) { CTX_OR_FAIL(ctx); if (!acc) return c->fail(CPIC_E_INVALID, "null"); return c->xfer_acc(const_cast<void*>(acc), true); }
int cpic_download_accumulators(cpic_ctx* ctx, void* acc) { CTX_OR_FAIL(ctx); if (!acc) return c->fail(CPIC_E_INVALID, "null"); return c->xfer_acc(acc, false); }

int cpic_load_interpolator_array(cpic_ctx* ctx) { CTX_OR_FAIL(ctx); return c->load_interpolator(); }
int cpic_initialize_interpolator(cpic_ctx* ctx) { CTX_OR_FAIL(ctx); return c->initialize_interpolator(); }
int cpic_clear_accumulator_array(cpic_ctx* ctx) { CTX_OR_FAIL(ctx); return c->clear_accumulator(); }
int cpic_push(cpic_ctx* ctx, const cpic_consts* k) { CTX_HOSTNP(ctx); if (!k) return c->fail(CPIC_E_INVALID, "push: null consts"); return c->push(*k); }
int cpic_contribute(cpic_ctx* ctx) { CTX_OR_FAIL(ctx); return CPIC_OK; }
int cpic_unload_accumulator_array(cpic_ctx* ctx, const cpic_consts* k) { CTX_OR_FAIL(ctx); if (!k) return c->fail(CPIC_E_INVALID, "unload: null consts"); return c->unload_accumulator(*k); }
int cpic_advance_b(cpic_ctx* ctx, double px, double py, double pz) { CTX_OR_FAIL(ctx); return c->advance_b(px, py, pz); }
int cpic_advance_e(cpic_ctx* ctx, double px, double py, double pz, double dt_eps0) { CTX_OR_FAIL(ctx); return c->advance_e(px, py, pz, dt_eps0); }
int cpic_uncenter_particles(cpic_ctx* ctx, double qdt_2mc) { CTX_HOSTNP(ctx); return c->uncenter(qdt_2mc); }
int cpic_update_ghosts(cpic_ctx* ctx, int which) { CTX_OR_FAIL(ctx); return c->update_ghosts(which); }
int cpic_sort_particles(cpic_ctx* ctx) { CTX_HOSTNP(ctx); return c->sort(); }
int cpic_push_reorder(cpic_ctx* ctx, const cpic_consts* k) { CTX_OR_FAIL(ctx); if (!k) return c->fail(CPIC_E_INVALID, "push_reorder: null consts"); return c->push_reorder(*k); }

int cpic_init_uniform_plasma(cpic_ctx* ctx, int64_t first, int64_t count, int32_t gnx, int32_t gny, int32_t gnz,
                             int32_t nppc, int32_t z0, uint64_t seed, double vthx, double vthy, double vthz,
                             double weight) {
    CTX_HOSTNP(ctx);
    if (nppc < 1 || gnx < 1 || gny < 1 || gnz < 1 || first < 0) return c->fail(CPIC_E_INVALID, "init_uniform_plasma: bad arguments");
    const long long total = (long long)gnx * gny * gnz * nppc;
    if (first + count > total) return c->fail(CPIC_E_INVALID, "init_uniform_plasma: slice exceeds the global particle list");
    // the slice must lie inside this context's z range [z0, z0+nz)
    const long long per_plane = (long long)gnx * gny * nppc;
    if (count > 0 && (first / per_plane < z0 || (first + count - 1) / per_plane >= (long long)z0 + c->g.nz))
        return c->fail(CPIC_E_INVALID, "init_uniform_plasma: slice lies outside this context's z-planes");
    UniformPlasmaArgs a;
    a.first = first; a.count = count; a.gnx = gnx; a.gny = gny; a.gnz = gnz; a.nppc = nppc; a.z0 = z0;
    a.lnx = c->g.nx; a.lny = c->g.ny; a.seed = seed; a.vth[0] = vthx; a.vth[1] = vthy; a.vth[2] = vthz; a.weight = weight;
    return c->init_uniform(a);
}

int cpic_energies(cpic_ctx* ctx, double* e_energy, double* b_energy) {
    CTX_OR_FAIL(ctx);
    double* d = c->energy_scratch();
    int rc = c->energies_async(d);
    if (rc) return rc;
    double h[2] = {0, 0};
    if ((rc = c->cuda(cudaMemcpyAsync(h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream), "D2H energies"))) return rc;
    if ((rc = c->cuda(cudaStreamSynchronize(c->stream), "energies"))) return rc;
    // reference: e_tot*0.5f*dV with dV = 1 (src/fields.h:584-585); B energy is 0 for ES
    if (e_energy) *e_energy = 0.5 * h[0];
    if (b_energy) *b_energy = (c->prm.solver == CPIC_SOLVER_EM) ? 0.5 * h[1] : 0.0;
    return CPIC_OK;
}

int cpic_kinetic_energy(cpic_ctx* ctx, double* out) {
    CTX_HOSTNP(ctx);
    if (!out) return c->fail(CPIC_E_INVALID, "kinetic_energy: null");
    double* d = c->energy_scratch();
    int rc = c->kinetic_async(d);
    if (rc) return rc;
    double h = 0.0;
    if ((rc = c->cuda(cudaMemcpyAsync(&h, d, sizeof h, cudaMemcpyDeviceToHost, c->stream), "D2H kinetic energy"))) return rc;
    if ((rc = c->cuda(cudaStreamSynchronize(c->stream), "kinetic_energy"))) return rc;
    *out = h;
    return CPIC_OK;
}

int cpic_step(cpic_ctx* ctx, const cpic_consts* k, int64_t nsteps, int32_t sort_interval, double* energies) {
    CTX_OR_FAIL(ctx);
    { const int rc_np_ = c->sync_np(); if (rc_np_) return rc_np_; }
    if (!k || nsteps < 0 || sort_interval < CPIC_SORT_FUSED) return c->fail(CPIC_E_INVALID, "step: bad arguments");
    int rc = CPIC_OK;
    double* en = nullptr;
    if (energies && nsteps > 0)
        if ((rc = c->cuda(cudaMalloc(&en, (size_t)nsteps * 2 * sizeof(double)), "cudaMalloc(energies)"))) return rc;
    c->rec(6);
    const double hx = (c->prm.real_bytes == 4) ? (double)(0.5f * (float)k->px) : 0.5 * k->px;
    const double hy = (c->prm.real_bytes == 4) ? (double)(0.5f * (float)k->py) : 0.5 * k->py;
    const double hz = (c->prm.real_bytes == 4) ? (double)(0.5f * (float)k->pz) : 0.5 * k->pz;
    const bool prof = c->prof_on;
    if (prof) {
        while ((long long)c->prof_ev.size() < 5 * nsteps) {
            cudaEvent_t e;
            if ((rc = c->cuda(cudaEventCreate(&e), "cudaEventCreate"))) return rc;
            c->prof_ev.push_back(e);
        }
        c->prof_steps = nsteps;
    }
    for (int64_t s = 0; s < nsteps && !rc; ++s) {
        // example/example.cpp:221-266, plus the optional sort of :224-228
        if (prof) cudaEventRecord(c->prof_ev[5 * s + 0], c->stream);
        // CPIC_SORT_FUSED = "keep the store cell-ordered the cheapest way": the reordering push, except on grids of a few
        // hundred cells, where its slot claims would serialise on a handful of cursors and a counting sort every 8
        // steps is 3x cheaper (BASELINE configs[1]: 16.1 vs 5.7 ms/step, profiles/r03_probe_c2_*)
        const bool fallback = sort_interval == CPIC_SORT_FUSED && c->few_cells();
        const bool fused = sort_interval == CPIC_SORT_FUSED && !fallback;
        if (fused) rc = c->prepare_reorder();        // histogram (first step only) + scan of the cell counts
        else if (fallback) { if (c->fb_steps++ % 8 == 0) rc = c->sort(); }      // (counted across calls: one-step calls too)
        else if (sort_interval > 0 && s % sort_interval == 0) rc = c->sort();
        if (prof) cudaEventRecord(c->prof_ev[5 * s + 1], c->stream);
        if (!rc) rc = c->load_interpolator();
        if (!rc) rc = c->clear_accumulator();
        if (prof) cudaEventRecord(c->prof_ev[5 * s + 2], c->stream);
        c->want_hist = fallback ? (c->fb_steps % 8 == 0) : (sort_interval > 0 && (s + 1) % sort_interval == 0);   // the next step starts with a sort
        if (!rc) rc = fused ? c->push_reorder(*k) : c->push(*k);
        if (prof) cudaEventRecord(c->prof_ev[5 * s + 3], c->stream);
        if (!rc) rc = c->unload_accumulator(*k);
        if (!rc) rc = c->advance_b(hx, hy, hz);
        if (!rc) rc = c->advance_e(k->px, k->py, k->pz, k->dt_eps0);
        if (!rc) rc = c->advance_b(hx, hy, hz);
        if (!rc && en) rc = c->energies_async(en + 2 * s);
        if (prof) cudaEventRecord(c->prof_ev[5 * s + 4], c->stream);
    }
    c->rec(7);
    c->ev_valid[3] = true;
    if (!rc && en) {
        rc = c->cuda(cudaMemcpyAsync(energies, en, (size_t)nsteps * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream), "D2H energies");
        if (!rc) rc = c->cuda(cudaStreamSynchronize(c->stream), "step");
        for (int64_t s = 0; s < nsteps && !rc; ++s) {
            energies[2 * s] *= 0.5;
            energies[2 * s + 1] = (c->prm.solver == CPIC_SOLVER_EM) ? 0.5 * energies[2 * s + 1] : 0.0;
        }
    }
    if (en) { cudaStreamSynchronize(c->stream); cudaFree(en); }
    return rc;
}

int cpic_step_host(cpic_ctx* ctx, const cpic_consts* k, const void* const in[8], void* const out[8], int64_t n,
                   const void* const fields_in[9], void* const fields_out[9], double* energies) {
    CTX_HOSTNP(ctx);
    if (!k || !in || !fields_in) return c->fail(CPIC_E_INVALID, "step_host: null argument");
    for (int m = 0; m < 9; ++m) if (!fields_in[m]) return c->fail(CPIC_E_INVALID, "step_host: null field member %d", m);
    if (n > 0) for (int m = 0; m < 8; ++m) if (!in[m]) return c->fail(CPIC_E_INVALID, "step_host: null particle member %d", m);
    return c->step_host(*k, in, out, n, fields_in, fields_out, energies);
}

int cpic_push_stats_get(cpic_ctx* ctx, cpic_push_stats* out) {
    CTX_OR_FAIL(ctx);
    if (!out) return c->fail(CPIC_E_INVALID, "push_stats_get: null");
    unsigned long long h[8];
    int rc = c->cuda(cudaMemcpyAsync(h, c->stats_dev(), sizeof h, cudaMemcpyDeviceToHost, c->stream), "D2H stats");
    if (rc) return rc;
    if ((rc = c->cuda(cudaStreamSynchronize(c->stream), "push_stats"))) return rc;
    out->movers = (int64_t)h[0]; out->crossings = (int64_t)h[1];
    for (int i = 0; i < 6; ++i) out->wraps[i] = (int64_t)h[2 + i];
    return CPIC_OK;
}

int cpic_device_ptr(cpic_ctx* ctx, int which, void** ptr, int64_t* count, int64_t* stride) {
    CTX_OR_FAIL(ctx);
    if (!ptr) return c->fail(CPIC_E_INVALID, "device_ptr: null");
    return c->device_ptr(which, ptr, count, stride);
}

int cpic_set_stream(cpic_ctx* ctx, void* cuda_stream) {
    CTX_OR_FAIL(ctx);
    cudaStreamSynchronize(c->stream);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    c->own_stream = false;
    return CPIC_OK;
}

int cpic_set_num_particles(cpic_ctx* ctx, int64_t n) {
    CTX_HOSTNP(ctx);
    if (n < 0 || n > c->prm.max_particles) return c->fail(CPIC_E_CAPACITY, "set_num_particles: %lld out of range", (long long)n);
    c->np = n;
    c->hist_valid = false; c->cursor_valid = false; c->seg_valid = false; c->leavers_valid = false; c->ghost_clean = false;
    return CPIC_OK;
}

int cpic_advance_b_stencil(cpic_ctx* ctx, double px, double py, double pz) { CTX_OR_FAIL(ctx); return c->stencil_only(0, px, py, pz, 0.0); }
int cpic_advance_e_stencil(cpic_ctx* ctx, double px, double py, double pz, double dt_eps0) { CTX_OR_FAIL(ctx); return c->stencil_only(1, px, py, pz, dt_eps0); }

int cpic_set_axis_periodic(cpic_ctx* ctx, int32_t px, int32_t py, int32_t pz) {
    CTX_OR_FAIL(ctx);
    c->g.per = (px ? 1 : 0) | (py ? 2 : 0) | (pz ? 4 : 0);
    return CPIC_OK;
}

int cpic_extract_z_leavers(cpic_ctx* ctx, void* lo_buf, void* hi_buf, int64_t capacity, int64_t* n_lo, int64_t* n_hi,
                           int32_t rebase_lo, int32_t rebase_hi) {
    CTX_HOSTNP(ctx);
    if (!lo_buf || !hi_buf || !n_lo || !n_hi) return c->fail(CPIC_E_INVALID, "extract_z_leavers: null argument");
    long long a = 0, b = 0;
    int rc = c->extract_z(lo_buf, hi_buf, capacity, &a, &b, rebase_lo, rebase_hi);
    *n_lo = a; *n_hi = b;
    return rc;
}

int cpic_slab_extract_async(cpic_ctx* ctx, void* lo_buf, void* hi_buf, int64_t capacity, int64_t* counts_dev,
                            int32_t rebase_lo, int32_t rebase_hi) {
    CTX_OR_FAIL(ctx);
    if (!lo_buf || !hi_buf || !counts_dev) return c->fail(CPIC_E_INVALID, "slab_extract_async: null argument");
    return c->slab_extract_async(lo_buf, hi_buf, capacity, reinterpret_cast<long long*>(counts_dev), rebase_lo, rebase_hi);
}

int cpic_slab_append_async(cpic_ctx* ctx, const void* buf, int64_t capacity, const int64_t* count_dev) {
    CTX_OR_FAIL(ctx);
    if (!buf || !count_dev || capacity < 1) return c->fail(CPIC_E_INVALID, "slab_append_async: bad argument");
    return c->slab_append_async(buf, capacity, reinterpret_cast<const long long*>(count_dev));
}

int cpic_append_particles_device(cpic_ctx* ctx, const void* buf, int64_t capacity, int64_t n) {
    CTX_HOSTNP(ctx);
    if (!buf && n > 0) return c->fail(CPIC_E_INVALID, "append_particles_device: null buffer");
    return c->append_device(buf, capacity, n);
}

int cpic_set_modes(cpic_ctx* ctx, int32_t fp_mode, int32_t deposit_mode) {
    CTX_HOSTNP(ctx);
    if ((fp_mode != CPIC_FP_STRICT && fp_mode != CPIC_FP_CONTRACT) || deposit_mode < 0 || deposit_mode > CPIC_DEPOSIT_ORDERED)
        return c->fail(CPIC_E_INVALID, "set_modes: bad mode");
    c->prm.fp_mode = fp_mode;
    c->prm.deposit_mode = deposit_mode;
    return CPIC_OK;
}

int cpic_enable_push_stats(cpic_ctx* ctx, int32_t on) {
    CTX_OR_FAIL(ctx);
    c->want_stats = on != 0;
    return CPIC_OK;
}

int cpic_enable_step_profile(cpic_ctx* ctx, int32_t on) {
    CTX_OR_FAIL(ctx);
    c->prof_on = on != 0;
    return CPIC_OK;
}

int cpic_step_profile(cpic_ctx* ctx, double ms_out[4], int64_t* steps) {
    CTX_OR_FAIL(ctx);
    if (!ms_out) return c->fail(CPIC_E_INVALID, "step_profile: null");
    if (!c->prof_on || c->prof_steps == 0) return c->fail(CPIC_E_INVALID, "step_profile: no profiled cpic_step yet");
    int rc = c->cuda(cudaStreamSynchronize(c->stream), "step_profile");
    if (rc) return rc;
    double sum[4] = {0, 0, 0, 0};
    for (long long s = 0; s < c->prof_steps; ++s)
        for (int ph = 0; ph < 4; ++ph) {
            float f = 0.f;
            if ((rc = c->cuda(cudaEventElapsedTime(&f, c->prof_ev[5 * s + ph], c->prof_ev[5 * s + ph + 1]), "cudaEventElapsedTime"))) return rc;
            sum[ph] += f;
        }
    for (int ph = 0; ph < 4; ++ph) ms_out[ph] = sum[ph];
    if (steps) *steps = c->prof_steps;
    return CPIC_OK;
}

int cpic_last_ms(cpic_ctx* ctx, int what, double* ms) {
    CTX_OR_FAIL(ctx);
    if (!ms || what < 0 || what > 3) return c->fail(CPIC_E_INVALID, "last_ms: bad arguments");
    if (!c->ev_valid[what]) return c->fail(CPIC_E_INVALID, "last_ms: nothing recorded for %d", what);
    int rc = c->cuda(cudaEventSynchronize(c->ev[2 * what + 1]), "cudaEventSynchronize");
    if (rc) return rc;
    float f = 0.f;
    if ((rc = c->cuda(cudaEventElapsedTime(&f, c->ev[2 * what], c->ev[2 * what + 1]), "cudaEventElapsedTime"))) return rc;
    *ms = f;
    return CPIC_OK;
}

int cpic_launch_count(cpic_ctx* ctx, int64_t* launches) {
    CTX_OR_FAIL(ctx);
    if (!launches) return c->fail(CPIC_E_INVALID, "launch_count: null");
    *launches = c->launches;
    return CPIC_OK;
}

}  // extern "C"

#include "cpic_mgpu.cuh"
