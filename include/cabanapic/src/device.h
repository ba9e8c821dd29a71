// cabanapic_b200 C++ host facade -- the runtime behind the device-backed arrays.
//
// One cpic_ctx (include/cabanapic_b200.h) per process, created on first use from the global `deck`
// (the reference's hot-path functions read `deck` the same way, src/fields.h:20,112).  Every array
// kind has at most one live instance in the reference's driver (example/example.cpp:121-144), which
// is what a context holds; the facade keeps that model.
//
// Residency protocol: need_on_device(x) uploads x's host mirror if the device copy is stale;
// device_wrote(x) marks the mirror stale; x.host_access() (called by Cabana::slice<>() on a
// device-backed array) downloads if the mirror is stale and then assumes the host may write.
// In the steady-state time loop nothing touches the mirrors, so nothing crosses PCIe.
#ifndef CABANAPIC_B200_DEVICE_H
#define CABANAPIC_B200_DEVICE_H

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cabanapic_b200.h"
#include "cabanapic_b200_mgpu.h"

namespace cabanapic {

class Runtime {
   public:
    static Runtime& get() {
        static Runtime r;
        return r;
    }
    cpic_ctx* ctx() {
        if (!ctx_) create();
        return ctx_;
    }
    bool created() const { return ctx_ != nullptr; }
    void check(int rc, const char* what) {
        if (rc == CPIC_OK) return;
        std::fprintf(stderr, "cabanapic_b200: %s failed (%d): %s\n", what, rc, cpic_last_error(ctx_));
        std::exit(1);      // the reference has no error channel either (exit(1), src/fields.h:24,116)
    }
    void destroy() {
        if (mgpu_) cpic_mgpu_destroy(mgpu_);
        else if (ctx_) cpic_destroy(ctx_);
        ctx_ = nullptr;
        mgpu_ = nullptr;
    }
    int sort_interval = 0;       // CPIC_SORT_INTERVAL: Cabana::sortByKey cadence (example/example.cpp:224-228)
    long pushes = 0;
    // Multi-GPU, replicated mode (one process per GPU: CPIC_WORLD, CPIC_RANK, CPIC_MGPU_ID_FILE in the environment): every
    // process runs the deck's initialisers for the whole box, keeps particles [lo, hi) of the list on its GPU, and
    // Kokkos::Experimental::contribute -- the call where the reference marks the spot (example/example.cpp:248-254) --
    // becomes an ncclAllReduce of the accumulator; the field solve is replicated.
    int world = 1, rank = 0;
    cpic_mgpu* mgpu() { ctx(); return mgpu_; }
    // the slice of the particle list this process owns
    void share(size_t n, size_t& lo, size_t& hi) const { lo = n * (size_t)rank / (size_t)world; hi = n * (size_t)(rank + 1) / (size_t)world; }
    // Keep the store cell-ordered with the reordering push (cpic_push_reorder) when it pays: large 3-D problems.  Small
    // decks keep the caller's particle order (it is observable through dump_particles: partloc).  CPIC_KEEP_ORDER=1 /
    // CPIC_REORDER=1 force either.
    bool reorder(size_t np) const {
        if (const char* e = std::getenv("CPIC_KEEP_ORDER")) if (std::atoi(e)) return false;
        if (const char* e = std::getenv("CPIC_REORDER")) if (std::atoi(e)) return true;
        return sort_interval == 0 && np >= (size_t)1 << 20 && (size_t)deck.num_cells > 1024;
    }
    ~Runtime() { destroy(); }

    // ---- residency ------------------------------------------------------------------------
    void need_on_device(const particle_list_t& p) {
        Residency& r = p.residency();
        if (r.device_valid) return;
        const auto& h = p.host();
        size_t lo, hi;
        share(p.size(), lo, hi);
        check(cpic_upload_particles(ctx(), h.template member_data<0>() + lo, h.template member_data<1>() + lo, h.template member_data<2>() + lo,
                                    h.template member_data<3>() + lo, h.template member_data<4>() + lo, h.template member_data<5>() + lo,
                                    h.template member_data<6>() + lo, h.template member_data<7>() + lo, (int64_t)(hi - lo)),
              "cpic_upload_particles");
        r.device_valid = true;
    }
    void need_on_device(const field_array_t& f) {
        Residency& r = f.residency();
        if (r.device_valid) return;
        const void* m[9];
        field_ptrs(f, m);
        check(cpic_upload_fields(ctx(), m), "cpic_upload_fields");
        r.device_valid = true;
    }
    void need_on_device(const interpolator_array_t& a) {
        Residency& r = a.residency();
        if (r.device_valid) return;
        std::vector<real_t> t(a.size() * 18);
        gather_interp(a, t.data());
        check(cpic_upload_interpolators(ctx(), t.data()), "cpic_upload_interpolators");
        r.device_valid = true;
    }
    void need_on_device(const accumulator_array_t& a) {
        Residency& r = a.residency();
        if (r.device_valid) return;
        check(cpic_upload_accumulators(ctx(), a.data()), "cpic_upload_accumulators");
        r.device_valid = true;
    }
    template <class A>
    void device_wrote(const A& a) {
        a.residency().device_valid = true;
        a.residency().host_valid = false;
    }

    void refresh_host(const particle_list_t& p) {
        const auto& h = p.host();
        int64_t n = 0;
        size_t lo, hi;
        share(p.size(), lo, hi);      // (multi-GPU: this process's slice of the mirror; the rest keeps its last host value)
        check(cpic_download_particles(ctx(), h.template member_data<0>() + lo, h.template member_data<1>() + lo, h.template member_data<2>() + lo,
                                      h.template member_data<3>() + lo, h.template member_data<4>() + lo, h.template member_data<5>() + lo,
                                      h.template member_data<6>() + lo, h.template member_data<7>() + lo, (int64_t)(hi - lo), &n),
              "cpic_download_particles");
    }
    void refresh_host(const field_array_t& f) {
        const void* m[9];
        void* w[9];
        field_ptrs(f, m);
        for (int k = 0; k < 9; ++k) w[k] = const_cast<void*>(m[k]);
        check(cpic_download_fields(ctx(), w), "cpic_download_fields");
    }
    void refresh_host(const interpolator_array_t& a) {
        std::vector<real_t> t(a.size() * 18);
        check(cpic_download_interpolators(ctx(), t.data()), "cpic_download_interpolators");
        scatter_interp(a, t.data());
    }
    void refresh_host(const accumulator_array_t& a) {
        check(cpic_download_accumulators(ctx(), a.data()), "cpic_download_accumulators");
    }

   private:
    Runtime() {
        if (const char* e = std::getenv("CPIC_SORT_INTERVAL")) sort_interval = std::atoi(e);
        if (const char* e = std::getenv("CPIC_WORLD")) world = std::atoi(e) > 1 ? std::atoi(e) : 1;
        if (const char* e = std::getenv("CPIC_RANK")) rank = std::atoi(e);
        if (rank < 0 || rank >= world) { std::fprintf(stderr, "cabanapic_b200: CPIC_RANK out of range\n"); std::exit(1); }
    }
    void create() {
        cpic_params p{};
        p.nx = (int32_t)deck.nx; p.ny = (int32_t)deck.ny; p.nz = (int32_t)deck.nz; p.ng = (int32_t)deck.num_ghosts;
        p.real_bytes = (int32_t)sizeof(real_t);
#ifdef ES_FIELD_SOLVER
        p.solver = CPIC_SOLVER_ES_1D;
#else
        p.solver = CPIC_SOLVER_EM;
#endif
        p.boundary = deck.BOUNDARY_TYPE == Boundary::Periodic ? CPIC_BOUNDARY_PERIODIC : CPIC_BOUNDARY_REFLECT;
        if (const char* e = std::getenv("CPIC_DEVICE")) p.device = std::atoi(e);
        else if (world > 1) p.device = rank;
        p.fp_mode = CPIC_FP_STRICT;
        if (const char* e = std::getenv("CPIC_FP_CONTRACT")) p.fp_mode = std::atoi(e) ? CPIC_FP_CONTRACT : CPIC_FP_STRICT;
        p.deposit_mode = CPIC_DEPOSIT_AUTO;
        p.max_particles = deck.num_particles > 0 ? deck.num_particles : 0;
        p.enable_sort = 1;
        if (world > 1) {
            size_t lo, hi;
            share((size_t)p.max_particles, lo, hi);
            p.max_particles = (int64_t)(hi - lo);
            const char* idf = std::getenv("CPIC_MGPU_ID_FILE");
            unsigned char id[CPIC_MGPU_ID_BYTES];
            int rc = cpic_mgpu_bootstrap_file(idf ? idf : "/tmp/cabanapic_b200.nccl_id", rank, world, 120.0, id);
            if (rc == CPIC_OK) rc = cpic_mgpu_create(&p, rank, world, id, CPIC_MGPU_REPLICATED, 0, &mgpu_);
            if (rc != CPIC_OK) {
                std::fprintf(stderr, "cabanapic_b200: multi-GPU start-up failed (%d): %s\n", rc, cpic_mgpu_last_error(nullptr));
                std::exit(1);
            }
            ctx_ = cpic_mgpu_context(mgpu_);
            return;
        }
        const int rc = cpic_create(&p, &ctx_);
        if (rc != CPIC_OK) {
            std::fprintf(stderr, "cabanapic_b200: cpic_create failed (%d): %s\n", rc, cpic_last_error(nullptr));
            std::exit(1);
        }
    }
    template <std::size_t... I>
    static void field_ptrs_impl(const field_array_t& f, const void** m, std::index_sequence<I...>) {
        const void* t[] = {f.host().template member_data<I>()...};
        for (int k = 0; k < 9; ++k) m[k] = t[k];
    }
    static void field_ptrs(const field_array_t& f, const void** m) { field_ptrs_impl(f, m, std::make_index_sequence<9>{}); }
    template <std::size_t... I>
    static void gather_interp_impl(const interpolator_array_t& a, real_t* out, std::index_sequence<I...>) {
        const real_t* col[] = {a.host().template member_data<I>()...};
        for (std::size_t c = 0; c < a.size(); ++c)
            for (int k = 0; k < 18; ++k) out[c * 18 + k] = col[k][c];
    }
    static void gather_interp(const interpolator_array_t& a, real_t* out) { gather_interp_impl(a, out, std::make_index_sequence<18>{}); }
    template <std::size_t... I>
    static void scatter_interp_impl(const interpolator_array_t& a, const real_t* in, std::index_sequence<I...>) {
        real_t* col[] = {a.host().template member_data<I>()...};
        for (std::size_t c = 0; c < a.size(); ++c)
            for (int k = 0; k < 18; ++k) col[k][c] = in[c * 18 + k];
    }
    static void scatter_interp(const interpolator_array_t& a, const real_t* in) { scatter_interp_impl(a, in, std::make_index_sequence<18>{}); }

    cpic_ctx* ctx_ = nullptr;
    cpic_mgpu* mgpu_ = nullptr;
};

template <ArrayKind K, class Members, int VL>
inline void DeviceBacked<K, Members, VL>::host_access() const {
    Residency& r = *res_;
    if (!r.host_valid) {
        Runtime::get().refresh_host(*this);
        r.host_valid = true;
    }
    r.device_valid = false;      // the caller holds a writable view of the mirror
}
inline void Accumulators::host_access() const {
    Residency& r = *res_;
    if (!r.host_valid) {
        Runtime::get().refresh_host(*this);
        r.host_valid = true;
    }
    r.device_valid = false;
}

}  // namespace cabanapic

// Host views of device-backed arrays: these overloads are chosen over the generic
// Cabana::slice / deep_copy because the parameter type is more specialised.
namespace Cabana {
template <std::size_t M, cabanapic::ArrayKind K, class Members, int VL>
inline Slice<typename cabanapic::DeviceBacked<K, Members, VL>::base::template member_t<M>, VL> slice(
    const cabanapic::DeviceBacked<K, Members, VL>& a, const std::string& label = "") {
    using base = typename cabanapic::DeviceBacked<K, Members, VL>::base;
    a.host_access();
    return slice<M, base>(a.host(), label);
}
template <class Dst, cabanapic::ArrayKind K, class Members, int VL>
inline void deep_copy(Dst& dst, const cabanapic::DeviceBacked<K, Members, VL>& src) {
    using base = typename cabanapic::DeviceBacked<K, Members, VL>::base;
    src.host_access();
    deep_copy<Dst, base>(dst, src.host());
}
}  // namespace Cabana

namespace Kokkos {
namespace Experimental {
inline cabanapic::ScatterHandle create_scatter_view(const cabanapic::Accumulators& a) { return cabanapic::ScatterHandle{a}; }
// contribute(): where the reference sums thread-private duplicates (example/example.cpp:248).  On one GPU the deposit
// is already complete; with CPIC_WORLD > 1 processes (replicated mode) it is the ncclAllReduce of the accumulator.
inline void contribute(const cabanapic::Accumulators&, const cabanapic::ScatterHandle&) {
    cabanapic::Runtime& rt = cabanapic::Runtime::get();
    if (rt.mgpu()) {
        if (cpic_mgpu_reduce_accumulator(rt.mgpu()) != CPIC_OK) {
            std::fprintf(stderr, "cabanapic_b200: cpic_mgpu_reduce_accumulator failed: %s\n", cpic_mgpu_last_error(rt.mgpu()));
            std::exit(1);
        }
        return;
    }
    rt.check(cpic_contribute(rt.ctx()), "cpic_contribute");
}
}  // namespace Experimental
}  // namespace Kokkos

#endif  // CABANAPIC_B200_DEVICE_H
