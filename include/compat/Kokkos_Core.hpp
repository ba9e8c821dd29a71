// Host-only stand-in for the handful of Kokkos names CabanaPIC touches.
//
// CabanaPIC's arithmetic lives entirely in its own sources; Kokkos supplies
// storage (View, ScatterView) and loop dispatch (parallel_for/reduce over
// Range / MDRange policies).  Kokkos itself is not installable offline, so this
// header provides just those names with plain host semantics:
//   * serial by default (deterministic, used to pin parity), or
//   * OpenMP when compiled with -fopenmp -DCPIC_COMPAT_OPENMP (ScatterView then
//     keeps one duplicate per thread and `contribute` sums them, which is what
//     Kokkos' OpenMP backend does; used as the CPU timing baseline).
//
// Names covered: View<T*[A][B]>, Experimental::ScatterView,
// create_scatter_view, contribute, reset_except, RangePolicy, MDRangePolicy,
// Rank, parallel_for (4 call forms), parallel_reduce (Range + MDRange<3>),
// ScopeGuard, HostSpace, Default(Host)ExecutionSpace, KOKKOS_LAMBDA,
// KOKKOS_INLINE_FUNCTION.
#ifndef CPIC_COMPAT_KOKKOS_CORE_HPP
#define CPIC_COMPAT_KOKKOS_CORE_HPP

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <fstream>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <typeinfo>
#include <utility>
#include <type_traits>
#include <vector>

#if defined(CPIC_COMPAT_OPENMP) && defined(_OPENMP)
#include <omp.h>
#define CPIC_COMPAT_PARALLEL 1
#else
#define CPIC_COMPAT_PARALLEL 0
#endif

#define KOKKOS_LAMBDA [=]
#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_FORCEINLINE_FUNCTION inline

namespace Kokkos {

struct HostSpace {
    using memory_space = HostSpace;
    static const char* name() { return "Host"; }
};

struct HostExec {
    using memory_space = HostSpace;
    using execution_space = HostExec;
    static const char* name() { return CPIC_COMPAT_PARALLEL ? "compat-OpenMP" : "compat-Serial"; }
    static int concurrency() {
#if CPIC_COMPAT_PARALLEL
        return omp_get_max_threads();
#else
        return 1;
#endif
    }
};
using DefaultExecutionSpace = HostExec;
using DefaultHostExecutionSpace = HostExec;
using Serial = HostExec;
using OpenMP = HostExec;

struct ScopeGuard {
    ScopeGuard() {}
    ScopeGuard(int&, char**) {}
};
inline void initialize() {}
inline void initialize(int&, char**) {}
inline void finalize() {}
inline void fence() {}
inline void fence(const std::string&) {}

// ---------------------------------------------------------------- policies
template <class Exec = DefaultExecutionSpace>
struct RangePolicy {
    long lo, hi;
    RangePolicy(long b, long e) : lo(b), hi(e) {}
};

template <unsigned N>
struct Rank {
    static constexpr unsigned rank = N;
};

template <class R, class... Rest>
struct MDRangePolicy {
    static constexpr unsigned rank = R::rank;
    long lo[R::rank], hi[R::rank];
    template <class A, class B>
    MDRangePolicy(std::initializer_list<A> l, std::initializer_list<B> h) {
        unsigned k = 0;
        for (auto v : l) { if (k < rank) lo[k++] = (long)v; }
        k = 0;
        for (auto v : h) { if (k < rank) hi[k++] = (long)v; }
    }
};

namespace Impl {
template <class F>
inline void run_range(long lo, long hi, const F& f) {
#if CPIC_COMPAT_PARALLEL
#pragma omp parallel for schedule(static)
#endif
    for (long i = lo; i < hi; ++i) f((int)i);
}
template <class R, class F>
inline typename std::enable_if<R::rank == 2>::type run_md(const MDRangePolicy<R>& p, const F& f) {
#if CPIC_COMPAT_PARALLEL
#pragma omp parallel for schedule(static)
#endif
    for (long a = p.lo[0]; a < p.hi[0]; ++a)
        for (long b = p.lo[1]; b < p.hi[1]; ++b) f((int)a, (int)b);
}
template <class R, class F>
inline typename std::enable_if<R::rank == 3>::type run_md(const MDRangePolicy<R>& p, const F& f) {
#if CPIC_COMPAT_PARALLEL
#pragma omp parallel for collapse(2) schedule(static)
#endif
    for (long a = p.lo[0]; a < p.hi[0]; ++a)
        for (long b = p.lo[1]; b < p.hi[1]; ++b)
            for (long c = p.lo[2]; c < p.hi[2]; ++c) f((int)a, (int)b, (int)c);
}
}  // namespace Impl

// parallel_for( label, policy|N, f )   and   parallel_for( policy|N, f [, label] )
template <class E, class F>
inline void parallel_for(const std::string&, const RangePolicy<E>& p, const F& f) { Impl::run_range(p.lo, p.hi, f); }
template <class E, class F>
inline void parallel_for(const RangePolicy<E>& p, const F& f, const std::string& = "") { Impl::run_range(p.lo, p.hi, f); }
template <class R, class F>
inline void parallel_for(const std::string&, const MDRangePolicy<R>& p, const F& f) { Impl::run_md(p, f); }
template <class R, class F>
inline void parallel_for(const MDRangePolicy<R>& p, const F& f, const std::string& = "") { Impl::run_md(p, f); }
template <class I, class F, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
inline void parallel_for(const std::string&, I n, const F& f) { Impl::run_range(0, (long)n, f); }
template <class I, class F, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
inline void parallel_for(I n, const F& f, const std::string& = "") { Impl::run_range(0, (long)n, f); }

// parallel_reduce( label, policy, f, result ): sum reduction in T, index order.
// (In OpenMP mode each thread keeps a partial that is combined in thread order.)
template <class E, class F, class T>
inline void parallel_reduce(const std::string&, const RangePolicy<E>& p, const F& f, T& result) {
    T total = T(0);
#if CPIC_COMPAT_PARALLEL
#pragma omp parallel
    {
        T part = T(0);
#pragma omp for schedule(static) nowait
        for (long i = p.lo; i < p.hi; ++i) f((int)i, part);
#pragma omp critical
        total += part;
    }
#else
    for (long i = p.lo; i < p.hi; ++i) f((int)i, total);
#endif
    result = total;
}
template <class R, class F, class T>
inline typename std::enable_if<R::rank == 3>::type parallel_reduce(const std::string&, const MDRangePolicy<R>& p,
                                                                   const F& f, T& result) {
    T total = T(0);
#if CPIC_COMPAT_PARALLEL
#pragma omp parallel
    {
        T part = T(0);
#pragma omp for collapse(2) schedule(static) nowait
        for (long a = p.lo[0]; a < p.hi[0]; ++a)
            for (long b = p.lo[1]; b < p.hi[1]; ++b)
                for (long c = p.lo[2]; c < p.hi[2]; ++c) f((int)a, (int)b, (int)c, part);
#pragma omp critical
        total += part;
    }
#else
    for (long a = p.lo[0]; a < p.hi[0]; ++a)
        for (long b = p.lo[1]; b < p.hi[1]; ++b)
            for (long c = p.lo[2]; c < p.hi[2]; ++c) f((int)a, (int)b, (int)c, total);
#endif
    result = total;
}

// -------------------------------------------------------------------- View
template <class DataType, class... Props>
class View;

// Only the shape CabanaPIC uses: one runtime extent, two compile-time extents,
// row-major (LayoutRight, Kokkos' host default), zero-initialised, ref-counted.
template <class T, std::size_t A, std::size_t B, class... Props>
class View<T* [A][B], Props...> {
   public:
    using value_type = T;
    View() : n_(0) {}
    View(const std::string& label, std::size_t n) : label_(label), n_(n), buf_(new std::vector<T>(n * A * B, T(0))) {}
    T& operator()(std::size_t i, std::size_t j, std::size_t k) const { return (*buf_)[(i * A + j) * B + k]; }
    std::size_t extent(int d) const { return d == 0 ? n_ : (d == 1 ? A : B); }
    std::size_t size() const { return n_ * A * B; }
    T* data() const { return buf_ ? buf_->data() : nullptr; }
    const std::string& label() const { return label_; }

   private:
    std::string label_;
    std::size_t n_;
    std::shared_ptr<std::vector<T>> buf_;
};

namespace Experimental {

struct ScatterSum {};

template <class DataType, class... Props>
class ScatterView;

template <class T, std::size_t A, std::size_t B, class... Props>
class ScatterView<T* [A][B], Props...> {
   public:
    using view_type = View<T* [A][B]>;

    // Accessor handed to each loop body by access(): `acc(i,j,k) += v`.
    struct Access {
        T* base;
        T& operator()(std::size_t i, std::size_t j, std::size_t k) const { return base[(i * A + j) * B + k]; }
    };

    ScatterView() {}
    explicit ScatterView(const view_type& v) : target_(v) {
#if CPIC_COMPAT_PARALLEL
        ndup_ = omp_get_max_threads();
        dup_.reset(new std::vector<T>(std::size_t(ndup_) * v.size(), T(0)));
#endif
    }

    Access access() const {
#if CPIC_COMPAT_PARALLEL
        return Access{dup_->data() + std::size_t(omp_get_thread_num()) * target_.size()};
#else
        return Access{target_.data()};
#endif
    }

    // Sum the per-thread duplicates into `dest` (no-op when not duplicated).
    void contribute_into(const view_type& dest) const {
#if CPIC_COMPAT_PARALLEL
        const std::size_t n = dest.size();
        T* out = dest.data();
        const T* d = dup_->data();
        const int nd = ndup_;
#pragma omp parallel for schedule(static)
        for (long e = 0; e < (long)n; ++e) {
            T s = out[e];
            for (int t = 0; t < nd; ++t) s += d[std::size_t(t) * n + e];
            out[e] = s;
        }
#else
        (void)dest;
#endif
    }

    // Zero the duplicates unless they alias `keep` (the non-duplicated case).
    void reset_except(const view_type& keep) {
#if CPIC_COMPAT_PARALLEL
        (void)keep;
        const std::size_t n = dup_->size();
        T* d = dup_->data();
#pragma omp parallel for schedule(static)
        for (long e = 0; e < (long)n; ++e) d[e] = T(0);
#else
        (void)keep;
#endif
    }

   private:
    view_type target_;
#if CPIC_COMPAT_PARALLEL
    int ndup_ = 1;
    std::shared_ptr<std::vector<T>> dup_;
#endif
};

template <class T, std::size_t A, std::size_t B, class... P>
inline ScatterView<T* [A][B]> create_scatter_view(const View<T* [A][B], P...>& v) {
    return ScatterView<T* [A][B]>(View<T* [A][B]>(v));
}

template <class T, std::size_t A, std::size_t B>
inline void contribute(const View<T* [A][B]>& dest, const ScatterView<T* [A][B]>& src) {
    src.contribute_into(dest);
}

}  // namespace Experimental
}  // namespace Kokkos

#endif  // CPIC_COMPAT_KOKKOS_CORE_HPP
