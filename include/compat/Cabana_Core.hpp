// Host-only stand-in for the Cabana names CabanaPIC touches:
// MemberTypes, AoSoA (size, vector_length, host_mirror_type), slice<M>() with
// operator()(i) and access(s,i), SimdPolicy, simd_parallel_for, deep_copy.
//
// Storage is one contiguous array per member (struct-of-arrays); decks and the
// reference kernels only ever see slices, so the physical layout is free.
// `vector_length` only fixes how a flat index splits into (s,i) = (k/VL, k%VL).
#ifndef CPIC_COMPAT_CABANA_CORE_HPP
#define CPIC_COMPAT_CABANA_CORE_HPP

#include <Kokkos_Core.hpp>

#include <tuple>
#include <utility>

namespace Cabana {

template <class... Ts>
struct MemberTypes {
    static constexpr std::size_t size = sizeof...(Ts);
};

// A view of one member of an AoSoA: pointer semantics, so a slice captured by
// value in a lambda still writes through to the container.
template <class T, int VL>
class Slice {
   public:
    using value_type = T;
    static constexpr int vector_length = VL;
    Slice() : p_(nullptr), n_(0) {}
    Slice(T* p, std::size_t n) : p_(p), n_(n) {}
    T& operator()(std::size_t i) const { return p_[i]; }
    T& access(std::size_t s, std::size_t i) const { return p_[s * VL + i]; }
    std::size_t size() const { return n_; }
    std::size_t numSoA() const { return (n_ + VL - 1) / VL; }
    T* data() const { return p_; }

   private:
    T* p_;
    std::size_t n_;
};

template <class Members, class MemorySpace, int VectorLength = 16>
class AoSoA;

template <class... Ts, class MemorySpace, int VectorLength>
class AoSoA<MemberTypes<Ts...>, MemorySpace, VectorLength> {
   public:
    using member_types = MemberTypes<Ts...>;
    using memory_space = MemorySpace;
    using host_mirror_type = AoSoA<MemberTypes<Ts...>, Kokkos::HostSpace, VectorLength>;
    static constexpr int vector_length = VectorLength;
    static constexpr std::size_t number_of_members = sizeof...(Ts);
    template <std::size_t M>
    using member_t = typename std::tuple_element<M, std::tuple<Ts...>>::type;

    AoSoA() : n_(0) {}
    AoSoA(const std::string& label, std::size_t n) : label_(label), n_(n) { allocate(n); }
    explicit AoSoA(std::size_t n) : n_(n) { allocate(n); }

    std::size_t size() const { return n_; }
    std::size_t numSoA() const { return (n_ + VectorLength - 1) / VectorLength; }
    const std::string& label() const { return label_; }

    void resize(std::size_t n) {
        // storage is padded to whole tiles so access(s,i) of a tail tile stays in bounds
        const std::size_t padded = ((n + VectorLength - 1) / VectorLength) * VectorLength;
        resize_impl(padded, std::index_sequence_for<Ts...>{});
        n_ = n;
    }

    template <std::size_t M>
    member_t<M>* member_data() const {
        return std::get<M>(cols_)->data();
    }

   private:
    void allocate(std::size_t n) {
        const std::size_t padded = ((n + VectorLength - 1) / VectorLength) * VectorLength;
        cols_ = std::make_tuple(std::make_shared<std::vector<Ts>>(padded, Ts(0))...);
    }
    template <std::size_t... I>
    void resize_impl(std::size_t padded, std::index_sequence<I...>) {
        int dummy[] = {(std::get<I>(cols_)->resize(padded), 0)...};
        (void)dummy;
    }

    std::string label_;
    std::size_t n_;
    std::tuple<std::shared_ptr<std::vector<Ts>>...> cols_;
};

template <std::size_t M, class AoSoA_t>
inline Slice<typename AoSoA_t::template member_t<M>, AoSoA_t::vector_length> slice(const AoSoA_t& a,
                                                                                   const std::string& = "") {
    return Slice<typename AoSoA_t::template member_t<M>, AoSoA_t::vector_length>(a.template member_data<M>(),
                                                                                a.size());
}

namespace Impl {
template <class Dst, class Src, std::size_t... I>
inline void copy_members(Dst& dst, const Src& src, std::index_sequence<I...>) {
    const std::size_t n = src.size();
    int dummy[] = {(std::memcpy(dst.template member_data<I>(), src.template member_data<I>(),
                                n * sizeof(typename Src::template member_t<I>)),
                    0)...};
    (void)dummy;
}
}  // namespace Impl

template <class Dst, class Src>
inline void deep_copy(Dst& dst, const Src& src) {
    Impl::copy_members(dst, src, std::make_index_sequence<Src::number_of_members>{});
}

template <int VL, class Exec = Kokkos::DefaultExecutionSpace>
struct SimdPolicy {
    std::size_t lo, hi;
    SimdPolicy(std::size_t b, std::size_t e) : lo(b), hi(e) {}
};

// f(s, i) for every flat index in [lo, hi), tile-major / lane-minor.  Serial
// builds walk strictly in index order (dioctron_3d's rand() stream relies on it).
template <int VL, class Exec, class F>
inline void simd_parallel_for(const SimdPolicy<VL, Exec>& p, const F& f, const std::string& = "") {
    const long s_lo = (long)(p.lo / VL), s_hi = (long)((p.hi + VL - 1) / VL);
#if CPIC_COMPAT_PARALLEL
#pragma omp parallel for schedule(static)
#endif
    for (long s = s_lo; s < s_hi; ++s) {
        const std::size_t k0 = std::size_t(s) * VL;
        const int i_lo = (k0 < p.lo) ? int(p.lo - k0) : 0;
        const int i_hi = (k0 + VL > p.hi) ? int(p.hi - k0) : VL;
        for (int i = i_lo; i < i_hi; ++i) f((int)s, i);
    }
}

}  // namespace Cabana

#endif  // CPIC_COMPAT_CABANA_CORE_HPP
