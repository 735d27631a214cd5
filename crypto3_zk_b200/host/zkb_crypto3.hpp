// Host-side C++ templates that re-create, on top of the C ABI (include/zkb200.h), the entities
// crypto3-zk's headers name at its hot-path call sites:
//
//   nil::crypto3::math::evaluation_domain<F>, make_evaluation_domain<F>, calculate_domain_set<F>,
//   multiply_by_coset, polynomial_dfs<V>                           (un-vendored crypto3-math)
//   nil::crypto3::algebra::multiexp<Method>, multiexp_with_mixed_addition<Method>,
//   policies::multiexp_method_BDLO12 / multiexp_method_bos_coster  (un-vendored crypto3-algebra)
//   nil::crypto3::zk::commitments::detail::fold_polynomial (dfs)   (fold_polynomial.hpp:68-93)
//   nil::crypto3::zk::algorithms::precommit (container<polynomial_dfs>) -> tree with root()
//                                                                  (basic_fri.hpp:445-496, lpc.hpp:101-106)
//
// Same names, argument meaning and error behaviour (std::invalid_argument for size errors,
// std::runtime_error for device errors).  The work is done by the CUDA library; there is no CPU
// fallback - the first call that needs the device throws when no GPU is present.
//
// Field elements are kept in Montgomery form on 64-bit limbs with R = 2^256 / 2^384, which is what
// upstream's modular_adaptor stores.  The transforms and the FRI fold are linear
// in the data, so vectors go to the device as they are (Montgomery in, Montgomery out, no conversion
// pass); MSM scalars/points and Merkle-leaf inputs are converted to the ABI's canonical form on the way in.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/zkb200.h"
#include "../csrc/zkb_field.cuh"   // Fp2<B> (generic over the base type; compiled for the host here)
#include "../csrc/zkb_hostfield.h"

namespace nil {
namespace crypto3 {

namespace zkb_detail {
inline void check(int status, zkb_ctx *ctx, const char *what) {
    if (status == ZKB_OK) return;
    std::string msg = std::string(what) + ": " + zkb_status_string(status);
    if (ctx && *zkb_ctx_last_error(ctx)) msg += std::string(" (") + zkb_ctx_last_error(ctx) + ")";
    if (status == ZKB_ERR_INVALID_ARGUMENT || status == ZKB_ERR_DOMAIN_TOO_LARGE) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
// one lazily created context per host thread (the reference's domains are not thread-safe either)
inline zkb_ctx *context() {
    struct holder {
        zkb_ctx *c = nullptr;
        ~holder() { if (c) zkb_ctx_destroy(c); }
    };
    static thread_local holder h;
    if (!h.c) check(zkb_ctx_create(0, &h.c), nullptr, "zkb_ctx_create");
    return h.c;
}
inline int log2_exact(std::size_t m) {
    if (m == 0 || (m & (m - 1))) throw std::invalid_argument("domain size must be a power of two (basic_radix2_domain)");
    int l = 0;
    while ((std::size_t(1) << l) < m) l++;
    return l;
}
}  // namespace zkb_detail

// ========================================================================================== algebra
namespace algebra {
namespace fields {

template <class Params, int FieldId>
struct zkb_field {
    typedef ::zkb::HostFp<Params> backend;
    static constexpr int field_id = FieldId;
    static constexpr int limbs32 = Params::N;
    static constexpr std::size_t modulus_bits = Params::BITS;
    static constexpr std::size_t number_bits = Params::BITS;

    struct value_type {
        typedef zkb_field field_type;
        backend data;  // Montgomery form
        value_type() : data(backend::zero()) {}
        value_type(std::uint64_t v) {
            backend t = backend::zero();
            t.v[0] = v;
            data = t.to_mont();
        }
        static value_type zero() { return value_type(); }
        static value_type one() { value_type r; r.data = backend::one(); return r; }
        static value_type from_canonical_limbs(const std::uint32_t *l) { value_type r; r.data = backend::from_limbs32(l).to_mont(); return r; }
        void to_canonical_limbs(std::uint32_t *l) const { data.from_mont().to_limbs32(l); }
        bool is_zero() const { return data.is_zero(); }
        bool is_one() const { return data == backend::one(); }
        value_type operator+(const value_type &o) const { value_type r; r.data = data + o.data; return r; }
        value_type operator-(const value_type &o) const { value_type r; r.data = data - o.data; return r; }
        value_type operator*(const value_type &o) const { value_type r; r.data = data * o.data; return r; }
        value_type operator-() const { value_type r; r.data = data.neg(); return r; }
        value_type &operator+=(const value_type &o) { data = data + o.data; return *this; }
        value_type &operator-=(const value_type &o) { data = data - o.data; return *this; }
        value_type &operator*=(const value_type &o) { data = data * o.data; return *this; }
        bool operator==(const value_type &o) const { return data == o.data; }
        bool operator!=(const value_type &o) const { return !(data == o.data); }
        value_type inversed() const { value_type r; r.data = data.inverse(); return r; }
        value_type squared() const { return *this * *this; }
        value_type pow(std::uint64_t e) const {
            value_type r = one(), b = *this;
            for (; e; e >>= 1) {
                if (e & 1) r *= b;
                b *= b;
            }
            return r;
        }
    };
};

template <std::size_t> struct bls12_fr;
template <> struct bls12_fr<381> : zkb_field<::zkb::params::Bls12381Fr, ZKB_FIELD_BLS12_381_FR> {};
template <std::size_t> struct bls12_fq;
template <> struct bls12_fq<381> : zkb_field<::zkb::params::Bls12381Fq, ZKB_FIELD_BLS12_381_FQ> {};
template <std::size_t> struct alt_bn128_fr;
template <> struct alt_bn128_fr<254> : zkb_field<::zkb::params::Bn254Fr, ZKB_FIELD_BN254_FR> {};
template <std::size_t> struct alt_bn128_fq;
template <> struct alt_bn128_fq<254> : zkb_field<::zkb::params::Bn254Fq, ZKB_FIELD_BN254_FQ> {};
struct pallas_base_field : zkb_field<::zkb::params::PallasFp, ZKB_FIELD_PALLAS_FP> {};
struct pallas_scalar_field : zkb_field<::zkb::params::PallasFq, ZKB_FIELD_PALLAS_FQ> {};

// Fq2 = Fq[u]/(u^2 + 1): coordinate field of the G2 groups (`fields::fp2<...>` upstream); canonical limbs
// are c0 || c1.  Only what the group value type below needs.
template <class BaseField>
struct zkb_field2 {
    typedef ::zkb::Fp2<typename BaseField::backend> backend;
    typedef BaseField underlying_field_type;
    static constexpr int limbs32 = 2 * BaseField::limbs32;
    struct value_type {
        typedef zkb_field2 field_type;
        backend data;  // Montgomery form
        value_type() : data(backend::zero()) {}
        static value_type zero() { return value_type(); }
        static value_type one() { value_type r; r.data = backend::one(); return r; }
        static value_type from_canonical_limbs(const std::uint32_t *l) {
            value_type r;
            r.data.c0 = BaseField::backend::from_limbs32(l).to_mont();
            r.data.c1 = BaseField::backend::from_limbs32(l + BaseField::limbs32).to_mont();
            return r;
        }
        void to_canonical_limbs(std::uint32_t *l) const {
            data.c0.from_mont().to_limbs32(l);
            data.c1.from_mont().to_limbs32(l + BaseField::limbs32);
        }
        bool is_zero() const { return data.is_zero(); }
        value_type operator+(const value_type &o) const { value_type r; r.data = data + o.data; return r; }
        value_type operator-(const value_type &o) const { value_type r; r.data = data - o.data; return r; }
        value_type operator*(const value_type &o) const { value_type r; r.data = data * o.data; return r; }
        value_type operator-() const { value_type r; r.data = data.neg(); return r; }
        bool operator==(const value_type &o) const { return data == o.data; }
        bool operator!=(const value_type &o) const { return !(data == o.data); }
        value_type inversed() const { value_type r; r.data = data.inverse(); return r; }
        value_type squared() const { value_type r; r.data = data.sqr(); return r; }
    };
};

// arithmetic_params<F>::multiplicative_generator as used for the coset shift (r1cs_to_qap.hpp:266-269)
template <class FieldType>
struct arithmetic_params {
    static typename FieldType::value_type multiplicative_generator_value() {
        std::uint32_t l[12] = {0};
        zkb_field_generator(FieldType::field_id, l);
        return FieldType::value_type::from_canonical_limbs(l);
    }
    static std::size_t two_adicity() { return (std::size_t)zkb_field_two_adicity(FieldType::field_id); }
};
}  // namespace fields

namespace curves {
template <class BaseField, class ScalarField, int CurveId>
struct zkb_curve_g1 {
    typedef BaseField base_field_type;
    typedef ScalarField scalar_field_type;
    static constexpr int curve_id = CurveId;
    // Jacobian (X, Y, Z) like the reference's default G1 element; infinity <=> Z == 0
    struct value_type {
        typedef zkb_curve_g1 group_type;
        typename BaseField::value_type X, Y, Z;
        static value_type zero() {
            value_type r;
            r.X = BaseField::value_type::one(); r.Y = BaseField::value_type::one();
            return r;
        }
        static value_type from_affine(const typename BaseField::value_type &x, const typename BaseField::value_type &y) {
            value_type r;
            r.X = x; r.Y = y; r.Z = BaseField::value_type::one();
            return r;
        }
        // the group generator, `value_type::one()` in the reference (kzg.hpp:102,112)
        static value_type one() {
            std::uint32_t l[48] = {0};
            zkb_curve_generator(CurveId, l);
            return from_affine(BaseField::value_type::from_canonical_limbs(l),
                               BaseField::value_type::from_canonical_limbs(l + BaseField::limbs32));
        }
        bool is_zero() const { return Z.is_zero(); }
        value_type to_affine() const {
            if (is_zero()) return zero();
            auto zi = Z.inversed(), zi2 = zi * zi;
            return from_affine(X * zi2, Y * zi2 * zi);
        }
        bool operator==(const value_type &o) const {  // projective equality, like upstream
            if (is_zero() || o.is_zero()) return is_zero() && o.is_zero();
            auto z1 = Z * Z, z2 = o.Z * o.Z;
            return X * z2 == o.X * z1 && Y * z2 * o.Z == o.Y * z1 * Z;
        }
        bool operator!=(const value_type &o) const { return !(*this == o); }
    };
};
template <std::size_t> struct bls12;
template <> struct bls12<381> {
    typedef fields::bls12_fq<381> base_field_type;
    typedef fields::bls12_fr<381> scalar_field_type;
    template <class...> using g1_type = zkb_curve_g1<base_field_type, scalar_field_type, ZKB_CURVE_BLS12_381_G1>;
    // G2 over Fq2 (B_query of r1cs_gg_ppzksnark, prover.hpp:113-119; v-keys of ipp2): same Jacobian value type
    template <class...> using g2_type = zkb_curve_g1<fields::zkb_field2<base_field_type>, scalar_field_type, ZKB_CURVE_BLS12_381_G2>;
};
template <std::size_t> struct alt_bn128;
template <> struct alt_bn128<254> {
    typedef fields::alt_bn128_fq<254> base_field_type;
    typedef fields::alt_bn128_fr<254> scalar_field_type;
    template <class...> using g1_type = zkb_curve_g1<base_field_type, scalar_field_type, ZKB_CURVE_BN254_G1>;
    template <class...> using g2_type = zkb_curve_g1<fields::zkb_field2<base_field_type>, scalar_field_type, ZKB_CURVE_BN254_G2>;
};
struct pallas {
    typedef fields::pallas_base_field base_field_type;
    typedef fields::pallas_scalar_field scalar_field_type;
    template <class...> using g1_type = zkb_curve_g1<base_field_type, scalar_field_type, ZKB_CURVE_PALLAS>;
};
}  // namespace curves

namespace policies {
struct multiexp_method_BDLO12 {};
struct multiexp_method_bos_coster {};
struct multiexp_method_naive_plain {};
}  // namespace policies

// A base vector kept resident on the GPU: a KZG commitment key (kzg.hpp:100-118) or a Groth16 query
// vector (proving_key.hpp:43-55) is created once and used by many multiexps.
template <class GroupType>
class multiexp_bases {
    zkb_msm_bases *h = nullptr;
    std::size_t n = 0;

public:
    multiexp_bases() = default;
    template <class BaseIt>
    multiexp_bases(BaseIt first, BaseIt last) {
        constexpr int CL = GroupType::base_field_type::limbs32;
        n = (std::size_t)std::distance(first, last);
        std::vector<std::uint32_t> buf(n * 2 * CL, 0);
        std::size_t i = 0;
        for (BaseIt it = first; it != last; ++it, ++i) {
            if (it->is_zero()) continue;  // the all-zero encoding is the point at infinity
            auto a = it->to_affine();
            a.X.to_canonical_limbs(&buf[(2 * i) * CL]);
            a.Y.to_canonical_limbs(&buf[(2 * i + 1) * CL]);
        }
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_msm_bases_create(ctx, GroupType::curve_id, n, buf.data(), ZKB_MEM_HOST, nullptr, &h), ctx,
                          "zkb_msm_bases_create");
    }
    multiexp_bases(const multiexp_bases &) = delete;
    multiexp_bases &operator=(const multiexp_bases &) = delete;
    multiexp_bases(multiexp_bases &&o) noexcept : h(o.h), n(o.n) { o.h = nullptr; }
    ~multiexp_bases() { if (h) zkb_msm_bases_free(h); }
    std::size_t size() const { return n; }
    // one-off window table for a key that serves many multiexps (zkb_msm_bases_precompute)
    void precompute(int window_bits = 0, std::uint64_t max_bytes = 8ull << 30) {
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_msm_bases_precompute(ctx, h, window_bits, max_bytes, nullptr), ctx, "zkb_msm_bases_precompute");
    }

    // sum_i scalars[i] * bases[offset + i]
    template <class ScalarIt>
    typename GroupType::value_type multiexp(std::size_t offset, ScalarIt s0, ScalarIt s1) const {
        typedef typename GroupType::base_field_type BF;
        constexpr int CL = BF::limbs32;
        std::size_t cnt = (std::size_t)std::distance(s0, s1);
        if (offset + cnt > n) throw std::invalid_argument("multiexp: scalar range longer than the base range");
        std::vector<std::uint32_t> sc(cnt * 8 + 8);
        std::size_t i = 0;
        for (ScalarIt it = s0; it != s1; ++it, ++i) it->to_canonical_limbs(&sc[8 * i]);
        std::uint32_t res[2 * 24] = {0};
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_msm(ctx, h, offset, cnt, sc.data(), ZKB_MEM_HOST, res, nullptr), ctx, "zkb_msm");
        bool inf = true;
        for (int k = 0; k < 2 * CL; k++) inf = inf && res[k] == 0;
        if (inf) return GroupType::value_type::zero();
        return GroupType::value_type::from_affine(BF::value_type::from_canonical_limbs(res),
                                                  BF::value_type::from_canonical_limbs(res + CL));
    }
};

// algebra::multiexp<Method>(bases_begin, bases_end, scalars_begin, scalars_end, chunks)
// (kzg.hpp:146; r1cs_gg_ppzksnark/prover.hpp:125-131).  The result is method independent; `chunks`
// (OpenMP chunking upstream) has no meaning on the device.  Ranges must have equal length (the
// callers assert it, kzg.hpp:145).
template <class Method, class BaseIt, class ScalarIt>
typename std::iterator_traits<BaseIt>::value_type multiexp(BaseIt b0, BaseIt b1, ScalarIt s0, ScalarIt s1,
                                                           std::size_t /*chunks*/) {
    typedef typename std::iterator_traits<BaseIt>::value_type G;
    if (std::distance(b0, b1) != std::distance(s0, s1)) throw std::invalid_argument("multiexp: ranges differ in length");
    multiexp_bases<typename G::group_type> B(b0, b1);
    return B.multiexp(0, s0, s1);
}
// multiexp_with_mixed_addition (prover.hpp:108-114, :133-139): upstream drops zero scalars and adds unit
// scalars directly before running the bucket method; on the device zero digits are skipped and a unit
// scalar is exactly one mixed addition, so the same entry point serves it.
template <class Method, class BaseIt, class ScalarIt>
typename std::iterator_traits<BaseIt>::value_type multiexp_with_mixed_addition(BaseIt b0, BaseIt b1, ScalarIt s0,
                                                                               ScalarIt s1, std::size_t chunks) {
    return multiexp<Method>(b0, b1, s0, s1, chunks);
}
}  // namespace algebra

// ========================================================================================== math
namespace math {

template <class FieldType>
class evaluation_domain {
public:
    typedef typename FieldType::value_type value_type;
    std::size_t m;
    explicit evaluation_domain(std::size_t m) : m(m) {}
    virtual ~evaluation_domain() {}
    std::size_t size() const { return m; }
    virtual void fft(std::vector<value_type> &a) = 0;
    virtual void inverse_fft(std::vector<value_type> &a) = 0;
    virtual value_type get_domain_element(std::size_t idx) = 0;
    virtual std::vector<value_type> evaluate_all_lagrange_polynomials(const value_type &t) = 0;
    virtual value_type compute_vanishing_polynomial(const value_type &t) = 0;
    virtual void add_poly_z(const value_type &coeff, std::vector<value_type> &H) = 0;
    virtual void divide_by_z_on_coset(std::vector<value_type> &P) = 0;
};

// m = 2^k <= 2^s.  fft/inverse_fft run on the GPU (zkb_ntt); the O(1)/O(m) scalar helpers stay on the host.
template <class FieldType>
class basic_radix2_domain : public evaluation_domain<FieldType> {
public:
    typedef typename FieldType::value_type value_type;
    int log_m;
    value_type omega;

    explicit basic_radix2_domain(std::size_t m) : evaluation_domain<FieldType>(m) {
        log_m = zkb_detail::log2_exact(m);
        std::uint32_t l[8];
        int s = zkb_field_unity_root(FieldType::field_id, log_m, l);
        if (s != ZKB_OK) throw std::invalid_argument("basic_radix2_domain: size exceeds the field's two-adicity");
        omega = value_type::from_canonical_limbs(l);
    }

    void transform(std::vector<value_type> &a, int inverse, const value_type *shift) {
        if (a.size() != this->m) {
            if (a.size() > this->m) throw std::invalid_argument("basic_radix2: expected a.size() <= this->m");
            a.resize(this->m, value_type::zero());   // upstream zero-pads short inputs
        }
        std::uint32_t sh[8];
        if (shift) shift->to_canonical_limbs(sh);
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_ntt(ctx, FieldType::field_id, log_m, 1, a.data(), a.data(), inverse, shift ? sh : nullptr,
                                  ZKB_MEM_HOST, nullptr), ctx, "zkb_ntt");
    }
    void fft(std::vector<value_type> &a) override { transform(a, 0, nullptr); }
    void inverse_fft(std::vector<value_type> &a) override { transform(a, 1, nullptr); }
    // fused forms of  multiply_by_coset(a, g); fft(a)  and  inverse_fft(a); multiply_by_coset(a, g^-1)
    void coset_fft(std::vector<value_type> &a, const value_type &g) { transform(a, 0, &g); }
    void inverse_coset_fft(std::vector<value_type> &a, const value_type &g) { transform(a, 1, &g); }

    value_type get_domain_element(std::size_t idx) override { return omega.pow(idx); }
    value_type compute_vanishing_polynomial(const value_type &t) override { return t.pow(this->m) - value_type::one(); }
    void add_poly_z(const value_type &coeff, std::vector<value_type> &H) override {
        if (H.size() != this->m + 1) throw std::invalid_argument("basic_radix2: expected H.size() == this->m+1");
        H[this->m] += coeff;
        H[0] -= coeff;
    }
    void divide_by_z_on_coset(std::vector<value_type> &P) override {
        value_type g = algebra::fields::arithmetic_params<FieldType>::multiplicative_generator_value();
        value_type zi = compute_vanishing_polynomial(g).inversed();
        for (std::size_t i = 0; i < this->m; i++) P[i] *= zi;
    }
    std::vector<value_type> evaluate_all_lagrange_polynomials(const value_type &t) override {
        const std::size_t m = this->m;
        std::vector<value_type> u(m, value_type::zero());
        if (m == 1) { u[0] = value_type::one(); return u; }
        if (t.pow(m) == value_type::one()) {
            value_type w = value_type::one();
            for (std::size_t i = 0; i < m; i++) {
                if (w == t) { u[i] = value_type::one(); return u; }
                w *= omega;
            }
        }
        value_type Z = t.pow(m) - value_type::one();
        value_type l = Z * value_type(m).inversed();
        value_type r = value_type::one();
        for (std::size_t i = 0; i < m; i++) {
            u[i] = l * (t - r).inversed();
            l *= omega;
            r *= omega;
        }
        return u;
    }
};

// Only power-of-two sizes are served (upstream picks extended/step radix-2 domains otherwise).
template <class FieldType>
std::shared_ptr<evaluation_domain<FieldType>> make_evaluation_domain(std::size_t m) {
    return std::make_shared<basic_radix2_domain<FieldType>>(m);
}

// D[i] of size 2^(max_log - i)  (basic_fri.hpp:162,179; pinned by test/commitment/fri.cpp:122-123)
template <class FieldType>
std::vector<std::shared_ptr<evaluation_domain<FieldType>>> calculate_domain_set(std::size_t max_log, std::size_t set_size) {
    std::vector<std::shared_ptr<evaluation_domain<FieldType>>> D(set_size);
    for (std::size_t i = 0; i < set_size; i++) D[i] = make_evaluation_domain<FieldType>(std::size_t(1) << (max_log - i));
    return D;
}

// a[i] *= g^i   (r1cs_to_qap.hpp:266).  Host loop like upstream; callers that follow it with fft() can
// use basic_radix2_domain::coset_fft to get both in one device pass.
template <class V>
void multiply_by_coset(std::vector<V> &a, const V &g) {
    V u = g;
    for (std::size_t i = 1; i < a.size(); i++) {
        a[i] *= u;
        u *= g;
    }
}

// polynomial in evaluation form on the 2^k subgroup, natural order
template <class V>
class polynomial_dfs {
    typedef typename V::field_type FieldType;
    std::vector<V> val;
    std::size_t _d = 0;

public:
    typedef V value_type;
    polynomial_dfs() : val(1, V::zero()) {}
    polynomial_dfs(std::size_t d, std::size_t n, const V &x) : val(n, x), _d(d) { zkb_detail::log2_exact(n); }
    template <class It>
    polynomial_dfs(std::size_t d, It first, It last) : val(first, last), _d(d) { zkb_detail::log2_exact(val.size()); }
    polynomial_dfs(std::size_t d, std::vector<V> v) : val(std::move(v)), _d(d) { zkb_detail::log2_exact(val.size()); }

    std::size_t size() const { return val.size(); }
    std::size_t degree() const { return _d; }
    V &operator[](std::size_t i) { return val[i]; }
    const V &operator[](std::size_t i) const { return val[i]; }
    typename std::vector<V>::const_iterator begin() const { return val.begin(); }
    typename std::vector<V>::const_iterator end() const { return val.end(); }
    const std::vector<V> &data() const { return val; }
    bool operator==(const polynomial_dfs &o) const { return val == o.val && _d == o._d; }

    // basic_fri.hpp:369-371,451-455: inverse_fft on the own-size domain, zero-pad, fft on the new one
    void resize(std::size_t sz, std::shared_ptr<evaluation_domain<FieldType>> = nullptr,
                std::shared_ptr<evaluation_domain<FieldType>> = nullptr) {
        if (sz == val.size()) return;
        int lo = zkb_detail::log2_exact(sz);
        if (val.size() == 1) { val.assign(sz, val[0]); return; }
        int li = zkb_detail::log2_exact(val.size());
        zkb_ctx *ctx = zkb_detail::context();
        if (lo > li) {
            std::vector<V> out(sz);
            zkb_detail::check(zkb_lde(ctx, FieldType::field_id, li, lo, 1, val.data(), out.data(), ZKB_MEM_HOST, nullptr), ctx, "zkb_lde");
            val.swap(out);
        } else {  // shrink: coefficients, truncate, evaluate
            zkb_detail::check(zkb_ntt(ctx, FieldType::field_id, li, 1, val.data(), val.data(), 1, nullptr, ZKB_MEM_HOST, nullptr), ctx, "zkb_ntt");
            val.resize(sz);
            zkb_detail::check(zkb_ntt(ctx, FieldType::field_id, lo, 1, val.data(), val.data(), 0, nullptr, ZKB_MEM_HOST, nullptr), ctx, "zkb_ntt");
        }
    }
    template <class Container>
    void from_coefficients(const Container &c) {
        std::size_t n = 1;
        while (n < c.size()) n <<= 1;
        val.assign(c.begin(), c.end());
        val.resize(n, V::zero());
        _d = c.size() ? c.size() - 1 : 0;
        basic_radix2_domain<FieldType>(n).fft(val);
    }
    std::vector<V> coefficients() const {
        std::vector<V> c(val);
        basic_radix2_domain<FieldType>(c.size()).inverse_fft(c);
        return c;
    }
    V evaluate(const V &x) const {
        std::vector<V> c = coefficients();
        V r = V::zero();
        for (std::size_t i = c.size(); i-- > 0;) r = r * x + c[i];
        return r;
    }
};
}  // namespace math

// ========================================================================================== zk
namespace zk {
namespace commitments {
namespace detail {
// fold_polynomial.hpp:68-93 (dfs form); linear in f, so Montgomery-form data is folded as is.
template <class FieldType>
math::polynomial_dfs<typename FieldType::value_type> fold_polynomial(math::polynomial_dfs<typename FieldType::value_type> &f,
                                                                     const typename FieldType::value_type &alpha,
                                                                     std::shared_ptr<math::evaluation_domain<FieldType>> domain) {
    typedef typename FieldType::value_type V;
    if (f.size() != domain->size()) throw std::invalid_argument("fold_polynomial: f.size() != domain->size()");
    int log_n = zkb_detail::log2_exact(domain->size());
    std::vector<V> out(domain->size() / 2);
    std::uint32_t a[8];
    alpha.to_canonical_limbs(a);
    zkb_ctx *ctx = zkb_detail::context();
    zkb_detail::check(zkb_fri_fold(ctx, FieldType::field_id, log_n, f.data().data(), a, out.data(), ZKB_MEM_HOST, nullptr), ctx,
                      "zkb_fri_fold");
    return math::polynomial_dfs<V>(domain->size() / 2 - 1, std::move(out));
}
}  // namespace detail

// ---- knowledge commitments (Groth16 B_query: pairs (G2, G1) that share one scalar) ----------------------
}  // namespace commitments
}  // namespace zk
namespace container {
// crypto3-containers' sparse_vector as the call sites use it (indices sorted ascending, values aligned)
template <class T>
struct sparse_vector {
    std::vector<std::size_t> indices;
    std::vector<typename T::value_type> values;
    std::size_t domain_size_ = 0;
    std::size_t size() const { return indices.size(); }
};
}  // namespace container
namespace zk {
namespace commitments {
namespace detail {
// element_kc<T1, T2> (detail/polynomial/element_knowledge_commitment.hpp:54-186): the pair (g, h)
template <class T1, class T2>
struct element_kc {
    typename T1::value_type g;
    typename T2::value_type h;
    element_kc() : g(T1::value_type::zero()), h(T2::value_type::zero()) {}
    element_kc(const typename T1::value_type &g_, const typename T2::value_type &h_) : g(g_), h(h_) {}
    static element_kc zero() { return element_kc(); }
    bool is_zero() const { return g.is_zero() && h.is_zero(); }
    bool operator==(const element_kc &o) const { return g == o.g && h == o.h; }
    bool operator!=(const element_kc &o) const { return !(*this == o); }
};
}  // namespace detail
template <class T1, class T2>
struct knowledge_commitment {
    typedef T1 type1;
    typedef T2 type2;
    typedef detail::element_kc<T1, T2> value_type;
};
template <class T1, class T2>
using knowledge_commitment_vector = container::sparse_vector<knowledge_commitment<T1, T2>>;

// kc_multiexp_with_mixed_addition (knowledge_commitment_multiexp.hpp:57-108): the entries of `vec` whose
// index lies in [min_idx, max_idx) take the scalar scalar_start[index - min_idx]; upstream skips zero
// scalars, adds unit scalars directly and runs multiexp<Method> over the rest on the (g, h) pairs.  Here the
// same selection feeds two device MSMs (one per group) that share the gathered scalars; zero digits are
// skipped and a unit scalar is a single mixed addition on the device.
template <class MultiexpMethod, class T1, class T2, class InputFieldIterator>
typename knowledge_commitment<T1, T2>::value_type kc_multiexp_with_mixed_addition(
    const knowledge_commitment_vector<T1, T2> &vec, const std::size_t min_idx, const std::size_t max_idx,
    InputFieldIterator scalar_start, InputFieldIterator scalar_end, const std::size_t chunks) {
    typedef typename std::iterator_traits<InputFieldIterator>::value_type field_value_type;
    const std::size_t scalar_length = (std::size_t)std::distance(scalar_start, scalar_end);
    if (scalar_length > vec.domain_size_) throw std::invalid_argument("kc_multiexp: more scalars than the vector's domain");
    auto index_it = std::lower_bound(vec.indices.begin(), vec.indices.end(), min_idx);
    auto value_it = vec.values.begin() + (index_it - vec.indices.begin());
    std::vector<field_value_type> p;
    std::vector<typename T1::value_type> g;
    std::vector<typename T2::value_type> h;
    for (; index_it != vec.indices.end() && *index_it < max_idx; ++index_it, ++value_it) {
        const std::size_t scalar_position = *index_it - min_idx;
        if (scalar_position >= scalar_length) throw std::invalid_argument("kc_multiexp: index outside the scalar range");
        p.push_back(*(scalar_start + scalar_position));
        g.push_back(value_it->g);
        h.push_back(value_it->h);
    }
    return typename knowledge_commitment<T1, T2>::value_type(
        algebra::multiexp<MultiexpMethod>(g.begin(), g.end(), p.begin(), p.end(), chunks),
        algebra::multiexp<MultiexpMethod>(h.begin(), h.end(), p.begin(), p.end(), chunks));
}

// the precommitment (containers::merkle_tree<Hash,2>) kept on the device
template <int HashId>
class device_merkle_tree {
    zkb_merkle_tree *h = nullptr;
    std::vector<std::uint8_t> _root;

public:
    device_merkle_tree() = default;
    device_merkle_tree(zkb_merkle_tree *t, std::vector<std::uint8_t> r) : h(t), _root(std::move(r)) {}
    device_merkle_tree(const device_merkle_tree &) = delete;
    device_merkle_tree(device_merkle_tree &&o) noexcept : h(o.h), _root(std::move(o._root)) { o.h = nullptr; }
    ~device_merkle_tree() { if (h) zkb_merkle_free(h); }
    const std::vector<std::uint8_t> &root() const { return _root; }
    std::size_t leaves() const { return (std::size_t)zkb_merkle_leaves(h); }
    // merkle_proof<Hash,2>(tree, idx): sibling digests, leaf level first
    std::vector<std::vector<std::uint8_t>> path(std::size_t idx) const {
        std::size_t depth = 0, db = (std::size_t)zkb_merkle_digest_bytes(HashId);
        for (std::size_t n = leaves(); n > 1; n >>= 1) depth++;
        std::vector<std::uint8_t> raw(depth * db + 1);
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_merkle_path(ctx, h, idx, raw.data()), ctx, "zkb_merkle_path");
        std::vector<std::vector<std::uint8_t>> p(depth);
        for (std::size_t i = 0; i < depth; i++) p[i].assign(raw.begin() + i * db, raw.begin() + (i + 1) * db);
        return p;
    }
};
}  // namespace commitments

namespace algorithms {
// precommit<FRI>(container<polynomial_dfs>, D, fri_step) (basic_fri.hpp:445-496): every polynomial is
// resized to |D|, leaves packed as in :466-492 and hashed into a binary Merkle tree.  All polynomials of the
// container must have the same size (they do at every call site: lpc.hpp:101-106 commits one batch).
template <class FieldType, int HashId, class Container>
commitments::device_merkle_tree<HashId> precommit(const Container &polys, std::shared_ptr<math::evaluation_domain<FieldType>> D,
                                                  std::size_t fri_step) {
    typedef typename FieldType::value_type V;
    if (polys.size() == 0) throw std::invalid_argument("precommit: empty polynomial list");
    const std::size_t n_in = polys[0].size();
    int log_in = zkb_detail::log2_exact(n_in), log_out = zkb_detail::log2_exact(D->size());
    // leaves are hashes of canonical big-endian integers: leave Montgomery form on the way in
    std::vector<std::uint32_t> buf(polys.size() * n_in * 8);
    for (std::size_t p = 0; p < polys.size(); p++) {
        if (polys[p].size() != n_in) throw std::invalid_argument("precommit: polynomials of different sizes");
        for (std::size_t i = 0; i < n_in; i++) polys[p][i].to_canonical_limbs(&buf[(p * n_in + i) * 8]);
    }
    zkb_ctx *ctx = zkb_detail::context();
    std::vector<std::uint8_t> root((std::size_t)zkb_merkle_digest_bytes(HashId));
    zkb_merkle_tree *t = nullptr;
    zkb_detail::check(zkb_lpc_commit(ctx, FieldType::field_id, HashId, log_in < 1 ? 1 : log_in, log_out, (int)fri_step,
                                     (std::uint32_t)polys.size(), buf.data(), ZKB_MEM_HOST, root.data(), &t, nullptr),
                      ctx, "zkb_lpc_commit");
    (void)sizeof(V);
    return commitments::device_merkle_tree<HashId>(t, std::move(root));
}
}  // namespace algorithms
}  // namespace zk
}  // namespace crypto3
}  // namespace nil
