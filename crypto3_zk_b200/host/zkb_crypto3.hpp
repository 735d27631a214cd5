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
//   nil::marshalling::curve_element_serializer<bls12<381>>, verifier_input_{,de}serializer_tvm (proof)
//                                                                  (r1cs_gg_ppzksnark/marshalling.hpp, host only)
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
#include <map>
#include <array>
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/zkb200.h"
#include "../csrc/zkb_field.cuh"   // Fp2<B> (generic over the base type; compiled for the host here)
#include "../csrc/zkb_curve.cuh"   // XYZZ<F> group formulas, shared with the device (host tails: proof assembly)
#include "../csrc/zkb_hostfield.h"

namespace nil {
namespace crypto3 {

namespace zkb_detail {
inline void check(int status, zkb_ctx *ctx, const char *what) {
    if (status == ZKB_OK) return;
    std::string msg = std::string(what) + ": " + zkb_status_string(status);
    if (ctx && *zkb_ctx_last_error(ctx)) msg += std::string(" (") + zkb_ctx_last_error(ctx) + ")";
    if (ctx) zkb_ctx_clear_error(ctx);   // consumed: a later failure that records no message must not show this one
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
        // this^e, e given as little-endian 32-bit limbs
        value_type pow_limbs(const std::uint32_t *e, int n) const {
            value_type r = one();
            for (int i = n - 1; i >= 0; i--)
                for (int b = 31; b >= 0; b--) {
                    r *= r;
                    if ((e[i] >> b) & 1) r *= *this;
                }
            return r;
        }
        bool operator<(const value_type &o) const {   // canonical integer order (std::map / std::find keys)
            std::uint32_t a[limbs32], b[limbs32];
            to_canonical_limbs(a);
            o.to_canonical_limbs(b);
            for (int i = limbs32 - 1; i >= 0; i--)
                if (a[i] != b[i]) return a[i] < b[i];
            return false;
        }
    };
    // (p - 1) >> k as limbs
    static void modulus_minus_one_shifted(int k, std::uint32_t *out) {
        std::uint32_t m[limbs32];
        for (int i = 0; i < limbs32; i++) m[i] = Params::mod(i);
        m[0] -= 1;   // p is odd
        for (int i = 0; i < limbs32; i++) {
            int src = i + k / 32, sh = k % 32;
            std::uint64_t lo = src < limbs32 ? m[src] : 0, hi = src + 1 < limbs32 ? m[src + 1] : 0;
            out[i] = (std::uint32_t)(sh ? ((lo >> sh) | (hi << (32 - sh))) : lo);
        }
    }
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

        // Host group law for the O(1) tails the reference runs on the host too (proof assembly prover.hpp:141-157,
        // kzg.hpp:113-117): the XYZZ formulas of csrc/zkb_curve.cuh on the host field.  (X, Y, Z) <-> XYZZ without
        // an inversion: ZZ = Z^2, ZZZ = Z^3 and back (X ZZ, Y ZZZ, ZZ).
        typedef ::zkb::XYZZ<typename BaseField::backend> xyzz_type;
        xyzz_type to_xyzz() const {
            if (is_zero()) return xyzz_type::infinity();
            xyzz_type r;
            r.X = X.data; r.Y = Y.data;
            r.ZZ = Z.data * Z.data;
            r.ZZZ = r.ZZ * Z.data;
            return r;
        }
        static value_type from_xyzz(const xyzz_type &p) {
            if (p.is_infinity()) return zero();
            value_type r;
            r.X.data = p.X * p.ZZ; r.Y.data = p.Y * p.ZZZ; r.Z.data = p.ZZ;
            return r;
        }
        value_type operator+(const value_type &o) const {
            xyzz_type a = to_xyzz();
            a.add(o.to_xyzz());
            return from_xyzz(a);
        }
        value_type operator-() const {
            value_type r = *this;
            r.Y = -r.Y;
            return r;
        }
        value_type operator-(const value_type &o) const { return *this + (-o); }
        value_type &operator+=(const value_type &o) { *this = *this + o; return *this; }
        value_type doubled() const { return from_xyzz(to_xyzz().dbl()); }
        // k * P, double-and-add over the canonical limbs of k (a handful of these per proof)
        friend value_type operator*(const typename ScalarField::value_type &k, const value_type &p) {
            std::uint32_t l[ScalarField::limbs32];
            k.to_canonical_limbs(l);
            xyzz_type acc = xyzz_type::infinity(), base = p.to_xyzz();
            for (int i = ScalarField::limbs32 - 1; i >= 0; i--)
                for (int b = 31; b >= 0; b--) {
                    acc = acc.dbl();
                    if ((l[i] >> b) & 1) acc.add(base);
                }
            return from_xyzz(acc);
        }
        friend value_type operator*(const value_type &p, const typename ScalarField::value_type &k) { return k * p; }
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
            if (it->Z == GroupType::base_field_type::value_type::one()) {   // already affine (generator / deserialiser output): no inversion
                it->X.to_canonical_limbs(&buf[(2 * i) * CL]);
                it->Y.to_canonical_limbs(&buf[(2 * i + 1) * CL]);
                continue;
            }
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
    multiexp_bases &operator=(multiexp_bases &&o) noexcept {
        if (this != &o) {
            if (h) zkb_msm_bases_free(h);
            h = o.h; n = o.n;
            o.h = nullptr;
        }
        return *this;
    }
    ~multiexp_bases() { if (h) zkb_msm_bases_free(h); }
    zkb_msm_bases *handle() const { return h; }
    std::size_t size() const { return n; }
    // one-off window table for a key that serves many multiexps (zkb_msm_bases_precompute)
    void precompute(int window_bits = 0, std::uint64_t max_bytes = 8ull << 30) {
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_msm_bases_precompute(ctx, h, window_bits, max_bytes, nullptr), ctx, "zkb_msm_bases_precompute");
    }

    // sum_i scalars[i] * bases[offset + i] for scalars already laid out as canonical limbs: in a DEVICE buffer (the
    // witness map leaves the assignment and H there) or in host memory
    typename GroupType::value_type multiexp_raw(std::size_t offset, std::size_t cnt, const void *scalars, int mem) const {
        typedef typename GroupType::base_field_type BF;
        constexpr int CL = BF::limbs32;
        if (offset + cnt > n) throw std::invalid_argument("multiexp: scalar range longer than the base range");
        std::uint32_t res[2 * 24] = {0};
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_msm(ctx, h, offset, cnt, scalars, mem, res, nullptr), ctx, "zkb_msm");
        bool inf = true;
        for (int k = 0; k < 2 * CL; k++) inf = inf && res[k] == 0;
        if (inf) return GroupType::value_type::zero();
        return GroupType::value_type::from_affine(BF::value_type::from_canonical_limbs(res),
                                                  BF::value_type::from_canonical_limbs(res + CL));
    }
    typename GroupType::value_type multiexp_device(std::size_t offset, std::size_t cnt, const void *d_scalars) const {
        return multiexp_raw(offset, cnt, d_scalars, ZKB_MEM_DEVICE);
    }
    typename GroupType::value_type multiexp_limbs(std::size_t offset, std::size_t cnt, const std::uint32_t *scalars) const {
        return multiexp_raw(offset, cnt, scalars, ZKB_MEM_HOST);
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

// Fixed-base batch exponentiation of the Groth16 generator (r1cs_gg_ppzksnark/generator.hpp:167-225): upstream builds a
// window table on the host (get_window_table) and calls batch_exp(scalar_size, window, table, v); here the table lives on
// the device inside zkb_batch_exp (32 byte windows), so window_table only carries the base point.
template <class GroupType>
struct window_table {
    typename GroupType::value_type base;
};
template <class GroupType>
std::size_t get_exp_window_size(std::size_t) { return 8; }
template <class GroupType>
window_table<GroupType> get_window_table(std::size_t /*scalar_size*/, std::size_t /*window*/, const typename GroupType::value_type &g) {
    window_table<GroupType> t;
    t.base = g;
    return t;
}
// v[i] * table.base for every i (zero scalars give the group's zero)
template <class GroupType, class FieldType>
std::vector<typename GroupType::value_type> batch_exp(std::size_t /*scalar_size*/, std::size_t /*window*/,
                                                      const window_table<GroupType> &table,
                                                      const std::vector<typename FieldType::value_type> &v) {
    typedef typename GroupType::base_field_type BF;
    constexpr int CL = BF::limbs32;
    std::vector<typename GroupType::value_type> out(v.size(), GroupType::value_type::zero());
    if (v.empty() || table.base.is_zero()) return out;
    std::uint32_t base[2 * 24] = {0};
    auto a = table.base.to_affine();
    a.X.to_canonical_limbs(base);
    a.Y.to_canonical_limbs(base + CL);
    std::vector<std::uint32_t> sc(v.size() * 8), res(v.size() * 2 * CL);
    for (std::size_t i = 0; i < v.size(); i++) v[i].to_canonical_limbs(&sc[8 * i]);
    zkb_ctx *ctx = zkb_detail::context();
    zkb_detail::check(zkb_batch_exp(ctx, GroupType::curve_id, v.size(), base, sc.data(), res.data(), ZKB_MEM_HOST, nullptr), ctx,
                      "zkb_batch_exp");
    for (std::size_t i = 0; i < v.size(); i++) {
        bool inf = true;
        for (int k = 0; k < 2 * CL; k++) inf = inf && res[i * 2 * CL + k] == 0;
        if (!inf)
            out[i] = GroupType::value_type::from_affine(BF::value_type::from_canonical_limbs(&res[i * 2 * CL]),
                                                        BF::value_type::from_canonical_limbs(&res[i * 2 * CL + CL]));
    }
    return out;
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
    // the lambda paths of a FRI query phase (basic_fri.hpp:846-862) with one device gather
    std::vector<std::vector<std::vector<std::uint8_t>>> paths(const std::vector<std::uint64_t> &idx) const {
        std::size_t depth = 0, db = (std::size_t)zkb_merkle_digest_bytes(HashId);
        for (std::size_t n = leaves(); n > 1; n >>= 1) depth++;
        std::vector<std::uint8_t> raw(idx.size() * depth * db + 1);
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_merkle_paths(ctx, h, (std::uint32_t)idx.size(), idx.data(), raw.data(), nullptr), ctx, "zkb_merkle_paths");
        std::vector<std::vector<std::vector<std::uint8_t>>> out(idx.size(), std::vector<std::vector<std::uint8_t>>(depth));
        for (std::size_t q = 0; q < idx.size(); q++)
            for (std::size_t i = 0; i < depth; i++)
                out[q][i].assign(raw.begin() + (q * depth + i) * db, raw.begin() + (q * depth + i + 1) * db);
        return out;
    }
};
}  // namespace commitments

}  // namespace zk

// ========================================================================================== hashes (host side of the transcript)
// The transcript stays on the host, as in the reference: a few hundred bytes per proof.  Tag types carry the ABI hash id so
// that the same tag selects the device Merkle hash (zkb_lpc_commit) and the device grinding kernel (zkb_pow_grind).
namespace hashes {
namespace zkb_detail_hash {
inline std::uint64_t rol(std::uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
inline void keccak_f(std::uint64_t a[25]) {
    static const std::uint64_t RC[24] = {
        0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
        0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
        0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
        0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int r = 0; r < 24; r++) {
        std::uint64_t c[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) {
            std::uint64_t d = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
            for (int y = 0; y < 5; y++) a[x + 5 * y] ^= d;
        }
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], ROT[x + 5 * y]);
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= RC[r];
    }
}
// original Keccak padding (0x01 .. 0x80), as crypto3-hash keccak_1600
inline std::vector<std::uint8_t> keccak(const std::uint8_t *data, std::size_t n, std::size_t digest_bytes) {
    const std::size_t rate = 200 - 2 * digest_bytes;
    std::uint64_t a[25] = {0};
    std::vector<std::uint8_t> m(data, data + n);
    m.push_back(0x01);
    while (m.size() % rate) m.push_back(0);
    m.back() |= 0x80;
    for (std::size_t off = 0; off < m.size(); off += rate) {
        for (std::size_t i = 0; i < rate; i++) a[i / 8] ^= (std::uint64_t)m[off + i] << (8 * (i % 8));
        keccak_f(a);
    }
    std::vector<std::uint8_t> out(digest_bytes);
    for (std::size_t i = 0; i < digest_bytes; i++) out[i] = (std::uint8_t)(a[i / 8] >> (8 * (i % 8)));
    return out;
}
inline std::uint32_t ror(std::uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
inline std::vector<std::uint8_t> sha256(const std::uint8_t *data, std::size_t n) {
    static const std::uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    std::uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    std::vector<std::uint8_t> m(data, data + n);
    m.push_back(0x80);
    while (m.size() % 64 != 56) m.push_back(0);
    for (int i = 7; i >= 0; i--) m.push_back((std::uint8_t)(((std::uint64_t)n * 8) >> (8 * i)));
    for (std::size_t off = 0; off < m.size(); off += 64) {
        std::uint32_t w[64];
        for (int i = 0; i < 16; i++)
            w[i] = ((std::uint32_t)m[off + 4 * i] << 24) | ((std::uint32_t)m[off + 4 * i + 1] << 16) | ((std::uint32_t)m[off + 4 * i + 2] << 8) | m[off + 4 * i + 3];
        for (int i = 16; i < 64; i++)
            w[i] = w[i - 16] + (ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] + (ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10));
        std::uint32_t v[8];
        for (int i = 0; i < 8; i++) v[i] = h[i];
        for (int i = 0; i < 64; i++) {
            std::uint32_t t1 = v[7] + (ror(v[4], 6) ^ ror(v[4], 11) ^ ror(v[4], 25)) + ((v[4] & v[5]) ^ (~v[4] & v[6])) + K[i] + w[i];
            std::uint32_t t2 = (ror(v[0], 2) ^ ror(v[0], 13) ^ ror(v[0], 22)) + ((v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]));
            for (int k = 7; k > 0; k--) v[k] = v[k - 1];
            v[4] += t1;
            v[0] = t1 + t2;
        }
        for (int i = 0; i < 8; i++) h[i] += v[i];
    }
    std::vector<std::uint8_t> out(32);
    for (int i = 0; i < 8; i++)
        for (int b = 0; b < 4; b++) out[4 * i + b] = (std::uint8_t)(h[i] >> (24 - 8 * b));
    return out;
}
}  // namespace zkb_detail_hash

template <std::size_t Bits> struct keccak_1600;
template <> struct keccak_1600<256> {
    static constexpr int hash_id = ZKB_HASH_KECCAK_256;
    static constexpr std::size_t digest_bytes = 32;
    static std::vector<std::uint8_t> hash(const std::uint8_t *d, std::size_t n) { return zkb_detail_hash::keccak(d, n, 32); }
};
template <> struct keccak_1600<512> {
    static constexpr int hash_id = ZKB_HASH_KECCAK_512;
    static constexpr std::size_t digest_bytes = 64;
    static std::vector<std::uint8_t> hash(const std::uint8_t *d, std::size_t n) { return zkb_detail_hash::keccak(d, n, 64); }
};
template <std::size_t Bits> struct sha2;
template <> struct sha2<256> {
    static constexpr int hash_id = ZKB_HASH_SHA2_256;
    static constexpr std::size_t digest_bytes = 32;
    static std::vector<std::uint8_t> hash(const std::uint8_t *d, std::size_t n) { return zkb_detail_hash::sha256(d, n); }
};
}  // namespace hashes

namespace zk {
namespace transcript {
// fiat_shamir_heuristic_sequential<Hash> (zk/transcript/fiat_shamir.hpp:131-199): state = H(init); absorbing is
// state = H(state || data); challenge<Field>() is state = H(state) read as a big-endian integer into the field;
// int_challenge<Integral>() the low bits of the same.  Known answers: test/transcript/transcript.cpp:50-64.
template <class Hash>
class fiat_shamir_heuristic_sequential {
    std::vector<std::uint8_t> _state;

public:
    typedef Hash hash_type;
    fiat_shamir_heuristic_sequential() {
        const std::uint8_t zero = 0;
        _state = Hash::hash(&zero, 1);
    }
    template <class Range>
    explicit fiat_shamir_heuristic_sequential(const Range &r) {
        std::vector<std::uint8_t> d(r.begin(), r.end());
        _state = Hash::hash(d.data(), d.size());
    }
    const std::vector<std::uint8_t> &state() const { return _state; }
    template <class Range>
    void operator()(const Range &r) {
        std::vector<std::uint8_t> d(_state);
        d.insert(d.end(), r.begin(), r.end());
        _state = Hash::hash(d.data(), d.size());
    }
    template <class FieldType>
    typename FieldType::value_type challenge() {
        typedef typename FieldType::backend B;
        _state = Hash::hash(_state.data(), _state.size());
        // big-endian digest = sum_k chunk_k 2^(256 k), chunk_0 the last 32 bytes; Montgomery products accept any 256-bit input
        typename FieldType::value_type acc, radix = FieldType::value_type::one(), scale = FieldType::value_type::one();
        const typename FieldType::value_type two32(std::uint64_t(1) << 32);
        for (int i = 0; i < 8; i++) radix *= two32;   // 2^256 mod p
        for (std::size_t off = _state.size(); off > 0; off -= 32) {
            std::uint32_t l[FieldType::limbs32] = {0};
            for (int i = 0; i < 32; i++) l[i / 4] |= (std::uint32_t)_state[off - 1 - i] << (8 * (i % 4));
            typename FieldType::value_type chunk;
            chunk.data = B::from_limbs32(l).to_mont();
            acc += chunk * scale;
            scale *= radix;
        }
        return acc;
    }
    template <class Integral>
    Integral int_challenge() {
        _state = Hash::hash(_state.data(), _state.size());
        Integral v = 0;
        for (std::size_t i = _state.size() - sizeof(Integral); i < _state.size(); i++) v = (Integral)((v << 8) | _state[i]);
        return v;
    }
};
}  // namespace transcript

namespace commitments {
// proof_of_work<TranscriptHashType, std::uint32_t> (zk/commitments/detail/polynomial/proof_of_work.hpp:40-82): the search
// runs on the device (zkb_pow_grind, one nonce per thread) from nonce 0 instead of std::rand(); verify is the reference's.
template <class TranscriptHashType, class OutType = std::uint32_t>
class proof_of_work {
public:
    typedef TranscriptHashType transcript_hash_type;
    typedef transcript::fiat_shamir_heuristic_sequential<transcript_hash_type> transcript_type;
    typedef OutType output_type;
    static std::vector<std::uint8_t> be32(output_type v) {
        return {std::uint8_t(v >> 24), std::uint8_t(v >> 16), std::uint8_t(v >> 8), std::uint8_t(v)};
    }
    static output_type generate(transcript_type &transcript, OutType mask = 0xFFFF) {
        zkb_ctx *ctx = zkb_detail::context();
        std::uint32_t nonce = 0;
        zkb_detail::check(zkb_pow_grind(ctx, TranscriptHashType::hash_id, transcript.state().data(), 0, (std::uint32_t)mask, &nonce, nullptr),
                          ctx, "zkb_pow_grind");
        transcript(be32(nonce));
        (void)transcript.template int_challenge<output_type>();
        return (output_type)nonce;
    }
    static bool verify(transcript_type &transcript, output_type proof_of_work, OutType mask = 0xFFFF) {
        transcript(be32(proof_of_work));
        output_type result = transcript.template int_challenge<output_type>();
        return (result & mask) == 0;
    }
};
}  // namespace commitments

// ------------------------------------------------------------------------------------------ lpc_commitment_scheme
namespace zkb_detail_lpc {
// RAII device buffer over zkb_buf_alloc / zkb_buf_free
struct device_buffer {
    void *p = nullptr;
    std::size_t bytes = 0;
    device_buffer() = default;
    explicit device_buffer(std::size_t n) : bytes(n) {
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_buf_alloc(ctx, n, &p), ctx, "zkb_buf_alloc");
    }
    device_buffer(const device_buffer &) = delete;
    device_buffer &operator=(const device_buffer &) = delete;
    device_buffer(device_buffer &&o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; }
    device_buffer &operator=(device_buffer &&o) noexcept {
        if (this != &o) { release(); p = o.p; bytes = o.bytes; o.p = nullptr; }
        return *this;
    }
    void release() { if (p) zkb_buf_free(zkb_detail::context(), p); p = nullptr; }
    ~device_buffer() { release(); }
};
}  // namespace zkb_detail_lpc

namespace commitments {
// basic_batched_fri::params_type (basic_fri.hpp:151-229): the fields the prover reads
struct fri_params_type {
    std::vector<std::size_t> step_list;
    std::size_t degree_log = 0, lambda = 40, expand_factor = 2;
    bool use_grinding = false;
    std::uint32_t grinding_parameter = 0xFFFF;
    std::size_t log_d0() const { return degree_log + expand_factor; }
    std::size_t r() const { std::size_t s = 0; for (auto v : step_list) s += v; return s; }
    // params_type(max_step = 1, degree_log, lambda, expand_factor) (basic_fri.hpp:150-166): degree_log - 1 steps of 1
    static fri_params_type with_max_step_one(std::size_t degree_log, std::size_t lambda, std::size_t expand_factor,
                                             bool use_grinding = false, std::uint32_t grinding_parameter = 0xFFFF) {
        fri_params_type q;
        q.step_list.assign(degree_log - 1, 1);
        q.degree_log = degree_log; q.lambda = lambda; q.expand_factor = expand_factor;
        q.use_grinding = use_grinding; q.grinding_parameter = grinding_parameter;
        return q;
    }
};

// lpc_commitment_scheme (zk/commitments/polynomial/lpc.hpp:66-200) over its base polys_evaluator
// (zk/commitments/batched_commitment.hpp:60-250), with zk::algorithms::proof_eval<FRI> (basic_fri.hpp:670-923) behind
// proof_eval(): commit phase, grinding and query phase.  Polynomials are polynomial_dfs values; after commit() a batch
// lives in a device buffer, every per-element operation is one ABI call, the transcript and the proof stay on the host.
// The same sequence as crypto3_zk_b200/lpc.py (which tests/test_gpu_flows.py checks bit for bit against the oracle).
template <class FieldType, class MerkleHash, class TranscriptHash>
class lpc_commitment_scheme {
public:
    typedef FieldType field_type;
    typedef typename FieldType::value_type value_type;
    typedef math::polynomial_dfs<value_type> poly_type;
    typedef transcript::fiat_shamir_heuristic_sequential<TranscriptHash> transcript_type;
    typedef std::vector<std::uint8_t> commitment_type;
    typedef device_merkle_tree<MerkleHash::hash_id> precommitment_type;
    struct merkle_proof_type {
        std::size_t index = 0;
        std::vector<commitment_type> path;
        commitment_type root;
    };
    typedef std::vector<std::array<value_type, 2>> polynomial_values_type;
    struct initial_proof_type { std::vector<polynomial_values_type> values; merkle_proof_type p; };
    struct round_proof_type { polynomial_values_type y; merkle_proof_type p; };
    struct query_proof_type { std::map<std::size_t, initial_proof_type> initial_proof; std::vector<round_proof_type> round_proofs; };
    struct fri_proof_type {
        std::vector<commitment_type> fri_roots;
        std::vector<value_type> final_polynomial;
        std::vector<query_proof_type> query_proofs;
        std::uint32_t proof_of_work = 0;
    };
    struct proof_type { std::map<std::size_t, std::vector<std::vector<value_type>>> z; fri_proof_type fri_proof; };

private:
    typedef zkb_detail_lpc::device_buffer dbuf;
    static constexpr int fid = FieldType::field_id;
    fri_params_type _fri;
    std::map<std::size_t, std::vector<poly_type>> _polys;
    std::map<std::size_t, std::vector<std::vector<value_type>>> _points;
    std::map<std::size_t, precommitment_type> _trees;
    std::map<std::size_t, bool> _fixed, _locked;
    std::map<std::size_t, dbuf> _dev, _coef;           // evaluations / coefficients of a batch on the device
    std::map<std::size_t, std::size_t> _n;
    value_type _etha;
    std::map<std::size_t, std::vector<value_type>> _fixed_values;
    std::map<std::size_t, std::vector<std::vector<value_type>>> _z;

    static zkb_ctx *ctx() { return zkb_detail::context(); }
    static void ck(int st, const char *what) { zkb_detail::check(st, ctx(), what); }
    static void limbs(const value_type &v, std::uint32_t *l) { v.to_canonical_limbs(l); }
    static value_type omega(std::size_t log_n) {
        std::uint32_t l[8];
        ck(zkb_field_unity_root(fid, (int)log_n, l), "zkb_field_unity_root");
        return value_type::from_canonical_limbs(l);
    }
    // batch k as canonical limbs on the device (leaves hash canonical integers; the transforms are linear anyway)
    void upload(std::size_t k) {
        if (_dev.count(k)) return;
        const auto &ps = _polys.at(k);
        const std::size_t n = ps[0].size();
        std::vector<std::uint32_t> buf(ps.size() * n * 8);
        for (std::size_t i = 0; i < ps.size(); i++) {
            if (ps[i].size() != n) throw std::invalid_argument("lpc: polynomials of one batch must have the same size");
            for (std::size_t j = 0; j < n; j++) ps[i][j].to_canonical_limbs(&buf[(i * n + j) * 8]);
        }
        dbuf d(buf.size() * 4);
        ck(zkb_buf_copy(ctx(), d.p, ZKB_MEM_DEVICE, buf.data(), ZKB_MEM_HOST, buf.size() * 4, nullptr), "zkb_buf_copy");
        _dev[k] = std::move(d);
        _n[k] = n;
    }
    std::vector<value_type> unique_points() const {
        std::vector<value_type> out;
        for (const auto &kv : _points)
            for (const auto &pts : kv.second)
                for (const auto &x : pts)
                    if (std::find(out.begin(), out.end(), x) == out.end()) out.push_back(x);
        return out;
    }
    struct cb_state { transcript_type *tr; };
    static int on_root(void *user, std::uint32_t, const std::uint8_t *root, std::uint32_t root_bytes, std::uint32_t count,
                       std::uint32_t *alphas_out) {
        try {
            cb_state *s = static_cast<cb_state *>(user);
            (*s->tr)(std::vector<std::uint8_t>(root, root + root_bytes));
            for (std::uint32_t k = 0; k < count; k++) limbs(s->tr->template challenge<FieldType>(), alphas_out + 8 * k);
            return 0;
        } catch (...) {
            return 1;
        }
    }
    static std::vector<std::pair<std::size_t, std::size_t>> s_indices(std::size_t x_index, std::size_t domain_size, std::size_t fri_step) {
        // index pairs of calculate_s (basic_fri.hpp:583-617), each ordered (min, max)
        const std::size_t half = domain_size / 2;
        std::vector<std::size_t> idx = {x_index};
        std::size_t base = domain_size / 4, prev = 1;
        while (idx.size() < (std::size_t(1) << fri_step) / 2) {
            for (std::size_t j = 0; j < prev; j++) idx.push_back((base + idx[j]) % domain_size);
            base /= 2;
            prev <<= 1;
        }
        std::vector<std::pair<std::size_t, std::size_t>> out;
        for (auto a : idx) {
            std::size_t b = (a + half) % domain_size;
            out.emplace_back(std::min(a, b), std::max(a, b));
        }
        return out;
    }
    static std::size_t folded_index(std::size_t x_index, std::size_t domain_size, std::size_t fri_step) {
        for (std::size_t i = 0; i < fri_step; i++) { domain_size /= 2; x_index %= domain_size; }
        return x_index;
    }
    // index of x in the 2^log_n subgroup (the reference searches linearly, basic_fri.hpp:780-786)
    static std::size_t domain_index(const value_type &x, std::size_t log_n) {
        const value_type w_inv = omega(log_n).inversed();
        std::size_t e = 0;
        for (std::size_t i = 0; i < log_n; i++) {
            value_type t = x * w_inv.pow(e);
            for (std::size_t k = 0; k + 1 + i < log_n; k++) t *= t;
            if (!t.is_one()) e |= std::size_t(1) << i;
        }
        return e;
    }
    static value_type horner(const std::vector<value_type> &c, const value_type &x) {
        value_type r = value_type::zero();
        for (std::size_t i = c.size(); i-- > 0;) r = r * x + c[i];
        return r;
    }
    static merkle_proof_type make_proof(std::size_t index, std::vector<commitment_type> path, const commitment_type &root) {
        merkle_proof_type p;
        p.index = index; p.path = std::move(path); p.root = root;
        return p;
    }

public:
    explicit lpc_commitment_scheme(const fri_params_type &fri_params) : _fri(fri_params) {
        if (_fri.r() > _fri.log_d0()) throw std::invalid_argument("lpc: sum(step_list) exceeds log2 |D[0]|");
    }
    const fri_params_type &get_commitment_params() const { return _fri; }

    // ---- polys_evaluator interface (batched_commitment.hpp:196-250)
    void append_to_batch(std::size_t index, const poly_type &poly) {
        if (_locked[index]) throw std::logic_error("lpc: batch is already committed");
        _polys[index].push_back(poly);
        _points[index].emplace_back();
    }
    template <class Container>
    void append_to_batch(std::size_t index, const Container &polys) { for (const auto &q : polys) append_to_batch(index, q); }
    // a whole batch that already lives on the device ([count][n] canonical limbs, e.g. columns an argument builder left
    // there): copied into the scheme's own buffer, never through the host
    void append_device_batch(std::size_t index, const void *device_polys, std::size_t count, std::size_t n) {
        if (_locked[index] || _polys.count(index)) throw std::logic_error("lpc: batch already exists");
        zkb_detail::log2_exact(n);
        dbuf d(count * n * 32);
        ck(zkb_buf_copy(ctx(), d.p, ZKB_MEM_DEVICE, device_polys, ZKB_MEM_DEVICE, count * n * 32, nullptr), "zkb_buf_copy");
        _dev[index] = std::move(d);
        _n[index] = n;
        _polys[index].assign(count, poly_type());      // placeholders: only their number is read once the batch is uploaded
        _points[index].assign(count, {});
    }
    bool has_batch(std::size_t index) const { return _polys.count(index) != 0; }
    void append_eval_point(std::size_t batch, const value_type &point) { for (auto &pts : _points.at(batch)) pts.push_back(point); }
    void append_eval_point(std::size_t batch, std::size_t poly, const value_type &point) { _points.at(batch).at(poly).push_back(point); }
    const std::map<std::size_t, std::vector<std::vector<value_type>>> &get_z() const { return _z; }

    // ---- lpc.hpp:101-111
    commitment_type commit(std::size_t index) {
        _locked[index] = true;
        upload(index);
        const std::size_t n = _n.at(index);
        std::vector<std::uint8_t> root(MerkleHash::digest_bytes);
        zkb_merkle_tree *t = nullptr;
        ck(zkb_lpc_commit(ctx(), fid, MerkleHash::hash_id, zkb_detail::log2_exact(n), (int)_fri.log_d0(), (int)_fri.step_list.front(),
                          (std::uint32_t)_polys.at(index).size(), _dev.at(index).p, ZKB_MEM_DEVICE, root.data(), &t, nullptr),
           "zkb_lpc_commit");
        _trees.erase(index);
        _trees.emplace(index, precommitment_type(t, root));
        return root;
    }
    void mark_batch_as_fixed(std::size_t index) { _fixed[index] = true; }
    // etha from the transcript; the values of the fixed batches at etha are computed here with the same device evaluation
    void setup(transcript_type &transcript) {
        _etha = transcript.template challenge<FieldType>();
        for (const auto &kv : _fixed) {
            if (!kv.second) continue;
            const std::size_t k = kv.first;
            upload(k);
            const std::size_t n = _n.at(k), cnt = _polys.at(k).size();
            std::uint32_t pt[8];
            limbs(_etha, pt);
            std::vector<std::uint32_t> out(cnt * 8);
            ck(zkb_poly_evaluate(ctx(), fid, ZKB_POLY_DFS, n, (std::uint32_t)cnt, _dev.at(k).p, ZKB_MEM_DEVICE, 1, pt, out.data(), nullptr),
               "zkb_poly_evaluate");
            std::vector<value_type> v(cnt);
            for (std::size_t i = 0; i < cnt; i++) v[i] = value_type::from_canonical_limbs(&out[8 * i]);
            _fixed_values[k] = v;
        }
    }
    const value_type &etha() const { return _etha; }
    const std::map<std::size_t, std::vector<value_type>> &fixed_values() const { return _fixed_values; }

    // ---- eval_polys (batched_commitment.hpp:176-190): one inverse transform per batch, all points of the batch at once
    void eval_polys() {
        _z.clear();
        for (const auto &kv : _polys) {
            const std::size_t k = kv.first, cnt = kv.second.size();
            upload(k);
            const std::size_t n = _n.at(k);
            dbuf co(cnt * n * 32);
            ck(zkb_ntt(ctx(), fid, zkb_detail::log2_exact(n), (std::uint32_t)cnt, _dev.at(k).p, co.p, 1, nullptr, ZKB_MEM_DEVICE, nullptr), "zkb_ntt");
            std::vector<value_type> uni;
            for (const auto &pts : _points.at(k))
                for (const auto &x : pts)
                    if (std::find(uni.begin(), uni.end(), x) == uni.end()) uni.push_back(x);
            _z[k].assign(cnt, {});
            if (!uni.empty()) {
                std::vector<std::uint32_t> pl(uni.size() * 8), out(cnt * uni.size() * 8);
                for (std::size_t j = 0; j < uni.size(); j++) limbs(uni[j], &pl[8 * j]);
                ck(zkb_poly_evaluate(ctx(), fid, ZKB_POLY_COEFFICIENTS, n, (std::uint32_t)cnt, co.p, ZKB_MEM_DEVICE, (std::uint32_t)uni.size(),
                                     pl.data(), out.data(), nullptr), "zkb_poly_evaluate");
                for (std::size_t i = 0; i < cnt; i++)
                    for (const auto &x : _points.at(k)[i]) {
                        std::size_t j = std::find(uni.begin(), uni.end(), x) - uni.begin();
                        _z[k][i].push_back(value_type::from_canonical_limbs(&out[(i * uni.size() + j) * 8]));
                    }
            }
            _coef[k] = std::move(co);
        }
    }

    // ---- proof_eval (lpc.hpp:113-200 -> basic_fri.hpp:670-923)
    proof_type proof_eval(transcript_type &transcript) {
        eval_polys();
        for (const auto &kv : _trees) transcript(kv.second.root());
        const value_type theta = transcript.template challenge<FieldType>();
        value_type theta_acc = value_type::one();
        std::size_t n_max = 0;
        for (const auto &kv : _n) n_max = std::max(n_max, kv.second);
        dbuf combined(n_max * 32), numer(n_max * 32), quot(n_max * 32);
        ck(zkb_buf_zero(ctx(), combined.p, n_max * 32, nullptr), "zkb_buf_zero");
        struct term { std::size_t k; std::vector<std::uint32_t> scalars; value_type constant; };
        auto add_quotient = [&](const value_type &point, const std::vector<term> &terms) {
            ck(zkb_buf_zero(ctx(), numer.p, n_max * 32, nullptr), "zkb_buf_zero");
            for (const auto &t : terms) {
                std::uint32_t c[8];
                limbs(t.constant, c);
                ck(zkb_poly_lincomb(ctx(), fid, _n.at(t.k), (std::uint32_t)_polys.at(t.k).size(), _coef.at(t.k).p, t.scalars.data(), c,
                                    numer.p, 1, nullptr), "zkb_poly_lincomb");
            }
            std::uint32_t pt[8], rem[8];
            limbs(point, pt);
            ck(zkb_poly_div_linear(ctx(), fid, n_max, numer.p, pt, quot.p, rem, nullptr), "zkb_poly_div_linear");
            ck(zkb_vec(ctx(), fid, ZKB_VEC_ADD, n_max, combined.p, quot.p, nullptr, nullptr, combined.p, ZKB_MEM_DEVICE, nullptr), "zkb_vec");
        };
        for (const auto &point : unique_points()) {
            std::vector<term> terms;
            for (const auto &kv : _polys) {
                const std::size_t k = kv.first;
                term t;
                t.k = k;
                t.scalars.assign(kv.second.size() * 8, 0);
                t.constant = value_type::zero();
                bool used = false;
                for (std::size_t i = 0; i < kv.second.size(); i++) {
                    const auto &pts = _points.at(k)[i];
                    auto it = std::find(pts.begin(), pts.end(), point);
                    if (it == pts.end()) continue;
                    limbs(theta_acc, &t.scalars[8 * i]);
                    t.constant += _z.at(k)[i][it - pts.begin()] * theta_acc;
                    theta_acc *= theta;
                    used = true;
                }
                if (used) terms.push_back(std::move(t));
            }
            add_quotient(point, terms);
        }
        for (const auto &kv : _polys) {
            const std::size_t k = kv.first;
            if (!_fixed.count(k) || !_fixed.at(k)) continue;
            term t;
            t.k = k;
            t.scalars.assign(kv.second.size() * 8, 0);
            t.constant = value_type::zero();
            for (std::size_t i = 0; i < kv.second.size(); i++) {
                limbs(theta_acc, &t.scalars[8 * i]);
                t.constant += _fixed_values.at(k)[i] * theta_acc;
                theta_acc *= theta;
            }
            add_quotient(_etha, {t});
        }
        // combined_Q.from_coefficients + precommit's resize to D[0]: one forward transform of the zero-padded coefficients
        const std::size_t log_d0 = _fri.log_d0(), d0 = std::size_t(1) << log_d0;
        dbuf q_d0(d0 * 32);
        ck(zkb_buf_zero(ctx(), q_d0.p, d0 * 32, nullptr), "zkb_buf_zero");
        ck(zkb_buf_copy(ctx(), q_d0.p, ZKB_MEM_DEVICE, combined.p, ZKB_MEM_DEVICE, n_max * 32, nullptr), "zkb_buf_copy");
        ck(zkb_ntt(ctx(), fid, (int)log_d0, 1, q_d0.p, q_d0.p, 0, nullptr, ZKB_MEM_DEVICE, nullptr), "zkb_ntt");
        // ---- commit phase (basic_fri.hpp:706-742)
        const auto &steps = _fri.step_list;
        const std::size_t rounds = steps.size(), total = _fri.r(), db = MerkleHash::digest_bytes;
        std::vector<std::uint32_t> step32(steps.begin(), steps.end()), alphas(total * 8), final_l((std::size_t(1) << (log_d0 - total)) * 8);
        std::vector<std::uint8_t> roots(rounds * db);
        std::vector<zkb_merkle_tree *> trees(rounds, nullptr);
        std::vector<std::pair<std::size_t, std::size_t>> fs_off;   // (offset, log size) of fs[i+1]
        std::size_t acc = log_d0, off = 0;
        for (auto s_ : steps) { acc -= s_; fs_off.emplace_back(off, acc); off += std::size_t(1) << acc; }
        dbuf fs(off * 32);
        cb_state cbs{&transcript};
        ck(zkb_fri_commit_phase(ctx(), fid, MerkleHash::hash_id, (int)log_d0, q_d0.p, ZKB_MEM_DEVICE, step32.data(), (std::uint32_t)rounds,
                                &on_root, &cbs, roots.data(), trees.data(), fs.p, alphas.data(), final_l.data(), nullptr),
           "zkb_fri_commit_phase");
        std::vector<device_merkle_tree<MerkleHash::hash_id>> fri_trees;
        proof_type proof;
        for (std::size_t i = 0; i < rounds; i++) {
            commitment_type r(roots.begin() + i * db, roots.begin() + (i + 1) * db);
            fri_trees.emplace_back(trees[i], r);
            proof.fri_proof.fri_roots.push_back(r);
        }
        for (std::size_t i = 0; i < final_l.size() / 8; i++) proof.fri_proof.final_polynomial.push_back(value_type::from_canonical_limbs(&final_l[8 * i]));
        proof.z = _z;
        // ---- grinding (basic_fri.hpp:744-747)
        if (_fri.use_grinding)
            proof.fri_proof.proof_of_work = proof_of_work<TranscriptHash>::generate(transcript, _fri.grinding_parameter);
        // ---- query phase (basic_fri.hpp:749-915): the challenges depend on nothing that is opened, so draw them all first
        if (steps.back() != 1) throw std::invalid_argument("lpc: step_list must end with 1 (check_step_list)");
        const std::size_t lam = _fri.lambda;
        std::vector<std::size_t> x_idx0(lam);
        std::uint32_t expo[8];
        FieldType::modulus_minus_one_shifted((int)log_d0, expo);
        for (std::size_t q = 0; q < lam; q++)
            x_idx0[q] = domain_index(transcript.template challenge<FieldType>().pow_limbs(expo, 8), log_d0);
        const value_type w0 = omega(log_d0);
        std::vector<std::vector<std::pair<std::size_t, std::size_t>>> init_pairs(lam);
        std::vector<std::size_t> flat, leaf0(lam);
        for (std::size_t q = 0; q < lam; q++) {
            init_pairs[q] = s_indices(x_idx0[q], d0, steps[0]);
            for (const auto &pr : init_pairs[q]) flat.push_back(pr.first);
            leaf0[q] = folded_index(x_idx0[q], d0, steps[0]);
        }
        std::sort(flat.begin(), flat.end());
        flat.erase(std::unique(flat.begin(), flat.end()), flat.end());
        std::vector<std::uint32_t> pts(flat.size() * 8);
        for (std::size_t j = 0; j < flat.size(); j++) limbs(w0.pow(flat[j]), &pts[8 * j]);
        proof.fri_proof.query_proofs.assign(lam, query_proof_type());
        for (const auto &kv : _polys) {
            const std::size_t k = kv.first, cnt = kv.second.size(), n = _n.at(k);
            std::vector<std::uint32_t> vals(cnt * flat.size() * 16);   // [poly][point][z, -z]
            if (n == d0) {   // already on D[0]: the values themselves (basic_fri.hpp:812-818)
                std::vector<std::uint64_t> idx;
                for (std::size_t i = 0; i < cnt; i++)
                    for (auto a : flat) { idx.push_back(i * n + a); idx.push_back(i * n + a + d0 / 2); }
                ck(zkb_gather(ctx(), _dev.at(k).p, (std::uint32_t)idx.size(), idx.data(), vals.data(), nullptr), "zkb_gather");
            } else {         // coefficient form at the 2 lambda points (:819-834), z and -z from one pass
                ck(zkb_poly_evaluate_pm(ctx(), fid, n, (std::uint32_t)cnt, _coef.at(k).p, (std::uint32_t)flat.size(), pts.data(), vals.data(), nullptr),
                   "zkb_poly_evaluate_pm");
            }
            std::vector<std::uint64_t> l64(leaf0.begin(), leaf0.end());
            auto paths = _trees.at(k).paths(l64);
            for (std::size_t q = 0; q < lam; q++) {
                initial_proof_type ip;
                ip.values.assign(cnt, {});
                for (std::size_t i = 0; i < cnt; i++)
                    for (const auto &pr : init_pairs[q]) {
                        std::size_t j = std::lower_bound(flat.begin(), flat.end(), pr.first) - flat.begin();
                        const std::uint32_t *v = &vals[(i * flat.size() + j) * 16];
                        ip.values[i].push_back({value_type::from_canonical_limbs(v), value_type::from_canonical_limbs(v + 8)});
                    }
                ip.p = make_proof(leaf0[q], paths[q], _trees.at(k).root());
                proof.fri_proof.query_proofs[q].initial_proof.emplace(k, std::move(ip));
            }
        }
        // round proofs: paths of every fri tree, y from the retained fs (one gather), last round from final_polynomial
        struct slot { std::size_t q, round, j, side; };
        std::vector<std::uint64_t> gidx;
        std::vector<slot> gslot;
        std::vector<std::size_t> xs(x_idx0);
        std::size_t t = 0;
        for (std::size_t q = 0; q < lam; q++) proof.fri_proof.query_proofs[q].round_proofs.assign(rounds, round_proof_type());
        for (std::size_t i = 0; i < rounds; i++) {
            const std::size_t size_t_ = std::size_t(1) << (log_d0 - t);
            std::vector<std::uint64_t> leaves(lam);
            for (std::size_t q = 0; q < lam; q++) { xs[q] %= size_t_; leaves[q] = folded_index(xs[q], size_t_, steps[i]); }
            auto paths = fri_trees[i].paths(leaves);
            t += steps[i];
            const std::size_t size_n = std::size_t(1) << (log_d0 - t);
            for (std::size_t q = 0; q < lam; q++) {
                round_proof_type &rp = proof.fri_proof.query_proofs[q].round_proofs[i];
                rp.p = make_proof((std::size_t)leaves[q], paths[q], proof.fri_proof.fri_roots[i]);
                if (i + 1 < rounds) {
                    auto pairs = s_indices(xs[q] % size_n, size_n, steps[i + 1]);
                    rp.y.assign(pairs.size(), {value_type::zero(), value_type::zero()});
                    for (std::size_t j = 0; j < pairs.size(); j++) {
                        gidx.push_back(fs_off[i].first + pairs[j].first);  gslot.push_back({q, i, j, 0});
                        gidx.push_back(fs_off[i].first + pairs[j].second); gslot.push_back({q, i, j, 1});
                    }
                } else {
                    const std::size_t size_p = std::size_t(1) << (log_d0 - t + 1);   // D[t-1]
                    const std::size_t xq = xs[q] % size_p;
                    value_type x = omega(log_d0 - t + 1).pow(xq);
                    x = x * x;
                    const std::size_t ind = (xq % (size_p / 2) < size_p / 4) ? 0 : 1;
                    rp.y.assign(1, {value_type::zero(), value_type::zero()});
                    rp.y[0][ind] = horner(proof.fri_proof.final_polynomial, x);
                    rp.y[0][1 - ind] = horner(proof.fri_proof.final_polynomial, -x);
                }
            }
            if (i + 1 < rounds)
                for (std::size_t q = 0; q < lam; q++) xs[q] %= size_n;
        }
        if (!gidx.empty()) {
            std::vector<std::uint32_t> g(gidx.size() * 8);
            ck(zkb_gather(ctx(), fs.p, (std::uint32_t)gidx.size(), gidx.data(), g.data(), nullptr), "zkb_gather");
            for (std::size_t n_ = 0; n_ < gslot.size(); n_++)
                proof.fri_proof.query_proofs[gslot[n_].q].round_proofs[gslot[n_].round].y[gslot[n_].j][gslot[n_].side] =
                    value_type::from_canonical_limbs(&g[8 * n_]);
        }
        return proof;
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

// ---- Groth16 wire format, BLS12-381 (r1cs_gg_ppzksnark/marshalling.hpp; host only, no device call) ----
// `curve_element_serializer<bls12<381>>` is crypto3-algebra's (un-vendored): the ZCash encoding - x big-endian
// (G2: x.c1 then x.c0), bit 7 of byte 0 = compressed, bit 6 = infinity, bit 5 = y is the larger of (y, -y)
// (Fp2: by c1, by c0 when c1 == 0).  The proof is g_A | g_B | g_C (marshalling.hpp:784-828, :1236-1256).
// crypto3_zk_b200/marshalling.py is the same format for the Python layer (keys and verifier input included);
// tests/test_cpp_host.py compares the two byte for byte.
namespace nil {
namespace crypto3 {
namespace zk {
namespace snark {
// proof.hpp:42-95
template <class CurveType>
struct r1cs_gg_ppzksnark_proof {
    typename CurveType::template g1_type<>::value_type g_A;
    typename CurveType::template g2_type<>::value_type g_B;
    typename CurveType::template g1_type<>::value_type g_C;
    r1cs_gg_ppzksnark_proof()
        : g_A(CurveType::template g1_type<>::value_type::one()), g_B(CurveType::template g2_type<>::value_type::one()),
          g_C(CurveType::template g1_type<>::value_type::one()) {}
    r1cs_gg_ppzksnark_proof(const typename CurveType::template g1_type<>::value_type &a,
                            const typename CurveType::template g2_type<>::value_type &b,
                            const typename CurveType::template g1_type<>::value_type &c)
        : g_A(a), g_B(b), g_C(c) {}
    bool operator==(const r1cs_gg_ppzksnark_proof &o) const { return g_A == o.g_A && g_B == o.g_B && g_C == o.g_C; }
};
template <class CurveType>
struct r1cs_gg_ppzksnark {
    typedef CurveType curve_type;
    typedef r1cs_gg_ppzksnark_proof<CurveType> proof_type;
    typedef std::vector<typename CurveType::scalar_field_type::value_type> primary_input_type;
};
}  // namespace snark
}  // namespace zk
}  // namespace crypto3

namespace marshalling {
enum class status_type { success, not_enough_data, invalid_msg_data };

template <class CurveType>
struct curve_element_serializer;

template <>
struct curve_element_serializer<crypto3::algebra::curves::bls12<381>> {
    typedef crypto3::algebra::curves::bls12<381> curve_type;
    typedef curve_type::base_field_type fq_type;
    typedef crypto3::algebra::fields::zkb_field2<fq_type> fq2_type;
    typedef curve_type::g1_type<>::value_type g1_value_type;
    typedef curve_type::g2_type<>::value_type g2_value_type;
    static constexpr std::size_t sizeof_field_element = 48;
    typedef std::array<std::uint8_t, sizeof_field_element> compressed_g1_octets;
    typedef std::array<std::uint8_t, 2 * sizeof_field_element> compressed_g2_octets;
    static constexpr std::uint8_t C_bit = 0x80, I_bit = 0x40, S_bit = 0x20;

    static compressed_g1_octets point_to_octets_compress(const g1_value_type &point) {
        compressed_g1_octets out{};
        if (point.is_zero()) {
            out[0] = C_bit | I_bit;
            return out;
        }
        g1_value_type a = point.to_affine();
        std::uint32_t x[12], y[12];
        a.X.to_canonical_limbs(x);
        a.Y.to_canonical_limbs(y);
        put_be(x, out.data());
        out[0] |= C_bit | (sign(y) ? S_bit : 0);
        return out;
    }
    static compressed_g2_octets point_to_octets_compress(const g2_value_type &point) {
        compressed_g2_octets out{};
        if (point.is_zero()) {
            out[0] = C_bit | I_bit;
            return out;
        }
        g2_value_type a = point.to_affine();
        std::uint32_t x[24], y[24];
        a.X.to_canonical_limbs(x);
        a.Y.to_canonical_limbs(y);
        put_be(x + 12, out.data());
        put_be(x, out.data() + sizeof_field_element);
        out[0] |= C_bit | (sign2(y) ? S_bit : 0);
        return out;
    }
    // throws std::invalid_argument for octets that are not the compressed form of a curve point
    static g1_value_type octets_to_g1_point(const compressed_g1_octets &in) {
        std::uint8_t flags = in[0];
        if (!(flags & C_bit)) throw std::invalid_argument("G1 point is not in compressed form");
        std::uint32_t xl[12];
        get_be(in.data(), xl);
        xl[11] &= 0x1FFFFFFFu;
        if (flags & I_bit) {
            if (!all_zero(xl, 12) || (flags & S_bit)) throw std::invalid_argument("non-zero payload in a point at infinity");
            return g1_value_type::zero();
        }
        if (!reduced(xl)) throw std::invalid_argument("x coordinate is not reduced");
        typedef fq_type::value_type V;
        V x = V::from_canonical_limbs(xl), rhs = x * x * x + V(4u);
        std::uint32_t e[12];
        fq_type::modulus_minus_one_shifted(2, e);   // (p - 3) / 4
        add_one(e);                                 // (p + 1) / 4
        V y = rhs.pow_limbs(e, 12);
        if (y * y != rhs) throw std::invalid_argument("x is not the abscissa of a curve point");
        std::uint32_t yl[12];
        y.to_canonical_limbs(yl);
        if (sign(yl) != bool(flags & S_bit)) y = -y;
        return g1_value_type::from_affine(x, y);
    }
    static g2_value_type octets_to_g2_point(const compressed_g2_octets &in) {
        std::uint8_t flags = in[0];
        if (!(flags & C_bit)) throw std::invalid_argument("G2 point is not in compressed form");
        std::uint32_t xl[24];
        get_be(in.data(), xl + 12);
        get_be(in.data() + sizeof_field_element, xl);
        xl[23] &= 0x1FFFFFFFu;
        if (flags & I_bit) {
            if (!all_zero(xl, 24) || (flags & S_bit)) throw std::invalid_argument("non-zero payload in a point at infinity");
            return g2_value_type::zero();
        }
        if (!reduced(xl) || !reduced(xl + 12)) throw std::invalid_argument("x coordinate is not reduced");
        typedef fq2_type::value_type V2;
        std::uint32_t bl[24] = {0};
        bl[0] = 4; bl[12] = 4;
        V2 x = V2::from_canonical_limbs(xl), rhs = x * x * x + V2::from_canonical_limbs(bl), y;
        if (!sqrt2(rhs, y)) throw std::invalid_argument("x is not the abscissa of a curve point");
        std::uint32_t yl[24];
        y.to_canonical_limbs(yl);
        if (sign2(yl) != bool(flags & S_bit)) y = -y;
        return g2_value_type::from_affine(x, y);
    }

private:
    static void put_be(const std::uint32_t *l, std::uint8_t *out) {
        for (int i = 0; i < 48; i++) out[i] = (std::uint8_t)(l[(47 - i) / 4] >> (8 * ((47 - i) % 4)));
    }
    static void get_be(const std::uint8_t *in, std::uint32_t *l) {
        for (int i = 0; i < 12; i++) l[i] = 0;
        for (int i = 0; i < 48; i++) l[(47 - i) / 4] |= (std::uint32_t)in[i] << (8 * ((47 - i) % 4));
    }
    static bool all_zero(const std::uint32_t *l, int n) {
        for (int i = 0; i < n; i++)
            if (l[i]) return false;
        return true;
    }
    static void add_one(std::uint32_t *l) {
        for (int i = 0; i < 12 && ++l[i] == 0; i++) {}
    }
    // a < p
    static bool reduced(const std::uint32_t *a) {
        std::uint32_t m[12];
        fq_type::modulus_minus_one_shifted(0, m);   // p - 1
        for (int i = 11; i >= 0; i--)
            if (a[i] != m[i]) return a[i] < m[i];
        return true;                                // a == p - 1
    }
    // y > (p - 1) / 2
    static bool sign(const std::uint32_t *y) {
        std::uint32_t h[12];
        fq_type::modulus_minus_one_shifted(1, h);
        for (int i = 11; i >= 0; i--)
            if (y[i] != h[i]) return y[i] > h[i];
        return false;
    }
    static bool sign2(const std::uint32_t *y) { return all_zero(y + 12, 12) ? sign(y) : sign(y + 12); }
    static fq2_type::value_type pow2(fq2_type::value_type b, const std::uint32_t *e) {
        fq2_type::value_type r = fq2_type::value_type::one();
        for (int i = 11; i >= 0; i--)
            for (int k = 31; k >= 0; k--) {
                r = r * r;
                if ((e[i] >> k) & 1) r = r * b;
            }
        return r;
    }
    // square root in Fq2 for p = 3 mod 4
    static bool sqrt2(const fq2_type::value_type &a, fq2_type::value_type &out) {
        typedef fq2_type::value_type V2;
        if (a.is_zero()) {
            out = a;
            return true;
        }
        std::uint32_t e[12], ul[24] = {0};
        fq_type::modulus_minus_one_shifted(2, e);   // (p - 3) / 4
        V2 a1 = pow2(a, e), alpha = a1 * a1 * a, x0 = a1 * a, x;
        if (alpha == -V2::one()) {
            ul[12] = 1;
            x = V2::from_canonical_limbs(ul) * x0;
        } else {
            fq_type::modulus_minus_one_shifted(1, e);   // (p - 1) / 2
            x = pow2(V2::one() + alpha, e) * x0;
        }
        if (x * x != a) return false;
        out = x;
        return true;
    }
};

template <class ProofSystem>
struct verifier_input_serializer_tvm;
template <class ProofSystem>
struct verifier_input_deserializer_tvm;

template <>
struct verifier_input_serializer_tvm<crypto3::zk::snark::r1cs_gg_ppzksnark<crypto3::algebra::curves::bls12<381>>> {
    typedef crypto3::algebra::curves::bls12<381> CurveType;
    typedef crypto3::zk::snark::r1cs_gg_ppzksnark<CurveType> scheme_type;
    typedef std::uint8_t chunk_type;
    static constexpr std::size_t g1_byteblob_size = curve_element_serializer<CurveType>::sizeof_field_element;
    static constexpr std::size_t g2_byteblob_size = 2 * curve_element_serializer<CurveType>::sizeof_field_element;
    static constexpr std::size_t std_size_t_byteblob_size = 4;
    static constexpr std::size_t fr_byteblob_size = (CurveType::scalar_field_type::modulus_bits + 7) / 8;
    // marshalling.hpp:975-985: 4 bytes, most significant first
    static void std_size_t_process(std::size_t v, std::vector<chunk_type> &out) {
        for (int i = 3; i >= 0; i--) out.push_back((chunk_type)(v >> (8 * i)));
    }
    // marshalling.hpp:921-935: fixed width, least significant byte first
    static void field_type_process(const CurveType::scalar_field_type::value_type &v, std::vector<chunk_type> &out) {
        std::uint32_t l[CurveType::scalar_field_type::limbs32];
        v.to_canonical_limbs(l);
        for (std::size_t i = 0; i < fr_byteblob_size; i++) out.push_back((chunk_type)(l[i / 4] >> (8 * (i % 4))));
    }
    // marshalling.hpp:1210-1234
    static std::vector<chunk_type> process(const scheme_type::primary_input_type &pi) {
        std::vector<chunk_type> out;
        out.reserve(std_size_t_byteblob_size + pi.size() * fr_byteblob_size);
        std_size_t_process(pi.size(), out);
        for (const auto &v : pi) field_type_process(v, out);
        return out;
    }
    // marshalling.hpp:1236-1256
    static std::vector<chunk_type> process(const scheme_type::proof_type &pr) {
        std::vector<chunk_type> out;
        out.reserve(2 * g1_byteblob_size + g2_byteblob_size);
        auto a = curve_element_serializer<CurveType>::point_to_octets_compress(pr.g_A);
        auto b = curve_element_serializer<CurveType>::point_to_octets_compress(pr.g_B);
        auto c = curve_element_serializer<CurveType>::point_to_octets_compress(pr.g_C);
        out.insert(out.end(), a.begin(), a.end());
        out.insert(out.end(), b.begin(), b.end());
        out.insert(out.end(), c.begin(), c.end());
        return out;
    }
};

template <>
struct verifier_input_deserializer_tvm<crypto3::zk::snark::r1cs_gg_ppzksnark<crypto3::algebra::curves::bls12<381>>> {
    typedef crypto3::algebra::curves::bls12<381> CurveType;
    typedef crypto3::zk::snark::r1cs_gg_ppzksnark<CurveType> scheme_type;
    typedef std::uint8_t chunk_type;
    static constexpr std::size_t g1_byteblob_size = curve_element_serializer<CurveType>::sizeof_field_element;
    static constexpr std::size_t g2_byteblob_size = 2 * curve_element_serializer<CurveType>::sizeof_field_element;
    static constexpr std::size_t std_size_t_byteblob_size = 4;
    static constexpr std::size_t fr_byteblob_size = (CurveType::scalar_field_type::modulus_bits + 7) / 8;
    // marshalling.hpp:465-491
    static std::size_t std_size_t_process(std::vector<chunk_type>::const_iterator read_iter_begin,
                                          std::vector<chunk_type>::const_iterator read_iter_end, status_type &processingStatus) {
        processingStatus = status_type::success;
        if ((std::size_t)std::distance(read_iter_begin, read_iter_end) < std_size_t_byteblob_size) {
            processingStatus = status_type::not_enough_data;
            return 0;
        }
        std::size_t v = 0;
        for (std::size_t i = 0; i < std_size_t_byteblob_size; i++) v = (v << 8) | read_iter_begin[i];
        return v;
    }
    // marshalling.hpp:122-144; a value that is not reduced gives invalid_msg_data
    static CurveType::scalar_field_type::value_type field_type_process(std::vector<chunk_type>::const_iterator read_iter_begin,
                                                                       std::vector<chunk_type>::const_iterator read_iter_end,
                                                                       status_type &processingStatus) {
        typedef CurveType::scalar_field_type F;
        processingStatus = status_type::success;
        if ((std::size_t)std::distance(read_iter_begin, read_iter_end) < fr_byteblob_size) {
            processingStatus = status_type::not_enough_data;
            return F::value_type::zero();
        }
        std::uint32_t l[F::limbs32] = {0}, m[F::limbs32];
        for (std::size_t i = 0; i < fr_byteblob_size; i++) l[i / 4] |= (std::uint32_t)read_iter_begin[i] << (8 * (i % 4));
        F::modulus_minus_one_shifted(0, m);   // p - 1
        for (int i = F::limbs32 - 1; i >= 0; i--)
            if (l[i] != m[i]) {
                if (l[i] > m[i]) {
                    processingStatus = status_type::invalid_msg_data;
                    return F::value_type::zero();
                }
                break;
            }
        return F::value_type::from_canonical_limbs(l);
    }
    // marshalling.hpp:740-782
    static scheme_type::primary_input_type primary_input_process(std::vector<chunk_type>::const_iterator read_iter_begin,
                                                                 std::vector<chunk_type>::const_iterator read_iter_end,
                                                                 status_type &processingStatus) {
        std::size_t pi_count = std_size_t_process(read_iter_begin, read_iter_end, processingStatus);
        if (processingStatus != status_type::success) return {};
        if ((std::size_t)std::distance(read_iter_begin, read_iter_end) < std_size_t_byteblob_size + pi_count * fr_byteblob_size) {
            processingStatus = status_type::not_enough_data;
            return {};
        }
        scheme_type::primary_input_type pi(pi_count);
        for (std::size_t i = 0; i < pi_count; i++) {
            pi[i] = field_type_process(read_iter_begin + std_size_t_byteblob_size + i * fr_byteblob_size,
                                       read_iter_begin + std_size_t_byteblob_size + (i + 1) * fr_byteblob_size, processingStatus);
            if (processingStatus != status_type::success) return {};
        }
        return pi;
    }
    // marshalling.hpp:784-828; octets that are no curve point give invalid_msg_data
    static scheme_type::proof_type proof_process(std::vector<chunk_type>::const_iterator read_iter_begin,
                                                 std::vector<chunk_type>::const_iterator read_iter_end,
                                                 status_type &processingStatus) {
        if ((std::size_t)std::distance(read_iter_begin, read_iter_end) < 2 * g1_byteblob_size + g2_byteblob_size) {
            processingStatus = status_type::not_enough_data;
            return {};
        }
        curve_element_serializer<CurveType>::compressed_g1_octets a, c;
        curve_element_serializer<CurveType>::compressed_g2_octets b;
        std::copy(read_iter_begin, read_iter_begin + g1_byteblob_size, a.begin());
        std::copy(read_iter_begin + g1_byteblob_size, read_iter_begin + g1_byteblob_size + g2_byteblob_size, b.begin());
        std::copy(read_iter_begin + g1_byteblob_size + g2_byteblob_size,
                  read_iter_begin + 2 * g1_byteblob_size + g2_byteblob_size, c.begin());
        try {
            scheme_type::proof_type pr(curve_element_serializer<CurveType>::octets_to_g1_point(a),
                                       curve_element_serializer<CurveType>::octets_to_g2_point(b),
                                       curve_element_serializer<CurveType>::octets_to_g1_point(c));
            processingStatus = status_type::success;
            return pr;
        } catch (const std::invalid_argument &) {
            processingStatus = status_type::invalid_msg_data;
            return {};
        }
    }
};
}  // namespace marshalling
}  // namespace nil
