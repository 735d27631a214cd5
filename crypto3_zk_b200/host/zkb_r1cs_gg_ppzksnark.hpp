// zkb_r1cs_gg_ppzksnark.hpp - host templates of the Groth16 proving path over the C ABI (include/zkb200.h).
//
// Re-creates, with the reference's names, argument meaning and order of operations:
//   math::linear_variable / linear_combination            (crypto3-math, as used by r1cs.hpp:60-63)
//   zk::snark::r1cs_constraint, r1cs_constraint_system    zk/snark/arithmetization/constraint_satisfaction_problems/r1cs.hpp:60-230
//   zk::snark::qap_witness                                zk/snark/arithmetization/arithmetic_programs/qap.hpp
//   zk::snark::reductions::r1cs_to_qap<F>::witness_map    zk/snark/reductions/r1cs_to_qap.hpp:219-325
//   zk::snark::r1cs_gg_ppzksnark_proving_key              zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/proving_key.hpp:38-120
//   zk::snark::r1cs_gg_ppzksnark_prover<Curve>::process   zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/prover.hpp:73-158
//
// Where the work runs: the rows <a_i, x>, <b_i, x>, <c_i, x> (zkb_sparse_matvec), the seven transforms (zkb_ntt), the
// pointwise passes (zkb_vec) and the five multiexps (zkb_msm) are device calls; vectors stay in device buffers between
// them (the polynomial H never visits the host inside process()).  The key's query vectors and the three CSR matrices
// of the constraint system are uploaded once per key (device_cache, built on first use) - a proving key is long-lived.
// The host keeps what the reference keeps on the host: the proof assembly prover.hpp:141-157 (a handful of group
// operations) and the d1/d2/d3 patch of witness_map (zero in the prover).
//
// Field elements cross the boundary as raw Montgomery limbs and are rescaled on the device (x R^-1 / x R are linear
// maps: one zkb_poly_lincomb pass), so a 2^22-element assignment costs a memcpy on the host, not 4 M conversions.
#ifndef ZKB_R1CS_GG_PPZKSNARK_HPP
#define ZKB_R1CS_GG_PPZKSNARK_HPP

#include <random>

#include "zkb_crypto3.hpp"

namespace nil {
namespace crypto3 {

// ========================================================================================== math
namespace math {
// linear_variable<F>: index 0 is the constant 1, index i >= 1 the i-th variable (r1cs.hpp:119-124)
template <class FieldType>
struct linear_variable {
    typedef FieldType field_type;
    std::size_t index = 0;
    linear_variable() = default;
    linear_variable(std::size_t i) : index(i) {}
};
template <class VariableType>
struct linear_term {
    typedef typename VariableType::field_type::value_type value_type;
    std::size_t index = 0;
    value_type coeff;
    linear_term() : coeff(value_type::one()) {}
    linear_term(const VariableType &v) : index(v.index), coeff(value_type::one()) {}
    linear_term(const VariableType &v, const value_type &c) : index(v.index), coeff(c) {}
    linear_term(std::size_t i, const value_type &c) : index(i), coeff(c) {}
};
template <class VariableType>
struct linear_combination {
    typedef typename VariableType::field_type::value_type value_type;
    std::vector<linear_term<VariableType>> terms;
    linear_combination() = default;
    linear_combination(const value_type &c) { terms.emplace_back(0, c); }
    linear_combination(const VariableType &v) { terms.emplace_back(v); }
    linear_combination(const linear_term<VariableType> &t) { terms.push_back(t); }
    void add_term(std::size_t index, const value_type &c) { terms.emplace_back(index, c); }
    // sum coeff * (index == 0 ? 1 : assignment[index - 1])
    value_type evaluate(const std::vector<value_type> &assignment) const {
        value_type acc = value_type::zero();
        for (const auto &t : terms) acc += t.index == 0 ? t.coeff : t.coeff * assignment[t.index - 1];
        return acc;
    }
    bool is_valid(std::size_t num_variables) const {
        for (const auto &t : terms)
            if (t.index > num_variables) return false;
        return true;
    }
};
}  // namespace math

namespace zk {
namespace snark {

template <class FieldType>
using r1cs_primary_input = std::vector<typename FieldType::value_type>;
template <class FieldType>
using r1cs_auxiliary_input = std::vector<typename FieldType::value_type>;
template <class FieldType>
using r1cs_variable_assignment = std::vector<typename FieldType::value_type>;

template <class FieldType, class VariableType = math::linear_variable<FieldType>>
struct r1cs_constraint {
    typedef FieldType field_type;
    math::linear_combination<VariableType> a, b, c;
    r1cs_constraint() = default;
    r1cs_constraint(const math::linear_combination<VariableType> &a_, const math::linear_combination<VariableType> &b_,
                    const math::linear_combination<VariableType> &c_) : a(a_), b(b_), c(c_) {}
};

namespace zkb_detail_r1cs {
// a device buffer of field elements (RAII over zkb_buf_alloc)
struct dev_elems {
    void *p = nullptr;
    std::size_t n = 0;
    dev_elems() = default;
    explicit dev_elems(std::size_t count) : n(count) {
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_buf_alloc(ctx, (std::uint64_t)count * 32, &p), ctx, "zkb_buf_alloc");
    }
    dev_elems(const dev_elems &) = delete;
    dev_elems &operator=(const dev_elems &) = delete;
    dev_elems(dev_elems &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; }
    dev_elems &operator=(dev_elems &&o) noexcept {
        if (this != &o) {
            if (p) zkb_buf_free(zkb_detail::context(), p);
            p = o.p; n = o.n;
            o.p = nullptr;
        }
        return *this;
    }
    ~dev_elems() { if (p) zkb_buf_free(zkb_detail::context(), p); }
    char *at(std::size_t i) const { return (char *)p + i * 32; }
};
// the three sides of the constraint system as device CSR matrices (uploaded once)
struct device_matrices {
    zkb_sparse_matrix *m[3] = {nullptr, nullptr, nullptr};
    ~device_matrices() { for (auto *x : m) if (x) zkb_sparse_matrix_free(x); }
};
}  // namespace zkb_detail_r1cs

template <class FieldType>
struct r1cs_constraint_system {
    typedef FieldType field_type;
    std::size_t primary_input_size = 0;
    std::size_t auxiliary_input_size = 0;
    std::vector<r1cs_constraint<FieldType>> constraints;

    std::size_t num_inputs() const { return primary_input_size; }
    std::size_t num_variables() const { return primary_input_size + auxiliary_input_size; }
    std::size_t num_constraints() const { return constraints.size(); }
    bool is_valid() const {
        if (num_inputs() > num_variables()) return false;
        for (const auto &k : constraints)
            if (!(k.a.is_valid(num_variables()) && k.b.is_valid(num_variables()) && k.c.is_valid(num_variables()))) return false;
        return true;
    }
    // r1cs.hpp:164-189 (host loop, like the reference; the device flow re-checks the same products for free, see
    // r1cs_to_qap::witness_map_device)
    bool is_satisfied(const r1cs_primary_input<FieldType> &primary_input, const r1cs_auxiliary_input<FieldType> &auxiliary_input) const {
        if (primary_input.size() != num_inputs() || primary_input.size() + auxiliary_input.size() != num_variables()) return false;
        r1cs_variable_assignment<FieldType> full = primary_input;
        full.insert(full.end(), auxiliary_input.begin(), auxiliary_input.end());
        for (const auto &k : constraints)
            if (k.a.evaluate(full) * k.b.evaluate(full) != k.c.evaluate(full)) return false;
        return true;
    }
    void add_constraint(const r1cs_constraint<FieldType> &c) {
        constraints.emplace_back(c);
        _dev.reset();
    }

    // device CSR matrices of the a / b / c sides, built on first use; add_constraint() drops them (code that edits
    // `constraints` directly calls invalidate_device_cache())
    void invalidate_device_cache() const { _dev.reset(); }
    const zkb_detail_r1cs::device_matrices &device() const {
        if (_dev) return *_dev;
        auto d = std::make_shared<zkb_detail_r1cs::device_matrices>();
        zkb_ctx *ctx = zkb_detail::context();
        const std::size_t nc = constraints.size();
        for (int side = 0; side < 3; side++) {
            std::vector<std::uint64_t> row_ptr(nc + 1, 0);
            std::vector<std::uint32_t> col, val;
            for (std::size_t i = 0; i < nc; i++) {
                const auto &lc = side == 0 ? constraints[i].a : side == 1 ? constraints[i].b : constraints[i].c;
                for (const auto &t : lc.terms) {
                    col.push_back((std::uint32_t)t.index);
                    val.resize(val.size() + 8);
                    t.coeff.to_canonical_limbs(&val[val.size() - 8]);
                }
                row_ptr[i + 1] = col.size();
            }
            zkb_detail::check(zkb_sparse_matrix_create(ctx, FieldType::field_id, nc, num_variables() + 1, row_ptr.data(),
                                                       col.empty() ? nullptr : col.data(), val.empty() ? nullptr : val.data(), nullptr,
                                                       &d->m[side]), ctx, "zkb_sparse_matrix_create");
        }
        _dev = d;
        return *_dev;
    }

private:
    mutable std::shared_ptr<zkb_detail_r1cs::device_matrices> _dev;
};

// qap.hpp: what witness_map returns
template <class FieldType>
struct qap_witness {
    std::size_t num_variables = 0, degree = 0, num_inputs = 0;
    typename FieldType::value_type d1, d2, d3;
    std::vector<typename FieldType::value_type> coefficients_for_ABCs;   // the full variable assignment
    std::vector<typename FieldType::value_type> coefficients_for_H;      // degree + 1 entries
};

namespace reductions {
template <class FieldType>
struct r1cs_to_qap {
    typedef typename FieldType::value_type value_type;

    // R^-1 and R (canonical limbs): the device rescales raw Montgomery limbs with one linear pass
    static void mont_scalars(std::uint32_t *r_inv, std::uint32_t *r) {
        value_type t;                         // data = 1 read as a Montgomery representation: the element R^-1
        std::uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
        t.data = FieldType::backend::from_limbs32(one);
        t.to_canonical_limbs(r_inv);
        value_type::one().data.to_limbs32(r);   // Montgomery one = R mod p
    }
    static void scale_on_device(const void *d_in, void *d_out, std::size_t n, const std::uint32_t *scalar) {
        zkb_ctx *ctx = zkb_detail::context();
        zkb_detail::check(zkb_poly_lincomb(ctx, FieldType::field_id, n, 1, d_in, scalar, nullptr, d_out, 0, nullptr), ctx, "zkb_poly_lincomb");
    }

    // The device part shared by witness_map and the prover: x = (1, primary, auxiliary) -> coefficients of H on the
    // device (m elements, canonical limbs).  Order of operations = r1cs_to_qap.hpp:229-315 with d1 = d2 = d3 = 0.
    // x_canonical_out (nv + 1 elements, device) receives the assignment in canonical limbs (the multiexp scalars).
    // ab_coeff_host: when non-null, receives the coefficient forms of A and B (2 m host elements) for the d-patch.
    static zkb_detail_r1cs::dev_elems witness_map_device(const r1cs_constraint_system<FieldType> &cs,
                                                         const r1cs_primary_input<FieldType> &primary_input,
                                                         const r1cs_auxiliary_input<FieldType> &auxiliary_input,
                                                         zkb_detail_r1cs::dev_elems &x_canonical_out, std::size_t &m_out,
                                                         std::vector<value_type> *ab_coeff_host, bool check_satisfied) {
        using zkb_detail_r1cs::dev_elems;
        zkb_ctx *ctx = zkb_detail::context();
        const std::size_t nc = cs.num_constraints(), ni = cs.num_inputs(), nv = cs.num_variables();
        if (primary_input.size() != ni || primary_input.size() + auxiliary_input.size() != nv)
            throw std::invalid_argument("witness_map: assignment size does not match the constraint system");
        auto domain = math::make_evaluation_domain<FieldType>(nc + ni + 1);   // power-of-two sizes only
        const std::size_t m = domain->m;
        if (m != nc + ni + 1) throw std::invalid_argument("witness_map: #constraints + #inputs + 1 must be a power of two (basic_radix2_domain)");
        const int log_m = zkb_detail::log2_exact(m);
        m_out = m;
        std::uint32_t r_inv[8], r_one[8];
        mont_scalars(r_inv, r_one);
        // full assignment with the constant 1 in front, raw Montgomery limbs -> device -> canonical
        std::vector<value_type> x(1, value_type::one());
        x.insert(x.end(), primary_input.begin(), primary_input.end());
        x.insert(x.end(), auxiliary_input.begin(), auxiliary_input.end());
        dev_elems xd(nv + 1);
        zkb_detail::check(zkb_buf_copy(ctx, xd.p, ZKB_MEM_DEVICE, x.data(), ZKB_MEM_HOST, (std::uint64_t)(nv + 1) * 32, nullptr), ctx, "zkb_buf_copy");
        scale_on_device(xd.p, xd.p, nv + 1, r_inv);
        // aA | aB | aC, m elements each
        dev_elems abc(3 * m);
        zkb_detail::check(zkb_buf_zero(ctx, abc.p, (std::uint64_t)3 * m * 32, nullptr), ctx, "zkb_buf_zero");
        const auto &mats = cs.device();
        for (int k = 0; k < 3; k++)
            zkb_detail::check(zkb_sparse_matvec(ctx, mats.m[k], xd.p, ZKB_MEM_DEVICE, abc.at(k * m), nullptr), ctx, "zkb_sparse_matvec");
        if (check_satisfied) {   // cs.is_satisfied (:225): <a,x><b,x> - <c,x> must vanish on every constraint row
            dev_elems resid(nc ? nc : 1);
            std::uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
            zkb_detail::check(zkb_vec(ctx, FieldType::field_id, ZKB_VEC_MUL_SUB_SCALE, nc, abc.at(0), abc.at(m), abc.at(2 * m), one, resid.p,
                                      ZKB_MEM_DEVICE, nullptr), ctx, "zkb_vec");
            std::vector<std::uint32_t> h(nc * 8);
            zkb_detail::check(zkb_buf_copy(ctx, h.data(), ZKB_MEM_HOST, resid.p, ZKB_MEM_DEVICE, (std::uint64_t)nc * 32, nullptr), ctx, "zkb_buf_copy");
            for (std::uint32_t w : h)
                if (w) throw std::invalid_argument("witness_map: the assignment does not satisfy the constraint system");
        }
        // the additional constraints input_i * 0 = 0 (:239-242): aA[nc + i] = x[i], i <= ni
        zkb_detail::check(zkb_buf_copy(ctx, abc.at(nc), ZKB_MEM_DEVICE, xd.p, ZKB_MEM_DEVICE, (std::uint64_t)(ni + 1) * 32, nullptr), ctx, "zkb_buf_copy");
        // inverse_fft(aA), inverse_fft(aB), inverse_fft(aC)  (:250,:252,:293)
        zkb_detail::check(zkb_ntt(ctx, FieldType::field_id, log_m, 3, abc.p, abc.p, 1, nullptr, ZKB_MEM_DEVICE, nullptr), ctx, "zkb_ntt");
        if (ab_coeff_host) {
            ab_coeff_host->resize(2 * m);
            dev_elems tmp(2 * m);
            scale_on_device(abc.p, tmp.p, 2 * m, r_one);    // canonical -> Montgomery limbs for the host value type
            zkb_detail::check(zkb_buf_copy(ctx, ab_coeff_host->data(), ZKB_MEM_HOST, tmp.p, ZKB_MEM_DEVICE, (std::uint64_t)2 * m * 32, nullptr), ctx, "zkb_buf_copy");
        }
        // multiply_by_coset(g) + fft on all three (:266-276, :295-299)
        std::uint32_t g[8];
        algebra::fields::arithmetic_params<FieldType>::multiplicative_generator_value().to_canonical_limbs(g);
        zkb_detail::check(zkb_ntt(ctx, FieldType::field_id, log_m, 3, abc.p, abc.p, 0, g, ZKB_MEM_DEVICE, nullptr), ctx, "zkb_ntt");
        // H_tmp = (aA * aB - aC) / Z(g)  (:283-308; divide_by_z_on_coset multiplies by 1 / (g^m - 1))
        std::uint32_t zinv[8];
        domain->compute_vanishing_polynomial(algebra::fields::arithmetic_params<FieldType>::multiplicative_generator_value()).inversed().to_canonical_limbs(zinv);
        dev_elems h(m);
        zkb_detail::check(zkb_vec(ctx, FieldType::field_id, ZKB_VEC_MUL_SUB_SCALE, m, abc.at(0), abc.at(m), abc.at(2 * m), zinv, h.p,
                                  ZKB_MEM_DEVICE, nullptr), ctx, "zkb_vec");
        // inverse_fft + multiply_by_coset(g^-1)  (:310-315)
        zkb_detail::check(zkb_ntt(ctx, FieldType::field_id, log_m, 1, h.p, h.p, 1, g, ZKB_MEM_DEVICE, nullptr), ctx, "zkb_ntt");
        x_canonical_out = std::move(xd);
        return h;
    }

    // r1cs_to_qap.hpp:219-325
    static qap_witness<FieldType> witness_map(const r1cs_constraint_system<FieldType> &cs, const r1cs_primary_input<FieldType> &primary_input,
                                              const r1cs_auxiliary_input<FieldType> &auxiliary_input, const value_type &d1,
                                              const value_type &d2, const value_type &d3) {
        zkb_ctx *ctx = zkb_detail::context();
        const bool patch = !(d1.is_zero() && d2.is_zero() && d3.is_zero());
        std::vector<value_type> ab;
        zkb_detail_r1cs::dev_elems xd;
        std::size_t m = 0;
        auto h = witness_map_device(cs, primary_input, auxiliary_input, xd, m, patch ? &ab : nullptr, true);
        std::uint32_t r_inv[8], r_one[8];
        mont_scalars(r_inv, r_one);
        scale_on_device(h.p, h.p, m, r_one);
        qap_witness<FieldType> w;
        w.num_variables = cs.num_variables();
        w.degree = m;
        w.num_inputs = cs.num_inputs();
        w.d1 = d1; w.d2 = d2; w.d3 = d3;
        w.coefficients_for_ABCs = primary_input;
        w.coefficients_for_ABCs.insert(w.coefficients_for_ABCs.end(), auxiliary_input.begin(), auxiliary_input.end());
        w.coefficients_for_H.assign(m + 1, value_type::zero());
        zkb_detail::check(zkb_buf_copy(ctx, w.coefficients_for_H.data(), ZKB_MEM_HOST, h.p, ZKB_MEM_DEVICE, (std::uint64_t)m * 32, nullptr), ctx, "zkb_buf_copy");
        if (patch) {   // (d2 A + d1 B - d3) + d1 d2 Z  (:256-263)
            for (std::size_t i = 0; i < m; i++) w.coefficients_for_H[i] += d2 * ab[i] + d1 * ab[m + i];
            w.coefficients_for_H[0] -= d3;
            w.coefficients_for_H[m] += d1 * d2;    // add_poly_z
            w.coefficients_for_H[0] -= d1 * d2;
        }
        return w;
    }
};
}  // namespace reductions

// proving_key.hpp:38-120.  The members are the reference's; device() uploads the query vectors once.
template <class CurveType>
struct r1cs_gg_ppzksnark_proving_key {
    typedef CurveType curve_type;
    typedef typename CurveType::scalar_field_type scalar_field_type;
    typedef typename CurveType::template g1_type<> g1_type;
    typedef typename CurveType::template g2_type<> g2_type;
    typedef r1cs_constraint_system<scalar_field_type> constraint_system_type;

    typename g1_type::value_type alpha_g1, beta_g1;
    typename g2_type::value_type beta_g2;
    typename g1_type::value_type delta_g1;
    typename g2_type::value_type delta_g2;
    std::vector<typename g1_type::value_type> A_query;
    commitments::knowledge_commitment_vector<g2_type, g1_type> B_query;
    std::vector<typename g1_type::value_type> H_query;
    std::vector<typename g1_type::value_type> L_query;
    constraint_system_type constraint_system;

    struct device_cache {
        algebra::multiexp_bases<g1_type> A, B_g1, H, L;
        algebra::multiexp_bases<g2_type> B_g2;
    };
    // window_tables: also build the window tables (zkb_msm_bases_precompute) - worth it for a key that serves many proofs
    const device_cache &device(bool window_tables = false) const {
        if (!_dev) {
            auto d = std::make_shared<device_cache>();
            d->A = algebra::multiexp_bases<g1_type>(A_query.begin(), A_query.end());
            d->H = algebra::multiexp_bases<g1_type>(H_query.begin(), H_query.end());
            d->L = algebra::multiexp_bases<g1_type>(L_query.begin(), L_query.end());
            std::vector<typename g1_type::value_type> h;
            std::vector<typename g2_type::value_type> g;
            for (const auto &v : B_query.values) { g.push_back(v.g); h.push_back(v.h); }
            d->B_g1 = algebra::multiexp_bases<g1_type>(h.begin(), h.end());
            d->B_g2 = algebra::multiexp_bases<g2_type>(g.begin(), g.end());
            _dev = d;
        }
        if (window_tables && !_tables) {
            _dev->A.precompute(); _dev->B_g1.precompute(); _dev->H.precompute(); _dev->L.precompute(); _dev->B_g2.precompute();
            _tables = true;
        }
        return *_dev;
    }
    void invalidate_device_cache() const { _dev.reset(); _tables = false; }

private:
    mutable std::shared_ptr<device_cache> _dev;
    mutable bool _tables = false;
};

// prover.hpp:52-158
template <class CurveType>
class r1cs_gg_ppzksnark_prover {
    typedef typename CurveType::scalar_field_type scalar_field_type;
    typedef typename CurveType::template g1_type<> g1_type;
    typedef typename CurveType::template g2_type<> g2_type;
    typedef typename scalar_field_type::value_type scalar;

public:
    typedef r1cs_primary_input<scalar_field_type> primary_input_type;
    typedef r1cs_auxiliary_input<scalar_field_type> auxiliary_input_type;
    typedef r1cs_gg_ppzksnark_proving_key<CurveType> proving_key_type;
    typedef r1cs_gg_ppzksnark_proof<CurveType> proof_type;

    // the reference draws r and s with algebra::random_element (:91-92)
    static proof_type process(const proving_key_type &proving_key, const primary_input_type &primary_input,
                              const auxiliary_input_type &auxiliary_input) {
        std::random_device rd;
        std::uint32_t l[2][8];
        for (auto &v : l) {
            for (auto &w : v) w = rd();
            v[7] &= 0x0FFFFFFFu;     // < every scalar modulus here (uniform up to a 2^-2 truncation of the range)
        }
        return process(proving_key, primary_input, auxiliary_input, scalar::from_canonical_limbs(l[0]), scalar::from_canonical_limbs(l[1]));
    }

    // same with the zero-knowledge randomness passed in (tests, and callers with their own generator)
    static proof_type process(const proving_key_type &pk, const primary_input_type &primary_input, const auxiliary_input_type &auxiliary_input,
                              const scalar &r, const scalar &s) {
        typedef reductions::r1cs_to_qap<scalar_field_type> qap;
        zkb_ctx *ctx = zkb_detail::context();
        const auto &cs = pk.constraint_system;
        const std::size_t nv = cs.num_variables(), ni = cs.num_inputs();
        // BOOST_ASSERT(is_satisfied) (:77) is checked on the device inside the witness map, on the row values it computes anyway
        zkb_detail_r1cs::dev_elems xd;
        std::size_t m = 0;
        auto h = qap::witness_map_device(cs, primary_input, auxiliary_input, xd, m, nullptr, true);
        const auto &dk = pk.device();
        if (dk.A.size() < nv + 1 || dk.H.size() < m - 1 || dk.L.size() != nv - ni)
            throw std::invalid_argument("r1cs_gg_ppzksnark_prover: query sizes do not match the constraint system");
        // A: const_padded_assignment[0 .. nv]; L: const_padded_assignment[ni + 1 .. nv]; H: coefficients_for_H[0 .. m - 2]
        auto evaluation_At = dk.A.multiexp_device(0, nv + 1, xd.p);
        auto evaluation_Ht = dk.H.multiexp_device(0, m - 1, h.p);
        auto evaluation_Lt = dk.L.multiexp_device(0, nv - ni, xd.at(ni + 1));
        // B: the entries of the sparse B_query take the scalar at their index (kc_multiexp_with_mixed_addition, :113-119)
        std::vector<std::uint32_t> xh((nv + 1) * 8), sb(pk.B_query.indices.size() * 8 + 8);
        zkb_detail::check(zkb_buf_copy(ctx, xh.data(), ZKB_MEM_HOST, xd.p, ZKB_MEM_DEVICE, (std::uint64_t)(nv + 1) * 32, nullptr), ctx, "zkb_buf_copy");
        for (std::size_t j = 0; j < pk.B_query.indices.size(); j++) {
            if (pk.B_query.indices[j] > nv) throw std::invalid_argument("r1cs_gg_ppzksnark_prover: B_query index outside the assignment");
            std::memcpy(&sb[8 * j], &xh[8 * pk.B_query.indices[j]], 32);
        }
        auto evaluation_Bt_g = dk.B_g2.multiexp_limbs(0, pk.B_query.indices.size(), sb.data());
        auto evaluation_Bt_h = dk.B_g1.multiexp_limbs(0, pk.B_query.indices.size(), sb.data());
        // :141-157
        auto g1_A = pk.alpha_g1 + evaluation_At + r * pk.delta_g1;
        auto g1_B = pk.beta_g1 + evaluation_Bt_h + s * pk.delta_g1;
        auto g2_B = pk.beta_g2 + evaluation_Bt_g + s * pk.delta_g2;
        auto g1_C = evaluation_Ht + evaluation_Lt + s * g1_A + r * g1_B - (r * s) * pk.delta_g1;
        return proof_type(g1_A, g2_B, g1_C);
    }
};

}  // namespace snark
}  // namespace zk
}  // namespace crypto3
}  // namespace nil

#endif  // ZKB_R1CS_GG_PPZKSNARK_HPP
