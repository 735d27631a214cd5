// Placeholder prover, commitment side, as C++ host templates over the C ABI: every
// column stays in a device buffer between the commits, the argument polynomials are postfix programs for zkb_expr_eval.
//
// Mirrors zk/snark/systems/plonk/placeholder/prover.hpp:133-217 (the order of commits and challenges),
// permutation_argument.hpp:95-215 (V_P, its parts when max_quotient_chunks is set, F_0..F_2), gates_argument.hpp:133-217
// (F_7), prover.hpp:220-283 (quotient, split) and :346-416 (evaluation points).  The same sequence as
// crypto3_zk_b200/placeholder.py, which tests/test_gpu_placeholder.py checks against the oracle and the verifier's
// identity; tests/cpp/host_api_test.cpp prints this driver's commitments / challenge / opened values and
// tests/test_cpp_host.py compares them with the Python driver's on the same circuit.
//
// Not the reference's proof object: the transcript is seeded by the caller (upstream absorbs a hash of the constraint
// system first, prover.hpp:126-131) and lookup gates are not supported here (F_3..F_6 = 0, prover.hpp:286-297).
#pragma once
#include <set>
#include "zkb_crypto3.hpp"

namespace nil {
namespace crypto3 {
namespace zk {
namespace snark {

// math::expression over plonk_variable (column of the assignment table, rotation), restricted to + - * and constants
template <class V>
class plonk_expression {
public:
    enum kind_t { COL, CONST, ADD, SUB, MUL };

private:
    struct node {
        kind_t kind;
        std::uint32_t column = 0;
        std::int32_t rotation = 0;
        V value = V::zero();
        std::shared_ptr<node> a, b;
    };
    std::shared_ptr<node> _n;
    static plonk_expression make(kind_t k, const plonk_expression &x, const plonk_expression &y) {
        plonk_expression e;
        e._n = std::make_shared<node>();
        e._n->kind = k; e._n->a = x._n; e._n->b = y._n;
        return e;
    }
    static std::size_t degree_of(const node &n) {
        switch (n.kind) {
            case COL: return 1;
            case CONST: return 0;
            case MUL: return degree_of(*n.a) + degree_of(*n.b);
            default: return std::max(degree_of(*n.a), degree_of(*n.b));
        }
    }
    static void emit(const node &n, std::vector<zkb_expr_instr> &prog, std::vector<V> &consts) {
        zkb_expr_instr in = {0, 0, 0};
        switch (n.kind) {
            case COL: in.op = ZKB_EXPR_PUSH_COL; in.a = n.column; in.b = n.rotation; break;
            case CONST: {
                std::size_t k = std::find(consts.begin(), consts.end(), n.value) - consts.begin();
                if (k == consts.size()) consts.push_back(n.value);
                in.op = ZKB_EXPR_PUSH_CONST; in.a = (std::uint32_t)k;
                break;
            }
            default:
                emit(*n.a, prog, consts);
                emit(*n.b, prog, consts);
                in.op = n.kind == ADD ? ZKB_EXPR_ADD : n.kind == SUB ? ZKB_EXPR_SUB : ZKB_EXPR_MUL;
        }
        prog.push_back(in);
    }
    template <class F>
    static void walk(const node &n, F &f) {
        if (n.kind == COL) f(n.column, n.rotation);
        else if (n.kind != CONST) { walk(*n.a, f); walk(*n.b, f); }
    }

public:
    plonk_expression() = default;
    static plonk_expression var(std::uint32_t column, std::int32_t rotation = 0) {
        plonk_expression e;
        e._n = std::make_shared<node>();
        e._n->kind = COL; e._n->column = column; e._n->rotation = rotation;
        return e;
    }
    static plonk_expression constant(const V &v) {
        plonk_expression e;
        e._n = std::make_shared<node>();
        e._n->kind = CONST; e._n->value = v;
        return e;
    }
    bool empty() const { return !_n; }
    friend plonk_expression operator+(const plonk_expression &x, const plonk_expression &y) { return make(ADD, x, y); }
    friend plonk_expression operator-(const plonk_expression &x, const plonk_expression &y) { return make(SUB, x, y); }
    friend plonk_expression operator*(const plonk_expression &x, const plonk_expression &y) { return make(MUL, x, y); }
    // expression_max_degree_visitor (gates_argument.hpp:161): every variable counts one
    std::size_t degree() const { return degree_of(*_n); }
    void compile(std::vector<zkb_expr_instr> &prog, std::vector<std::uint32_t> &const_limbs) const {
        std::vector<V> consts;
        prog.clear();
        emit(*_n, prog, consts);
        const_limbs.assign(consts.size() * 8, 0);
        for (std::size_t k = 0; k < consts.size(); k++) consts[k].to_canonical_limbs(&const_limbs[8 * k]);
    }
    template <class F>
    void for_each_variable(F f) const { walk(*_n, f); }
};

template <class V>
struct plonk_gate {
    std::size_t selector_index;
    std::vector<plonk_expression<V>> constraints;
};

// The preprocessed public side of a circuit with its columns on the device.  Table columns in this order:
// witness | public_input | constant | selector (plonk_table_description::global_index).
template <class FieldType>
struct placeholder_device_circuit {
    typedef typename FieldType::value_type value_type;
    typedef zkb_detail_lpc::device_buffer dbuf;
    std::size_t log_n = 0, witness_columns = 0, public_input_columns = 0, constant_columns = 0, selector_columns = 0;
    std::size_t usable_rows = 0, max_quotient_chunks = 0;
    std::vector<plonk_gate<value_type>> gates;
    std::vector<std::size_t> permuted_columns;       // global indices
    dbuf s_id, s_sigma;                              // [permuted][n]   (preprocessor.hpp:418-460)
    dbuf q_last, q_blind, lagrange_0;                // [n]             (preprocessor.hpp:463-476)
    dbuf constants, selectors;                       // [count][n]

    std::size_t n() const { return std::size_t(1) << log_n; }
    std::size_t table_width() const { return witness_columns + public_input_columns + constant_columns + selector_columns; }
    std::size_t selector_column(std::size_t k) const { return witness_columns + public_input_columns + constant_columns + k; }
    std::size_t max_gates_degree() const {
        std::size_t d = 0;
        for (const auto &g : gates)
            for (const auto &c : g.constraints) d = std::max(d, c.degree());
        return d;
    }
    // permutation_partitions_num (preprocessor.hpp:78-87)
    std::size_t permutation_parts() const {
        const std::size_t npc = permuted_columns.size();
        if (npc == 0) return 0;
        if (max_quotient_chunks == 0) return 1;
        return (npc + max_quotient_chunks - 2) / (max_quotient_chunks - 1);
    }
    // columns_rotations (preprocessor.hpp:363-383): a std::set per table column, 0 always inside
    std::vector<std::set<int>> columns_rotations() const {
        std::vector<std::set<int>> r(table_width());
        for (auto &s : r) s.insert(0);
        for (const auto &g : gates)
            for (const auto &c : g.constraints) c.for_each_variable([&](std::uint32_t col, std::int32_t rot) { r[col].insert(rot); });
        return r;
    }
    // split_polynomial_size (prover.hpp:227-246), no lookups
    std::size_t quotient_chunks() const {
        const std::size_t N = n();
        std::size_t size = std::max((permuted_columns.size() + 2) * (N - 1), (max_gates_degree() + 1) * (N - 1));
        size = (size + N - 1) / N;
        if (max_quotient_chunks && size > max_quotient_chunks) size = max_quotient_chunks;
        return size;
    }
};

enum { FIXED_VALUES_BATCH = 0, VARIABLE_VALUES_BATCH = 1, PERMUTATION_BATCH = 2, QUOTIENT_BATCH = 3, LOOKUP_BATCH = 4 };

template <class FieldType, class MerkleHash, class TranscriptHash>
class placeholder_prover {
public:
    typedef typename FieldType::value_type value_type;
    typedef plonk_expression<value_type> expr;
    typedef commitments::lpc_commitment_scheme<FieldType, MerkleHash, TranscriptHash> commitment_scheme_type;
    typedef typename commitment_scheme_type::transcript_type transcript_type;
    typedef typename commitment_scheme_type::commitment_type commitment_type;
    typedef placeholder_device_circuit<FieldType> circuit_type;
    typedef zkb_detail_lpc::device_buffer dbuf;
    static constexpr std::size_t f_parts = 8;

    struct proof_type {
        std::map<std::size_t, commitment_type> commitments;
        value_type challenge;
        typename commitment_scheme_type::proof_type eval_proof;
        std::size_t quotient_chunks = 0, log_extension = 0;
    };

private:
    static constexpr int fid = FieldType::field_id;
    static zkb_ctx *ctx() { return zkb_detail::context(); }
    static void ck(int st, const char *what) { zkb_detail::check(st, ctx(), what); }
    static void copy(dbuf &dst, std::size_t dst_elem, const dbuf &src, std::size_t src_elem, std::size_t count) {
        if (count == 0) return;
        ck(zkb_buf_copy(ctx(), (char *)dst.p + dst_elem * 32, ZKB_MEM_DEVICE, (const char *)src.p + src_elem * 32, ZKB_MEM_DEVICE, count * 32, nullptr),
           "zkb_buf_copy");
    }
    static void eval(const expr &e, std::size_t n, std::size_t ncols, const void *cols, void *out, std::size_t stride, std::size_t offset,
                     bool accumulate) {
        std::vector<zkb_expr_instr> prog;
        std::vector<std::uint32_t> consts;
        e.compile(prog, consts);
        ck(zkb_expr_eval(ctx(), fid, n, (std::uint32_t)ncols, cols, prog.data(), (std::uint32_t)prog.size(), consts.empty() ? nullptr : consts.data(),
                         (std::uint32_t)(consts.size() / 8), out, stride, offset, accumulate ? 1 : 0, nullptr),
           "zkb_expr_eval");
    }
    static expr product(const std::vector<expr> &f, std::size_t first, std::size_t last) {
        expr acc = f[first];
        for (std::size_t i = first + 1; i < last; i++) acc = acc * f[i];
        return acc;
    }
    static expr total(const std::vector<expr> &t) {
        expr acc = t[0];
        for (std::size_t i = 1; i < t.size(); i++) acc = acc + t[i];
        return acc;
    }

public:
    // The preprocessor's commitment part (preprocessor.hpp:481-489): identity | sigma | q_last | q_blind | constants |
    // selectors committed as the fixed batch, the root absorbed, etha drawn, the fixed values at etha recorded.
    static commitment_type preprocess(const circuit_type &c, commitment_scheme_type &scheme, transcript_type &transcript) {
        const std::size_t n = c.n(), npc = c.permuted_columns.size();
        const std::size_t count = 2 * npc + 2 + c.constant_columns + c.selector_columns;
        dbuf fixed(count * n * 32);
        std::size_t o = 0;
        copy(fixed, o, c.s_id, 0, npc * n); o += npc * n;
        copy(fixed, o, c.s_sigma, 0, npc * n); o += npc * n;
        copy(fixed, o, c.q_last, 0, n); o += n;
        copy(fixed, o, c.q_blind, 0, n); o += n;
        copy(fixed, o, c.constants, 0, c.constant_columns * n); o += c.constant_columns * n;
        copy(fixed, o, c.selectors, 0, c.selector_columns * n);
        scheme.append_device_batch(FIXED_VALUES_BATCH, fixed.p, count, n);
        scheme.mark_batch_as_fixed(FIXED_VALUES_BATCH);
        commitment_type root = scheme.commit(FIXED_VALUES_BATCH);
        transcript(root);
        scheme.setup(transcript);
        return root;
    }

    // witness: [witness_columns][n], public_input: [public_input_columns][n], canonical limbs on the device
    static proof_type process(const circuit_type &c, const dbuf &witness, const dbuf &public_input, commitment_scheme_type &scheme,
                              transcript_type &transcript) {
        proof_type proof;
        const std::size_t n = c.n(), log_n = c.log_n, tw = c.table_width(), npc = c.permuted_columns.size();
        const std::size_t nvar = c.witness_columns + c.public_input_columns, parts = c.permutation_parts();
        // columns the expressions see: the table, S_id, S_sigma, q_last, q_blind, L_0, V_P and the permutation parts
        const std::size_t c_sid = tw, c_ssg = tw + npc, c_qlast = tw + 2 * npc, c_qblind = c_qlast + 1, c_l0 = c_qlast + 2, c_vp = c_qlast + 3;
        const std::size_t n_base = c_vp, n_all = c_vp + std::max<std::size_t>(parts, 1);
        dbuf all(n_all * n * 32);
        ck(zkb_buf_zero(ctx(), (char *)all.p + n_base * n * 32, (n_all - n_base) * n * 32, nullptr), "zkb_buf_zero");
        std::size_t o = 0;
        copy(all, o, witness, 0, c.witness_columns * n); o += c.witness_columns * n;
        copy(all, o, public_input, 0, c.public_input_columns * n); o += c.public_input_columns * n;
        copy(all, o, c.constants, 0, c.constant_columns * n); o += c.constant_columns * n;
        copy(all, o, c.selectors, 0, c.selector_columns * n); o += c.selector_columns * n;
        copy(all, o, c.s_id, 0, npc * n); o += npc * n;
        copy(all, o, c.s_sigma, 0, npc * n); o += npc * n;
        copy(all, o, c.q_last, 0, n); o += n;
        copy(all, o, c.q_blind, 0, n); o += n;
        copy(all, o, c.lagrange_0, 0, n);
        // 2. witness and public-input columns (prover.hpp:141)
        scheme.append_device_batch(VARIABLE_VALUES_BATCH, all.p, nvar, n);
        proof.commitments[VARIABLE_VALUES_BATCH] = scheme.commit(VARIABLE_VALUES_BATCH);
        transcript(proof.commitments[VARIABLE_VALUES_BATCH]);
        // 4. permutation argument
        std::map<std::size_t, expr> F;
        const expr one = expr::constant(value_type::one());
        const expr mask = one - expr::var((std::uint32_t)c_qlast) - expr::var((std::uint32_t)c_qblind);
        if (npc) {
            const value_type beta = transcript.template challenge<FieldType>(), gamma = transcript.template challenge<FieldType>();
            std::uint32_t bl[8], gl[8];
            beta.to_canonical_limbs(bl);
            gamma.to_canonical_limbs(gl);
            dbuf cols(npc * n * 32);
            for (std::size_t i = 0; i < npc; i++) copy(cols, i * n, all, c.permuted_columns[i] * n, n);
            char *vp = (char *)all.p + c_vp * n * 32;
            ck(zkb_permutation_grand_product(ctx(), fid, n, (std::uint32_t)npc, cols.p, c.s_id.p, c.s_sigma.p, bl, gl, vp, nullptr),
               "zkb_permutation_grand_product");
            std::vector<expr> g_f, h_f;
            for (std::size_t i = 0; i < npc; i++) {
                const expr col = expr::var((std::uint32_t)c.permuted_columns[i]);
                g_f.push_back(expr::constant(beta) * expr::var((std::uint32_t)(c_sid + i)) + expr::constant(gamma) + col);
                h_f.push_back(expr::constant(beta) * expr::var((std::uint32_t)(c_ssg + i)) + expr::constant(gamma) + col);
            }
            const std::size_t group = c.max_quotient_chunks == 0 ? npc : c.max_quotient_chunks - 1;
            std::vector<expr> gs, hs;
            for (std::size_t i = 0; i < npc; i += group) {
                gs.push_back(product(g_f, i, std::min(npc, i + group)));
                hs.push_back(product(h_f, i, std::min(npc, i + group)));
            }
            if (gs.size() != parts) throw std::logic_error("placeholder: permutation parts");
            std::vector<value_type> perm_alphas;
            for (std::size_t i = 0; i + 1 < parts; i++) perm_alphas.push_back(transcript.template challenge<FieldType>());
            const expr VP = expr::var((std::uint32_t)c_vp), VP_shifted = expr::var((std::uint32_t)c_vp, 1);
            F[0] = (one - VP) * expr::var((std::uint32_t)c_l0);
            if (parts == 1) {
                F[1] = mask * (VP_shifted * hs[0] - VP * gs[0]);
            } else {
                // the running product after every part is a committed polynomial (permutation_argument.hpp:194-210)
                dbuf gv(n * 32), hv(n * 32);
                std::vector<expr> terms;
                std::size_t prev_c = c_vp;
                for (std::size_t i = 0; i + 1 < parts; i++) {
                    eval(gs[i], n, n_base, all.p, gv.p, 1, 0, false);
                    eval(hs[i], n, n_base, all.p, hv.p, 1, 0, false);
                    ck(zkb_vec(ctx(), fid, ZKB_VEC_MUL, n, (char *)all.p + prev_c * n * 32, gv.p, nullptr, nullptr, gv.p, ZKB_MEM_DEVICE, nullptr), "zkb_vec");
                    ck(zkb_batch_inverse(ctx(), fid, n, hv.p, hv.p, nullptr), "zkb_batch_inverse");
                    ck(zkb_vec(ctx(), fid, ZKB_VEC_MUL, n, gv.p, hv.p, nullptr, nullptr, gv.p, ZKB_MEM_DEVICE, nullptr), "zkb_vec");
                    const std::size_t cur_c = c_vp + 1 + i;
                    copy(all, cur_c * n, all, c_vp * n, n);                 // current_poly = V_P, then the usable rows
                    copy(all, cur_c * n, gv, 0, c.usable_rows);
                    terms.push_back(expr::constant(perm_alphas[i]) *
                                    (expr::var((std::uint32_t)prev_c) * gs[i] - expr::var((std::uint32_t)cur_c) * hs[i]));
                    prev_c = cur_c;
                }
                terms.push_back(expr::var((std::uint32_t)prev_c) * gs.back() - VP_shifted * hs.back());
                F[1] = total(terms) * (expr::var((std::uint32_t)c_qlast) + expr::var((std::uint32_t)c_qblind) - one);
            }
            F[2] = expr::var((std::uint32_t)c_qlast) * (VP * VP - VP);
            scheme.append_device_batch(PERMUTATION_BATCH, vp, parts, n);
            proof.commitments[PERMUTATION_BATCH] = scheme.commit(PERMUTATION_BATCH);
            transcript(proof.commitments[PERMUTATION_BATCH]);
        }
        // 6. circuit satisfiability (gates_argument.hpp:133-217)
        if (!c.gates.empty()) {
            const value_type theta = transcript.template challenge<FieldType>();
            value_type theta_acc = value_type::one();
            std::vector<expr> terms;
            for (const auto &g : c.gates) {
                std::vector<expr> inner;
                for (const auto &con : g.constraints) {
                    inner.push_back(con * expr::constant(theta_acc));
                    theta_acc *= theta;
                }
                terms.push_back(total(inner) * expr::var((std::uint32_t)c.selector_column(g.selector_index)));
            }
            F[7] = total(terms) * mask;
        }
        // 7. quotient (prover.hpp:220-283): alphas, sum alpha_i F_i on the extended domain, / Z, split, commit
        std::vector<value_type> alphas;
        for (std::size_t i = 0; i < f_parts; i++) alphas.push_back(transcript.template challenge<FieldType>());
        std::vector<expr> weighted;
        std::size_t deg = 1;
        for (const auto &kv : F) {
            weighted.push_back(kv.second * expr::constant(alphas[kv.first]));
            deg = std::max(deg, kv.second.degree());
        }
        std::size_t log_d = 1;
        while ((std::size_t(1) << log_d) * n <= deg * (n - 1)) log_d++;
        const std::size_t D = std::size_t(1) << log_d;
        dbuf f(n * D * 32);
        {
            dbuf coef(n_all * n * 32), work(n_all * n * 32);
            ck(zkb_ntt(ctx(), fid, (int)log_n, (std::uint32_t)n_all, all.p, coef.p, 1, nullptr, ZKB_MEM_DEVICE, nullptr), "zkb_ntt");
            std::uint32_t wl[8];
            ck(zkb_field_unity_root(fid, (int)(log_n + log_d), wl), "zkb_field_unity_root");
            const value_type w_ext = value_type::from_canonical_limbs(wl);
            for (std::size_t j = 0; j < D; j++) {
                const void *cos = all.p;                                  // coset 0 is the basic domain itself
                if (j) {
                    std::uint32_t sh[8];
                    w_ext.pow(j).to_canonical_limbs(sh);
                    ck(zkb_ntt(ctx(), fid, (int)log_n, (std::uint32_t)n_all, coef.p, work.p, 0, sh, ZKB_MEM_DEVICE, nullptr), "zkb_ntt");
                    cos = work.p;
                }
                for (std::size_t k = 0; k < weighted.size(); k++) eval(weighted[k], n, n_all, cos, f.p, D, j, k > 0);
            }
        }
        ck(zkb_ntt(ctx(), fid, (int)(log_n + log_d), 1, f.p, f.p, 1, nullptr, ZKB_MEM_DEVICE, nullptr), "zkb_ntt");
        const std::size_t nchunks = c.quotient_chunks();
        dbuf t_chunks(nchunks * n * 32);
        ck(zkb_quotient_split(ctx(), fid, (int)log_n, (int)(log_n + log_d), f.p, (std::uint32_t)nchunks, t_chunks.p, nullptr), "zkb_quotient_split");
        scheme.append_device_batch(QUOTIENT_BATCH, t_chunks.p, nchunks, n);
        proof.commitments[QUOTIENT_BATCH] = scheme.commit(QUOTIENT_BATCH);
        transcript(proof.commitments[QUOTIENT_BATCH]);
        proof.quotient_chunks = nchunks;
        proof.log_extension = log_d;
        // 8. evaluation points (generate_evaluation_points, prover.hpp:346-416) and the evaluation proof
        const value_type y = transcript.template challenge<FieldType>();
        proof.challenge = y;
        const value_type omega = math::basic_radix2_domain<FieldType>(n).get_domain_element(1);
        auto rotated = [&](int r) { return y * omega.pow((std::size_t)(((r % (long long)n) + (long long)n) % (long long)n)); };
        const auto rots = c.columns_rotations();
        for (std::size_t i = 0; i < nvar; i++)
            for (int r : rots[i]) scheme.append_eval_point(VARIABLE_VALUES_BATCH, i, rotated(r));
        if (npc) {
            scheme.append_eval_point(PERMUTATION_BATCH, y);
            scheme.append_eval_point(PERMUTATION_BATCH, 0, rotated(1));
        }
        scheme.append_eval_point(QUOTIENT_BATCH, y);
        if (scheme.has_batch(FIXED_VALUES_BATCH)) {
            const std::size_t start = 2 * npc + 2;
            for (std::size_t i = 0; i < start; i++) scheme.append_eval_point(FIXED_VALUES_BATCH, i, y);
            scheme.append_eval_point(FIXED_VALUES_BATCH, start - 2, rotated(1));
            scheme.append_eval_point(FIXED_VALUES_BATCH, start - 1, rotated(1));
            for (std::size_t ind = 0; ind < c.constant_columns + c.selector_columns; ind++)
                for (int r : rots[nvar + ind]) scheme.append_eval_point(FIXED_VALUES_BATCH, start + ind, rotated(r));
        }
        proof.eval_proof = scheme.proof_eval(transcript);
        return proof;
    }
};

}  // namespace snark
}  // namespace zk
}  // namespace crypto3
}  // namespace nil
