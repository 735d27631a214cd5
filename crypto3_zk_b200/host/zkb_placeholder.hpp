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
// system first, prover.hpp:126-131).  The lookup argument (lookup_argument.hpp:153-325) is included: compressed table
// values / inputs as expressions, zkb_lookup_sort, zkb_lookup_grand_product, F_3..F_6.
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
    // every variable read r rows further (math::polynomial_shift of the polynomial the expression denotes)
    plonk_expression shifted(std::int32_t r) const {
        if (_n->kind == COL) return var(_n->column, _n->rotation + r);
        if (_n->kind == CONST) return *this;
        plonk_expression a, b;
        a._n = _n->a; b._n = _n->b;
        return make(_n->kind, a.shifted(r), b.shifted(r));
    }
};

template <class V>
struct plonk_gate {
    std::size_t selector_index;
    std::vector<plonk_expression<V>> constraints;
};

// plonk_lookup_table / plonk_lookup_gate (arithmetization/plonk/lookup_table.hpp, lookup_gate.hpp) as the prover reads them
struct plonk_lookup_table {
    std::size_t tag_index;                                   // selector column of the table's rows
    std::vector<std::vector<std::size_t>> lookup_options;    // per option: the constant columns holding the table
};
template <class V>
struct plonk_lookup_constraint {
    std::size_t table_id;                                    // 1-based
    std::vector<plonk_expression<V>> lookup_input;
};
template <class V>
struct plonk_lookup_gate {
    std::size_t tag_index;                                   // selector column of the gate
    std::vector<plonk_lookup_constraint<V>> constraints;
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
    std::vector<plonk_lookup_table> lookup_tables;
    std::vector<plonk_lookup_gate<value_type>> lookup_gates;
    std::vector<std::size_t> permuted_columns;       // global indices
    dbuf s_id, s_sigma;                              // [permuted][n]   (preprocessor.hpp:418-460)
    dbuf q_last, q_blind, lagrange_0;                // [n]             (preprocessor.hpp:463-476)
    dbuf constants, selectors;                       // [count][n]

    std::size_t n() const { return std::size_t(1) << log_n; }
    std::size_t table_width() const { return witness_columns + public_input_columns + constant_columns + selector_columns; }
    std::size_t selector_column(std::size_t k) const { return witness_columns + public_input_columns + constant_columns + k; }
    std::size_t constant_column(std::size_t k) const { return witness_columns + public_input_columns + k; }
    bool lookups() const { return !lookup_gates.empty() && !lookup_tables.empty(); }
    std::size_t lookup_constraints_num() const {
        std::size_t r = 0;
        for (const auto &g : lookup_gates) r += g.constraints.size();
        return r;
    }
    std::size_t lookup_options_num() const {
        std::size_t r = 0;
        for (const auto &t : lookup_tables) r += t.lookup_options.size();
        return r;
    }
    static std::size_t inputs_degree(const plonk_lookup_constraint<value_type> &c) {
        std::size_t d = 0;
        for (const auto &e : c.lookup_input) d = std::max(d, e.degree());
        return d;
    }
    // plonk_constraint_system::lookup_poly_degree_bound (constraint_system.hpp:235-253)
    std::size_t lookup_poly_degree_bound() const {
        std::size_t d = 0;
        if (!lookup_gates.empty()) {
            for (const auto &g : lookup_gates)
                for (const auto &c : g.constraints) d += inputs_degree(c) + 1;
            for (const auto &t : lookup_tables) d += 3 * t.lookup_options.size();
        }
        return d;
    }
    // plonk_constraint_system::lookup_parts (constraint_system.hpp:256-300)
    std::vector<std::size_t> lookup_parts() const {
        if (max_quotient_chunks == 0) return {lookup_constraints_num() + lookup_options_num()};
        std::vector<std::size_t> parts;
        std::size_t chunk = 0, part = 0;
        for (const auto &g : lookup_gates)
            for (const auto &c : g.constraints) {
                const std::size_t d = inputs_degree(c);
                if (chunk + d + 1 >= max_quotient_chunks) { parts.push_back(part); chunk = 0; part = 0; }
                chunk += d + 1;
                part++;
            }
        for (const auto &t : lookup_tables)
            for (std::size_t o = 0; o < t.lookup_options.size(); o++) {
                if (chunk + 3 >= max_quotient_chunks) { parts.push_back(part); chunk = 0; part = 0; }
                chunk += 3;
                part++;
            }
        parts.push_back(part);
        for (auto v : parts)
            if (v == 0) throw std::invalid_argument("placeholder: max_quotient_chunks is too small for the lookup constraints");
        return parts;
    }
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
        if (!lookup_gates.empty()) {
            for (const auto &g : lookup_gates)
                for (const auto &c : g.constraints)
                    for (const auto &e : c.lookup_input) e.for_each_variable([&](std::uint32_t col, std::int32_t rot) { r[col].insert(rot); });
            for (const auto &t : lookup_tables) {             // tag and option columns are also read one row further
                r[selector_column(t.tag_index)].insert(1);
                for (const auto &o : t.lookup_options)
                    for (auto cidx : o) r[constant_column(cidx)].insert(1);
            }
        }
        return r;
    }
    // split_polynomial_size (prover.hpp:227-246)
    std::size_t quotient_chunks() const {
        const std::size_t N = n();
        std::size_t size = std::max((permuted_columns.size() + 2) * (N - 1), (max_gates_degree() + 1) * (N - 1));
        size = std::max(size, (lookup_poly_degree_bound() + 1) * (N - 1));
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
        const std::size_t nvar = c.witness_columns + c.public_input_columns, parts = c.permutation_parts(), usable = c.usable_rows;
        const bool lookup = c.lookups();
        const std::vector<std::size_t> lparts = lookup ? c.lookup_parts() : std::vector<std::size_t>();
        const std::size_t n_sorted = lookup ? c.lookup_constraints_num() + c.lookup_options_num() : 0;
        // columns the expressions see: the table, S_id, S_sigma, q_last, q_blind, L_0, then V_P and its parts, the sorted
        // lookup columns, V_L and its parts - one device buffer, [column][n]
        const std::size_t c_sid = tw, c_ssg = tw + npc, c_qlast = tw + 2 * npc, c_qblind = c_qlast + 1, c_l0 = c_qlast + 2;
        const std::size_t n_base = c_l0 + 1;
        const std::size_t c_vp = n_base, c_sorted = c_vp + parts, c_vl = c_sorted + n_sorted;
        const std::size_t n_all = std::max(c_vl + lparts.size(), n_base + 1);
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
        auto column_ptr = [&](std::size_t i) { return (char *)all.p + i * n * 32; };
        auto var = [](std::size_t i, std::int32_t r = 0) { return expr::var((std::uint32_t)i, r); };
        // 2. witness and public-input columns (prover.hpp:141)
        scheme.append_device_batch(VARIABLE_VALUES_BATCH, all.p, nvar, n);
        proof.commitments[VARIABLE_VALUES_BATCH] = scheme.commit(VARIABLE_VALUES_BATCH);
        transcript(proof.commitments[VARIABLE_VALUES_BATCH]);
        std::map<std::size_t, expr> F;
        const expr one = expr::constant(value_type::one());
        const expr mask = one - var(c_qlast) - var(c_qblind), neg_mask = var(c_qlast) + var(c_qblind) - one;
        dbuf gv(n * 32), hv(n * 32);
        // the part polynomials of a grand product cut into gs.size() parts (permutation_argument.hpp:194-215,
        // lookup_argument.hpp:262-289): current = previous * g_i / h_i on the usable rows, the rows beyond keep the grand
        // product's values; part i goes to column first_c + 1 + i.  Returns the bracket sum alpha_i (prev g_i - cur h_i) +
        // (prev g_last - shifted h_last).
        auto running_products = [&](std::size_t first_c, const std::vector<expr> &gs, const std::vector<expr> &hs,
                                    const std::vector<value_type> &alphas_, const expr &shifted) {
            std::vector<expr> terms;
            std::size_t prev_c = first_c;
            for (std::size_t i = 0; i + 1 < gs.size(); i++) {
                eval(gs[i], n, n_all, all.p, gv.p, 1, 0, false);
                eval(hs[i], n, n_all, all.p, hv.p, 1, 0, false);
                ck(zkb_vec(ctx(), fid, ZKB_VEC_MUL, n, column_ptr(prev_c), gv.p, nullptr, nullptr, gv.p, ZKB_MEM_DEVICE, nullptr), "zkb_vec");
                ck(zkb_batch_inverse(ctx(), fid, n, hv.p, hv.p, nullptr), "zkb_batch_inverse");
                ck(zkb_vec(ctx(), fid, ZKB_VEC_MUL, n, gv.p, hv.p, nullptr, nullptr, gv.p, ZKB_MEM_DEVICE, nullptr), "zkb_vec");
                const std::size_t cur_c = first_c + 1 + i;
                copy(all, cur_c * n, all, first_c * n, n);
                copy(all, cur_c * n, gv, 0, usable);
                terms.push_back(expr::constant(alphas_[i]) * (var(prev_c) * gs[i] - var(cur_c) * hs[i]));
                prev_c = cur_c;
            }
            terms.push_back(var(prev_c) * gs.back() - shifted * hs.back());
            return total(terms);
        };
        // 4. permutation argument (permutation_argument.hpp:95-215)
        if (npc) {
            const value_type beta = transcript.template challenge<FieldType>(), gamma = transcript.template challenge<FieldType>();
            std::uint32_t bl[8], gl[8];
            beta.to_canonical_limbs(bl);
            gamma.to_canonical_limbs(gl);
            dbuf cols(npc * n * 32);
            for (std::size_t i = 0; i < npc; i++) copy(cols, i * n, all, c.permuted_columns[i] * n, n);
            ck(zkb_permutation_grand_product(ctx(), fid, n, (std::uint32_t)npc, cols.p, c.s_id.p, c.s_sigma.p, bl, gl, column_ptr(c_vp), nullptr),
               "zkb_permutation_grand_product");
            std::vector<expr> g_f, h_f;
            for (std::size_t i = 0; i < npc; i++) {
                const expr col = var(c.permuted_columns[i]);
                g_f.push_back(expr::constant(beta) * var(c_sid + i) + expr::constant(gamma) + col);
                h_f.push_back(expr::constant(beta) * var(c_ssg + i) + expr::constant(gamma) + col);
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
            const expr VP = var(c_vp), VP_shifted = var(c_vp, 1);
            F[0] = (one - VP) * var(c_l0);
            if (parts == 1) F[1] = mask * (VP_shifted * hs[0] - VP * gs[0]);
            else F[1] = running_products(c_vp, gs, hs, perm_alphas, VP_shifted) * neg_mask;
            F[2] = var(c_qlast) * (VP * VP - VP);
        }
        // 5. lookup argument (lookup_argument.hpp:153-325)
        if (lookup) {
            const value_type theta = transcript.template challenge<FieldType>();
            std::vector<expr> value_exprs, input_exprs;
            for (std::size_t t_id = 0; t_id < c.lookup_tables.size(); t_id++) {       // prepare_lookup_value (:411-432)
                const auto &tab = c.lookup_tables[t_id];
                const expr tag = var(c.selector_column(tab.tag_index));
                for (const auto &option : tab.lookup_options) {
                    expr v = expr::constant(value_type(t_id + 1)) * tag;
                    value_type acc = theta;
                    for (auto cidx : option) {
                        v = v + expr::constant(acc) * tag * var(c.constant_column(cidx));
                        acc *= theta;
                    }
                    value_exprs.push_back(v * mask);
                }
            }
            for (const auto &gate : c.lookup_gates) {                                  // prepare_lookup_input (:434-486)
                const expr sel = var(c.selector_column(gate.tag_index));
                for (const auto &con : gate.constraints) {
                    expr l = sel * expr::constant(value_type(con.table_id));
                    value_type acc = theta;
                    for (const auto &e : con.lookup_input) {
                        l = l + expr::constant(acc) * sel * e;
                        acc *= theta;
                    }
                    input_exprs.push_back(l);
                }
            }
            const std::size_t nv = value_exprs.size(), ni = input_exprs.size();
            dbuf values(nv * n * 32), inputs(ni * n * 32);
            for (std::size_t i = 0; i < nv; i++) eval(value_exprs[i], n, n_all, all.p, (char *)values.p + i * n * 32, 1, 0, false);
            for (std::size_t i = 0; i < ni; i++) eval(input_exprs[i], n, n_all, all.p, (char *)inputs.p + i * n * 32, 1, 0, false);
            ck(zkb_lookup_sort(ctx(), fid, n, usable, (std::uint32_t)ni, inputs.p, (std::uint32_t)nv, values.p, column_ptr(c_sorted), nullptr),
               "zkb_lookup_sort");
            scheme.append_device_batch(LOOKUP_BATCH, column_ptr(c_sorted), n_sorted, n);
            proof.commitments[LOOKUP_BATCH] = scheme.commit(LOOKUP_BATCH);
            transcript(proof.commitments[LOOKUP_BATCH]);
            const value_type beta = transcript.template challenge<FieldType>(), gamma = transcript.template challenge<FieldType>();
            std::vector<value_type> lookup_alphas;
            for (std::size_t i = 0; i + 1 < lparts.size(); i++) lookup_alphas.push_back(transcript.template challenge<FieldType>());
            std::uint32_t bl[8], gl[8];
            beta.to_canonical_limbs(bl);
            gamma.to_canonical_limbs(gl);
            ck(zkb_lookup_grand_product(ctx(), fid, n, usable, (std::uint32_t)ni, inputs.p, (std::uint32_t)nv, values.p, (std::uint32_t)n_sorted,
                                        column_ptr(c_sorted), bl, gl, column_ptr(c_vl), nullptr),
               "zkb_lookup_grand_product");
            const value_type one_beta = value_type::one() + beta, part1 = one_beta * gamma;
            std::vector<expr> g_f, h_f;
            for (const auto &e : input_exprs) g_f.push_back(expr::constant(one_beta) * (expr::constant(gamma) + e));
            for (const auto &e : value_exprs) g_f.push_back(expr::constant(part1) + e + expr::constant(beta) * e.shifted(1));
            for (std::size_t i = 0; i < n_sorted; i++)
                h_f.push_back(expr::constant(part1) + var(c_sorted + i) + expr::constant(beta) * var(c_sorted + i, 1));
            std::vector<expr> gs, hs;
            std::size_t at = 0;
            for (auto sz : lparts) {
                gs.push_back(product(g_f, at, at + sz));
                hs.push_back(product(h_f, at, at + sz));
                at += sz;
            }
            if (at != g_f.size() || at != h_f.size()) throw std::logic_error("placeholder: lookup parts");
            const expr VL = var(c_vl), VL_shifted = var(c_vl, 1);
            F[3] = var(c_l0) * (one - VL);
            F[4] = var(c_qlast) * (VL * VL - VL);
            if (lparts.size() == 1) F[5] = (gs[0] * VL - hs[0] * VL_shifted) * neg_mask;
            else F[5] = running_products(c_vl, gs, hs, lookup_alphas, VL_shifted) * neg_mask;
            std::vector<expr> terms;
            for (std::size_t i = 0; i + 1 < n_sorted; i++) {
                const value_type a_i = transcript.template challenge<FieldType>();
                terms.push_back(expr::constant(a_i) * (var(c_sorted + i + 1) - var(c_sorted + i, (std::int32_t)usable)));
            }
            if (!terms.empty()) F[6] = total(terms) * var(c_l0);
        }
        // PERMUTATION_BATCH = V_P, its parts, V_L, its parts (prover.hpp:169-172)
        const std::size_t n_perm = parts + (lookup ? lparts.size() : 0);
        if (n_perm) {
            dbuf perm(n_perm * n * 32);
            copy(perm, 0, all, c_vp * n, parts * n);
            if (lookup) copy(perm, parts * n, all, c_vl * n, lparts.size() * n);
            scheme.append_device_batch(PERMUTATION_BATCH, perm.p, n_perm, n);
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
                terms.push_back(total(inner) * var(c.selector_column(g.selector_index)));
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
        auto rotated = [&](long long r) { return y * omega.pow((std::size_t)(((r % (long long)n) + (long long)n) % (long long)n)); };
        const auto rots = c.columns_rotations();
        for (std::size_t i = 0; i < nvar; i++)
            for (int r : rots[i]) scheme.append_eval_point(VARIABLE_VALUES_BATCH, i, rotated(r));
        if (n_perm) {
            scheme.append_eval_point(PERMUTATION_BATCH, y);
            if (npc) scheme.append_eval_point(PERMUTATION_BATCH, 0, rotated(1));
            if (lookup) {
                scheme.append_eval_point(PERMUTATION_BATCH, parts, rotated(1));
                scheme.append_eval_point(LOOKUP_BATCH, y);
                scheme.append_eval_point(LOOKUP_BATCH, rotated(1));
                scheme.append_eval_point(LOOKUP_BATCH, rotated((long long)usable));
            }
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
