"""Placeholder prover, commitment side, with every column resident on the device.

Mirrors the sequence of zk/snark/systems/plonk/placeholder/prover.hpp:133-217 around the commitment scheme:

    commit(VARIABLE_VALUES_BATCH)                      :141   -> transcript
    permutation argument: beta, gamma, V_P, F_0..F_2   permutation_argument.hpp:95-215 (permutation_parts == 1)
    (V_P cut into permutation_parts running products when max_quotient_chunks is set, :139-210)
    commit(PERMUTATION_BATCH)                          :170   -> transcript
    gates argument: theta, F_7                         gates_argument.hpp:133-217
    quotient: alpha_0..7, T = sum alpha_i F_i / Z      :260-283, split into chunks (:220-258)
    commit(QUOTIENT_BATCH)                             :202,:314-318 -> transcript
    challenge y, evaluation points, proof_eval         :205-213

The argument polynomials are expressions over the columns (oracle/placeholder.py uses the same nested-tuple form):
they are compiled to postfix programs and evaluated coset by coset on the extended domain by zkb_expr_eval; V_P is
zkb_permutation_grand_product; the quotient is one inverse transform, zkb_quotient_split and the chunk transforms.  The
fixed batch (identity / sigma permutation polynomials, q_last, q_blind, constants, selectors) is committed once per
circuit by the preprocessor (preprocessor.hpp:481-489) - `PlaceholderCircuit.preprocess`.

Not the reference's proof object: the transcript is seeded by the caller (the reference absorbs a hash of the
constraint system and the fixed-values commitment, prover.hpp:126-131), and lookups are off (constraint systems without
lookup gates: F_3..F_6 = 0, prover.hpp:286-297).
"""
import numpy as np

from . import capi
from .fields import FIELD_BY_NAME, omega
from .lpc import LpcCommitmentScheme

FIXED_VALUES_BATCH, VARIABLE_VALUES_BATCH, PERMUTATION_BATCH, QUOTIENT_BATCH, LOOKUP_BATCH = 0, 1, 2, 3, 4
F_PARTS = 8


# ---------------------------------------------------------------------------------------------------- expressions
def col(c, rot=0):
    return ("col", c, rot)


def const(v):
    return ("const", int(v))


def add(a, b):
    return ("add", a, b)


def sub(a, b):
    return ("sub", a, b)


def mul(a, b):
    return ("mul", a, b)


def product(factors):
    acc = factors[0]
    for f in factors[1:]:
        acc = mul(acc, f)
    return acc


def total(terms):
    acc = terms[0]
    for t in terms[1:]:
        acc = add(acc, t)
    return acc


def shift_expr(expr, r):
    """the expression with every column read r rows further (math::polynomial_shift of the polynomial it denotes)"""
    k = expr[0]
    if k == "col":
        return ("col", expr[1], expr[2] + r)
    if k == "const":
        return expr
    return (k,) + tuple(shift_expr(e, r) for e in expr[1:])


def degree(expr):
    """degree in units of (n - 1): every column counts 1 (expression_max_degree_visitor, gates_argument.hpp:161)"""
    k = expr[0]
    if k == "col":
        return 1
    if k == "const":
        return 0
    if k == "neg":
        return degree(expr[1])
    a, b = degree(expr[1]), degree(expr[2])
    return a + b if k == "mul" else max(a, b)


def compile_expr(expr, p, column_map=None):
    """nested tuples -> (postfix program for zkb_expr_eval, constants); column_map renumbers the columns"""
    prog, consts, index = [], [], {}

    def visit(e):
        k = e[0]
        if k == "col":
            prog.append((capi.EXPR_PUSH_COL, e[1] if column_map is None else column_map[e[1]], e[2]))
        elif k == "const":
            v = e[1] % p
            if v not in index:
                index[v] = len(consts)
                consts.append(v)
            prog.append((capi.EXPR_PUSH_CONST, index[v], 0))
        elif k == "neg":
            visit(e[1])
            prog.append((capi.EXPR_NEG, 0, 0))
        else:
            visit(e[1])
            visit(e[2])
            prog.append(({"add": capi.EXPR_ADD, "sub": capi.EXPR_SUB, "mul": capi.EXPR_MUL}[k], 0, 0))
    visit(expr)
    return prog, consts


def used_columns(exprs):
    """sorted column indices the expressions read"""
    cols = set()

    def visit(e):
        if e[0] == "col":
            cols.add(e[1])
        elif e[0] != "const":
            for sub_e in e[1:]:
                visit(sub_e)
    for e in exprs:
        visit(e)
    return sorted(cols)


def evaluate_on_extended_domain(ctx, field, columns, exprs, log_n, log_d, coefficients=None):
    """sum of the expressions as polynomial_dfs on the subgroup of size 2^(log_n + log_d): device tensor [2^(log_n+log_d), 8].
    columns: device tensor [ncols, 2^log_n, 8], evaluation form on the basic domain.  Coset by coset: one batched coset
    transform of the columns the expressions read, one zkb_expr_eval launch per expression, results interleaved with
    stride 2^log_d."""
    import torch
    F = FIELD_BY_NAME[field] if isinstance(field, str) else field
    n, D = 1 << log_n, 1 << log_d
    used = used_columns(exprs)
    if len(used) < columns.shape[0]:
        idx = torch.tensor(used, device=columns.device)
        columns = columns.index_select(0, idx)
        if coefficients is not None:
            coefficients = coefficients.index_select(0, idx)
    cmap = {c: i for i, c in enumerate(used)}
    if coefficients is None:
        coefficients = ctx.ntt(F.name, columns, log_n, inverse=True, out=torch.empty_like(columns))
    out = torch.empty((n * D, 8), dtype=torch.int32, device=columns.device)
    programs = [compile_expr(e, F.p, cmap) for e in exprs]
    w_ext = omega(F, log_n + log_d)
    work = torch.empty_like(coefficients)
    for j in range(D):
        if j == 0:
            cos = columns                      # coset 0 is the basic domain itself
        else:
            cos = ctx.ntt(F.name, coefficients, log_n, coset_shift=pow(w_ext, j, F.p), out=work)
        for k, (prog, consts) in enumerate(programs):
            ctx.expr_eval(F.name, cos, prog, consts, out, out_stride=D, out_offset=j, accumulate=k > 0)
    return out


# ---------------------------------------------------------------------------------------------------- circuit + prover
class PlaceholderCircuit:
    """The preprocessed public side of a circuit: column layout, gates, copy-constraint permutation and the special
    selectors.  Columns of the polynomial table in this order: witness | public_input | constant | selector.
    max_quotient_chunks (preprocessor.hpp:504,559-561): 0 = unbounded; otherwise the permutation product is cut into
    parts of max_quotient_chunks - 1 columns (permutation_partitions_num, preprocessor.hpp:78-87)."""

    def __init__(self, field, log_n, n_witness, n_public, n_constant, n_selector, gates, permuted_columns, s_id, s_sigma,
                 q_last, q_blind, lagrange_0, constants, selectors, usable_rows, max_quotient_chunks=0, lookup_tables=None,
                 lookup_gates=None):
        self.F = FIELD_BY_NAME[field] if isinstance(field, str) else field
        self.log_n, self.n = log_n, 1 << log_n
        self.n_witness, self.n_public, self.n_constant, self.n_selector = n_witness, n_public, n_constant, n_selector
        self.gates = gates                        # [(selector index, [constraint expr over table columns, ..]), ..]
        self.permuted_columns = list(permuted_columns)    # global indices into the table
        self.s_id, self.s_sigma = s_id, s_sigma   # device [len(permuted_columns), n, 8]
        self.q_last, self.q_blind, self.lagrange_0 = q_last, q_blind, lagrange_0    # device [n, 8]
        self.constants, self.selectors = constants, selectors                      # device [count, n, 8]
        self.usable_rows = usable_rows
        self.max_gates_degree = max([degree(c) for _, cs in gates for c in cs] + [0])
        self.max_quotient_chunks = max_quotient_chunks
        # lookup_tables: [(tag selector index, [option: [constant column index, ..], ..]), ..]  (plonk_lookup_table)
        # lookup_gates:  [(tag selector index, [(table id, 1-based, [input expr, ..]), ..]), ..] (plonk_lookup_gate)
        self.lookup_tables = list(lookup_tables or [])
        self.lookup_gates = list(lookup_gates or [])
        self._identity_ratios = None
        if max_quotient_chunks and max_quotient_chunks <= self.max_gates_degree:
            raise ValueError("max_quotient_chunks must exceed the gates' degree (preprocessor.hpp:559)")

    def identity_ratios(self, ctx):
        """S_id[i] = ratio_i * S_id[0] as polynomials when the identity polynomials are the preprocessor's
        (S_id[i][j] = delta^i omega^j, preprocessor.hpp:418-436): then the expressions read ONE identity column and the
        extended domain needs one transform for all of them.  Returns [ratio_i] (checked on the device once per
        circuit) or None when the columns are not of that form."""
        if self._identity_ratios is None:
            import numpy as np
            import torch
            p, npc = self.F.p, len(self.permuted_columns)
            first = [int.from_bytes(self.s_id[i, 0].cpu().numpy().view(np.uint32).tobytes(), "little") for i in range(npc)]
            ok = npc > 0 and first[0] % p != 0
            ratios = []
            if ok:
                inv0 = pow(first[0], p - 2, p)
                ratios = [v * inv0 % p for v in first]
                for i in range(1, npc):
                    if not torch.equal(ctx.poly_lincomb(self.F.name, self.s_id[0], self.n, [ratios[i]]), self.s_id[i]):
                        ok = False
                        break
            self._identity_ratios = ratios if ok else False
        return self._identity_ratios or None

    @property
    def table_width(self):
        return self.n_witness + self.n_public + self.n_constant + self.n_selector

    @property
    def permutation_parts(self):
        npc, mqc = len(self.permuted_columns), self.max_quotient_chunks
        if npc == 0:
            return 0
        return 1 if mqc == 0 else (npc + mqc - 2) // (mqc - 1)

    def selector_column(self, k):
        return self.n_witness + self.n_public + self.n_constant + k

    def constant_column(self, k):
        return self.n_witness + self.n_public + k

    # ---- lookup argument (lookup_argument.hpp:411-486, constraint_system.hpp:165-300)
    def lookup_value_exprs(self, theta, mask):
        """prepare_lookup_value: per table and option  mask * tag * (table id + sum_i theta^(i+1) constant_i)"""
        out = []
        for t_id, (tag, options) in enumerate(self.lookup_tables):
            tag_c = col(self.selector_column(tag))
            for option in options:
                v = mul(const(t_id + 1), tag_c)
                acc = theta
                for c in option:
                    v = add(v, mul(mul(const(acc), tag_c), col(self.constant_column(c))))
                    acc = acc * theta % self.F.p
                out.append(mul(v, mask))
        return out

    def lookup_input_exprs(self, theta):
        """prepare_lookup_input: per lookup constraint  selector * table id + sum_k theta^(k+1) selector * input_k"""
        out = []
        for tag, constraints in self.lookup_gates:
            sel = col(self.selector_column(tag))
            for table_id, inputs in constraints:
                l = mul(sel, const(table_id))
                acc = theta
                for e in inputs:
                    l = add(l, mul(mul(const(acc), sel), e))
                    acc = acc * theta % self.F.p
                out.append(l)
        return out

    def lookup_poly_degree_bound(self):
        d = 0
        if self.lookup_gates:
            for _, constraints in self.lookup_gates:
                for _, inputs in constraints:
                    d += max([degree(e) for e in inputs] + [0]) + 1
            for _, options in self.lookup_tables:
                d += 3 * len(options)
        return d

    def lookup_parts(self):
        """plonk_constraint_system::lookup_parts (constraint_system.hpp:256-300): how many sorted columns (= g / h factors,
        inputs first, then table options) go into every part of the lookup product"""
        n_inputs = sum(len(cs) for _, cs in self.lookup_gates)
        n_options = sum(len(o) for _, o in self.lookup_tables)
        mqc = self.max_quotient_chunks
        if mqc == 0:
            return [n_inputs + n_options]
        parts, chunk, part = [], 0, 0
        for _, constraints in self.lookup_gates:
            for _, inputs in constraints:
                d = max([degree(e) for e in inputs] + [0])
                if chunk + d + 1 >= mqc:
                    parts.append(part)
                    chunk, part = 0, 0
                chunk += d + 1
                part += 1
        for _, options in self.lookup_tables:
            for _ in options:
                if chunk + 3 >= mqc:
                    parts.append(part)
                    chunk, part = 0, 0
                chunk += 3
                part += 1
        parts.append(part)
        if 0 in parts:
            raise ValueError("max_quotient_chunks is too small for the lookup constraints (an empty lookup part)")
        return parts

    def columns_rotations(self):
        """per table column the sorted rotations the gates use, 0 always included (preprocessor.hpp:363-383, a std::set)"""
        rots = [{0} for _ in range(self.table_width)]

        def visit(e):
            if e[0] == "col":
                rots[e[1]].add(e[2])
            elif e[0] != "const":
                for sub_e in e[1:]:
                    visit(sub_e)
        for _, constraints in self.gates:
            for c in constraints:
                visit(c)
        if self.lookup_gates:
            for _, constraints in self.lookup_gates:
                for _, inputs in constraints:
                    for e in inputs:
                        visit(e)
            for tag, options in self.lookup_tables:          # tag and option columns are also read one row further
                rots[self.selector_column(tag)].add(1)
                for option in options:
                    for c in option:
                        rots[self.constant_column(c)].add(1)
        return [sorted(r) for r in rots]

    def fixed_batch(self):
        """identity | sigma permutation polynomials | q_last | q_blind | constants | selectors (preprocessor.hpp:481-489)"""
        import torch
        parts = [self.s_id, self.s_sigma, self.q_last.unsqueeze(0), self.q_blind.unsqueeze(0)]
        if self.n_constant:
            parts.append(self.constants)
        if self.n_selector:
            parts.append(self.selectors)
        return torch.cat(parts, dim=0)

    def quotient_chunks(self):
        """split_polynomial_size of prover.hpp:227-246"""
        n = self.n
        size = max((len(self.permuted_columns) + 2) * (n - 1), (self.lookup_poly_degree_bound() + 1) * (n - 1),
                   (self.max_gates_degree + 1) * (n - 1))
        size = (size + n - 1) // n
        if self.max_quotient_chunks and size > self.max_quotient_chunks:
            size = self.max_quotient_chunks
        return size


def placeholder_prove(ctx, circuit, hash_id, fri, witness, public_input, transcript, scheme=None, query=True, keep=None,
                      timings=None):
    """Runs the prover's commitment side for one assignment.  witness / public_input: device tensors [count, n, 8].
    Returns {"commitments": {batch: root}, "challenge": y, "eval_proof": lpc proof_eval result, ...}.  `keep` (a dict)
    receives the intermediate device tensors (V_P, F, T chunks) for tests; `timings` (a dict) receives wall-clock
    milliseconds per stage (each stage is followed by a device synchronize when it is given)."""
    import time
    import torch
    F, n, log_n = circuit.F, circuit.n, circuit.log_n
    p = F.p
    dev = witness.device
    commitments = {}
    t_last = [time.perf_counter()]

    def lap(name):
        if timings is not None:
            torch.cuda.synchronize()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + (now - t_last[0]) * 1e3
            t_last[0] = now

    if scheme is None:
        # the preprocessor's part (preprocessor.hpp:481-489): commit the fixed batch, seed the transcript with it
        # (prover.hpp:126-131), draw etha and record the fixed polynomials' values there
        scheme = LpcCommitmentScheme(ctx, F.name, hash_id, fri)
        scheme.append_to_batch(FIXED_VALUES_BATCH, circuit.fixed_batch())
        scheme.mark_batch_as_fixed(FIXED_VALUES_BATCH)
        commitments[FIXED_VALUES_BATCH] = scheme.commit(FIXED_VALUES_BATCH)
        transcript(commitments[FIXED_VALUES_BATCH])
        scheme.setup(transcript, None)
        scheme.fixed_batch_values(FIXED_VALUES_BATCH)
        lap("preprocess_fixed_batch")
    table = [witness, public_input]
    if circuit.n_constant:
        table.append(circuit.constants)
    if circuit.n_selector:
        table.append(circuit.selectors)
    tw = circuit.table_width
    npc = len(circuit.permuted_columns)
    usable = circuit.usable_rows
    # columns the expressions see: the table, then S_id, S_sigma, q_last, q_blind, L_0 (base_cols), then whatever the
    # arguments produce (V_P and its parts, V_L and its parts, the sorted lookup columns), numbered as they appear
    c_sid, c_ssg = tw, tw + npc
    c_qlast, c_qblind, c_l0 = tw + 2 * npc, tw + 2 * npc + 1, tw + 2 * npc + 2
    base_cols = torch.cat(table + [circuit.s_id, circuit.s_sigma, circuit.q_last.unsqueeze(0), circuit.q_blind.unsqueeze(0),
                                   circuit.lagrange_0.unsqueeze(0)], dim=0)
    n_base = base_cols.shape[0]
    extra, col_src = [], {}

    def new_col(t, src):
        """src = (batch, polynomial index in the batch): where a verifier finds this column's opened values"""
        extra.append(t)
        col_src[n_base + len(extra) - 1] = src
        return n_base + len(extra) - 1

    def column(i):
        return base_cols[i] if i < n_base else extra[i - n_base]

    def eval_basic(expr, out=None):
        """the expression on the basic domain (its values there are the reduce_dfs_polynomial_domain of upstream's product)"""
        used = used_columns([expr])
        cols = torch.stack([column(i) for i in used]) if used else torch.zeros((1, n, 8), dtype=torch.int32, device=dev)
        prog, consts = compile_expr(expr, p, {c: i for i, c in enumerate(used)})
        if out is None:
            out = torch.empty((n, 8), dtype=torch.int32, device=dev)
        return ctx.expr_eval(F.name, cols, prog, consts, out)

    gv = torch.empty((n, 8), dtype=torch.int32, device=dev)
    hv = torch.empty((n, 8), dtype=torch.int32, device=dev)

    def running_products(first_t, first_c, first_index, gs, hs, alphas_, shifted):
        """the part polynomials of a grand product cut into len(gs) parts (permutation_argument.hpp:194-215,
        lookup_argument.hpp:262-289): current = previous * g_i / h_i on the usable rows, rows beyond keep the grand
        product's values; returns (part tensors, the bracket sum alpha_i (prev g_i - cur h_i) + (prev g_last - shifted h_last))"""
        prev_t, prev_c, terms, made = first_t, first_c, [], []
        for i in range(len(gs) - 1):
            eval_basic(gs[i], gv)
            eval_basic(hs[i], hv)
            ctx.vec(F.name, capi.VEC_MUL, prev_t, gv, out=gv)
            ctx.batch_inverse(F.name, hv, out=hv)
            ctx.vec(F.name, capi.VEC_MUL, gv, hv, out=gv)
            cur = first_t.clone()
            cur[:usable] = gv[:usable]
            cur_c = new_col(cur, (PERMUTATION_BATCH, first_index + 1 + i))
            made.append(cur)
            terms.append(mul(const(alphas_[i]), sub(mul(col(prev_c), gs[i]), mul(col(cur_c), hs[i]))))
            prev_t, prev_c = cur, cur_c
        terms.append(sub(mul(col(prev_c), gs[-1]), mul(shifted, hs[-1])))
        return made, total(terms)

    # 2. witness and public-input columns
    scheme.append_to_batch(VARIABLE_VALUES_BATCH, base_cols[:circuit.n_witness + circuit.n_public])
    commitments[VARIABLE_VALUES_BATCH] = scheme.commit(VARIABLE_VALUES_BATCH)
    transcript(commitments[VARIABLE_VALUES_BATCH])
    lap("commit_variable_values")
    f_exprs = {}
    one = const(1)
    mask = sub(sub(one, col(c_qlast)), col(c_qblind))
    neg_mask = sub(add(col(c_qlast), col(c_qblind)), one)
    perm_polys = []
    # 4. permutation argument (permutation_argument.hpp:95-215)
    if npc:
        beta, gamma = transcript.challenge(p), transcript.challenge(p)
        cols = base_cols.index_select(0, torch.tensor(circuit.permuted_columns, device=dev))
        v_p = ctx.permutation_grand_product(F.name, cols, circuit.s_id, circuit.s_sigma, beta, gamma)
        c_vp = new_col(v_p, (PERMUTATION_BATCH, 0))
        perm_polys.append(v_p)
        ratios = circuit.identity_ratios(ctx)
        if ratios is None:
            g_f = [add(add(mul(const(beta), col(c_sid + i)), const(gamma)), col(circuit.permuted_columns[i])) for i in range(npc)]
        else:       # beta S_id[i] = (beta delta^i) S_id[0]: one identity column on the extended domain instead of npc
            g_f = [add(add(mul(const(beta * ratios[i] % p), col(c_sid)), const(gamma)), col(circuit.permuted_columns[i])) for i in range(npc)]
        h_f = [add(add(mul(const(beta), col(c_ssg + i)), const(gamma)), col(circuit.permuted_columns[i])) for i in range(npc)]
        group = npc if circuit.max_quotient_chunks == 0 else circuit.max_quotient_chunks - 1
        gs = [product(g_f[i:i + group]) for i in range(0, npc, group)]
        hs = [product(h_f[i:i + group]) for i in range(0, npc, group)]
        parts = len(gs)
        assert parts == circuit.permutation_parts
        perm_alphas = [transcript.challenge(p) for _ in range(parts - 1)]
        f_exprs[0] = mul(sub(one, col(c_vp)), col(c_l0))
        if parts == 1:
            f_exprs[1] = mul(mask, sub(mul(col(c_vp, 1), hs[0]), mul(col(c_vp), gs[0])))
        else:
            made, bracket = running_products(v_p, c_vp, 0, gs, hs, perm_alphas, col(c_vp, 1))
            perm_polys += made
            f_exprs[1] = mul(bracket, neg_mask)
        f_exprs[2] = mul(col(c_qlast), sub(mul(col(c_vp), col(c_vp)), col(c_vp)))
        lap("permutation_argument")
    # 5. lookup argument (lookup_argument.hpp:153-325)
    lookup = circuit.lookup_gates and circuit.lookup_tables
    v_l_index = None
    sorted_cols = []
    if lookup:
        theta = transcript.challenge(p)
        value_exprs, input_exprs = circuit.lookup_value_exprs(theta, mask), circuit.lookup_input_exprs(theta)
        values = torch.stack([eval_basic(e) for e in value_exprs])          # reduced to the basic domain (:176-186)
        inputs = torch.stack([eval_basic(e) for e in input_exprs])
        srt = ctx.lookup_sort(F.name, inputs, values, usable)
        scheme.append_to_batch(LOOKUP_BATCH, srt)
        commitments[LOOKUP_BATCH] = scheme.commit(LOOKUP_BATCH)
        transcript(commitments[LOOKUP_BATCH])
        sorted_cols = [new_col(srt[i], (LOOKUP_BATCH, i)) for i in range(srt.shape[0])]
        lbeta, lgamma = transcript.challenge(p), transcript.challenge(p)
        part_sizes = circuit.lookup_parts()
        lookup_alphas = [transcript.challenge(p) for _ in range(len(part_sizes) - 1)]
        v_l = ctx.lookup_grand_product(F.name, inputs, values, srt, lbeta, lgamma, usable)
        v_l_index = len(perm_polys)
        c_vl = new_col(v_l, (PERMUTATION_BATCH, v_l_index))
        perm_polys.append(v_l)
        one_beta, part1 = (1 + lbeta) % p, (1 + lbeta) * lgamma % p
        g_f = [mul(const(one_beta), add(const(lgamma), e)) for e in input_exprs]
        g_f += [add(add(const(part1), e), mul(const(lbeta), shift_expr(e, 1))) for e in value_exprs]
        h_f = [add(add(const(part1), col(c)), mul(const(lbeta), col(c, 1))) for c in sorted_cols]
        assert sum(part_sizes) == len(g_f) == len(h_f)
        gs, hs, o = [], [], 0
        for sz in part_sizes:
            gs.append(product(g_f[o:o + sz]))
            hs.append(product(h_f[o:o + sz]))
            o += sz
        f_exprs[3] = mul(col(c_l0), sub(one, col(c_vl)))
        f_exprs[4] = mul(col(c_qlast), sub(mul(col(c_vl), col(c_vl)), col(c_vl)))
        if len(part_sizes) == 1:
            f_exprs[5] = mul(sub(mul(gs[0], col(c_vl)), mul(hs[0], col(c_vl, 1))), neg_mask)
        else:
            made, bracket = running_products(v_l, c_vl, v_l_index, gs, hs, lookup_alphas, col(c_vl, 1))
            perm_polys += made
            f_exprs[5] = mul(bracket, neg_mask)
        terms = []
        for i in range(len(sorted_cols) - 1):          # sorted[i+1](X) = sorted[i](omega^usable X) at the first row (:291-299)
            a_i = transcript.challenge(p)
            terms.append(mul(const(a_i), sub(col(sorted_cols[i + 1]), col(sorted_cols[i], usable))))
        if terms:
            f_exprs[6] = mul(total(terms), col(c_l0))
        lap("lookup_argument")
    if perm_polys:
        scheme.append_to_batch(PERMUTATION_BATCH, torch.stack(perm_polys))
        commitments[PERMUTATION_BATCH] = scheme.commit(PERMUTATION_BATCH)
        transcript(commitments[PERMUTATION_BATCH])
        lap("commit_permutation")
    # 6. circuit satisfiability (gates_argument.hpp:133-217)
    if circuit.gates:
        theta_g = transcript.challenge(p)
        theta_acc, terms = 1, []
        for sel, constraints in circuit.gates:
            inner = []
            for c in constraints:
                inner.append(mul(c, const(theta_acc)))
                theta_acc = theta_acc * theta_g % p
            terms.append(mul(total(inner), col(circuit.selector_column(sel))))
        f_exprs[7] = mul(total(terms), mask)
    # 7. quotient: alphas, F consolidated, division by Z, split, commit (prover.hpp:220-283)
    alphas = [transcript.challenge(p) for _ in range(F_PARTS)]
    parts_exprs = [mul(f_exprs[i], const(alphas[i])) for i in sorted(f_exprs)]
    deg = max([degree(e) for e in f_exprs.values()] + [1])
    log_d = 1
    while (1 << log_d) * n <= deg * (n - 1):      # the extended domain must hold degree deg (n - 1)
        log_d += 1
    all_cols = torch.cat([base_cols] + ([torch.stack(extra)] if extra else [torch.zeros((1, n, 8), dtype=torch.int32, device=dev)]), dim=0)
    f_dfs = evaluate_on_extended_domain(ctx, F, all_cols, parts_exprs, log_n, log_d)
    lap("argument_polynomials_on_extended_domain")
    f_coeff = ctx.ntt(F.name, (f_dfs.clone() if keep is not None else f_dfs).unsqueeze(0), log_n + log_d, inverse=True)[0]   # in place
    nchunks = circuit.quotient_chunks()
    t_chunks = ctx.quotient_split(F.name, f_coeff, log_n, log_n + log_d, nchunks)
    lap("quotient")
    scheme.append_to_batch(QUOTIENT_BATCH, t_chunks)
    commitments[QUOTIENT_BATCH] = scheme.commit(QUOTIENT_BATCH)
    transcript(commitments[QUOTIENT_BATCH])
    lap("commit_quotient")
    # 8. evaluation points (generate_evaluation_points, prover.hpp:346-416) and the evaluation proof
    y = transcript.challenge(p)
    w = omega(F, log_n)

    def rotated(r):
        return y * pow(w, r % n, p) % p
    rots = circuit.columns_rotations()
    nvar = circuit.n_witness + circuit.n_public
    for i in range(nvar):
        for r in rots[i]:
            scheme.append_eval_point(VARIABLE_VALUES_BATCH, rotated(r), poly=i)
    if perm_polys:
        scheme.append_eval_point(PERMUTATION_BATCH, y)
        if npc:
            scheme.append_eval_point(PERMUTATION_BATCH, rotated(1), poly=0)
        if lookup:
            scheme.append_eval_point(PERMUTATION_BATCH, rotated(1), poly=v_l_index)
            scheme.append_eval_point(LOOKUP_BATCH, y)
            scheme.append_eval_point(LOOKUP_BATCH, rotated(1))
            scheme.append_eval_point(LOOKUP_BATCH, rotated(usable))
    scheme.append_eval_point(QUOTIENT_BATCH, y)
    if scheme.has_batch(FIXED_VALUES_BATCH):
        start = 2 * npc + 2
        for i in range(start):
            scheme.append_eval_point(FIXED_VALUES_BATCH, y, poly=i)
        scheme.append_eval_point(FIXED_VALUES_BATCH, rotated(1), poly=start - 2)
        scheme.append_eval_point(FIXED_VALUES_BATCH, rotated(1), poly=start - 1)
        for ind in range(circuit.n_constant + circuit.n_selector):
            for r in rots[nvar + ind]:
                scheme.append_eval_point(FIXED_VALUES_BATCH, rotated(r), poly=start + ind)
    if keep is not None:
        keep.update(v_p=perm_polys[0] if npc else None, perm_polys=perm_polys, f_dfs=f_dfs, f_coeff=f_coeff, t_chunks=t_chunks,
                    alphas=alphas, log_d=log_d, f_exprs=f_exprs, all_cols=all_cols, n_base=n_base, sorted_cols=sorted_cols,
                    v_l_index=v_l_index, column_source=dict(col_src))
    eval_proof = scheme.proof_eval(transcript, query=query)
    lap("proof_eval")
    return {"commitments": commitments, "challenge": y, "eval_proof": eval_proof, "quotient_chunks": nchunks, "log_d": log_d,
            "etha": scheme._etha, "fixed_values": scheme._fixed_values, "scheme": scheme, "rotations": rots}
