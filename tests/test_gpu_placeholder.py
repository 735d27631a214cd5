"""GPU parity of the Placeholder argument builders and of the resident-column prover flow (row a12),
through the C ABI: expression evaluation over the extended domain, quotient division + split, lookup sort, and
placeholder_prove checked by the verifier's identity at the challenge point.

Reference: zk/snark/systems/plonk/placeholder/gates_argument.hpp:76-217, permutation_argument.hpp:133-215,
lookup_argument.hpp:565-633, prover.hpp:133-283, verifier.hpp:233-330."""
import random

import numpy as np
import pytest

from oracle import fields, ntt, placeholder

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from crypto3_zk_b200 import Context
    c = Context(0)
    yield c
    c.close()


def to_arr(vals, limbs=8):
    return fields.ints_to_u32_array(vals, limbs)


def from_arr(a):
    return fields.u32_array_to_ints(np.asarray(a).reshape(-1, a.shape[-1]))


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


def cols_to_dev(cols):
    n = len(cols[0])
    return dev(to_arr([x for c in cols for x in c]).reshape(len(cols), n, 8))


def cols_from_dev(t):
    flat = from_arr(host(t))
    n = t.shape[-2]
    return [flat[i * n:(i + 1) * n] for i in range(len(flat) // n)]


# ------------------------------------------------------------------------------------------------ expression evaluation
@pytest.mark.parametrize("F", [fields.PALLAS_FP, fields.BLS12_381_FR, fields.BN254_FR], ids=lambda f: f.name)
@pytest.mark.parametrize("log_n", [2, 5])
def test_expr_eval_on_extended_domain_vs_oracle(ctx, F, log_n):
    """a degree-4 expression with rotations (+1, -1, +2), constants, neg and every operator: the device's coset-by-coset
    point evaluation equals the oracle's exact polynomial arithmetic, in evaluation form on the size-8n subgroup"""
    from crypto3_zk_b200 import placeholder as P
    n = 1 << log_n
    p = F.p
    cols = [fields.random_elements(F, n, 70 + i) for i in range(4)]
    c = P.col
    e1 = P.mul(P.mul(c(0), c(1, 1)), P.sub(c(2, -1), P.const(5)))
    e2 = ("neg", P.mul(P.add(c(3, 2), P.const(p - 1)), P.mul(c(0), P.mul(c(0, 1), c(2)))))
    e3 = P.add(P.const(7), c(1))
    got = from_arr(host(P.evaluate_on_extended_domain(ctx, F.name, cols_to_dev(cols), [e1, e2, e3], log_n, 3)))
    want = placeholder.expr_dfs(("add", ("add", e1, e2), e3), cols, F, 8 * n)
    assert got == want


def test_expr_eval_program_validation(ctx):
    import torch
    from crypto3_zk_b200 import capi
    F = fields.PALLAS_FP
    cols = dev(to_arr(fields.random_elements(F, 8, 1)).reshape(1, 8, 8))
    out = torch.zeros((8, 8), dtype=torch.int32, device="cuda")
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.expr_eval(F.name, cols, [(capi.EXPR_PUSH_COL, 1, 0)], [], out)                       # column out of range
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.expr_eval(F.name, cols, [(capi.EXPR_PUSH_COL, 0, 0), (capi.EXPR_ADD, 0, 0)], [], out)   # underflow
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.expr_eval(F.name, cols, [(capi.EXPR_PUSH_COL, 0, 0), (capi.EXPR_PUSH_COL, 0, 0)], [], out)   # two values left
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.expr_eval(F.name, cols, [(capi.EXPR_PUSH_CONST, 0, 0)], [F.p], out)                  # non-canonical constant
    with pytest.raises(capi.ZkbError):
        ctx.expr_eval(F.name, cols, [(capi.EXPR_PUSH_COL, 0, 0)] * 17 + [(capi.EXPR_ADD, 0, 0)] * 16, [], out)   # > 16 slots
    # a plain copy with a negative rotation and accumulate
    ctx.expr_eval(F.name, cols, [(capi.EXPR_PUSH_COL, 0, -3)], [], out)
    ctx.expr_eval(F.name, cols, [(capi.EXPR_PUSH_COL, 0, -3)], [], out, accumulate=True)
    x = from_arr(host(cols))
    assert from_arr(host(out)) == [2 * x[(i - 3) % 8] % F.p for i in range(8)]


# ------------------------------------------------------------------------------------------------ quotient
@pytest.mark.parametrize("log_n,log_ext,nchunks", [(3, 6, 7), (3, 6, 9), (4, 4, 1), (6, 8, 3), (10, 13, 7)])
def test_quotient_split_vs_oracle(ctx, log_n, log_ext, nchunks):
    """T = F / (X^n - 1) with the remainder dropped, chunks in evaluation form; more chunks than coefficients -> zero
    chunks; a random F (with remainder) and an exact multiple of Z"""
    F = fields.PALLAS_FQ
    p, n, E = F.p, 1 << log_n, 1 << log_ext
    f = fields.random_elements(F, E, 5 + log_ext)
    want, _ = placeholder.quotient_split(f, n, nchunks, F)
    got = cols_from_dev(ctx.quotient_split(F.name, dev(to_arr(f)), log_n, log_ext, nchunks))
    assert got == want
    if E > n:
        t = fields.random_elements(F, E - n, 9)
        fz = [0] * E
        for i, v in enumerate(t):
            fz[i + n] = (fz[i + n] + v) % p
            fz[i] = (fz[i] - v) % p
        got = cols_from_dev(ctx.quotient_split(F.name, dev(to_arr(fz)), log_n, log_ext, (E - n) // n))
        back = [v for ch in got for v in ntt.dfs_coefficients(ch, F)]
        assert back == t


# ------------------------------------------------------------------------------------------------ lookup sort
def _lookup_case(F, n, usable, n_inputs, n_values, distinct, seed, zero_runs=True, inner_zeros=True, pad_random=True):
    rnd = random.Random(seed)
    vals = [rnd.randrange(1, F.p) for _ in range(distinct)]
    table = []
    if zero_runs:
        table += [0] * rnd.randrange(1, 4)
    for k, v in enumerate(vals):
        table += [v] * rnd.randrange(1, 4)
        if zero_runs and inner_zeros and k == distinct // 2:
            table += [0, 0]
    table = (table + [0] * (n_values * usable))[:n_values * usable]
    if not zero_runs:                      # no trailing zeros either: repeat the last value
        table = [v if v else vals[-1] for v in table]
    values = []
    for i in range(n_values):
        col = table[i * usable:(i + 1) * usable]
        values.append(col + [rnd.randrange(F.p) if pad_random else 0 for _ in range(n - usable)])      # rows past the usable ones are ignored by the sort
    present = sorted(set(table))
    inputs = [[rnd.choice(present) for _ in range(usable)] + [rnd.randrange(F.p) for _ in range(n - usable)] for _ in range(n_inputs)]
    return inputs, values


@pytest.mark.parametrize("n,usable,ni,nv,distinct,zero_runs", [
    (8, 5, 1, 1, 2, True), (32, 27, 2, 1, 8, True), (32, 31, 1, 2, 20, False), (256, 250, 3, 2, 100, True),
    (4096, 4000, 2, 1, 1500, True), (1 << 14, (1 << 14) - 5, 4, 2, 9000, True), (64, 60, 0, 1, 20, True)])
def test_lookup_sort_vs_oracle(ctx, n, usable, ni, nv, distinct, zero_runs):
    """sort_polynomials (lookup_argument.hpp:565-633) bit for bit: zero handling (single leading zero, inner zero runs once,
    trailing zero run dropped), multi-column table, inputs counted into their table value, link cells at usable_rows"""
    F = fields.BLS12_381_FR
    inputs, values = _lookup_case(F, n, usable, ni, nv, distinct, 100 + n + ni)
    want = placeholder.sort_polynomials(inputs, values, n, usable)
    got = cols_from_dev(ctx.lookup_sort(F.name, cols_to_dev(inputs) if ni else None, cols_to_dev(values), usable))
    assert got == want


def test_lookup_sort_closes_the_grand_product(ctx):
    """device sort -> device V_L: the product closes at the last usable row (the relation the lookup argument proves)"""
    F = fields.PALLAS_FP
    n, usable = 1 << 12, (1 << 12) - 7
    inputs, values = _lookup_case(F, n, usable, 3, 1, 1200, 77, inner_zeros=False, pad_random=False)     # zeros as padding only and a zero tail (the table's selector is off there)
    d_in, d_val = cols_to_dev(inputs), cols_to_dev(values)
    srt = ctx.lookup_sort(F.name, d_in, d_val, usable)
    beta, gamma = fields.random_elements(F, 2, 6)
    V = from_arr(host(ctx.lookup_grand_product(F.name, d_in, d_val, srt, beta, gamma, usable)))
    assert V[0] == 1 and V[usable] == 1 and V[usable // 2] != 1


def test_lookup_sort_rejects_an_input_missing_from_the_table(ctx):
    from crypto3_zk_b200 import capi
    F = fields.PALLAS_FP
    n, usable = 64, 60
    inputs, values = _lookup_case(F, n, usable, 1, 1, 10, 5)
    # enough foreign values that the output overflows its capacity (the reference asserts on the first one)
    inputs[0][:usable] = [F.p - 1 - k for k in range(usable)]
    values[0][:usable] = [k + 1 for k in range(usable)]
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.lookup_sort(F.name, cols_to_dev(inputs), cols_to_dev(values), usable)


# ------------------------------------------------------------------------------------------------ the prover flow
class _Transcript:
    def __init__(self, hid, init):
        from crypto3_zk_b200.transcript import FiatShamirSequential
        self.t = FiatShamirSequential(hid, init)

    def __call__(self, data):
        self.t(data)

    def challenge(self, modulus):
        return self.t.challenge(modulus)

    def int_challenge(self, bits=32):
        return self.t.int_challenge(bits)

    @property
    def state(self):
        return self.t.state


@pytest.mark.parametrize("log_n,triples,mqc,lookup", [(4, 1, 0, False), (6, 2, 0, False), (6, 2, 4, False), (5, 3, 3, False), (10, 1, 0, False),
                                                     (10, 4, 4, False), (5, 1, 0, True), (6, 1, 5, True), (10, 2, 5, True)])
def test_placeholder_prove_resident_columns(ctx, log_n, triples, mqc, lookup):
    """placeholder_prover's commitment side with every column resident (prover.hpp:133-217).  Checked three ways:
    (1) V_P closes and the consolidated F equals the oracle's exact polynomial arithmetic (small sizes);
    (2) the quotient is exact (F vanishes on the basic domain) and T's chunks equal the oracle's division + split;
    (3) the verifier's identity sum alpha_i F_i(y) = Z(y) sum_k y^(n k) T_k(y) holds on the values the evaluation proof
        opens (placeholder/verifier.hpp:233-330), and the lpc proof verifies."""
    from crypto3_zk_b200 import workloads as W
    from crypto3_zk_b200 import placeholder as P
    from crypto3_zk_b200.lpc import FriParams
    F = fields.PALLAS_FP
    p, n = F.p, 1 << log_n
    circuit, witness, public = W.placeholder_chain_circuit(ctx, F.name, log_n, triples=triples, seed=log_n, max_quotient_chunks=mqc, lookup=lookup)
    fri = FriParams.with_max_step_one(log_n, 4, 3)
    tr = _Transcript(0, b"placeholder-test")
    keep = {}
    res = P.placeholder_prove(ctx, circuit, 0, fri, witness, public, tr, keep=keep)
    usable = circuit.usable_rows
    v_p = from_arr(host(keep["v_p"]))
    assert v_p[0] == 1 and v_p[usable] == 1
    all_cols = cols_from_dev(keep["all_cols"])
    alphas, f_exprs, log_d = keep["alphas"], keep["f_exprs"], keep["log_d"]
    total = None
    for i in sorted(f_exprs):
        term = ("mul", f_exprs[i], ("const", alphas[i]))
        total = term if total is None else ("add", total, term)
    f_coeff = from_arr(host(keep["f_coeff"]))
    if log_n <= 6:
        want = placeholder.expr_polynomial(total, all_cols, F)
        assert f_coeff == want + [0] * (len(f_coeff) - len(want))
        assert from_arr(host(keep["f_dfs"])) == placeholder.expr_dfs(total, all_cols, F, n << log_d)
    nchunks = res["quotient_chunks"]
    want_chunks, rem = placeholder.quotient_split(f_coeff, n, nchunks, F)
    assert rem == [0] * n, "F does not vanish on the basic domain"
    assert cols_from_dev(keep["t_chunks"]) == want_chunks
    # every coefficient of T is covered by the chunks (split_polynomial_size, prover.hpp:227-246)
    deg = max(i for i, v in enumerate(f_coeff) if v)
    assert deg - n < nchunks * n
    # (3) the verifier's identity at y on the opened values
    y = res["challenge"]
    z = res["eval_proof"]["z"]
    npc, tw = len(circuit.permuted_columns), circuit.table_width
    nvar = circuit.n_witness + circuit.n_public

    rots = res["rotations"]
    assert rots[0] == [0, 1] and rots[1] == [0]          # a_k is read on the next row by the rotation gate
    src = keep["column_source"]
    batch_points = {P.PERMUTATION_BATCH: [0, 1], P.LOOKUP_BATCH: [0, 1, usable]}   # rotations in the order they were appended

    def value_of(c, rot):
        if c < nvar:                                      # opened at y omega^r for the rotations the gates use, ascending
            return z[P.VARIABLE_VALUES_BATCH][c][rots[c].index(rot)]
        if c < tw:                                        # constants / selectors follow S_id, S_sigma, q_last, q_blind
            return z[P.FIXED_VALUES_BATCH][2 * npc + 2 + (c - nvar)][rots[c].index(rot)]
        k = c - tw
        if k < 2 * npc:                                   # S_id, S_sigma
            assert rot == 0
            return z[P.FIXED_VALUES_BATCH][k][0]
        if k < 2 * npc + 2:                               # q_last, q_blind: opened at y and y omega
            return z[P.FIXED_VALUES_BATCH][k][rot]
        if k == 2 * npc + 2:                              # L_0(y) = (y^n - 1) / (n (y - 1))
            assert rot == 0
            return (pow(y, n, p) - 1) * pow(n * (y - 1) % p, p - 2, p) % p
        batch, poly = src[c]                              # V_P, V_L and their parts; the sorted lookup columns
        return z[batch][poly][batch_points[batch].index(rot)]

    assert len(z[P.PERMUTATION_BATCH]) == circuit.permutation_parts + (len(circuit.lookup_parts()) if lookup else 0)
    if not lookup:
        assert nchunks == (mqc if mqc else npc + 2)
    else:
        assert {3, 4, 5, 6} <= set(keep["f_exprs"]) and len(keep["sorted_cols"]) == 2
        v_l = from_arr(host(keep["perm_polys"][keep["v_l_index"]]))
        assert v_l[0] == 1 and v_l[usable] == 1          # the lookup grand product closes (lookup_argument.hpp:224)

    lhs = placeholder.expr_at_point(total, value_of, p)
    t_y = sum(pow(y, n * k, p) * z[P.QUOTIENT_BATCH][k][0] for k in range(nchunks)) % p
    assert lhs == t_y * (pow(y, n, p) - 1) % p
    # the evaluation proof itself
    assert all(r == 0 for r in res["eval_proof"]["remainders"])
    if lookup:
        # a looked-up pair outside the table: the sort refuses it like upstream's assert
        from crypto3_zk_b200 import capi
        bad = witness.clone()
        bad[3 * triples + 1, 2, 0] += 1                   # v != u^2 on one row
        with pytest.raises(capi.ZkbInvalidArgument):
            P.placeholder_prove(ctx, circuit, 0, fri, bad, public, _Transcript(0, b"placeholder-test"), query=False)
    # a violated copy constraint: V_P no longer closes, the quotient has a remainder
    witness2 = witness.clone()
    witness2[2, 1, 0] ^= 1
    keep2 = {}
    P.placeholder_prove(ctx, circuit, 0, fri, witness2, public, _Transcript(0, b"placeholder-test"), keep=keep2, query=False)
    _, rem = placeholder.quotient_split(from_arr(host(keep2["f_coeff"])), n, nchunks, F)
    assert any(rem)
