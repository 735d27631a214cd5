"""GPU parity tests: the CUDA path, called through the C ABI (crypto3_zk_b200.api -> libzkb200.so),
against the CPU oracle on the same seeded inputs, plus size-independent properties at full sizes.
Bit-exact comparisons everywhere (integer arithmetic)."""
import json
import os
import random

import numpy as np
import pytest

from oracle import curves, fields, fri, hashes, ntt

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "bls12_381_ipp2.json")


@pytest.fixture(scope="module")
def ctx():
    from crypto3_zk_b200 import Context, build
    build.build()
    c = Context(0)
    yield c
    c.close()


def to_arr(vals, limbs=8):
    return fields.ints_to_u32_array(vals, limbs)


def from_arr(a):
    return fields.u32_array_to_ints(np.asarray(a).reshape(-1, np.asarray(a).shape[-1]))


def dev(a):
    import torch
    return torch.from_numpy(a.view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


# ------------------------------------------------------------------------------------------ NTT
@pytest.mark.parametrize("F", fields.NTT_FIELDS, ids=lambda f: f.name)
@pytest.mark.parametrize("log_n,batch", [(0, 2), (1, 1), (2, 3), (3, 11), (5, 8), (8, 9), (9, 2), (10, 3), (12, 2), (13, 1)])
def test_ntt_vs_oracle(ctx, F, log_n, batch):
    n = 1 << log_n
    polys = [fields.random_elements(F, n, 1000 * log_n + b) for b in range(batch)]
    a = to_arr([v for p in polys for v in p]).reshape(batch, n, 8)
    want = []
    for p in polys:
        w = list(p)
        ntt.EvaluationDomain(F, n).fft(w)
        want.append(w)
    flat_want = [v for w in want for v in w]
    # host buffers through the ABI
    got = ctx.ntt(F.name, a.copy(), log_n)
    assert from_arr(got) == flat_want
    # device-resident buffers, in place, then inverse back
    d = dev(a)
    ctx.ntt(F.name, d, log_n)
    assert from_arr(host(d)) == flat_want
    ctx.ntt(F.name, d, log_n, inverse=True)
    assert from_arr(host(d)) == [v for p in polys for v in p]


@pytest.mark.parametrize("F", [fields.BLS12_381_FR, fields.BN254_FR], ids=lambda f: f.name)
@pytest.mark.parametrize("log_n", [4, 8, 10, 12])
def test_coset_ntt_vs_oracle(ctx, F, log_n):
    """multiply_by_coset(a, g); fft(a)  and  inverse_fft(a); multiply_by_coset(a, g^-1)
    (r1cs_to_qap.hpp:266-276, 310-315)."""
    n = 1 << log_n
    p = fields.random_elements(F, n, 77 + log_n)
    want = ntt.coset_fft(p, F, F.g)
    got = ctx.ntt(F.name, to_arr(p).reshape(1, n, 8), log_n, coset_shift=F.g)
    assert from_arr(got) == want
    back = ctx.ntt(F.name, got, log_n, inverse=True, coset_shift=F.g)
    assert from_arr(back) == p
    assert from_arr(back) == ntt.coset_inverse_fft(want, F, F.g)


@pytest.mark.parametrize("log_n", [16, 17, 20, 22, 24])
def test_ntt_large_properties(ctx, log_n):
    """Full-size transforms: delta -> geometric sequence (exact closed form), sparse-input spot
    checks against the DFT definition, and inverse(forward(x)) == x on random data."""
    import torch
    F = fields.BLS12_381_FR if log_n != 22 else fields.BN254_FR
    n = 1 << log_n
    w = F.omega(log_n)
    rnd = random.Random(log_n)
    # sparse input: 4 non-zeros
    pos = [0, 1, rnd.randrange(n), n - 1]
    val = [rnd.randrange(F.p) for _ in pos]
    a = np.zeros((1, n, 8), dtype=np.uint32)
    for j, v in zip(pos, val):
        a[0, j] = to_arr([v])[0]
    d = dev(a)
    ctx.ntt(F.name, d, log_n)
    out = host(d)[0]
    ks = [0, 1, 2, n // 2, n - 1] + [rnd.randrange(n) for _ in range(40)]
    for k in ks:
        want = sum(v * pow(w, j * k, F.p) for j, v in zip(pos, val)) % F.p
        assert from_arr(out[k:k + 1])[0] == want, k
    # random data round trip (device generated: 253-bit values are < every modulus here)
    g = torch.Generator(device="cuda").manual_seed(log_n)
    x = torch.randint(-2**31, 2**31 - 1, (1, n, 8), dtype=torch.int32, device="cuda", generator=g)
    x[..., 7] &= 0x0FFFFFFF
    y = x.clone()
    ctx.ntt(F.name, y, log_n)
    assert not torch.equal(x, y)
    ctx.ntt(F.name, y, log_n, inverse=True)
    assert torch.equal(x, y)


def test_ntt_errors(ctx):
    from crypto3_zk_b200 import capi
    a = np.zeros((1, 4, 8), dtype=np.uint32)
    with pytest.raises(ValueError):                       # std::invalid_argument analogue
        ctx.ntt("bn254_fr", np.zeros((1, 8), dtype=np.uint32), 29)   # beyond two-adicity 28
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.ntt("bls12_381_fr", a, 3)                     # size mismatch
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.ntt("bls12_381_fq", a, 2)                     # not an NTT field
    # empty batch is a no-op
    ctx.ntt("bls12_381_fr", np.zeros((0, 4, 8), dtype=np.uint32), 2)


# ------------------------------------------------------------------------------------------ LDE
@pytest.mark.parametrize("F", [fields.PALLAS_FQ, fields.PALLAS_FP, fields.BLS12_381_FR], ids=lambda f: f.name)
@pytest.mark.parametrize("log_in,log_out,batch", [(1, 1, 2), (1, 4, 3), (3, 6, 9), (7, 10, 2), (9, 12, 3), (4, 13, 1)])
def test_lde_vs_oracle(ctx, F, log_in, log_out, batch):
    """polynomial_dfs::resize (basic_fri.hpp:451-455)."""
    polys = [fields.random_elements(F, 1 << log_in, 31 * log_out + b) for b in range(batch)]
    a = to_arr([v for p in polys for v in p]).reshape(batch, 1 << log_in, 8)
    got = ctx.lde(F.name, a, log_in, log_out)
    want = [v for p in polys for v in ntt.dfs_resize(p, F, 1 << log_out)]
    assert from_arr(got) == want
    got_d = ctx.lde(F.name, dev(a), log_in, log_out)
    assert from_arr(host(got_d)) == want


def test_lde_config2_shape_properties(ctx):
    """BASELINE config #2 geometry (2^20 -> 2^23, Pallas scalar field), reduced batch: the input
    evaluations reappear at stride 8, and a random off-subgroup point of the extended domain equals
    Horner evaluation of the coefficients."""
    import torch
    F = fields.PALLAS_FQ
    log_in, log_out, batch = 20, 23, 3
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randint(-2**31, 2**31 - 1, (batch, 1 << log_in, 8), dtype=torch.int32, device="cuda", generator=g)
    x[..., 7] &= 0x0FFFFFFF
    y = ctx.lde(F.name, x, log_in, log_out)
    assert torch.equal(y[:, ::8, :], x)
    coeffs = x[1:2].clone()
    ctx.ntt(F.name, coeffs, log_in, inverse=True)
    c = from_arr(host(coeffs)[0])
    idx = 8 * 12345 + 3
    pt = pow(F.omega(log_out), idx, F.p)
    acc = 0
    for v in reversed(c):
        acc = (acc * pt + v) % F.p
    assert from_arr(host(y[1, idx:idx + 1]))[0] == acc


# ------------------------------------------------------------------------------------------ pointwise + fold
def test_vec_ops(ctx):
    from crypto3_zk_b200 import capi
    F = fields.BN254_FR
    n = 1000
    a, b, c = (fields.random_elements(F, n, s) for s in (1, 2, 3))
    s = fields.random_elements(F, 1, 4)[0]
    A, B, C = (to_arr(v) for v in (a, b, c))
    assert from_arr(ctx.vec(F.name, capi.VEC_MUL, A, B)) == [x * y % F.p for x, y in zip(a, b)]
    assert from_arr(ctx.vec(F.name, capi.VEC_SUB, A, B)) == [(x - y) % F.p for x, y in zip(a, b)]
    assert from_arr(ctx.vec(F.name, capi.VEC_ADD, A, B)) == [(x + y) % F.p for x, y in zip(a, b)]
    got = ctx.vec(F.name, capi.VEC_MUL_SUB_SCALE, dev(A), dev(B), dev(C), scalar=s)
    assert from_arr(host(got)) == [(x * y - z) * s % F.p for x, y, z in zip(a, b, c)]


@pytest.mark.parametrize("F", [fields.BN254_FR, fields.BLS12_381_FR], ids=lambda f: f.name)
def test_witness_map_flow_vs_oracle(ctx, F):
    """r1cs_to_qap<F>::witness_map (zk/snark/reductions/r1cs_to_qap.hpp:219-325), the FFT part of the Groth16
    prover: 3 iFFT, 3 coset FFT, A o B - C, divide_by_z_on_coset, coset iFFT, on the device in the reference's
    order of operations, against the oracle's literal restatement.  Evaluation vectors stay on the device."""
    from crypto3_zk_b200 import capi
    p, g = F.p, F.g
    for log_m in (6, 11):
        m = 1 << log_m
        aA, aB, aC = (fields.random_elements(F, m, 50 + k + log_m) for k in range(3))
        d1, d2, d3 = fields.random_elements(F, 3, 60 + log_m)
        # ---- oracle (CPU)
        dom = ntt.EvaluationDomain(F, m)
        oA, oB, oC = list(aA), list(aB), list(aC)
        dom.inverse_fft(oA)
        dom.inverse_fft(oB)
        coeffs = [(d2 * x + d1 * y) % p for x, y in zip(oA, oB)] + [0]
        coeffs[0] = (coeffs[0] - d3) % p
        dom.add_poly_z(d1 * d2 % p, coeffs)
        oA, oB = ntt.coset_fft(oA, F, g), ntt.coset_fft(oB, F, g)
        dom.inverse_fft(oC)
        oC = ntt.coset_fft(oC, F, g)
        H = [(x * y - z) % p for x, y, z in zip(oA, oB, oC)]
        dom.divide_by_z_on_coset(H)
        H = ntt.coset_inverse_fft(H, F, g)
        want = [(c + h) % p for c, h in zip(coeffs, H)] + [coeffs[m]]
        # ---- device
        dA, dB, dC = (dev(to_arr(v).reshape(1, m, 8)) for v in (aA, aB, aC))
        cA, cB, cC = (ctx.ntt(F.name, d, log_m, inverse=True) for d in (dA, dB, dC))
        gA, gB = from_arr(host(cA)), from_arr(host(cB))
        eA, eB, eC = (ctx.ntt(F.name, c, log_m, coset_shift=g) for c in (cA, cB, cC))
        zinv = F.inv((pow(g, m, p) - 1) % p)
        Ht = ctx.vec(F.name, capi.VEC_MUL_SUB_SCALE, eA, eB, eC, scalar=zinv)
        Hc = from_arr(host(ctx.ntt(F.name, Ht, log_m, inverse=True, coset_shift=g)))
        got = [(d2 * x + d1 * y) % p for x, y in zip(gA, gB)] + [0]
        got[0] = (got[0] - d3) % p
        got[m] = (got[m] + d1 * d2) % p
        got[0] = (got[0] - d1 * d2) % p
        got = [(c + h) % p for c, h in zip(got, Hc)] + [got[m]]
        assert got == want


@pytest.mark.parametrize("F", fields.NTT_FIELDS, ids=lambda f: f.name)
@pytest.mark.parametrize("log_n", [1, 2, 6, 11])
def test_fri_fold_vs_oracle(ctx, F, log_n):
    """fold_polynomial dfs form (fold_polynomial.hpp:68-93)."""
    f = fields.random_elements(F, 1 << log_n, 5 + log_n)
    alpha = fields.random_elements(F, 1, 99)[0]
    got = ctx.fri_fold(F.name, to_arr(f), log_n, alpha)
    assert from_arr(got) == fri.fold_polynomial_dfs(f, alpha, F)


def test_fri_fold_large_matches_coefficient_fold(ctx):
    """Size-independent identity (test/commitment/fold_polynomial.cpp): folding evaluations equals
    evaluating the coefficient fold f_even + alpha f_odd on the squared domain."""
    import torch
    F = fields.PALLAS_FP
    log_n = 18
    n = 1 << log_n
    g = torch.Generator(device="cuda").manual_seed(1)
    co = torch.randint(-2**31, 2**31 - 1, (1, n, 8), dtype=torch.int32, device="cuda", generator=g)
    co[..., 7] &= 0x0FFFFFFF
    alpha = fields.random_elements(F, 1, 5)[0]
    ev = co.clone()
    ctx.ntt(F.name, ev, log_n)
    folded = ctx.fri_fold(F.name, ev[0], log_n, alpha)
    c = from_arr(host(co)[0])
    fc = fri.fold_polynomial_coeffs(c, alpha, F.p)
    want = dev(to_arr(fc).reshape(1, n // 2, 8))
    ctx.ntt(F.name, want, log_n - 1)
    assert torch.equal(want[0], folded)


# ------------------------------------------------------------------------------------------ LPC commit / Merkle
HASHES = [(0, hashes.keccak256), (1, hashes.sha256), (2, hashes.keccak512)]


@pytest.mark.parametrize("hid,h", HASHES, ids=["keccak256", "sha256", "keccak512"])
@pytest.mark.parametrize("log_in,log_out,step,batch", [(2, 3, 1, 1), (3, 5, 1, 3), (4, 7, 2, 2), (4, 7, 3, 5), (5, 8, 1, 9), (3, 3, 3, 1)])
def test_lpc_commit_vs_oracle(ctx, hid, h, log_in, log_out, step, batch):
    """precommit + root (basic_fri.hpp:445-496, lpc.hpp:101-106) and Merkle paths."""
    F = fields.PALLAS_FP
    polys = [fields.random_elements(F, 1 << log_in, 17 * batch + b) for b in range(batch)]
    levels, ext = fri.precommit(polys, F, 1 << log_out, step, h)
    a = to_arr([v for p in polys for v in p]).reshape(batch, 1 << log_in, 8)
    tree = ctx.lpc_commit(F.name, hid, a, log_in, log_out, step, keep_tree=True)
    assert tree.root() == levels[-1][0]
    assert tree.leaves == len(levels[0])
    for idx in {0, 1, tree.leaves - 1, tree.leaves // 2}:
        if idx < tree.leaves and tree.leaves > 1:
            assert tree.path(idx) == fri.merkle_proof(levels, idx)
    tree.free()
    # merkle-only entry on already extended evaluations (device resident)
    e = dev(to_arr([v for p in ext for v in p]).reshape(batch, 1 << log_out, 8))
    assert ctx.merkle_commit(F.name, hid, e, log_out, step) == levels[-1][0]
    assert ctx.lpc_commit(F.name, hid, dev(a), log_in, log_out, step) == levels[-1][0]


def test_lpc_commit_many_polys_bls(ctx):
    """More polynomials than one Keccak block can hold per element boundary pattern (rate 136 B is not
    a multiple of 32 B): 37 polynomials, BLS12-381 Fr."""
    F = fields.BLS12_381_FR
    polys = [fields.random_elements(F, 8, 900 + b) for b in range(37)]
    a = to_arr([v for p in polys for v in p]).reshape(37, 8, 8)
    for hid, h in HASHES:
        levels, _ = fri.precommit(polys, F, 32, 2, h)
        assert ctx.lpc_commit(F.name, hid, a, 3, 5, 2) == levels[-1][0]


# ------------------------------------------------------------------------------------------ MSM
def enc_points(C, pts):
    """Affine points -> [n, 2, coord_limbs] u32 (G2: every coordinate is c0 || c1)."""
    n = C.coord_limbs32
    if isinstance(C.F.zero, tuple):
        vals = []
        for P in pts:
            vals += [0, 0, 0, 0] if P is None else [P[0][0], P[0][1], P[1][0], P[1][1]]
        return to_arr(vals, n // 2).reshape(len(pts), 2, n)
    vals = []
    for P in pts:
        vals += [0, 0] if P is None else [P[0], P[1]]
    return to_arr(vals, n).reshape(len(pts), 2, n)


def _i(x):
    return int(x, 16) if isinstance(x, str) else int(x)


def test_msm_golden_bellperson_vectors(ctx):
    """The reference's literal BLS12-381 vectors (conformity.cpp:864-930, :1065-1884) through the GPU."""
    g = json.load(open(GOLD))
    C = curves.BLS12_381_G1
    p = fields.BLS12_381_FR.p
    gi = g["gipa"]
    c = [(_i(x), _i(y)) for x, y in gi["c"]]
    r = [_i(x) for x in gi["r"]]
    for rnd in range(3):
        split = len(c) // 2
        bl = ctx.msm_bases(C.name, enc_points(C, c[split:]))
        br = ctx.msm_bases(C.name, enc_points(C, c[:split]))
        zl = ctx.multiexp(bl, to_arr(r[:split]))
        zr = ctx.multiexp(br, to_arr(r[split:]))
        exp = gi["z_c"][rnd]
        assert zl == (_i(exp[0][0]), _i(exp[0][1]))
        assert zr == (_i(exp[1][0]), _i(exp[1][1]))
        x, xinv = _i(gi["ch"][rnd]), _i(gi["ch_inv"][rnd])
        c = [C.add(c[i], C.mul(c[split + i], x)) for i in range(split)]
        r = [(r[i] + r[split + i] * xinv) % p for i in range(split)]
    # prove_commitment_w: MSM of size 16 over powers of alpha / beta
    pc = g["prove_commitment"]
    n = pc["n"]
    alpha, beta, z, r_shift = _i(pc["alpha"]), _i(pc["beta"]), _i(pc["kzg_challenge"]), _i(pc["r_shift"])
    tr = [_i(t) for t in pc["tr"]]
    co, pw = [1], r_shift
    for x in tr:
        co += [cc * (x * pw % p) % p for cc in co]
        pw = pw * pw % p
    fw = [0] * n + co
    quo = [0] * (2 * n)
    carry = 0
    for i in range(2 * n - 1, 0, -1):
        carry = (fw[i] + carry * z) % p
        quo[i - 1] = carry
    for s, exp in ((alpha, pc["comm_w"][0]), (beta, pc["comm_w"][1])):
        srs = [C.mul(C.gen, pow(s, i, p)) for i in range(2 * n)]
        b = ctx.msm_bases(C.name, enc_points(C, srs))
        assert ctx.multiexp(b, to_arr(quo)) == (_i(exp[0]), _i(exp[1]))
    # prove_commitment_v: the same on G2 (f_v with r_shift = 1 over h^(alpha^i), conformity.cpp:876-892)
    C2 = curves.BLS12_381_G2
    co, pw = [1], 1
    for x in tr:
        co += [cc * (x * pw % p) % p for cc in co]
        pw = pw * pw % p
    quo_v, carry = [0] * n, 0
    for i in range(n - 1, 0, -1):
        carry = (co[i] + carry * z) % p
        quo_v[i - 1] = carry
    for s, exp in ((alpha, pc["comm_v"][0]), (beta, pc["comm_v"][1])):
        srs = [C2.mul(C2.gen, pow(s, i, p)) for i in range(n)]
        b = ctx.msm_bases(C2.name, enc_points(C2, srs))
        want = ((_i(exp[0][0]), _i(exp[0][1])), (_i(exp[1][0]), _i(exp[1][1])))
        assert ctx.multiexp(b, to_arr(quo_v)) == want


@pytest.mark.parametrize("C", [curves.BLS12_381_G2, curves.BN254_G2], ids=lambda c: c.name)
def test_msm_g2_vs_oracle(ctx, C):
    """G2 MSM (B_query of the Groth16 prover, prover.hpp:113-119; ipp2 v-keys): random points, special
    scalars, repeated / opposite / infinite points, sub-ranges, 0/1-heavy assignments."""
    r = C.scalar_field.p
    rnd = random.Random(5)
    for n in (1, 3, 64, 700):
        pts = C.random_points(n, 30 + n)
        sc = fields.random_elements(C.scalar_field, n, 40 + n)
        b = ctx.msm_bases(C.name, enc_points(C, pts))
        assert ctx.multiexp(b, to_arr(sc)) == C.msm_bdlo12(pts, sc)
        if n >= 64:
            assert ctx.multiexp(b, to_arr(sc[5:25]), offset=5, n=20) == C.msm_naive(pts[5:25], sc[5:25])
    pts = C.random_points(12, 9)
    P, Q = pts[0], pts[1]
    pts2 = [P] * 5 + [C.neg(P), None, Q, None, P]
    sc2 = [5, 5, 7, 1, r - 1, 5, 9, 3, 0, (1 << 254) % r]
    b2 = ctx.msm_bases(C.name, enc_points(C, pts2))
    assert ctx.multiexp(b2, to_arr(sc2)) == C.msm_naive(pts2, sc2)
    b3 = ctx.msm_bases(C.name, enc_points(C, [P, C.neg(P)]))
    assert ctx.multiexp(b3, to_arr([77, 77])) is None
    n = 1200
    ptsn = C.random_points(n, 8)
    scn = [rnd.choice([0, 1, 1, 1, rnd.randrange(r)]) for _ in range(n)]
    bn = ctx.msm_bases(C.name, enc_points(C, ptsn))
    assert ctx.multiexp(bn, to_arr(scn)) == C.msm_with_mixed_addition(ptsn, scn)
    from crypto3_zk_b200 import msm_combine
    p0 = ctx.multiexp_partial(bn, to_arr(scn[:500]), offset=0, n=500)
    p1 = ctx.multiexp_partial(bn, to_arr(scn[500:]), offset=500, n=700)
    assert msm_combine(C.name, [p0, p1]) == C.msm_with_mixed_addition(ptsn, scn)


@pytest.mark.parametrize("C", [curves.BLS12_381_G1, curves.BN254_G1, curves.PALLAS], ids=lambda c: c.name)
@pytest.mark.parametrize("n", [1, 2, 7, 100, 1500])
def test_msm_vs_oracle(ctx, C, n):
    pts = C.random_points(n, 10 + n)
    sc = fields.random_elements(C.scalar_field, n, 20 + n)
    b = ctx.msm_bases(C.name, enc_points(C, pts))
    assert ctx.multiexp(b, to_arr(sc)) == C.msm_bdlo12(pts, sc)
    # sub-range of resident bases (prover.hpp:133-139 passes iterator sub-ranges)
    if n >= 7:
        assert ctx.multiexp(b, to_arr(sc[2:6]), offset=2, n=4) == C.msm_naive(pts[2:6], sc[2:6])
    # the same through a window table (zkb_msm_bases_precompute), default and explicit window sizes
    want = C.msm_bdlo12(pts, sc)
    for wb in (0, 5, 11):
        bt = ctx.msm_bases(C.name, enc_points(C, pts)).precompute(wb)
        assert ctx.multiexp(bt, to_arr(sc)) == want
        if n >= 7:
            assert ctx.multiexp(bt, to_arr(sc[2:6]), offset=2, n=4) == C.msm_naive(pts[2:6], sc[2:6])
        bt.free()


@pytest.mark.parametrize("C", [curves.BLS12_381_G1, curves.BN254_G1, curves.PALLAS], ids=lambda c: c.name)
def test_msm_edge_cases(ctx, C):
    r = C.scalar_field.p
    pts = C.random_points(40, 3)
    rnd = random.Random(4)
    # special scalars: 0, 1, 2, r-1, r-2, 2^k boundaries of the signed windows
    sc = [0, 1, 2, r - 1, r - 2, (1 << 15), (1 << 15) + 1, (1 << 16) - 1, (1 << 16), (1 << 254) % r, (r - 1) // 2]
    sc += [rnd.randrange(r) for _ in range(40 - len(sc))]
    b = ctx.msm_bases(C.name, enc_points(C, pts))
    assert ctx.multiexp(b, to_arr(sc)) == C.msm_naive(pts, sc)
    # all-zero scalars -> infinity ; empty -> infinity
    assert ctx.multiexp(b, to_arr([0] * 40)) is None
    assert ctx.multiexp(b, np.zeros((0, 8), dtype=np.uint32), n=0) is None
    # repeated points (doubling inside a bucket), P and -P with equal scalars (cancellation), infinity inputs
    P, Q = pts[0], pts[1]
    pts2 = [P] * 6 + [C.neg(P)] + [None, Q, None]
    sc2 = [5, 5, 5, 7, 7, 1, 5, 9, 3, 0]
    b2 = ctx.msm_bases(C.name, enc_points(C, pts2))
    assert ctx.multiexp(b2, to_arr(sc2)) == C.msm_naive(pts2, sc2)
    pts3 = [P, C.neg(P)]
    b3 = ctx.msm_bases(C.name, enc_points(C, pts3))
    assert ctx.multiexp(b3, to_arr([123456789, 123456789])) is None
    # 0/1-heavy Groth16-style assignment: most scalars 0 or 1 (one huge bucket -> task splitting)
    n = 3000
    ptsn = C.random_points(n, 8)
    scn = [rnd.choice([0, 1, 1, 1, rnd.randrange(r)]) for _ in range(n)]
    bn = ctx.msm_bases(C.name, enc_points(C, ptsn))
    assert ctx.multiexp(bn, to_arr(scn)) == C.msm_with_mixed_addition(ptsn, scn)
    # all points equal, all scalars equal
    pe = [P] * 700
    be = ctx.msm_bases(C.name, enc_points(C, pe))
    assert ctx.multiexp(be, to_arr([r - 3] * 700)) == C.mul(P, 700 * (r - 3))



def _grid_msm_case(ctx, C, log_n, scalars_fn, seed, table=0):
    """MSM over n = 2^log_n grid points P_i = A[i % m] + B[i / m] (built on the device) checked through the
    size-independent identity  sum s_i P_i = sum_a (sum_{i%m=a} s_i) A_a + sum_b (sum_{i/m=b} s_i) B_b,
    whose right-hand side is an (m + n/m)-point MSM the oracle does on the CPU."""
    r = C.scalar_field.p
    n, m = 1 << log_n, 256
    nb = n // m
    rnd = random.Random(seed)

    def progression(count):   # P_0 + k D by a running addition (cheap in Python, distinct w.h.p.)
        cur, D, out = C.mul(C.gen, rnd.randrange(1, r)), C.mul(C.gen, rnd.randrange(1, r)), []
        for _ in range(count):
            out.append(cur)
            cur = C.add(cur, D)
        return out
    A, Bt = progression(m), progression(nb)
    pts = ctx.grid_points(C.name, n, enc_points(C, A), enc_points(C, Bt))
    bases = ctx.msm_bases(C.name, pts)
    sc = scalars_fn(n, rnd, r)
    arr = np.zeros((n, 8), dtype=np.uint32)
    for k in range(8):
        arr[:, k] = [(v >> (32 * k)) & 0xFFFFFFFF for v in sc]
    got = ctx.multiexp(bases, dev(arr))
    if table:   # same sum through the window table, whole range and a sub-range that starts inside it
        bases.precompute(table)
        assert ctx.multiexp(bases, dev(arr)) == got
        h = n // 2 + 3
        from crypto3_zk_b200 import msm_combine
        p0 = ctx.multiexp_partial(bases, dev(arr[:h]), offset=0, n=h)
        p1 = ctx.multiexp_partial(bases, dev(arr[h:]), offset=h, n=n - h)
        assert msm_combine(C.name, [p0, p1]) == got
    sa, sb = [0] * m, [0] * nb
    for i, v in enumerate(sc):
        sa[i % m] += v
        sb[i // m] += v
    want = C.msm_bdlo12(A + Bt, [v % r for v in sa + sb])
    assert got == want
    bases.free()


@pytest.mark.parametrize("kind", ["uniform", "all_equal", "zero_one", "top_window"])
def test_msm_large_grid_identity(ctx, kind):
    """2^17..2^20-point BLS12-381 MSMs: uniform 255-bit scalars (the top signed window holds only a few
    significant bits, so its buckets are split into many tasks), all-equal scalars (one bucket per
    window), a 0/1-heavy Groth16-style assignment (knowledge_commitment_multiexp.hpp:88-101) and scalars
    that differ only in the top window."""
    C = curves.BLS12_381_G1
    fns = {
        "uniform": lambda n, rnd, r: [rnd.randrange(r) for _ in range(n)],
        "all_equal": lambda n, rnd, r: [r - 12345] * n,
        "zero_one": lambda n, rnd, r: [rnd.choice([0, 1, 1, 1, 1, 1, 1, rnd.randrange(r)]) for _ in range(n)],
        "top_window": lambda n, rnd, r: [((rnd.randrange(7) << 252) + 99) % r for _ in range(n)],
    }
    _grid_msm_case(ctx, C, 20 if kind == "uniform" else 17, fns[kind], 77, table={"uniform": 20, "all_equal": 13}.get(kind, 17))


def test_kzg_commit_identity_2p16(ctx):
    """BASELINE config #1: KZG commit of a degree-2^16 polynomial over BLS12-381 (kzg.hpp:143-148):
    with commitment_key[i] = alpha^i G the commitment equals f(alpha) G (test/commitment/kzg.cpp:97)."""
    C = curves.BLS12_381_G1
    r = C.scalar_field.p
    n = 1 << 16
    alpha = 7
    # SRS by a running scalar multiplication would be slow in Python: use points k_i G with known k_i
    rnd = random.Random(1)
    k0, step = rnd.randrange(r), rnd.randrange(r)
    cur, stepj = C.j_from_affine(C.mul(C.gen, k0)), C.j_from_affine(C.mul(C.gen, step))
    js = []
    for _ in range(n):
        js.append(cur)
        cur = C.j_add(cur, stepj)
    pts = C.batch_to_affine(js)
    f = fields.random_elements(C.scalar_field, n, 3)
    b = ctx.msm_bases(C.name, enc_points(C, pts))
    got = ctx.multiexp(b, to_arr(f))
    k = sum(fi * ((k0 + i * step) % r) for i, fi in enumerate(f)) % r
    assert got == C.mul(C.gen, k)
    # true KZG shape on a short key: commit == f(alpha) G
    m = 64
    srs = [C.mul(C.gen, pow(alpha, i, r)) for i in range(m)]
    bs = ctx.msm_bases(C.name, enc_points(C, srs))
    fa = sum(fi * pow(alpha, i, r) for i, fi in enumerate(f[:m])) % r
    assert ctx.multiexp(bs, to_arr(f[:m])) == C.mul(C.gen, fa)
    # point-sharded partial sums (multi-GPU layout, one device): halves combine to the same point
    from crypto3_zk_b200 import msm_combine
    p0 = ctx.multiexp_partial(b, to_arr(f[:n // 2]), offset=0, n=n // 2)
    p1 = ctx.multiexp_partial(b, to_arr(f[n // 2:]), offset=n // 2, n=n // 2)
    assert msm_combine(C.name, [p0, p1]) == got


def test_field_mul_microbench_runs(ctx):
    r = ctx.bench_field_mul("bls12_381_fq", 148 * 4, 256, 256)
    assert r > 1e9
