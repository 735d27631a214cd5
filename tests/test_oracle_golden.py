"""Pins the oracle against every literal vector the reference's own tests hold for the hot path
(SURVEY.md 8(c)).  CPU only."""
import hashlib
import json
import os

import pytest

from oracle import curves, fields, fri, hashes, ntt

G = os.path.join(os.path.dirname(__file__), "golden", "bls12_381_ipp2.json")


def _i(x):
    return int(x, 16) if isinstance(x, str) else int(x)


@pytest.fixture(scope="module")
def gold():
    return json.load(open(G))


def test_fr_products_bls381_polynomial(gold):
    """conformity.cpp:578-862 : polynomial_coefficients_from_transcript + product-form evaluation
    (ipp2/prover.hpp:99-158)."""
    g = gold["polynomial"]
    p = fields.BLS12_381_FR.p
    r = _i(g["r_shift"])
    co = [1]
    pw = r
    for x in map(_i, g["tr"]):
        co += [c * (x * pw % p) % p for c in co]
        pw = pw * pw % p
    assert co == [_i(c) for c in g["coeffs"]]
    z = _i(g["kzg_challenge"])
    pz = z * r % p
    res = 1
    for x in map(_i, g["tr"]):
        res = res * (1 + x * pz) % p
        pz = pz * pz % p
    assert res == _i(g["eval"])


def _quotient(f, fz, z, p):
    """(f - f(z)) / (X - z) by synthetic division, padded to len(f)."""
    n = len(f)
    q = [0] * n
    carry = 0
    for i in range(n - 1, 0, -1):
        carry = (f[i] + carry * z) % p
        q[i - 1] = carry
    assert (f[0] + carry * z - fz) % p == 0
    return q


def test_msm_g1_g2_prove_commitment(gold):
    """conformity.cpp:864-930 : MSM of size 16 over [alpha^i]G (ipp2/srs.hpp:42-55) on G1 and G2."""
    g = gold["prove_commitment"]
    Fr = fields.BLS12_381_FR
    p = Fr.p
    n = g["n"]
    alpha, beta, z = _i(g["alpha"]), _i(g["beta"]), _i(g["kzg_challenge"])
    tr = [_i(t) for t in g["tr"]]

    def coeffs(r_shift):
        co, pw = [1], r_shift
        for x in tr:
            co += [c * (x * pw % p) % p for c in co]
            pw = pw * pw % p
        return co

    def prod_eval(r_shift):
        pz, res = z * r_shift % p, 1
        for x in tr:
            res = res * (1 + x * pz) % p
            pz = pz * pz % p
        return res

    # --- w (G1): f_w = X^n f(X) with r_shift
    r_shift = _i(g["r_shift"])
    fw = [0] * n + coeffs(r_shift)
    fwz = prod_eval(r_shift) * pow(z, n, p) % p
    q = _quotient(fw, fwz, z, p)
    C = curves.BLS12_381_G1
    for s, exp in ((alpha, g["comm_w"][0]), (beta, g["comm_w"][1])):
        srs = [C.mul(C.gen, pow(s, i, p)) for i in range(2 * n)]
        want = (_i(exp[0]), _i(exp[1]))
        assert C.msm_naive(srs, q) == want
        assert C.msm_bdlo12(srs, q) == want
        assert C.msm_bdlo12(srs, q, c=3) == want
    # --- v (G2): f_v with r_shift = 1, SRS over G2 (prove_commitment_v uses the first n powers
    # after specialize(): h_alpha_powers[0:n])
    fv = coeffs(1)
    fvz = prod_eval(1)
    qv = _quotient(fv, fvz, z, p)
    C2 = curves.BLS12_381_G2
    assert C2.is_on_curve(C2.gen)
    for s, exp in ((alpha, g["comm_v"][0]), (beta, g["comm_v"][1])):
        srs = [C2.mul(C2.gen, pow(s, i, p)) for i in range(n)]
        want = ((_i(exp[0][0]), _i(exp[0][1])), (_i(exp[1][0]), _i(exp[1][1])))
        assert C2.msm_naive(srs, qv) == want
        assert C2.msm_bdlo12(srs, qv, c=4) == want


def test_msm_g1_gipa_rounds(gold):
    """conformity.cpp:1065-1884 : zc_l = MSM(c[n':], r[:n']), zc_r = MSM(c[:n'], r[n':]) then
    compress with the literal challenges (ipp2/prover.hpp:384-391,417-423)."""
    g = gold["gipa"]
    C = curves.BLS12_381_G1
    p = fields.BLS12_381_FR.p
    c = [(_i(x), _i(y)) for x, y in g["c"]]
    r = [_i(x) for x in g["r"]]
    assert all(C.is_on_curve(P) for P in c)
    for rnd in range(3):
        split = len(c) // 2
        zl = C.msm_naive(c[split:], r[:split])
        zr = C.msm_naive(c[:split], r[split:])
        assert zl == C.msm_bdlo12(c[split:], r[:split])
        exp = g["z_c"][rnd]
        assert zl == (_i(exp[0][0]), _i(exp[0][1]))
        assert zr == (_i(exp[1][0]), _i(exp[1][1]))
        x, xinv = _i(g["ch"][rnd]), _i(g["ch_inv"][rnd])
        assert x * xinv % p == 1
        c = [C.add(c[i], C.mul(c[split + i], x)) for i in range(split)]
        r = [(r[i] + r[split + i] * xinv) % p for i in range(split)]
    assert c[0] == (_i(g["final_c"][0]), _i(g["final_c"][1]))


def test_keccak_transcript_kat():
    """test/transcript/transcript.cpp:50-64 (keccak_1600<256>, BN254 Fr) pins the padding variant."""
    tr = hashes.FiatShamirSequential(hashes.keccak256, bytes(range(10)))
    F = fields.BN254_FR
    assert tr.challenge(F) == 0xe858ba005424eabd6d97de7e930779def59a85c1a9ff7e8a5d001cdb07f6e4
    assert tr.challenge(F) == 0xf61f38f58a55b3bbee0480fc5ec3cf8df81603579f4f7134f764bfd3ca5938b
    assert tr.challenge(F) == 0x4f6b97a9bc99d6996fab5e03d1cd0b418a9b3c97ed64cca070e15777e7cc99a
    assert tr.challenge(F) == 0x2414ddf7ecff246500beb2c01b0c5912a400bc3cdca6d7f24bd2bd4987b21e04
    assert tr.challenge(F) == 0x10bfe2f4a414eec551dda5fd9899e9b46e327648b4fa564ed0517b6a99396aec


def test_keccak_permutation_vs_hashlib():
    for n in (0, 1, 135, 136, 137, 500):
        d = bytes((7 * i + 3) & 0xFF for i in range(n))
        assert hashes.sha3_256_via_own_permutation(d) == hashlib.sha3_256(d).digest()
    # Keccak-256("") and Keccak-512("") published values
    assert hashes.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert hashes.keccak512(b"").hex().startswith("0eab42de4c3ceb9235fc91acffe746b29c29a8c366b7c60e4e67c466f36a4304")


def test_kzg_identity():
    """test/commitment/kzg.cpp:83-97 : commit({-1,1,2,3}, alpha=10) == 3209*G (any curve)."""
    for C in (curves.BLS12_381_G1, curves.BN254_G1, curves.PALLAS):
        r = C.scalar_field.p
        srs = [C.mul(C.gen, pow(10, i, r)) for i in range(16)]
        f = [r - 1, 1, 2, 3]
        assert C.msm_bdlo12(srs[:4], f) == C.mul(C.gen, 3209)
        assert C.msm_with_mixed_addition(srs[:4], f) == C.mul(C.gen, 3209)


def test_generators_and_orders():
    for C in curves.CURVES.values():
        assert C.is_on_curve(C.gen), C.name
        assert C.mul(C.gen, C.scalar_field.p - 1) == C.neg(C.gen), C.name


def test_fri_domain_structure():
    """test/commitment/fri.cpp:122-123 : D[1].m == D[0].m/2, D[1].w == D[0].w^2."""
    for F in fields.NTT_FIELDS:
        D = ntt.calculate_domain_set(F, 7, 3)
        assert D[1].m == D[0].m // 2
        assert D[1].get_domain_element(1) == D[0].get_domain_element(1) ** 2 % F.p
        assert pow(D[0].omega, D[0].m, F.p) == 1 and pow(D[0].omega, D[0].m // 2, F.p) == F.p - 1


def test_fft_is_dft_and_roundtrip():
    for F in fields.NTT_FIELDS:
        a = fields.random_elements(F, 16, 5)
        d = ntt.EvaluationDomain(F, 16)
        b = list(a)
        d.fft(b)
        assert b == ntt.dft_naive(a, d.omega, F.p)
        d.inverse_fft(b)
        assert b == a


def test_fold_identity():
    """test/commitment/fold_polynomial.cpp:52-135 : fold(f,a)(w^2) equals the line through
    (w, f(w)), (-w, f(-w)) evaluated at a; and dfs fold == coefficient fold."""
    F = fields.PALLAS_FP
    p = F.p
    n = 16
    coeffs = fields.random_elements(F, n, 11)
    alpha = fields.random_elements(F, 1, 12)[0]
    ev = list(coeffs)
    d = ntt.EvaluationDomain(F, n)
    d.fft(ev)
    folded = fri.fold_polynomial_dfs(ev, alpha, F)
    fc = fri.fold_polynomial_coeffs(coeffs, alpha, p)
    d2 = ntt.EvaluationDomain(F, n // 2)
    ev2 = list(fc)
    d2.fft(ev2)
    assert folded == ev2
    w = d.get_domain_element(1)
    i = 3
    x = pow(w, i, p)
    y0, y1 = ev[i], ev[i + n // 2]            # f(x), f(-x)
    # interpolate{(x,y0),(-x,y1)}(alpha)
    lam = (y0 - y1) * F.inv(2 * x % p) % p
    val = (y0 + lam * (alpha - x)) % p
    assert folded[i] == val


def test_leaf_index_pattern():
    """SURVEY Appendix B.4 table (model of basic_fri.hpp:466-492)."""
    assert fri.leaf_indices(0, 16, 1) == [0, 8]
    assert fri.leaf_indices(1, 16, 1) == [1, 9]
    assert fri.leaf_indices(0, 16, 2) == [0, 8, 4, 12]
    assert fri.leaf_indices(1, 16, 2) == [1, 9, 5, 13]
    assert fri.leaf_indices(0, 16, 3) == [0, 8, 4, 12, 2, 10, 6, 14]
    assert fri.leaf_indices(1, 16, 3) == [1, 9, 5, 13, 3, 11, 7, 15]
