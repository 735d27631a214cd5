"""Host transcript of the product (crypto3_zk_b200/transcript.py) against the reference's known answers
(test/transcript/transcript.cpp:50-64) and against the oracle's independent Keccak; oracle restatements of the
FRI commit phase and of the LPC combined quotient against their defining identities.  CPU only."""
import os

from crypto3_zk_b200 import transcript as T
from oracle import fields, fri, hashes, lpc, ntt


def test_keccak_matches_oracle_and_hashlib():
    import hashlib
    for n in (0, 1, 31, 32, 64, 71, 72, 73, 135, 136, 137, 200, 300):
        d = os.urandom(n)
        assert T.keccak256(d) == hashes.keccak256(d)
        assert T.keccak512(d) == hashes.keccak512(d)
        assert T.sha256(d) == hashlib.sha256(d).digest() == hashes.sha256(d)


def test_transcript_reference_known_answers():
    F = fields.BN254_FR
    tr = T.FiatShamirSequential(0, bytes(range(10)))
    want = [0xe858ba005424eabd6d97de7e930779def59a85c1a9ff7e8a5d001cdb07f6e4,
            0xf61f38f58a55b3bbee0480fc5ec3cf8df81603579f4f7134f764bfd3ca5938b,
            0x4f6b97a9bc99d6996fab5e03d1cd0b418a9b3c97ed64cca070e15777e7cc99a,
            0x2414ddf7ecff246500beb2c01b0c5912a400bc3cdca6d7f24bd2bd4987b21e04,
            0x10bfe2f4a414eec551dda5fd9899e9b46e327648b4fa564ed0517b6a99396aec]
    assert [tr.challenge(F.p) for _ in range(2)] + tr.challenges(F.p, 3) == want
    o = hashes.FiatShamirSequential(hashes.keccak256, bytes(range(10)))
    assert [o.challenge(F) for _ in range(5)] == want


def test_transcript_absorb_matches_oracle():
    F = fields.PALLAS_FQ
    a, b = T.FiatShamirSequential(0, b"\x01\x02"), hashes.FiatShamirSequential(hashes.keccak256, b"\x01\x02")
    for blob in (b"root-1" * 5, os.urandom(32), os.urandom(64)):
        a(blob)
        b.absorb(blob)
        assert a.challenge(F.p) == b.challenge(F)


def test_oracle_commit_phase_final_polynomial_is_the_coefficient_fold():
    """fold identity (test/commitment/fold_polynomial.cpp:52-135) carried through the whole commit phase: the final
    polynomial equals the coefficient-form folds f_even + alpha f_odd applied with the same challenges."""
    F = fields.PALLAS_FP
    log_n, steps = 7, [2, 1, 1]
    co = fields.random_elements(F, 1 << 5, 3) + [0] * ((1 << log_n) - (1 << 5))
    f = list(co)
    ntt.EvaluationDomain(F, 1 << log_n).fft(f)
    res = fri.commit_phase(f, F, log_n, steps, hashes.keccak256, hashes.FiatShamirSequential(hashes.keccak256))
    c = co
    for a in res["alphas"]:
        c = fri.fold_polynomial_coeffs(c, a, F.p)
    assert res["final_polynomial"] == c
    assert len(res["roots"]) == 3 and [len(x) for x in res["fs"]] == [128, 32, 16, 8]


def test_oracle_combined_q_identity():
    """Q(x) (lpc.hpp:126-181) = sum over points of (sum theta^k (g(x) - z)) / (x - point) at a random x."""
    F = fields.BLS12_381_FR
    p = F.p
    polys = {0: [fields.random_elements(F, 16, 1), fields.random_elements(F, 16, 2)],
             1: [fields.random_elements(F, 16, 3)]}
    y = 1234567
    points = {0: [[y], [y, y * 3 % p]], 1: [[y * 3 % p]]}
    z = lpc.eval_polys(polys, points, F)
    theta = 987654321
    fixed_values = {0: [lpc.poly_eval(ntt.dfs_coefficients(q, F), 55, p) for q in polys[0]]}
    q, q_dfs = lpc.combined_q(polys, points, z, theta, F, fixed_batches=(0,), etha=55, fixed_values=fixed_values)
    x = 424242
    g = {k: [lpc.poly_eval(ntt.dfs_coefficients(v, F), x, p) for v in polys[k]] for k in polys}
    acc, want = 1, 0
    for pt in lpc.unique_points(points):
        num = 0
        for k in sorted(polys):
            for i in range(len(polys[k])):
                if pt in points[k][i]:
                    num = (num + acc * (g[k][i] - z[k][i][points[k][i].index(pt)])) % p
                    acc = acc * theta % p
        want = (want + num * F.inv((x - pt) % p)) % p
    num = 0
    for i in range(2):
        num = (num + acc * (g[0][i] - fixed_values[0][i])) % p
        acc = acc * theta % p
    want = (want + num * F.inv((x - 55) % p)) % p
    assert lpc.poly_eval(q, x, p) == want
    assert len(q_dfs) == 16
