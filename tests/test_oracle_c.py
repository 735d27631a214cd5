"""Cross-checks the C restatement (oracle/c, the timed CPU baseline) against the Python oracle,
which is itself pinned to the reference's literal vectors (tests/test_oracle_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import cref, curves, fields, fri, hashes, ntt


def arr(vals, limbs=8):
    return fields.ints_to_u32_array(vals, limbs)


def ints(a):
    return fields.u32_array_to_ints(np.asarray(a).reshape(-1, np.asarray(a).shape[-1]))


@pytest.mark.parametrize("F", fields.NTT_FIELDS, ids=lambda f: f.name)
def test_c_ntt_lde_fold(F):
    for log_n in (1, 4, 9):
        n = 1 << log_n
        p = fields.random_elements(F, 2 * n, log_n)
        a = arr(p).reshape(2, n, 8)
        cref.ntt(F.fid, a, log_n, threads=2)
        want = []
        for b in range(2):
            w = p[b * n:(b + 1) * n]
            ntt.EvaluationDomain(F, n).fft(w)
            want += w
        assert ints(a) == want
        cref.ntt(F.fid, a, log_n, inverse=True)
        assert ints(a) == p
        c = arr(p[:n]).reshape(1, n, 8)
        cref.ntt(F.fid, c, log_n, shift=F.g)
        assert ints(c) == ntt.coset_fft(p[:n], F, F.g)
        cref.ntt(F.fid, c, log_n, inverse=True, shift=F.g)
        assert ints(c) == p[:n]
    p = fields.random_elements(F, 16, 3)
    out, _ = cref.lde(F.fid, arr(p).reshape(1, 16, 8), 4, 7)
    assert ints(out) == ntt.dfs_resize(p, F, 128)
    f = fields.random_elements(F, 64, 4)
    got, _ = cref.fri_fold(F.fid, arr(f), 6, 12345)
    assert ints(got) == fri.fold_polynomial_dfs(f, 12345, F)


def test_c_hashes():
    for n in (0, 1, 55, 56, 63, 64, 71, 72, 135, 136, 137, 300, 4096):
        d = bytes((i * 31 + n) & 0xFF for i in range(n))
        assert cref.hash_bytes(0, d) == hashes.keccak256(d)
        assert cref.hash_bytes(1, d) == hashlib.sha256(d).digest()
        assert cref.hash_bytes(2, d) == hashes.keccak512(d)


@pytest.mark.parametrize("hid,h", [(0, hashes.keccak256), (1, hashes.sha256), (2, hashes.keccak512)])
def test_c_lpc_commit(hid, h):
    F = fields.PALLAS_FP
    for log_in, log_out, step, batch in ((3, 5, 1, 3), (4, 7, 3, 2), (3, 6, 2, 5)):
        polys = [fields.random_elements(F, 1 << log_in, 5 + b) for b in range(batch)]
        root, _, _ = cref.lpc_commit(F.fid, hid, arr([v for p in polys for v in p]), log_in, log_out, step, threads=2)
        assert root == fri.lpc_commit(polys, F, log_in, log_out - log_in, step, h)


@pytest.mark.parametrize("C", [curves.BLS12_381_G1, curves.BN254_G1, curves.PALLAS], ids=lambda c: c.name)
def test_c_msm(C):
    n = C.coord_limbs32
    for cnt, threads in ((1, 1), (5, 1), (200, 3)):
        pts = C.random_points(cnt, cnt) + [None]
        sc = fields.random_elements(C.scalar_field, cnt, 1) + [5]
        pa = arr([c for P in pts for c in (P if P else (0, 0))], n).reshape(-1, 2, n)
        got, _ = cref.msm(C.cid, pa, arr(sc), threads=threads)
        assert got == C.msm_naive(pts, sc)


def test_c_msm_golden():
    """reference literal vector (conformity.cpp:1065-1884, first z_c pair) through the C port"""
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bls12_381_ipp2.json")))["gipa"]
    I = lambda x: int(x, 16) if isinstance(x, str) else int(x)
    c = [(I(x), I(y)) for x, y in g["c"]]
    r = [I(x) for x in g["r"]]
    pa = arr([v for P in c[4:] for v in P], 12).reshape(-1, 2, 12)
    got, _ = cref.msm(0, pa, arr(r[:4]))
    assert got == (I(g["z_c"][0][0][0]), I(g["z_c"][0][0][1]))
