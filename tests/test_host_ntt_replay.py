"""CPU replay of the NTT pass kernels (same plan/tables/phase functions as the device path) against
the oracle's radix-2 domain.  Covers 1/2/3/4-pass plans (small tile radix builds), inverse, coset
and zero-padded (LDE) inputs, ragged batches.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import fields, ntt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "crypto3_zk_b200", "csrc")


def _build(maxlogr):
    so = os.path.join(ROOT, "crypto3_zk_b200", "libzkb_hosttest_r%d.so" % maxlogr)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".cpp"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DZKB_NTT_MAX_LOG_R=%d" % maxlogr,
                               "-x", "c++", os.path.join(CSRC, "host_selftest.cpp"), "-o", so])
    return ctypes.CDLL(so)


@pytest.fixture(scope="module")
def libs():
    return {r: _build(r) for r in (3, 4, 8)}


def run(lib, F, log_n, polys, inverse=False, shift=None, log_n_in=None):
    log_n_in = log_n if log_n_in is None else log_n_in
    batch = len(polys)
    a = fields.ints_to_u32_array([v for p in polys for v in p], 8)
    out = np.zeros((batch << log_n, 8), dtype=np.uint32)
    sh = None
    if shift is not None:
        sh = (ctypes.c_uint32 * 8)(*fields.to_limbs32(shift, 8))
    rc = lib.zkb_host_ntt(F.fid, log_n, log_n_in, int(inverse), sh, batch,
                          a.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    vals = fields.u32_array_to_ints(out)
    n = 1 << log_n
    return [vals[i * n:(i + 1) * n] for i in range(batch)]


def oracle(F, log_n, poly, inverse=False, shift=None):
    a = list(poly)
    d = ntt.EvaluationDomain(F, 1 << log_n)
    if inverse:
        d.inverse_fft(a)
        if shift is not None:
            ntt.multiply_by_coset(a, F.inv(shift), F.p)
    else:
        if shift is not None:
            ntt.multiply_by_coset(a, shift, F.p)
        d.fft(a)
    return a


CASES = [(8, 1), (8, 2), (8, 3), (8, 5), (8, 8), (8, 9), (8, 10), (8, 13),
         (4, 7), (4, 9), (4, 10), (4, 11), (4, 12), (4, 13), (3, 6), (3, 9), (3, 12)]


@pytest.mark.parametrize("maxlogr,log_n", CASES)
def test_forward_inverse(libs, maxlogr, log_n):
    F = fields.NTT_FIELDS[(maxlogr + log_n) % 4]
    n = 1 << log_n
    batch = 3 if log_n <= 8 else 2
    polys = [fields.random_elements(F, n, 100 * log_n + b) for b in range(batch)]
    got = run(libs[maxlogr], F, log_n, polys)
    for b in range(batch):
        assert got[b] == oracle(F, log_n, polys[b]), "forward"
    back = run(libs[maxlogr], F, log_n, got, inverse=True)
    assert back == polys, "inverse"


@pytest.mark.parametrize("maxlogr,log_n", [(8, 4), (8, 10), (3, 9), (4, 12)])
def test_coset(libs, maxlogr, log_n):
    F = fields.NTT_FIELDS[log_n % 4]
    polys = [fields.random_elements(F, 1 << log_n, 7 + log_n)]
    g = F.g
    got = run(libs[maxlogr], F, log_n, polys, shift=g)
    assert got[0] == oracle(F, log_n, polys[0], shift=g)
    back = run(libs[maxlogr], F, log_n, got, inverse=True, shift=g)
    assert back == polys


@pytest.mark.parametrize("maxlogr,log_in,log_out", [(8, 3, 6), (8, 7, 10), (3, 6, 9), (4, 9, 12), (3, 2, 9), (8, 1, 11)])
def test_zero_padded_input(libs, maxlogr, log_in, log_out):
    """Forward transform of a 2^log_in coefficient vector on the 2^log_out domain (second half of
    polynomial_dfs::resize, basic_fri.hpp:451-455)."""
    F = fields.PALLAS_FQ
    polys = [fields.random_elements(F, 1 << log_in, 3 + b) for b in range(2)]
    got = run(libs[maxlogr], F, log_out, polys, log_n_in=log_in)
    for b in range(2):
        a = list(polys[b])
        ntt.EvaluationDomain(F, 1 << log_out).fft(a)
        assert got[b] == a


@pytest.mark.parametrize("maxlogr,log_in,log_out", [(8, 3, 6), (8, 7, 10), (3, 6, 9), (4, 9, 12), (8, 0, 5), (8, 8, 11)])
def test_zero_padded_coset_input(libs, maxlogr, log_in, log_out):
    """Zero-padded input with a coset pre-scale: the replicas of the live rows are written after the scaling."""
    F = fields.BLS12_381_FR
    polys = [fields.random_elements(F, 1 << log_in, 11 + b) for b in range(3)]
    got = run(libs[maxlogr], F, log_out, polys, log_n_in=log_in, shift=F.g)
    for b in range(3):
        a = list(polys[b]) + [0] * ((1 << log_out) - (1 << log_in))
        assert got[b] == oracle(F, log_out, a, shift=F.g)


@pytest.mark.parametrize("maxlogr,log_in,log_out", [(8, 7, 10), (8, 6, 10), (8, 9, 12), (4, 6, 10), (4, 6, 9), (4, 8, 12),
                                                     (3, 6, 9), (3, 3, 9), (8, 5, 7), (8, 3, 6), (4, 9, 13),
                                                     (8, 10, 13), (8, 11, 14), (8, 12, 16), (8, 15, 18)])
def test_lde_with_known_outputs(libs, maxlogr, log_in, log_out):
    """lde_device as the runtime drives it (iNTT, then the zero-padded forward transform that takes the outputs at
    multiples of the blow-up from the input evaluations) against polynomial_dfs::resize; 2/3/4-pass plans, blow-ups
    4..64 (below 8 the known-output path is off), with the work buffer poisoned."""
    F = fields.NTT_FIELDS[(log_in + log_out) % 4]
    lib = libs[maxlogr]
    batch = 3 if log_out <= 14 else 2      # (the last cases are the ones where the known-output path is on: 2- and 3-pass
                                           # plans, periods of 7 and 15 kept k_1, first radix 64 / 128 / 256)
    polys = [fields.random_elements(F, 1 << log_in, 50 + b) for b in range(batch)]
    a = fields.ints_to_u32_array([v for p in polys for v in p], 8)
    out = np.zeros((batch << log_out, 8), dtype=np.uint32)
    rc = lib.zkb_host_lde(F.fid, log_in, log_out, batch, a.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    vals = fields.u32_array_to_ints(out)
    n = 1 << log_out
    for b in range(batch):
        assert vals[b * n:(b + 1) * n] == ntt.dfs_resize(polys[b], F, n)


def test_ragged_batch(libs):
    F = fields.BN254_FR
    polys = [fields.random_elements(F, 16, b) for b in range(11)]   # 11 = 8 + 3 columns
    got = run(libs[8], F, 4, polys)
    for b in range(11):
        assert got[b] == oracle(F, 4, polys[b])
