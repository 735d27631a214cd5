"""The device arithmetic headers (zkb_field.cuh / zkb_curve.cuh) compiled with g++ (PTX carry
primitives emulated) against Python big integers.  CPU only: this is how the kernels' arithmetic is
validated in the GPU-less build container; the same code paths run in the kernels."""
import ctypes
import os
import random
import subprocess

import pytest

from oracle import curves, fields

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "crypto3_zk_b200", "csrc")
SO = os.path.join(ROOT, "crypto3_zk_b200", "libzkb_hosttest.so")


@pytest.fixture(scope="module")
def lib():
    srcs = [os.path.join(CSRC, f) for f in ("host_selftest.cpp", "zkb_field.cuh", "zkb_curve.cuh",
                                             "zkb_ptx.cuh", "zkb_params.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                               srcs[0], "-o", SO])
    return ctypes.CDLL(SO)


def _call(fn, ident, op, a, b, n):
    A = (ctypes.c_uint32 * n)(*a)
    B = (ctypes.c_uint32 * n)(*b)
    O = (ctypes.c_uint32 * n)()
    assert fn(ident, op, A, B, O) == 0
    return list(O)


def _edge_values(p):
    return [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (1 << (p.bit_length() - 1)) - 1, 1 << (p.bit_length() - 1),
            0xFFFFFFFF, (1 << 64) - 1, p - 0xFFFFFFFF]


@pytest.mark.parametrize("F", list(fields.FIELDS.values()), ids=lambda f: f.name)
def test_field_ops(lib, F):
    n, p = F.limbs32, F.p
    R = pow(2, 32 * n, p)
    Rinv = pow(R, -1, p)
    rnd = random.Random(F.fid)
    vals = _edge_values(p) + [rnd.randrange(p) for _ in range(40)]
    pairs = [(a, b) for a in vals[:11] for b in vals[:11]] + \
            [(rnd.choice(vals), rnd.randrange(p)) for _ in range(300)]

    def f(op, a, b):
        return fields.from_limbs32(_call(lib.zkb_host_field_op, F.fid, op,
                                         fields.to_limbs32(a, n), fields.to_limbs32(b, n), n))
    for a, b in pairs:
        assert f(0, a, b) == a * b * Rinv % p, ("montmul", hex(a), hex(b))
        assert f(1, a, b) == (a + b) % p
        assert f(2, a, b) == (a - b) % p
        assert f(7, a, b) == a * b % p
    for a in vals:
        assert f(3, a, 0) == (-a) % p
        assert f(4, a, 0) == a * R % p
        assert f(5, a, 0) == a * Rinv % p
    for a in vals[1:6] + vals[11:16]:
        am = a * R % p
        assert f(6, am, 0) == pow(a, -1, p) * R % p


@pytest.mark.parametrize("C", [curves.BLS12_381_G1, curves.BN254_G1, curves.PALLAS], ids=lambda c: c.name)
def test_curve_ops(lib, C):
    n = C.coord_limbs32
    pts = C.random_points(6, 3) + [None]

    def enc(P):
        if P is None:
            return [0] * (2 * n)
        return fields.to_limbs32(P[0], n) + fields.to_limbs32(P[1], n)

    def dec(l):
        x, y = fields.from_limbs32(l[:n]), fields.from_limbs32(l[n:])
        return None if x == 0 and y == 0 else (x, y)

    def f(op, P, Q):
        return dec(_call(lib.zkb_host_curve_op, C.cid, op, enc(P), enc(Q), 2 * n))
    cases = [(P, Q) for P in pts for Q in pts]
    cases += [(pts[0], C.neg(pts[0]))]
    for P, Q in cases:
        assert f(0, P, Q) == C.add(P, Q)
        assert f(1, P, Q) == C.add(P, Q)
        assert f(2, P, Q) == C.add(P, P)
        assert f(3, P, Q) == C.add(P, C.add(C.add(P, P), Q))
