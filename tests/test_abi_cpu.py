"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/zkb200.h declares,
its host-only descriptors agree with the oracle, and compute calls fail loudly without a device."""
import ctypes
import os
import re

import pytest

from oracle import fields as ofields

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from crypto3_zk_b200 import build, capi
    build.build()
    return capi.lib()


def test_header_symbols_exported(L):
    from crypto3_zk_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    declared = set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s


def test_field_descriptors_match_oracle(L):
    from crypto3_zk_b200 import api, fields
    for f in ofields.FIELDS.values():
        pf = fields.FIELD_BY_NAME[f.name]
        assert (pf.fid, pf.p, pf.two_adicity, pf.generator, pf.limbs32) == (f.fid, f.p, f.s, f.g, f.limbs32)
        assert L.zkb_field_limbs(f.fid) == f.limbs32
        assert L.zkb_field_two_adicity(f.fid) == f.s
        assert api.field_generator(f.fid) == f.g
    for f in ofields.NTT_FIELDS:
        for k in (0, 1, 5, 20, f.s):
            assert api.unity_root(f.fid, k) == f.omega(k)
        with pytest.raises(ValueError):
            api.unity_root(f.fid, f.s + 1)


def test_msm_combine_host(L):
    """zkb_msm_combine is host-only: adding XYZZ partials (Z = 1, Montgomery limbs) == oracle sum."""
    from crypto3_zk_b200 import api
    from oracle import curves
    for C in (curves.BLS12_381_G1, curves.BN254_G1, curves.PALLAS):
        n = C.coord_limbs32
        R = pow(2, 32 * n, C.base_field.p)
        pts = C.random_points(3, 9)
        rows = []
        for (x, y) in pts:
            rows.append(ofields.to_limbs32(x * R % C.base_field.p, n) + ofields.to_limbs32(y * R % C.base_field.p, n) +
                        ofields.to_limbs32(R % C.base_field.p, n) * 2)
        rows.append([0] * (4 * n))
        got = api.msm_combine(C.name, rows)
        assert got == C.add(C.add(pts[0], pts[1]), pts[2])
        assert api.msm_combine(C.name, [[0] * (4 * n)]) is None


def test_no_device_fails_loudly(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from crypto3_zk_b200 import Context, capi
    with pytest.raises(capi.ZkbError) as e:
        Context(0)
    assert e.value.status == capi.ERR_NO_DEVICE
