"""GPU parity of zkb_points_decompress: the compressed BLS12-381 point encoding of the Groth16 wire format
(r1cs_gg_ppzksnark/marshalling.hpp:97-198 readers; curve_element_serializer<bls12<381>> of crypto3-algebra), one square
root per point on the device, against the host reader of crypto3_zk_b200/marshalling.py (pinned on the published
generator encodings by tests/test_marshalling.py) and the oracle's curve arithmetic."""
import numpy as np
import pytest

from oracle import curves, fields

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from crypto3_zk_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _points(arr, cl, deg):
    from crypto3_zk_b200.api import _affine_from_limbs
    a = np.asarray(arr).view(np.uint32)
    return [_affine_from_limbs(a[i].reshape(-1), cl, deg) for i in range(a.shape[0])]


def test_g1_decompress_vs_host_reader(ctx):
    import torch
    from crypto3_zk_b200 import marshalling as m
    C = curves.BLS12_381_G1
    pts = C.random_points(200, 3) + [None, C.gen, C.neg(C.gen)]
    blob = b"".join(m.g1_to_bytes(p) for p in pts)
    assert sum(1 for i in range(len(pts)) if blob[48 * i] & 0x20) not in (0, len(pts))       # both sign flags occur
    got = ctx.points_decompress("bls12_381_g1", blob, len(pts))                               # host buffers
    assert _points(got, 12, 1) == pts
    d = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).cuda()
    got, st = ctx.points_decompress("bls12_381_g1", d, len(pts), status=True)                 # device buffers
    assert _points(got.cpu().numpy(), 12, 1) == pts
    assert st.cpu().tolist() == [1 if p is None else 0 for p in pts]
    # offset + count inside a larger buffer
    got = ctx.points_decompress("bls12_381_g1", b"\x00" * 7 + blob, 5, offset=7 + 48 * 2)
    assert _points(got, 12, 1) == pts[2:7]


def test_g2_decompress_vs_host_reader(ctx):
    from crypto3_zk_b200 import marshalling as m
    C = curves.BLS12_381_G2
    pts = C.random_points(60, 4) + [None, C.gen, C.neg(C.gen)]
    blob = b"".join(m.g2_to_bytes(p) for p in pts)
    got = ctx.points_decompress("bls12_381_g2", blob, len(pts))
    assert _points(got, 24, 2) == pts
    assert [m.g2_from_bytes(blob, 96 * i) for i in range(len(pts))] == pts


def test_interleaved_knowledge_commitment_layout(ctx):
    """a knowledge_commitment_vector stores (G2 | G1) pairs: both halves are read with stride 144"""
    from crypto3_zk_b200 import marshalling as m
    g2 = curves.BLS12_381_G2.random_points(9, 5)
    g1 = curves.BLS12_381_G1.random_points(9, 6)
    blob = b"".join(m.g2_to_bytes(a) + m.g1_to_bytes(b) for a, b in zip(g2, g1))
    assert _points(ctx.points_decompress("bls12_381_g2", blob, 9, stride=144), 24, 2) == g2
    assert _points(ctx.points_decompress("bls12_381_g1", blob, 9, offset=96, stride=144), 12, 1) == g1


def test_malformed_encodings(ctx):
    """the reader's invalid_msg_data cases: uncompressed form, infinity flag with payload, x >= p, x off the curve"""
    from crypto3_zk_b200 import capi
    from crypto3_zk_b200 import marshalling as m
    C = curves.BLS12_381_G1
    p = fields.BLS12_381_FQ.p
    good = m.g1_to_bytes(C.gen)
    not_compressed = bytes([good[0] & 0x7F]) + good[1:]
    bad_inf = bytes([0xC0]) + bytes(46) + b"\x01"
    sign_inf = bytes([0xE0]) + bytes(47)
    not_reduced = bytearray(p.to_bytes(48, "big"))
    not_reduced[0] |= 0x80
    x = 1
    while pow((x ** 3 + 4) % p, (p - 1) // 2, p) == 1:          # a non-residue: no y
        x += 1
    off_curve = bytearray(x.to_bytes(48, "big"))
    off_curve[0] |= 0x80
    cases = [good, not_compressed, bad_inf, sign_inf, bytes(not_reduced), bytes(off_curve), m.g1_to_bytes(None)]
    blob = b"".join(cases)
    pts, st = ctx.points_decompress("bls12_381_g1", blob, len(cases), status=True)
    assert list(st) == [0, 2, 3, 3, 4, 5, 1]
    assert _points(pts, 12, 1) == [C.gen, None, None, None, None, None, None]
    with pytest.raises(capi.ZkbInvalidArgument, match="point 1 "):
        ctx.points_decompress("bls12_381_g1", blob, len(cases))
    for bad in cases[1:6]:
        with pytest.raises(m.InvalidMsgData):
            m.g1_from_bytes(bad)
    # G2: x off the curve and a non-reduced c0
    C2 = curves.BLS12_381_G2
    g = bytearray(m.g2_to_bytes(C2.gen))
    k = 0
    while True:
        k += 1
        trial = bytearray(g)
        trial[95] = (trial[95] + k) & 0xFF
        try:
            m.g2_from_bytes(bytes(trial))
        except m.InvalidMsgData:
            break
    nr = bytearray(g)
    nr[48:96] = p.to_bytes(48, "big")
    pts, st = ctx.points_decompress("bls12_381_g2", bytes(g) + bytes(trial) + bytes(nr), 3, status=True)
    assert list(st) == [0, 5, 4] and _points(pts, 24, 2)[0] == C2.gen
    with pytest.raises(capi.ZkbError):
        ctx.points_decompress("bn254_g1", bytes(32 * 4), 1)        # no such encoding upstream


def test_decompress_a_large_vector_matches_the_msm_bases(ctx):
    """2^16 points: compress on the host from device-made points (x, sign of y), decompress on the device, same points"""
    import torch
    from crypto3_zk_b200 import workloads as W
    n = 1 << 16
    pts = W.curve_grid_points(ctx, "bls12_381_g1", n, seed=11)
    a = pts.cpu().numpy().view(np.uint32)                                    # [n, 2, 12] little-endian limbs
    p = fields.BLS12_381_FQ.p
    half = (p - 1) // 2
    xb = a[:, 0, ::-1].astype(">u4").tobytes()                               # big-endian x, 48 bytes per point
    blob = bytearray(xb)
    ys = fields.u32_array_to_ints(a[:, 1, :])
    for i, y in enumerate(ys):
        blob[48 * i] |= 0x80 | (0x20 if y > half else 0)
    d = torch.from_numpy(np.frombuffer(bytes(blob), dtype=np.uint8).copy()).cuda()
    got = ctx.points_decompress("bls12_381_g1", d, n)
    assert torch.equal(got, pts)
