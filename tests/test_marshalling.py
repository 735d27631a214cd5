"""Groth16 wire format (crypto3_zk_b200/marshalling.py; reference: r1cs_gg_ppzksnark/marshalling.hpp).

The compressed point encoding is crypto3-algebra's `curve_element_serializer<bls12<381>>` (not vendored); the
reference holds no byte vectors of it, so the encoding is pinned on the published ZCash encodings of the two
generators and on round trips against the oracle's curve arithmetic.  The reference's own test of this format
(test/systems/ppzksnark/r1cs_gg_ppzksnark/run_r1cs_gg_ppzksnark_tvm_marshalling.hpp:89-200) is the same kind of check:
generate, prove, serialise key / input / proof, deserialise, compare field by field - no byte vectors.
test_oracle_groth16_key_and_proof_through_the_wire mirrors it on CPU, tests/test_gpu_flows.py (wire leg of
test_groth16_prove_vs_oracle) with the device prover in the middle."""
import random

import pytest

from crypto3_zk_b200 import marshalling as m
from oracle import curves, fields

G1 = curves.BLS12_381_G1
G2 = curves.BLS12_381_G2

G1_GEN_HEX = ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
              "6c55e83ff97a1aeffb3af00adb22c6bb")
G2_GEN_HEX = ("93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049"
              "334cf11213945d57e5ac7d055d042b7e024aa2b2f08f0a91260805272dc51051"
              "c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")


def test_generator_encodings():
    assert m.g1_to_bytes(G1.gen).hex() == G1_GEN_HEX
    assert m.g2_to_bytes(G2.gen).hex() == G2_GEN_HEX
    assert m.g1_from_bytes(bytes.fromhex(G1_GEN_HEX)) == G1.gen
    assert m.g2_from_bytes(bytes.fromhex(G2_GEN_HEX)) == G2.gen
    # the negated generators differ in the sign bit only
    assert m.g1_to_bytes(G1.neg(G1.gen)).hex() == "b" + G1_GEN_HEX[1:]
    assert m.g2_to_bytes(G2.neg(G2.gen)).hex() == "b" + G2_GEN_HEX[1:]


def test_infinity():
    assert m.g1_to_bytes(None) == bytes([0xC0]) + bytes(47)
    assert m.g2_to_bytes(None) == bytes([0xC0]) + bytes(95)
    assert m.g1_from_bytes(m.g1_to_bytes(None)) is None
    assert m.g2_from_bytes(m.g2_to_bytes(None)) is None
    with pytest.raises(m.InvalidMsgData):
        m.g1_from_bytes(bytes([0xC0]) + bytes(46) + b"\x01")


def test_point_round_trips():
    p1 = G1.random_points(24, 5)
    p2 = G2.random_points(12, 6)
    for P in p1 + [G1.neg(Q) for Q in p1]:
        b = m.g1_to_bytes(P)
        assert len(b) == m.G1_BYTES and b[0] & 0x80
        assert m.g1_from_bytes(b) == P and G1.is_on_curve(m.g1_from_bytes(b))
    for P in p2 + [G2.neg(Q) for Q in p2]:
        b = m.g2_to_bytes(P)
        assert len(b) == m.G2_BYTES and b[0] & 0x80
        assert m.g2_from_bytes(b) == P


def test_point_rejects():
    with pytest.raises(m.NotEnoughData):
        m.g1_from_bytes(bytes(47))
    with pytest.raises(m.InvalidMsgData):          # uncompressed flag
        m.g1_from_bytes(bytes(48))
    # an x with no point: search the first one above the generator's x
    x = G1.gen[0]
    while pow((x ** 3 + 4) % m.P, (m.P - 1) // 2, m.P) == 1:
        x += 1
    bad = bytearray(x.to_bytes(48, "big"))
    bad[0] |= 0x80
    with pytest.raises(m.InvalidMsgData):
        m.g1_from_bytes(bytes(bad))
    with pytest.raises(m.InvalidMsgData):          # x >= p
        m.g1_from_bytes(bytes([0x9F]) + b"\xff" * 47)


def test_scalars():
    assert m.size_t_to_bytes(0x01020304) == bytes([1, 2, 3, 4])
    assert m.size_t_from_bytes(bytes([1, 2, 3, 4])) == 0x01020304
    assert m.fr_to_bytes(1) == b"\x01" + bytes(31)
    assert m.fr_to_bytes(m.R - 1)[-1] == 0x73
    with pytest.raises(m.InvalidMsgData):
        m.fr_from_bytes(m.R.to_bytes(32, "little"))
    with pytest.raises(m.NotEnoughData):
        m.fr_from_bytes(bytes(31))
    gt = tuple(tuple((6 * i + 2 * j + 1, 6 * i + 2 * j + 2) for j in range(3)) for i in range(2))
    b = m.gt_to_bytes(gt)
    assert len(b) == m.GT_BYTES == 576 and b[0] == 1 and b[48] == 2 and b[11 * 48] == 12
    assert m.gt_from_bytes(b) == gt


def _sample():
    rng = random.Random(11)
    pts = G1.random_points(6, 9)
    q = G2.random_points(3, 10)
    proof = (pts[0], q[0], pts[1])
    pi = [rng.randrange(m.R) for _ in range(3)]
    gt = tuple(tuple((rng.randrange(m.P), rng.randrange(m.P)) for _ in range(3)) for _ in range(2))
    vk = {"alpha_g1_beta_g2": gt, "gamma_g2": q[1], "delta_g2": q[2], "gamma_ABC_g1": (pts[2], pts[3:6])}
    return vk, pi, proof


def test_proof_layout():
    vk, pi, proof = _sample()
    b = m.proof_to_bytes(proof)
    assert len(b) == m.PROOF_BYTES == 192
    assert b[:48] == m.g1_to_bytes(proof[0]) and b[48:144] == m.g2_to_bytes(proof[1]) and b[144:] == m.g1_to_bytes(proof[2])
    assert m.proof_from_bytes(b) == proof
    with pytest.raises(m.NotEnoughData):
        m.proof_from_bytes(b[:-1])


def test_verifier_input_round_trip():
    vk, pi, proof = _sample()
    blob = m.verifier_input_to_bytes(vk, pi, proof)
    n = len(vk["gamma_ABC_g1"][1])
    # proof | count + Fr's | Fp12 + 2 G2 | first + (count, indices, values, domain)
    assert len(blob) == 192 + (4 + 32 * len(pi)) + (576 + 2 * 96) + (48 + 4 + 4 * n + 48 * n + 4)
    off = 192
    assert blob[off:off + 4] == (len(pi)).to_bytes(4, "big")
    assert blob[off + 4:off + 36] == pi[0].to_bytes(32, "little")
    off_av = 192 + 4 + 32 * len(pi) + 576 + 192 + 48
    assert blob[off_av:off_av + 4] == n.to_bytes(4, "big")
    assert [int.from_bytes(blob[off_av + 4 + 4 * i:off_av + 8 + 4 * i], "big") for i in range(n)] == list(range(n))
    assert blob[-4:] == n.to_bytes(4, "big")
    vk2, pi2, proof2 = m.verifier_input_from_bytes(blob)
    assert proof2 == proof and pi2 == pi
    assert vk2["alpha_g1_beta_g2"] == vk["alpha_g1_beta_g2"]
    assert vk2["gamma_g2"] == vk["gamma_g2"] and vk2["delta_g2"] == vk["delta_g2"]
    assert vk2["gamma_ABC_g1"] == (vk["gamma_ABC_g1"][0], list(vk["gamma_ABC_g1"][1]))
    for cut in (100, 192 + 3, 192 + 4 + 32 * len(pi) + 10, len(blob) - 1):
        with pytest.raises(m.NotEnoughData):
            m.verifier_input_from_bytes(blob[:cut])


def test_sparse_vector():
    pts = G1.random_points(3, 12)
    b = m.g1_sparse_vector_to_bytes([1, 4, 7], pts, 9)
    ind, val, dom, used = m.g1_sparse_vector_from_bytes(b)
    assert (ind, val, dom, used) == ([1, 4, 7], pts, 9, len(b))
    assert m.g1_sparse_vector_from_bytes(m.g1_sparse_vector_to_bytes([], [], 0))[:3] == ([], [], 0)


def _sample_pk():
    rng = random.Random(13)
    g1 = G1.random_points(12, 14)
    g2 = G2.random_points(5, 15)
    fr = lambda: rng.randrange(m.R)
    constraints = [([(0, 1), (2, fr())], [(1, fr())], [(3, fr()), (4, fr()), (1, fr())]),
                   ([], [(2, 1)], [(4, fr())]),
                   ([(1, fr())], [], [])]
    return dict(alpha_g1=g1[0], beta_g1=g1[1], beta_g2=g2[0], delta_g1=g1[2], delta_g2=g2[1],
                A_query=g1[3:6] + [None], B_indices=[0, 2, 3], B_g2=g2[2:5], B_g1=g1[6:9], B_domain_size=5,
                H_query=g1[9:11], L_query=g1[11:12], num_inputs=1, num_aux=3, constraints=constraints)


def test_proving_key_round_trip():
    pk = _sample_pk()
    body = m.proving_key_to_bytes(pk, pad=False)
    blob = m.proving_key_to_bytes(pk)
    con = [sum(len(s) * 36 + 4 for s in c) for c in pk["constraints"]]
    kc = (2 + 3) * 4 + 3 * 144
    assert len(body) == 3 * 48 + 2 * 96 + (4 + 4 * 48) + (4 + kc) + (4 + 2 * 48) + (4 + 48) + 12 + sum(4 + c for c in con)
    # the writer's buffer: twice its estimate, zero tail (marshalling.hpp:1119-1129)
    assert len(blob) == 2 * (3 * 48 + 2 * 96 + 4 * 48 + kc + 2 * 48 + 48 + 8 + sum(con))
    assert blob[:len(body)] == body and not any(blob[len(body):])
    off_b = 3 * 48 + 2 * 96 + 4 + 4 * 48
    assert blob[off_b:off_b + 4] == kc.to_bytes(4, "big") and blob[off_b + 4:off_b + 8] == (3).to_bytes(4, "big")
    for b in (body, blob):
        back = m.proving_key_from_bytes(b)
        assert back == pk
    cs = m.r1cs_constraint_system_to_bytes(1, 3, pk["constraints"])
    assert cs[:12] == bytes([0, 0, 0, 1, 0, 0, 0, 3, 0, 0, 0, 3]) and cs[12:16] == con[0].to_bytes(4, "big")
    assert m.r1cs_constraint_system_from_bytes(cs) == (1, 3, pk["constraints"], len(cs))
    with pytest.raises(m.NotEnoughData):
        m.proving_key_from_bytes(body[:-5])
    bad = bytearray(cs)
    bad[15] ^= 4
    with pytest.raises(m.MarshallingError):
        m.r1cs_constraint_system_from_bytes(bytes(bad))


def test_oracle_groth16_key_and_proof_through_the_wire():
    """A real (tiny) BLS12-381 Groth16 key and proof of the oracle prover survive the byte format unchanged."""
    from oracle import groth16
    F = fields.BLS12_381_FR
    cs, primary, aux = groth16.example_with_field_input(F, 13, 2, seed=5)
    t, alpha, beta, gamma, delta = fields.random_elements(F, 5, 21)
    pk = groth16.generator(cs, G1, G2, F, t, alpha, beta, gamma, delta)
    d = dict(alpha_g1=pk.alpha_g1, beta_g1=pk.beta_g1, beta_g2=pk.beta_g2, delta_g1=pk.delta_g1, delta_g2=pk.delta_g2,
             A_query=pk.A_query, B_indices=pk.B_indices, B_g2=pk.B_g2, B_g1=pk.B_g1, B_domain_size=len(pk.A_query),
             H_query=pk.H_query, L_query=pk.L_query, num_inputs=pk.cs.num_inputs, num_aux=pk.cs.num_aux,
             constraints=pk.cs.constraints)
    k = m.proving_key_from_bytes(m.proving_key_to_bytes(d))
    norm = lambda cons: [tuple([tuple(x) for x in side] for side in c) for c in cons]
    for key in d:
        if key == "constraints":
            assert norm(k[key]) == norm(d[key])
        else:
            assert k[key] == d[key], key
    r, s = fields.random_elements(F, 2, 22)
    proof = groth16.prove(pk, primary, aux, r, s, G1, G2, F)
    wire = m.proof_to_bytes(proof)
    assert len(wire) == 192 and m.proof_from_bytes(wire) == tuple(proof)
    assert m.primary_input_from_bytes(m.primary_input_to_bytes(primary)) == list(primary)
