"""Field / curve parameter table of the product (host side).

Mirrors crypto3-algebra's `fields::arithmetic_params<F>` / curve params as they are used at the
reference call sites (multiplicative_generator as coset shift: zk/snark/reductions/
r1cs_to_qap.hpp:266-269; unity_root via make_evaluation_domain: r1cs_to_qap.hpp:229).
The numeric ids are the ZKB_FIELD_* / ZKB_CURVE_* values of include/zkb200.h.
csrc/gen_params.py turns this table into csrc/zkb_params.cuh.
"""
from collections import namedtuple

FieldParams = namedtuple("FieldParams", "name fid p bits two_adicity generator limbs32")
# deg = extension degree of the coordinate field (2 for the G2 groups: coordinates (c0, c1) in Fq[u]/(u^2+1),
# c0 || c1 at the ABI); gen_x / gen_y are ints for deg 1 and (c0, c1) tuples for deg 2
CurveParams = namedtuple("CurveParams", "name cid base_field scalar_field b gen_x gen_y deg")

FIELDS = [
    FieldParams("bls12_381_fr", 0,
                0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 255, 32, 7, 8),
    FieldParams("bn254_fr", 1,
                21888242871839275222246405745257275088548364400416034343698204186575808495617, 254, 28, 5, 8),
    FieldParams("pallas_fp", 2,
                0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001, 255, 32, 5, 8),
    FieldParams("pallas_fq", 3,
                0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001, 255, 32, 5, 8),
    FieldParams("bls12_381_fq", 4,
                0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
                381, 1, 2, 12),
    FieldParams("bn254_fq", 5,
                21888242871839275222246405745257275088696311157297823662689037894645226208583, 254, 1, 3, 8),
]
FIELD_BY_NAME = {f.name: f for f in FIELDS}
FIELD_BY_ID = {f.fid: f for f in FIELDS}

CURVES = [
    CurveParams("bls12_381_g1", 0, "bls12_381_fq", "bls12_381_fr", 4,
                0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
                0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1, 1),
    CurveParams("bn254_g1", 1, "bn254_fq", "bn254_fr", 3, 1, 2, 1),
    CurveParams("pallas", 2, "pallas_fp", "pallas_fq", 5,
                0x40000000000000000000000000000000224698fc094cf91b992d30ed00000000, 2, 1),
    CurveParams("bls12_381_g2", 3, "bls12_381_fq", "bls12_381_fr", (4, 4),
                (0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
                 0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
                (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
                 0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be), 2),
    CurveParams("bn254_g2", 4, "bn254_fq", "bn254_fr", None,   # b = 3 / (9 + u)
                (10857046999023057135944570762232829481370756359578518086990519993285655852781,
                 11559732032986387107991004021392285783925812861821192530917403151452391805634),
                (8495653923123431417604973247489272438418190587263600148770280649306958101930,
                 4082367875863433681332203403145435568316851327593401208105741076214120093531), 2),
]
CURVE_BY_NAME = {c.name: c for c in CURVES}
CURVE_BY_ID = {c.cid: c for c in CURVES}


def coord_limbs(c):
    """u32 limbs per affine coordinate of curve `c` at the C ABI."""
    return FIELD_BY_NAME[c.base_field].limbs32 * c.deg


def root_of_unity(f):
    return pow(f.generator, (f.p - 1) >> f.two_adicity, f.p)


def omega(f, log_n):
    """unity_root<F>(2^log_n)."""
    if log_n > f.two_adicity:
        raise ValueError("domain of size 2^%d exceeds the field's two-adicity %d" % (log_n, f.two_adicity))
    return pow(root_of_unity(f), 1 << (f.two_adicity - log_n), f.p)
