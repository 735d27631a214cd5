"""Field / curve parameter table of the product (host side).

Mirrors crypto3-algebra's `fields::arithmetic_params<F>` / curve params as they are used at the
reference call sites (multiplicative_generator as coset shift: zk/snark/reductions/
r1cs_to_qap.hpp:266-269; unity_root via make_evaluation_domain: r1cs_to_qap.hpp:229).
The numeric ids are the ZKB_FIELD_* / ZKB_CURVE_* values of include/zkb200.h.
csrc/gen_params.py turns this table into csrc/zkb_params.cuh.
"""
from collections import namedtuple

FieldParams = namedtuple("FieldParams", "name fid p bits two_adicity generator limbs32")
CurveParams = namedtuple("CurveParams", "name cid base_field scalar_field b gen_x gen_y")

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
                0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
    CurveParams("bn254_g1", 1, "bn254_fq", "bn254_fr", 3, 1, 2),
    CurveParams("pallas", 2, "pallas_fp", "pallas_fq", 5,
                0x40000000000000000000000000000000224698fc094cf91b992d30ed00000000, 2),
]
CURVE_BY_NAME = {c.name: c for c in CURVES}
CURVE_BY_ID = {c.cid: c for c in CURVES}


def root_of_unity(f):
    return pow(f.generator, (f.p - 1) >> f.two_adicity, f.p)


def omega(f, log_n):
    """unity_root<F>(2^log_n)."""
    if log_n > f.two_adicity:
        raise ValueError("domain of size 2^%d exceeds the field's two-adicity %d" % (log_n, f.two_adicity))
    return pow(root_of_unity(f), 1 << (f.two_adicity - log_n), f.p)
