"""Field parameter table + plain big-int modular helpers (oracle; test infrastructure only).

Constants follow crypto3-algebra's `fields::arithmetic_params<F>` (un-vendored; libff lineage):
modulus, two-adicity s, multiplicative_generator g, root_of_unity = g^((p-1)/2^s).
In-repo corroboration: tests use `arithmetic_params<F>::multiplicative_generator` as the coset
shift (/root/reference/include/nil/crypto3/zk/snark/reductions/r1cs_to_qap.hpp:266-269) and the
bit widths `_cppui_modular254/255/381` (test/transcript/transcript.cpp:58,
r1cs_gg_ppzksnark_aggregation_conformity.cpp:202,215).
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class Field:
    name: str
    fid: int          # id used across the C ABI (include/zkb200.h ZKB_FIELD_*)
    p: int
    bits: int
    s: int            # two-adicity
    g: int            # multiplicative generator
    limbs32: int

    @property
    def root_of_unity(self):
        return pow(self.g, (self.p - 1) >> self.s, self.p)

    @property
    def nbytes(self):
        return (self.bits + 7) // 8

    def omega(self, log_n):
        """unity_root<F>(2^log_n) = root_of_unity^(2^(s-log_n)) (Appendix A.1)."""
        if log_n > self.s:
            raise ValueError("domain larger than two-adicity")
        return pow(self.root_of_unity, 1 << (self.s - log_n), self.p)

    def inv(self, a):
        return pow(a, self.p - 2, self.p)


BLS12_381_FR = Field("bls12_381_fr", 0,
                     0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001, 255, 32, 7, 8)
BN254_FR = Field("bn254_fr", 1,
                 21888242871839275222246405745257275088548364400416034343698204186575808495617, 254, 28, 5, 8)
PALLAS_FP = Field("pallas_fp", 2,   # Pallas base field (= Vesta scalar field)
                  0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001, 255, 32, 5, 8)
PALLAS_FQ = Field("pallas_fq", 3,   # Pallas scalar field (= Vesta base field)
                  0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001, 255, 32, 5, 8)
BLS12_381_FQ = Field("bls12_381_fq", 4,
                     0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab,
                     381, 1, 2, 12)
BN254_FQ = Field("bn254_fq", 5,
                 21888242871839275222246405745257275088696311157297823662689037894645226208583, 254, 1, 3, 8)

FIELDS = {f.name: f for f in (BLS12_381_FR, BN254_FR, PALLAS_FP, PALLAS_FQ, BLS12_381_FQ, BN254_FQ)}
FIELDS_BY_ID = {f.fid: f for f in FIELDS.values()}
NTT_FIELDS = (BLS12_381_FR, BN254_FR, PALLAS_FP, PALLAS_FQ)


# --------------------------------------------------------------------------- limb packing helpers
def to_limbs32(x, n):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def from_limbs32(limbs):
    r = 0
    for i, l in enumerate(limbs):
        r |= int(l) << (32 * i)
    return r


def ints_to_u32_array(vals, limbs):
    """list[int] -> numpy uint32 array [len, limbs], little-endian limbs (the C-ABI layout)."""
    import numpy as np
    nbytes = limbs * 4
    buf = b"".join(int(v).to_bytes(nbytes, "little") for v in vals)
    return np.frombuffer(buf, dtype="<u4").reshape(len(vals), limbs).copy()


def u32_array_to_ints(arr):
    import numpy as np
    arr = np.ascontiguousarray(arr, dtype="<u4")
    limbs = arr.shape[-1]
    raw = arr.tobytes()
    nbytes = limbs * 4
    return [int.from_bytes(raw[i:i + nbytes], "little") for i in range(0, len(raw), nbytes)]


def random_elements(field, n, seed):
    """Deterministic synthetic field elements: 256 (or 384) random bits reduced mod p
    (PCG64(seed); bias <= 2^-128 for the 8-limb fields)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    raw = rng.integers(0, 1 << 32, size=(n, field.limbs32 + (4 if field.limbs32 == 12 else 0)), dtype=np.uint64)
    out = []
    for row in raw:
        v = 0
        for i, l in enumerate(row):
            v |= int(l) << (32 * i)
        out.append(v % field.p)
    return out
