"""Groth16 proof / verification-key wire format for BLS12-381 (host side only - no device work).

Mirrors `verifier_input_serializer_tvm` / `verifier_input_deserializer_tvm<r1cs_gg_ppzksnark<bls12<381>>>`
(zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/marshalling.hpp:100-890 reader, :897-1258 writer):

  std::size_t      4 bytes big-endian                           (:465-491, :975-985)
  Fr / Fp element  fixed width ceil(modulus_bits / 8) = 32 / 48 bytes, LEAST significant byte first
                   (export_bits(..., 8, false), :921-935)
  Fp2/Fp6/Fp12     the coefficients data[0], data[1], ... in order, recursively (:937-949)
  G1 / G2 point    48 / 96 bytes, `curve_element_serializer<bls12<381>>::point_to_octets_compress` (:951-973)
  proof            g_A (G1) | g_B (G2) | g_C (G1) = 192 bytes     (:784-828)
  primary input    count | count x Fr                            (:740-782)
  sparse_vector    count | count x index | count x G1 | domain_size   (:493-569)
  accumulation_vector  first (G1) | rest (sparse_vector)         (:571-598)
  verification key alpha_g1_beta_g2 (Fp12, 576 B) | gamma_g2 | delta_g2 | gamma_ABC_g1   (:600-653)
  verifier input   proof | primary input | verification key      (:830-890)
  linear_term      index | Fr;  linear_combination  count | terms (:204-258, :1020-1036)
  r1cs_constraint  byte size of (a | b | c) | a | b | c;  constraint system  #primary | #auxiliary | count | constraints
                                                                 (:260-372, :1038-1071)
  kc vector (G2,G1)  count | count x index | count x (G2 | G1) | domain_size     (:374-461, :1073-1115)
  proving key      alpha_g1 | beta_g1 | beta_g2 | delta_g1 | delta_g2 | count, A_query | byte size, B_query (kc vector)
                   | count, H_query | count, L_query | constraint system, zero-padded to the writer's buffer size
                                                                 (:656-738, :1117-1163)

`curve_element_serializer` lives in crypto3-algebra (not vendored in the reference).  It implements the ZCash
BLS12-381 encoding: x big-endian (for G2: x.c1 then x.c0), bit 7 of byte 0 = compressed, bit 6 = point at
infinity, bit 5 = y is the lexicographically larger of (y, -y) (for Fp2: compare c1, then c0 when c1 == 0).
The reference's tests hold no byte vector of this format; tests/test_marshalling.py pins it on the published
encodings of the two generators.

Error behaviour follows the reader: a short buffer raises `NotEnoughData` (status_type::not_enough_data), a
non-canonical field element or an x with no point on the curve raises `InvalidMsgData`
(status_type::invalid_msg_data).  Points are affine (x, y) integers ((c0, c1) pairs on G2), None = infinity, as
in groth16.py.
"""
import numpy as np

from .fields import FIELD_BY_NAME

P = FIELD_BY_NAME["bls12_381_fq"].p
R = FIELD_BY_NAME["bls12_381_fr"].p

SIZE_T_BYTES = 4
FR_BYTES = 32
FP_BYTES = 48
G1_BYTES = 48
G2_BYTES = 96
GT_BYTES = 2 * 3 * 2 * FP_BYTES
PROOF_BYTES = G1_BYTES + G2_BYTES + G1_BYTES

_C_BIT, _I_BIT, _S_BIT = 0x80, 0x40, 0x20
_HALF = (P - 1) // 2


class MarshallingError(ValueError):
    pass


class NotEnoughData(MarshallingError):
    """status_type::not_enough_data"""


class InvalidMsgData(MarshallingError):
    """status_type::invalid_msg_data"""


def _need(buf, off, n):
    if len(buf) - off < n:
        raise NotEnoughData("need %d bytes at offset %d, have %d" % (n, off, len(buf) - off))


# ---- Fp2 helpers for the decompression (u^2 = -1)
def _f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def _f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = _f2_mul(r, a)
        a = _f2_mul(a, a)
        e >>= 1
    return r


def _f2_sqrt(a):
    """square root in Fp2 for p = 3 mod 4, or None"""
    if a == (0, 0):
        return (0, 0)
    a1 = _f2_pow(a, (P - 3) // 4)
    alpha = _f2_mul(_f2_mul(a1, a1), a)
    x0 = _f2_mul(a1, a)
    if alpha == (P - 1, 0):
        x = _f2_mul((0, 1), x0)
    else:
        b = _f2_pow(((1 + alpha[0]) % P, alpha[1]), (P - 1) // 2)
        x = _f2_mul(b, x0)
    return x if _f2_mul(x, x) == (a[0] % P, a[1] % P) else None


def _sign_fp(v):
    return v > _HALF


def _sign_fp2(v):
    return _sign_fp(v[0]) if v[1] == 0 else _sign_fp(v[1])


# ---- scalars
def size_t_to_bytes(v):
    if not 0 <= v < 1 << 32:
        raise MarshallingError("std::size_t value does not fit the 4-byte wire field")
    return int(v).to_bytes(SIZE_T_BYTES, "big")


def size_t_from_bytes(buf, off=0):
    _need(buf, off, SIZE_T_BYTES)
    return int.from_bytes(buf[off:off + SIZE_T_BYTES], "big")


def field_to_bytes(v, modulus, width):
    return (int(v) % modulus).to_bytes(width, "little")


def field_from_bytes(buf, off, modulus, width):
    _need(buf, off, width)
    v = int.from_bytes(buf[off:off + width], "little")
    if v >= modulus:
        raise InvalidMsgData("field element is not reduced")
    return v


def fr_to_bytes(v):
    return field_to_bytes(v, R, FR_BYTES)


def fr_from_bytes(buf, off=0):
    return field_from_bytes(buf, off, R, FR_BYTES)


def gt_to_bytes(v):
    """v: the 12 Fp coefficients in data[] order: ((c00, c01), (c10, c11), (c20, c21)) x 2, flattened or nested."""
    flat = []

    def walk(x):
        if isinstance(x, (tuple, list)):
            for y in x:
                walk(y)
        else:
            flat.append(x)
    walk(v)
    if len(flat) != 12:
        raise MarshallingError("an Fp12 element has 12 coefficients")
    return b"".join(field_to_bytes(c, P, FP_BYTES) for c in flat)


def gt_from_bytes(buf, off=0):
    _need(buf, off, GT_BYTES)
    c = [field_from_bytes(buf, off + i * FP_BYTES, P, FP_BYTES) for i in range(12)]
    return tuple(tuple((c[6 * i + 2 * j], c[6 * i + 2 * j + 1]) for j in range(3)) for i in range(2))


# ---- points
def g1_to_bytes(pt):
    if pt is None:
        return bytes([_C_BIT | _I_BIT]) + bytes(G1_BYTES - 1)
    x, y = int(pt[0]) % P, int(pt[1]) % P
    out = bytearray(x.to_bytes(FP_BYTES, "big"))
    out[0] |= _C_BIT | (_S_BIT if _sign_fp(y) else 0)
    return bytes(out)


def g1_from_bytes(buf, off=0):
    _need(buf, off, G1_BYTES)
    b = bytes(buf[off:off + G1_BYTES])
    flags = b[0]
    if not flags & _C_BIT:
        raise InvalidMsgData("G1 point is not in compressed form")
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if flags & _I_BIT:
        if x or flags & _S_BIT:
            raise InvalidMsgData("non-zero payload in a point at infinity")
        return None
    if x >= P:
        raise InvalidMsgData("x coordinate is not reduced")
    rhs = (x * x * x + 4) % P
    y = pow(rhs, (P + 1) // 4, P)
    if y * y % P != rhs:
        raise InvalidMsgData("x is not the abscissa of a curve point")
    if _sign_fp(y) != bool(flags & _S_BIT):
        y = P - y
    return (x, y)


def g2_to_bytes(pt):
    if pt is None:
        return bytes([_C_BIT | _I_BIT]) + bytes(G2_BYTES - 1)
    (x0, x1), (y0, y1) = pt
    out = bytearray((int(x1) % P).to_bytes(FP_BYTES, "big") + (int(x0) % P).to_bytes(FP_BYTES, "big"))
    out[0] |= _C_BIT | (_S_BIT if _sign_fp2((int(y0) % P, int(y1) % P)) else 0)
    return bytes(out)


def g2_from_bytes(buf, off=0):
    _need(buf, off, G2_BYTES)
    b = bytes(buf[off:off + G2_BYTES])
    flags = b[0]
    if not flags & _C_BIT:
        raise InvalidMsgData("G2 point is not in compressed form")
    x1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:FP_BYTES], "big")
    x0 = int.from_bytes(b[FP_BYTES:], "big")
    if flags & _I_BIT:
        if x0 or x1 or flags & _S_BIT:
            raise InvalidMsgData("non-zero payload in a point at infinity")
        return None
    if x0 >= P or x1 >= P:
        raise InvalidMsgData("x coordinate is not reduced")
    x = (x0, x1)
    x3 = _f2_mul(_f2_mul(x, x), x)
    rhs = ((x3[0] + 4) % P, (x3[1] + 4) % P)
    y = _f2_sqrt(rhs)
    if y is None:
        raise InvalidMsgData("x is not the abscissa of a curve point")
    if _sign_fp2(y) != bool(flags & _S_BIT):
        y = ((-y[0]) % P, (-y[1]) % P)
    return (x, y)


# ---- containers
def g1_sparse_vector_to_bytes(indices, values, domain_size):
    if len(indices) != len(values):
        raise MarshallingError("sparse_vector: indices and values differ in length")
    return (size_t_to_bytes(len(values)) + b"".join(size_t_to_bytes(i) for i in indices)
            + b"".join(g1_to_bytes(v) for v in values) + size_t_to_bytes(domain_size))


def g1_sparse_vector_from_bytes(buf, off=0):
    """-> (indices, values, domain_size, bytes consumed)"""
    n = size_t_from_bytes(buf, off)
    total = SIZE_T_BYTES + n * SIZE_T_BYTES + n * G1_BYTES + SIZE_T_BYTES
    _need(buf, off, total)
    o = off + SIZE_T_BYTES
    indices = [size_t_from_bytes(buf, o + SIZE_T_BYTES * i) for i in range(n)]
    o += SIZE_T_BYTES * n
    values = [g1_from_bytes(buf, o + G1_BYTES * i) for i in range(n)]
    o += G1_BYTES * n
    return indices, values, size_t_from_bytes(buf, o), total


def g1_accumulation_vector_to_bytes(first, rest_values, rest_indices=None, domain_size=None):
    """accumulation_vector(first, rest): `rest` is dense in a verification key (indices 0..n-1, domain n)."""
    n = len(rest_values)
    return g1_to_bytes(first) + g1_sparse_vector_to_bytes(
        list(range(n)) if rest_indices is None else rest_indices, rest_values, n if domain_size is None else domain_size)


def g1_accumulation_vector_from_bytes(buf, off=0):
    """-> (first, (indices, values, domain_size), bytes consumed)"""
    first = g1_from_bytes(buf, off)
    ind, val, dom, used = g1_sparse_vector_from_bytes(buf, off + G1_BYTES)
    return first, (ind, val, dom), G1_BYTES + used


# ---- scheme objects
def proof_to_bytes(proof):
    """proof: (g_A, g_B, g_C) as returned by groth16.prove"""
    g_A, g_B, g_C = proof
    return g1_to_bytes(g_A) + g2_to_bytes(g_B) + g1_to_bytes(g_C)


def proof_from_bytes(buf, off=0):
    _need(buf, off, PROOF_BYTES)
    return (g1_from_bytes(buf, off), g2_from_bytes(buf, off + G1_BYTES), g1_from_bytes(buf, off + G1_BYTES + G2_BYTES))


def primary_input_to_bytes(pi):
    return size_t_to_bytes(len(pi)) + b"".join(fr_to_bytes(v) for v in pi)


def primary_input_from_bytes(buf, off=0):
    n = size_t_from_bytes(buf, off)
    _need(buf, off, SIZE_T_BYTES + n * FR_BYTES)
    return [fr_from_bytes(buf, off + SIZE_T_BYTES + i * FR_BYTES) for i in range(n)]


def verification_key_to_bytes(vk):
    """vk: dict(alpha_g1_beta_g2 = Fp12 coefficients, gamma_g2, delta_g2, gamma_ABC_g1 = (first, [rest...]))"""
    first, rest = vk["gamma_ABC_g1"]
    return (gt_to_bytes(vk["alpha_g1_beta_g2"]) + g2_to_bytes(vk["gamma_g2"]) + g2_to_bytes(vk["delta_g2"])
            + g1_accumulation_vector_to_bytes(first, rest))


def verification_key_from_bytes(buf, off=0):
    _need(buf, off, GT_BYTES + 2 * G2_BYTES)
    gt = gt_from_bytes(buf, off)
    gamma = g2_from_bytes(buf, off + GT_BYTES)
    delta = g2_from_bytes(buf, off + GT_BYTES + G2_BYTES)
    first, (ind, val, dom), _ = g1_accumulation_vector_from_bytes(buf, off + GT_BYTES + 2 * G2_BYTES)
    if ind != list(range(len(val))) or dom != len(val):
        raise InvalidMsgData("gamma_ABC_g1 is not a dense accumulation vector")
    return {"alpha_g1_beta_g2": gt, "gamma_g2": gamma, "delta_g2": delta, "gamma_ABC_g1": (first, val)}


# ---- proving key
def linear_combination_to_bytes(terms):
    """terms: [(variable index, coefficient), ...]"""
    return size_t_to_bytes(len(terms)) + b"".join(size_t_to_bytes(i) + fr_to_bytes(c) for i, c in terms)


def linear_combination_from_bytes(buf, off=0):
    """-> (terms, bytes consumed)"""
    n = size_t_from_bytes(buf, off)
    term = SIZE_T_BYTES + FR_BYTES
    _need(buf, off, SIZE_T_BYTES + n * term)
    o = off + SIZE_T_BYTES
    return ([(size_t_from_bytes(buf, o + i * term), fr_from_bytes(buf, o + i * term + SIZE_T_BYTES)) for i in range(n)],
            SIZE_T_BYTES + n * term)


def _constraint_bytes(con):
    return sum(len(side) * (SIZE_T_BYTES + FR_BYTES) + SIZE_T_BYTES for side in con)


def r1cs_constraint_to_bytes(con):
    """con: (a, b, c) linear combinations"""
    return size_t_to_bytes(_constraint_bytes(con)) + b"".join(linear_combination_to_bytes(side) for side in con)


def r1cs_constraint_system_to_bytes(num_inputs, num_aux, constraints):
    return (size_t_to_bytes(num_inputs) + size_t_to_bytes(num_aux) + size_t_to_bytes(len(constraints))
            + b"".join(r1cs_constraint_to_bytes(c) for c in constraints))


def r1cs_constraint_system_from_bytes(buf, off=0):
    """-> (num_inputs, num_aux, constraints, bytes consumed)"""
    num_inputs = size_t_from_bytes(buf, off)
    num_aux = size_t_from_bytes(buf, off + SIZE_T_BYTES)
    count = size_t_from_bytes(buf, off + 2 * SIZE_T_BYTES)
    o = off + 3 * SIZE_T_BYTES
    constraints = []
    for _ in range(count):
        total = size_t_from_bytes(buf, o)
        o += SIZE_T_BYTES
        _need(buf, o, total)
        sides, q = [], o
        for _k in range(3):
            terms, used = linear_combination_from_bytes(buf, q)
            sides.append(terms)
            q += used
        if q - o != total:
            raise InvalidMsgData("r1cs_constraint size field disagrees with its linear combinations")
        constraints.append(tuple(sides))
        o += total
    return num_inputs, num_aux, constraints, o - off


def _kc_vector_bytes(n):
    return (2 + n) * SIZE_T_BYTES + n * (G2_BYTES + G1_BYTES)


def kc_vector_to_bytes(indices, g2_values, g1_values, domain_size):
    """knowledge_commitment_vector<G2, G1> without its leading byte-size field"""
    if not len(indices) == len(g2_values) == len(g1_values):
        raise MarshallingError("knowledge_commitment_vector: indices and values differ in length")
    return (size_t_to_bytes(len(indices)) + b"".join(size_t_to_bytes(i) for i in indices)
            + b"".join(g2_to_bytes(g) + g1_to_bytes(h) for g, h in zip(g2_values, g1_values)) + size_t_to_bytes(domain_size))


def kc_vector_from_bytes(buf, off=0):
    """-> (indices, g2 values, g1 values, domain_size, bytes consumed)"""
    n = size_t_from_bytes(buf, off)
    total = _kc_vector_bytes(n)
    _need(buf, off, total)
    o = off + SIZE_T_BYTES
    indices = [size_t_from_bytes(buf, o + SIZE_T_BYTES * i) for i in range(n)]
    o += SIZE_T_BYTES * n
    pair = G2_BYTES + G1_BYTES
    g2v = [g2_from_bytes(buf, o + pair * i) for i in range(n)]
    g1v = [g1_from_bytes(buf, o + pair * i + G2_BYTES) for i in range(n)]
    return indices, g2v, g1v, size_t_from_bytes(buf, o + pair * n), total


def _g1_list_to_bytes(pts):
    return size_t_to_bytes(len(pts)) + b"".join(g1_to_bytes(p) for p in pts)


def _g1_list_from_bytes(buf, off):
    n = size_t_from_bytes(buf, off)
    _need(buf, off, SIZE_T_BYTES + n * G1_BYTES)
    return [g1_from_bytes(buf, off + SIZE_T_BYTES + i * G1_BYTES) for i in range(n)], SIZE_T_BYTES + n * G1_BYTES


def proving_key_to_bytes(pk, pad=True):
    """pk: dict(alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, A_query, B_indices, B_g2, B_g1, B_domain_size, H_query,
    L_query, num_inputs, num_aux, constraints) - the argument names of groth16.ProvingKey.  With `pad` the result has the
    length of the writer's buffer (twice its size estimate, marshalling.hpp:1119-1129; the tail is zero) so that it equals
    the reference's output byte for byte; the reader ignores the tail."""
    nb = len(pk["B_indices"])
    body = (g1_to_bytes(pk["alpha_g1"]) + g1_to_bytes(pk["beta_g1"]) + g2_to_bytes(pk["beta_g2"])
            + g1_to_bytes(pk["delta_g1"]) + g2_to_bytes(pk["delta_g2"])
            + _g1_list_to_bytes(pk["A_query"])
            + size_t_to_bytes(_kc_vector_bytes(nb))
            + kc_vector_to_bytes(pk["B_indices"], pk["B_g2"], pk["B_g1"], pk["B_domain_size"])
            + _g1_list_to_bytes(pk["H_query"]) + _g1_list_to_bytes(pk["L_query"])
            + r1cs_constraint_system_to_bytes(pk["num_inputs"], pk["num_aux"], pk["constraints"]))
    if not pad:
        return body
    estimate = (3 * G1_BYTES + 2 * G2_BYTES + len(pk["A_query"]) * G1_BYTES + _kc_vector_bytes(nb)
                + len(pk["H_query"]) * G1_BYTES + len(pk["L_query"]) * G1_BYTES + 2 * SIZE_T_BYTES
                + sum(_constraint_bytes(c) for c in pk["constraints"]))
    if len(body) > 2 * estimate:      # cannot happen: the unaccounted size fields are 4 bytes per counted item
        raise MarshallingError("proving key exceeds the writer's buffer")
    return body + bytes(2 * estimate - len(body))


def _device_points(ctx, curve, buf, off, n, stride=None):
    """n compressed points starting at byte `off` of the host blob -> device tensor [n, 2, limbs] (one square root per
    point on the GPU: zkb_points_decompress); a malformed encoding raises InvalidMsgData like the scalar readers"""
    import torch
    from . import capi
    width = G2_BYTES if curve.endswith("g2") else G1_BYTES
    stride = width if stride is None else stride
    if n == 0:
        return torch.empty((0, 2, width // 4), dtype=torch.int32, device="cuda:%d" % ctx.device)
    _need(buf, off, (n - 1) * stride + width)
    raw = np.frombuffer(buf, dtype=np.uint8, count=(n - 1) * stride + width, offset=off)
    octets = torch.from_numpy(raw.copy()).to("cuda:%d" % ctx.device)
    try:
        return ctx.points_decompress(curve, octets, n, stride=stride)
    except capi.ZkbInvalidArgument as e:
        raise InvalidMsgData(str(e))


def proving_key_from_bytes(buf, off=0, ctx=None):
    """With `ctx` the query vectors (A, B, H, L: one compressed point per variable / constraint) are decompressed on the
    device and returned as [n, 2, limbs] device tensors - the form groth16.ProvingKey takes; without it as lists of
    Python-integer points."""
    out = {}
    o = off
    for name, rd, n in (("alpha_g1", g1_from_bytes, G1_BYTES), ("beta_g1", g1_from_bytes, G1_BYTES),
                        ("beta_g2", g2_from_bytes, G2_BYTES), ("delta_g1", g1_from_bytes, G1_BYTES),
                        ("delta_g2", g2_from_bytes, G2_BYTES)):
        out[name] = rd(buf, o)
        o += n

    def g1_list(o):
        if ctx is None:
            return _g1_list_from_bytes(buf, o)
        n = size_t_from_bytes(buf, o)
        return _device_points(ctx, "bls12_381_g1", buf, o + SIZE_T_BYTES, n), SIZE_T_BYTES + n * G1_BYTES

    out["A_query"], used = g1_list(o)
    o += used
    total_b = size_t_from_bytes(buf, o)
    o += SIZE_T_BYTES
    _need(buf, o, total_b)
    if ctx is None:
        out["B_indices"], out["B_g2"], out["B_g1"], out["B_domain_size"], used = kc_vector_from_bytes(buf, o)
    else:
        n = size_t_from_bytes(buf, o)
        used = _kc_vector_bytes(n)
        _need(buf, o, used)
        idx = np.frombuffer(buf, dtype=">u4", count=n, offset=o + SIZE_T_BYTES).astype(np.int64)
        po = o + SIZE_T_BYTES * (1 + n)
        pair = G2_BYTES + G1_BYTES
        out["B_indices"] = idx
        out["B_g2"] = _device_points(ctx, "bls12_381_g2", buf, po, n, stride=pair)
        out["B_g1"] = _device_points(ctx, "bls12_381_g1", buf, po + G2_BYTES, n, stride=pair)
        out["B_domain_size"] = size_t_from_bytes(buf, po + pair * n)
    if used != total_b:
        raise InvalidMsgData("B_query size field disagrees with its contents")
    o += total_b
    out["H_query"], used = g1_list(o)
    o += used
    out["L_query"], used = g1_list(o)
    o += used
    out["num_inputs"], out["num_aux"], out["constraints"], _ = r1cs_constraint_system_from_bytes(buf, o)
    return out


def verifier_input_to_bytes(vk, pi, proof):
    """`verifier_input_serializer_tvm::process(vk, pi, proof)`: proof | primary input | verification key"""
    return proof_to_bytes(proof) + primary_input_to_bytes(pi) + verification_key_to_bytes(vk)


def verifier_input_from_bytes(buf):
    """`verifier_input_deserializer_tvm::verifier_input_process` -> (vk, pi, proof)"""
    proof = proof_from_bytes(buf, 0)
    pi = primary_input_from_bytes(buf, PROOF_BYTES)
    vk = verification_key_from_bytes(buf, PROOF_BYTES + SIZE_T_BYTES + FR_BYTES * len(pi))
    return vk, pi, proof
