"""Host-side Fiat-Shamir transcript that drives the device commit phase when the caller has none of its own.

Mirrors transcript::fiat_shamir_heuristic_sequential<Hash> (zk/transcript/fiat_shamir.hpp:131-188):
state = H(init); operator()(data): state = H(state || data); challenge<Field>(): state = H(state), the digest
read as a big-endian integer reduced into the field.  Pinned by the reference's known answers
(test/transcript/transcript.cpp:50-64) in tests/test_transcript.py.  The messages are 32-64 bytes per FRI round -
control plane, not a data path: the data-parallel hashing (leaves, tree nodes) runs in zkb_hash.cu.
"""
import hashlib

_MASK = (1 << 64) - 1
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]


def _rotations():
    rot = [[0] * 5 for _ in range(5)]
    x, y = 1, 0
    for t in range(24):
        rot[x][y] = ((t + 1) * (t + 2) // 2) % 64
        x, y = y, (2 * x + 3 * y) % 5
    return rot


_ROT = _rotations()


def _permute(s):
    """Keccak-f[1600] on 25 lanes, s[x + 5 y]."""
    for rc in _RC:
        c = [s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20] for x in range(5)]
        d = [c[(x - 1) % 5] ^ (((c[(x + 1) % 5] << 1) | (c[(x + 1) % 5] >> 63)) & _MASK) for x in range(5)]
        s = [s[i] ^ d[i % 5] for i in range(25)]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                v, r = s[x + 5 * y], _ROT[x][y]
                b[y + 5 * ((2 * x + 3 * y) % 5)] = ((v << r) | (v >> (64 - r))) & _MASK if r else v
        s = [b[i] ^ ((~b[(i % 5 + 1) % 5 + 5 * (i // 5)]) & b[(i % 5 + 2) % 5 + 5 * (i // 5)] & _MASK) for i in range(25)]
        s[0] ^= rc
    return s


def _keccak(data, digest_bytes):
    """Original Keccak (pad10*1 with the 0x01 domain byte), capacity = 2 x digest, as hashes::keccak_1600<bits>."""
    rate = 200 - 2 * digest_bytes
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0x00)
    msg[-1] |= 0x80
    s = [0] * 25
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            s[i] ^= int.from_bytes(msg[off + 8 * i:off + 8 * i + 8], "little")
        s = _permute(s)
    return b"".join(v.to_bytes(8, "little") for v in s)[:digest_bytes]


def keccak256(data):
    return _keccak(bytes(data), 32)


def keccak512(data):
    return _keccak(bytes(data), 64)


def sha256(data):
    return hashlib.sha256(bytes(data)).digest()


HASH_BY_ID = {0: keccak256, 1: sha256, 2: keccak512}


class FiatShamirSequential:
    def __init__(self, hash_id=0, init=b"\x00"):
        self.hash_id = hash_id
        self._h = HASH_BY_ID[hash_id]
        self.state = self._h(bytes(init))

    def __call__(self, data):
        self.state = self._h(self.state + bytes(data))

    absorb = __call__

    def challenge(self, modulus):
        self.state = self._h(self.state)
        return int.from_bytes(self.state, "big") % int(modulus)

    def int_challenge(self, bits=32):
        """int_challenge<Integral> (fiat_shamir.hpp:190-199): state = H(state); the low `bits` bits of the digest
        read as a big-endian integer."""
        self.state = self._h(self.state)
        return int.from_bytes(self.state, "big") & ((1 << bits) - 1)

    def challenges(self, modulus, count):
        return [self.challenge(modulus) for _ in range(count)]
