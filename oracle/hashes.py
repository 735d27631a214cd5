"""Hash restatements for the Merkle/transcript side of the LPC path (oracle; test infra only).

crypto3-hash is un-vendored.  `hashes::keccak_1600<256>` is original Keccak (pad 0x01, as in
Ethereum), not NIST SHA-3 (pad 0x06); this is PINNED by the reference's transcript KAT
(test/transcript/transcript.cpp:50-64) in tests/test_oracle_golden.py.  sha2<256> = FIPS 180-4
(hashlib).  The permutation is cross-checked against hashlib.sha3_256 (same Keccak-f[1600]).
"""
import hashlib

_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M if n else x


def keccak_f1600(A):
    """A: 5x5 list of 64-bit lanes indexed A[x][y]."""
    for rnd in range(24):
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= _RC[rnd]
    return A


def _keccak(data: bytes, rate: int, outlen: int, pad: int) -> bytes:
    A = [[0] * 5 for _ in range(5)]
    msg = bytearray(data)
    msg.append(pad)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    for off in range(0, len(msg), rate):
        blk = msg[off:off + rate]
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(blk[8 * i:8 * i + 8], "little")
        A = keccak_f1600(A)
    out = b"".join(A[i % 5][i // 5].to_bytes(8, "little") for i in range(rate // 8))
    return out[:outlen]


def keccak256(data: bytes) -> bytes:
    return _keccak(data, 136, 32, 0x01)


def keccak512(data: bytes) -> bytes:
    return _keccak(data, 72, 64, 0x01)


def sha3_256_via_own_permutation(data: bytes) -> bytes:
    return _keccak(data, 136, 32, 0x06)


def sha256(data: bytes) -> bytes:
    return hashlib.sha256(data).digest()


HASHES = {"keccak256": (0, keccak256, 32), "sha256": (1, sha256, 32), "keccak512": (2, keccak512, 64)}


# --------------------------------------------------------------------------- Fiat-Shamir transcript
class FiatShamirSequential:
    """fiat_shamir_heuristic_sequential<Hash> (zk/transcript/fiat_shamir.hpp:131-188):
    state = H(init); absorb: state = H(state || data); challenge: state = H(state),
    value = big-endian integer of the digest reduced into the field."""

    def __init__(self, h, init=b"\x00"):
        self.h = h
        self.state = h(bytes(init))

    def absorb(self, data: bytes):
        self.state = self.h(self.state + bytes(data))

    def challenge(self, field):
        self.state = self.h(self.state)
        return int.from_bytes(self.state, "big") % field.p

    def int_challenge(self, bits=32):
        """int_challenge<Integral> (fiat_shamir.hpp:190-199): state = H(state); raw_result &= ~Integral(0)."""
        self.state = self.h(self.state)
        return int.from_bytes(self.state, "big") & ((1 << bits) - 1)

    def copy(self):
        t = FiatShamirSequential(self.h)
        t.state = self.state
        return t
