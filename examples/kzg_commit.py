#!/usr/bin/env python3
"""KZG commitment of a random degree-(2^16 - 1) polynomial over BLS12-381 (BASELINE configs[0]; the reference's
test/commitment/kzg.cpp:75-101 at full size): the SRS [alpha^i] G is built on the GPU with the fixed-base batch
exponentiation, the commitment is one multiexp, and it must equal f(alpha) * G."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crypto3_zk_b200 import Context  # noqa: E402
from crypto3_zk_b200.api import _affine_from_limbs, _int_rows  # noqa: E402
from crypto3_zk_b200.fields import CURVE_BY_NAME, FIELD_BY_NAME  # noqa: E402


def main(log_n=16):
    ctx = Context(0)
    C, F = CURVE_BY_NAME["bls12_381_g1"], FIELD_BY_NAME["bls12_381_fr"]
    r, n, alpha = F.p, 1 << log_n, 7
    rng = np.random.Generator(np.random.PCG64(1))
    coeffs = [int.from_bytes(rng.bytes(32), "little") % r for _ in range(n)]
    powers, acc = [], 1
    for _ in range(n):
        powers.append(acc)
        acc = acc * alpha % r
    t0 = time.perf_counter()
    srs = ctx.batch_exp("bls12_381_g1", (C.gen_x, C.gen_y), np.ascontiguousarray(_int_rows(powers)))   # [alpha^i] G, affine
    t1 = time.perf_counter()
    key = ctx.msm_bases("bls12_381_g1", srs)                      # resident commitment key
    t2 = time.perf_counter()
    commitment = ctx.multiexp(key, np.ascontiguousarray(_int_rows(coeffs)))
    t3 = time.perf_counter()
    f_alpha = sum(c * p for c, p in zip(coeffs, powers)) % r
    want = ctx.batch_exp("bls12_381_g1", (C.gen_x, C.gen_y), np.ascontiguousarray(_int_rows([f_alpha])))
    want = _affine_from_limbs(want[0].reshape(-1), 12)
    assert commitment == want, "commit != f(alpha) * G"
    print("KZG commit of 2^%d coefficients: SRS %.1f ms, key upload %.1f ms, commit %.2f ms; commit == f(alpha) G" %
          (log_n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    ctx.close()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
