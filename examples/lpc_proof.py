#!/usr/bin/env python3
"""LPC / FRI commitment-scheme proof on the GPU (the reference's lpc_commitment_scheme, lpc.hpp:66-200): commit three
batches of columns, open them at two points, run the FRI commit phase, grinding and the query phase."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crypto3_zk_b200 import Context  # noqa: E402
from crypto3_zk_b200.fields import FIELD_BY_NAME, omega  # noqa: E402
from crypto3_zk_b200.lpc import FriParams, LpcCommitmentScheme  # noqa: E402
from crypto3_zk_b200.transcript import FiatShamirSequential  # noqa: E402


def main(rows_log=18, retain=True):
    ctx = Context(0)
    F = FIELD_BY_NAME["pallas_fp"]
    n = 1 << rows_log
    g = torch.Generator(device="cuda").manual_seed(1)
    cols = {}
    for k, cnt in {0: 8, 1: 6, 2: 2}.items():
        x = torch.randint(-2**31, 2**31 - 1, (cnt, n, 8), dtype=torch.int32, device="cuda", generator=g)
        x[..., 7] &= 0x0FFFFFFF                       # canonical: below the 255-bit modulus
        cols[k] = x
    fri = FriParams.with_max_step_one(rows_log, lambda_=20, expand_factor=3, use_grinding=True, grinding_parameter=0xFFFF)
    scheme = LpcCommitmentScheme(ctx, F.name, 0, fri, retain_lde=retain)      # hash 0 = keccak-256
    tr = FiatShamirSequential(0, b"example")
    t0 = time.perf_counter()
    for k in cols:
        scheme.append_to_batch(k, cols[k])
        tr(scheme.commit(k))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    y = tr.challenge(F.p)
    for k in cols:
        scheme.append_eval_point(k, y)
    scheme.append_eval_point(1, y * omega(F, rows_log) % F.p)
    proof = scheme.proof_eval(tr, query=True)["proof"]
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    fp = proof["fri_proof"]
    print("rows 2^%d, 16 columns, blow-up 8: commits %.1f ms, proof_eval %.1f ms (%d FRI rounds, %d queries, nonce %d)" %
          (rows_log, (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(fp["fri_roots"]), len(fp["query_proofs"]), fp["proof_of_work"]))
    ctx.close()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 18)
