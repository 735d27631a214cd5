#!/usr/bin/env python3
"""One pass over the kernels of the commitment-scheme side for ncu (see run_r1h.sh): LPC scheme with the complete proof_eval
(commit phase, grinding, query phase) at 2^16 rows, coset NTT forward / inverse, FRI fold, pointwise op, SHA-256 commit.
(The MSM kernels have their own captures: r1c_full_msm_*.txt, r1f_full_msm_acc.txt.)  Region = one call of each after a
warm-up call."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crypto3_zk_b200 import Context  # noqa: E402
from crypto3_zk_b200.lpc import FriParams, LpcCommitmentScheme  # noqa: E402
from crypto3_zk_b200.transcript import FiatShamirSequential  # noqa: E402
from profiles.prof_run import rand  # noqa: E402


def main():
    ctx = Context(0)
    rows_log = 16
    n = 1 << rows_log
    cols = {0: rand((6, n, 8), 1), 1: rand((5, n, 8), 2), 2: rand((1, n, 8), 3)}

    def lpc_flow():
        fri = FriParams.with_max_step_one(rows_log, 8, 3, use_grinding=True, grinding_parameter=0xFFFF)
        scheme = LpcCommitmentScheme(ctx, "pallas_fp", 0, fri)
        tr = FiatShamirSequential(0, b"prof")
        for k in cols:
            scheme.append_to_batch(k, cols[k])
        for k in cols:
            tr(scheme.commit(k))
        for k in cols:
            scheme.append_eval_point(k, 12345)
        scheme.append_eval_point(1, 67890)
        scheme.proof_eval(tr, query=True)

    x20 = rand((1, 1 << 20, 8), 4)
    f = rand((1 << 20, 8), 6)
    a, b = rand((1 << 20, 8), 7), rand((1 << 20, 8), 8)

    def everything():
        lpc_flow()
        ctx.ntt("bls12_381_fr", x20, 20, coset_shift=7)
        ctx.ntt("bls12_381_fr", x20, 20, inverse=True, coset_shift=7)
        ctx.fri_fold("pallas_fq", f, 20, 99)
        ctx.vec("bn254_fr", 2, a, b)
        ctx.lpc_commit("pallas_fq", 1, cols[1], rows_log, rows_log + 3, 2)

    everything()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    everything()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ctx.close()


if __name__ == "__main__":
    main()
