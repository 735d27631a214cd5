#!/usr/bin/env python3
"""Where the wall-clock of LpcCommitmentScheme.commit goes in the Placeholder flow (diagnostic)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crypto3_zk_b200 import Context, workloads as W
from crypto3_zk_b200.lpc import FriParams, LpcCommitmentScheme
from crypto3_zk_b200.transcript import FiatShamirSequential
from profiles.prof_run import rand

ctx = Context(0)
sizes = W.placeholder_batches()
cols = {k: rand((cnt, 1 << 20, 8), 300 + k) for k, cnt in sizes.items()}
orig_stack = LpcCommitmentScheme._batch_tensor
def timed_stack(self, index):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig_stack(self, index)
    torch.cuda.synchronize(); print("    stack %d: %.1f ms" % (index, (time.perf_counter() - t0) * 1e3), flush=True)
    return r
LpcCommitmentScheme._batch_tensor = timed_stack
orig_commit = ctx.lpc_commit
def timed_commit(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig_commit(*a, **k)
    torch.cuda.synchronize(); print("    zkb_lpc_commit: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
    return r
ctx.lpc_commit = timed_commit
for expand, hid in ((4, 2), (3, 0)):
    torch.cuda.empty_cache()
    fri = FriParams.with_max_step_one(20, 40, expand)
    for it in range(2):
        tr = FiatShamirSequential(0 if hid != 2 else 2, b"placeholder")
        scheme = LpcCommitmentScheme(ctx, "pallas_fp", hid, fri)
        for k in sizes:
            scheme.append_to_batch(k, cols[k])
        for k in sizes:
            torch.cuda.synchronize(); t0 = time.perf_counter()
            root = scheme.commit(k)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            tr(root)
            t2 = time.perf_counter()
            print("expand %d it %d batch %d (%d polys): commit %.1f ms, transcript %.1f ms" % (expand, it, k, sizes[k], (t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
        t0 = time.perf_counter()
        del scheme
        torch.cuda.synchronize()
        print("  del scheme %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
