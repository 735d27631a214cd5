#!/usr/bin/env python3
"""Wall-clock of successive lpc_commit calls with different batch sizes (diagnostic for the Placeholder flow)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crypto3_zk_b200 import Context
from profiles.prof_run import rand

ctx = Context(0)
for log_out in (24, 23):
    for it in range(3):
        for cnt in (65, 31, 1, 4):
            x = rand((cnt, 1 << 20, 8), cnt)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            tree = ctx.lpc_commit("pallas_fp", 0, x, 20, log_out, 1, keep_tree=True)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            root = tree.root()
            t2 = time.perf_counter()
            del tree
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            print("log_out %d it %d polys %2d: commit %.1f ms root %.1f ms free %.1f ms" % (log_out, it, cnt, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
