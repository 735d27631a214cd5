"""one tabled (or plain) MSM per size for an ncu launch list: python profiles/msm_once.py LOG [table_c|0] ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from crypto3_zk_b200 import Context
ctx = Context(0)
dev = torch.device("cuda", 0)
args = [int(a) for a in sys.argv[1:]]
for log_m, c in zip(args[0::2], args[1::2]):
    nm = 1 << log_m
    pts = bench.msm_points(torch, ctx, np, log_m)
    sc = bench.rand_elems(torch, (nm, 8), 13, dev)
    bases = ctx.msm_bases("bls12_381_g1", pts)
    if c:
        bases.precompute(c, 32 << 30)
    ctx.multiexp(bases, sc)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("msm_%d_%d" % (log_m, c))
    ctx.multiexp(bases, sc)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    bases.free()
