"""MSM tuning sweep (window bits x task cap) on one GPU: python profiles/quick_msm_tune.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from crypto3_zk_b200 import Context
ctx = Context(0)
dev = torch.device("cuda", 0)
out = {}
for log_m in [int(a) for a in sys.argv[1:]] or [16, 18, 20]:
    nm = 1 << log_m
    pts = bench.msm_points(torch, ctx, np, log_m)
    sc = bench.rand_elems(torch, (nm, 8), 13, dev)
    bases = ctx.msm_bases("bls12_381_g1", pts)
    ref = None
    for c in range(log_m - 5, log_m + 1):
        for cap in (0, 32, 256):
            os.environ["ZKB_MSM_C"] = str(c)
            if cap: os.environ["ZKB_MSM_CAP"] = str(cap)
            else: os.environ.pop("ZKB_MSM_CAP", None)
            r = ctx.multiexp(bases, sc)
            ref = ref or r
            assert r == ref
            out["2p%d plain c=%d cap=%d" % (log_m, c, cap)] = round(bench.time_cuda(torch, lambda: ctx.multiexp(bases, sc), 5, warmup=1), 3)
    bases.free()
    os.environ.pop("ZKB_MSM_C", None)
    for c in range(log_m - 1, min(23, log_m + 4)):
        bases = ctx.msm_bases("bls12_381_g1", pts)
        bases.precompute(c, 32 << 30)
        for cap in (0, 32, 256):
            if cap: os.environ["ZKB_MSM_CAP"] = str(cap)
            else: os.environ.pop("ZKB_MSM_CAP", None)
            assert ctx.multiexp(bases, sc) == ref
            out["2p%d table c=%d cap=%d" % (log_m, c, cap)] = round(bench.time_cuda(torch, lambda: ctx.multiexp(bases, sc), 5, warmup=1), 3)
        bases.free()
print(json.dumps(out, indent=1))
