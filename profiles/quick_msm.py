"""MSM timing sweep on one GPU (device-timed, bases resident): python profiles/quick_msm.py [log sizes...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from crypto3_zk_b200 import Context
ctx = Context(0)
dev = torch.device("cuda", 0)
out = {"imad_wide_per_s": ctx.bench_imad_wide(148 * 8, 256, 4096), "fq12_mul_per_s": ctx.bench_field_mul("bls12_381_fq", 148 * 8, 256, 1024)}
for log_m in [int(a) for a in sys.argv[1:]] or [16, 18, 20]:
    nm = 1 << log_m
    pts = bench.msm_points(torch, ctx, np, log_m)
    bases = ctx.msm_bases("bls12_381_g1", pts)
    sc = bench.rand_elems(torch, (nm, 8), 13, dev)
    l0 = ctx.kernel_launches()
    r0 = ctx.multiexp(bases, sc)
    e = {"launches": ctx.kernel_launches() - l0, "ms": bench.time_cuda(torch, lambda: ctx.multiexp(bases, sc), 10, warmup=3)}
    bases.precompute(max(8, min(22, log_m)), 32 << 30)
    r1 = ctx.multiexp(bases, sc)
    e["ms_table"] = bench.time_cuda(torch, lambda: ctx.multiexp(bases, sc), 10, warmup=3)
    e["same"] = r0 == r1
    out["2p%d" % log_m] = e
    bases.free()
print(json.dumps(out, indent=1))
