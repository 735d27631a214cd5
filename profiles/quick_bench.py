#!/usr/bin/env python3
"""Quick A/B timings on one B200 (CUDA events, inputs resident): field-mul peaks, coset NTT 2^24,
LDE 16 x (2^20 -> 2^23), LPC commit 16 polys, G1 MSM 2^k.  `ZKB200_LIB=path python profiles/quick_bench.py [what..]`."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crypto3_zk_b200 import Context  # noqa: E402
from crypto3_zk_b200.fields import CURVE_BY_NAME, FIELD_BY_NAME  # noqa: E402
from profiles.prof_run import rand  # noqa: E402


def tcuda(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def msm_setup(ctx, curve, log_m):
    C = CURVE_BY_NAME[curve]
    cl = 12 if "bls" in curve else 8
    gen = np.array([[(C.gen_x >> (32 * i)) & 0xFFFFFFFF for i in range(cl)],
                    [(C.gen_y >> (32 * i)) & 0xFFFFFFFF for i in range(cl)]], dtype=np.uint32).reshape(1, 2, cl)
    gb = ctx.msm_bases(curve, gen)
    nm, m = 1 << log_m, 1024
    nbt = (nm + m - 1) // m
    rng = np.random.Generator(np.random.PCG64(7))
    ks = rng.integers(0, 1 << 32, size=(m + nbt, 8), dtype=np.uint64).astype(np.uint32)
    ks[:, 7] &= 0x0FFFFFFF
    tabs = np.zeros((m + nbt, 2, cl), dtype=np.uint32)
    for i in range(m + nbt):
        pt = ctx.multiexp(gb, ks[i:i + 1])
        tabs[i, 0] = [(pt[0] >> (32 * k)) & 0xFFFFFFFF for k in range(cl)]
        tabs[i, 1] = [(pt[1] >> (32 * k)) & 0xFFFFFFFF for k in range(cl)]
    pts = ctx.grid_points(curve, nm, tabs[:m], tabs[m:])
    return ctx.msm_bases(curve, pts), rand((nm, 8), 13), pts


def main():
    what = sys.argv[1:] or ["mul", "ntt24", "lde", "lpc", "msm20"]
    ctx = Context(0)
    out = {"lib": os.environ.get("ZKB200_LIB", "default")}
    for w in what:
        if w == "mul":
            for f in ("bls12_381_fr", "bn254_fr", "pallas_fq", "bls12_381_fq", "bn254_fq"):
                out["mul_" + f] = ctx.bench_field_mul(f, 148 * 8, 256, 1024)
        elif w.startswith("ntt"):
            lg = int(w[3:])
            x = rand((1, 1 << lg, 8), 2)
            for f in ("bls12_381_fr", "pallas_fq", "bn254_fr"):
                if lg > FIELD_BY_NAME[f].two_adicity:
                    continue
                out["%s_%s_coset_ms" % (w, f)] = tcuda(lambda: ctx.ntt(f, x, lg, coset_shift=7))
                out["%s_%s_plain_ms" % (w, f)] = tcuda(lambda: ctx.ntt(f, x, lg))
            del x
        elif w == "lde":
            x = rand((16, 1 << 20, 8), 1)
            y = torch.empty((16, 1 << 23, 8), dtype=torch.int32, device="cuda")
            out["lde16_pallas_fq_ms"] = tcuda(lambda: ctx.lde("pallas_fq", x, 20, 23, out=y), 3, 1)
            del x, y
        elif w == "lpc":
            x = rand((16, 1 << 20, 8), 1)
            out["lpc16_keccak_ms"] = tcuda(lambda: ctx.lpc_commit("pallas_fq", 0, x, 20, 23, 1), 3, 1)
            out["lpc16_sha256_ms"] = tcuda(lambda: ctx.lpc_commit("pallas_fq", 1, x, 20, 23, 1), 3, 1)
            del x
        elif w.startswith("msm"):
            lg = int(w[3:])
            curve = "bls12_381_g1"
            bases, sc, pts = msm_setup(ctx, curve, lg)
            out["%s_%s_ms" % (w, curve)] = tcuda(lambda: ctx.multiexp(bases, sc), 5, 2)
            for wb in [int(x) for x in os.environ.get("MSM_TABLE_BITS", "").split(",") if x]:
                bt = ctx.msm_bases(curve, pts).precompute(wb, 64 << 30)
                out["%s_%s_table%d_ms" % (w, curve, wb)] = tcuda(lambda: ctx.multiexp(bt, sc), 5, 2)
                bt.free()
            bases.free()
    ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
