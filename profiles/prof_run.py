#!/usr/bin/env python3
"""Workloads for ncu (run under gpurun; see profiles/README.md for the exact commands).
Only the region between cudaProfilerStart/Stop is captured (ncu --profile-from-start off)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crypto3_zk_b200 import Context  # noqa: E402
from crypto3_zk_b200.fields import CURVE_BY_NAME  # noqa: E402


def rand(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randint(-2**31, 2**31 - 1, shape, dtype=torch.int32, device="cuda", generator=g)
    x[..., 7] &= 0x0FFFFFFF
    return x


def main():
    mode = sys.argv[1]
    ctx = Context(0)
    if mode == "lde":          # BASELINE config #2, reduced to 16 polynomials (one scratch chunk)
        x = rand((16, 1 << 20, 8), 1)
        y = torch.empty((16, 1 << 23, 8), dtype=torch.int32, device="cuda")
        fn = lambda: ctx.lde("pallas_fq", x, 20, 23, out=y)
    elif mode == "ntt24":
        x = rand((1, 1 << 24, 8), 2)
        fn = lambda: ctx.ntt("bls12_381_fr", x, 24, coset_shift=7)
    elif mode == "lpc":
        x = rand((16, 1 << 20, 8), 1)
        fn = lambda: ctx.lpc_commit("pallas_fq", int(sys.argv[2]) if len(sys.argv) > 2 else 0, x, 20, 23, 1)
    elif mode == "evalpm":     # query-phase openings: 31 polynomials of 2^20 coefficients at 40 (z, -z) pairs
        x = rand((31, 1 << 20, 8), 5)
        pts = [pow(7, 3 * i + 1, (1 << 61) - 1) for i in range(40)]
        fn = lambda: ctx.poly_evaluate_pm("pallas_fp", x, 1 << 20, pts)
    elif mode == "grind":
        st = bytes(range(32))
        fn = lambda: ctx.pow_grind(0, st, (1 << 24) - 1)
    elif mode == "msm":
        log_m = int(sys.argv[2]) if len(sys.argv) > 2 else 20
        C = CURVE_BY_NAME["bls12_381_g1"]
        gen = np.array([[(C.gen_x >> (32 * i)) & 0xFFFFFFFF for i in range(12)],
                        [(C.gen_y >> (32 * i)) & 0xFFFFFFFF for i in range(12)]], dtype=np.uint32).reshape(1, 2, 12)
        gb = ctx.msm_bases("bls12_381_g1", gen)
        nm, m = 1 << log_m, 1024
        nbt = (nm + m - 1) // m
        rng = np.random.Generator(np.random.PCG64(7))
        ks = rng.integers(0, 1 << 32, size=(m + nbt, 8), dtype=np.uint64).astype(np.uint32)
        ks[:, 7] &= 0x0FFFFFFF
        tabs = np.zeros((m + nbt, 2, 12), dtype=np.uint32)
        for i in range(m + nbt):
            pt = ctx.multiexp(gb, ks[i:i + 1])
            tabs[i, 0] = [(pt[0] >> (32 * k)) & 0xFFFFFFFF for k in range(12)]
            tabs[i, 1] = [(pt[1] >> (32 * k)) & 0xFFFFFFFF for k in range(12)]
        pts = ctx.grid_points("bls12_381_g1", nm, tabs[:m], tabs[m:])
        bases = ctx.msm_bases("bls12_381_g1", pts)
        if len(sys.argv) > 3:          # window table with this many bits (0 = automatic)
            bases.precompute(int(sys.argv[3]), 64 << 30)
        sc = rand((nm, 8), 13)
        fn = lambda: ctx.multiexp(bases, sc)
    else:
        raise SystemExit("mode?")
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ctx.close()


if __name__ == "__main__":
    main()
