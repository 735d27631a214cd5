#!/usr/bin/env python3
"""Workload for ncu over the round-2 kernels: the resident-column Placeholder prover at 2^18 rows (expr_eval_kernel,
quotient_div_kernel, the scans), a lookup sort of 2^20 rows, and 2^18 G1 / 2^14 G2 point decompressions.
ncu --profile-from-start off -k regex:... python profiles/prof_new_kernels.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crypto3_zk_b200 import Context, placeholder as P, workloads as W
from crypto3_zk_b200.lpc import FriParams
from crypto3_zk_b200.transcript import FiatShamirSequential
from crypto3_zk_b200 import marshalling as m

ctx = Context(0)
log_n = 18
circuit, witness, public = W.placeholder_chain_circuit(ctx, "pallas_fp", log_n, triples=10, seed=5, max_quotient_chunks=4)
fri = FriParams.with_max_step_one(log_n, 8, 3)
# lookup sort input: a table of 2^19 distinct values, each twice, and three input columns drawn from it
n, usable = 1 << 20, (1 << 20) - 8
g = torch.Generator(device="cuda").manual_seed(1)
vals = torch.randint(1, 2**31 - 1, (1 << 19, 8), dtype=torch.int32, device="cuda", generator=g)
vals[:, 7] &= 0x0FFFFFFF
table = torch.zeros((1, n, 8), dtype=torch.int32, device="cuda")
table[0, 8:8 + (1 << 20) - 16] = vals.repeat_interleave(2, dim=0)[:(1 << 20) - 16]
idx = torch.randint(0, (1 << 19) - 8, (3, n), device="cuda", generator=g)
inputs = vals[idx]
# compressed points
pts = W.curve_grid_points(ctx, "bls12_381_g1", 1 << 18, seed=11)
a = pts.cpu().numpy().view(np.uint32)
blob = np.frombuffer(a[:, 0, ::-1].astype(">u4").tobytes(), dtype=np.uint8).reshape(-1, 48).copy()
blob[:, 0] |= 0x80
d1 = torch.from_numpy(blob.reshape(-1)).cuda()
from oracle import curves
g2 = curves.BLS12_381_G2.random_points(64, 4)
d2 = torch.from_numpy(np.frombuffer(b"".join(m.g2_to_bytes(p) for p in g2) * 256, dtype=np.uint8).copy()).cuda()

def run():
    if len(sys.argv) < 2 or sys.argv[1] != "keys":
        P.placeholder_prove(ctx, circuit, 0, fri, witness, public, FiatShamirSequential(0, b"prof"), query=False)
    ctx.lookup_sort("pallas_fp", inputs, table, usable)
    ctx.points_decompress("bls12_381_g1", d1, 1 << 18, status=True)
    ctx.points_decompress("bls12_381_g2", d2, 1 << 14)

run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ctx.close()
