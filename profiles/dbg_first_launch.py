"""first-call cost of the point decompression kernels (module load / local-memory set-up), then steady state"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from crypto3_zk_b200 import Context, marshalling as m
from oracle import curves
ctx = Context(0)
torch.cuda.synchronize()
g1 = m.g1_to_bytes(curves.BLS12_381_G1.gen) * 4
g2 = m.g2_to_bytes(curves.BLS12_381_G2.gen) * 4
for name, curve, blob in (("g1", "bls12_381_g1", g1), ("g2", "bls12_381_g2", g2), ("g1", "bls12_381_g1", g1), ("g2", "bls12_381_g2", g2)):
    t0 = time.perf_counter()
    ctx.points_decompress(curve, blob, 4)
    print(name, "ms", round((time.perf_counter() - t0) * 1e3, 2))
