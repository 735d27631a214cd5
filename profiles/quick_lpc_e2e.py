"""LPC commit of config #2 from device tensors and from pinned host buffers (zkb_lpc_commit mem = HOST): same root, times"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from crypto3_zk_b200 import Context
ctx = Context(0)
dev = torch.device("cuda", 0)
x = bench.rand_elems(torch, (64, 1 << 20, 8), 1, dev)
hx, hxn = bench.pinned_like(torch, np, x)
out = {}
for hid, name in ((0, "keccak256"), (1, "sha256")):
    r_dev = ctx.lpc_commit("pallas_fq", hid, x, 20, 23, 1)
    r_host = ctx.lpc_commit("pallas_fq", hid, hxn, 20, 23, 1)
    out[name] = {"same_root": r_dev == r_host, "device_ms": bench.time_cuda(torch, lambda: ctx.lpc_commit("pallas_fq", hid, x, 20, 23, 1), 3, warmup=1),
                 "host_buffers_ms": bench.time_wall(torch, lambda: ctx.lpc_commit("pallas_fq", hid, hxn, 20, 23, 1), 3, warmup=1)}
# ragged: 5 polynomials of 2^12 from a plain (unpinned) numpy array
a = np.random.Generator(np.random.PCG64(3)).integers(0, 1 << 32, size=(5, 1 << 12, 8), dtype=np.uint64).astype(np.uint32)
a[..., 7] &= 0x0FFFFFFF
out["small_unpinned_same_root"] = ctx.lpc_commit("pallas_fq", 0, a, 12, 15, 2) == ctx.lpc_commit("pallas_fq", 0, torch.from_numpy(a.view(np.int32)).to(dev), 12, 15, 2)
print(json.dumps(out, indent=1))
