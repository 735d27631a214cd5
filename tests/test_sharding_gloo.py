"""N>1 host path on CPU: world_size-2 gloo processes shard an MSM by point range, all-gather the XYZZ
partials and combine them with the product's zkb_msm_combine.  (Per-rank partials come from the oracle here
because there is no GPU in this container; on the GPU box tests/test_gpu_multi.py uses the kernels.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import curves, fields


def test_shard_range():
    from crypto3_zk_b200.sharding import shard_range
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            got = [shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (o1, c1), (o2, _) in zip(got, got[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from crypto3_zk_b200.sharding import allgather_combine, shard_range
    C = curves.BN254_G1
    n, cl = 37, C.coord_limbs32
    pts = C.random_points(n, 5)
    sc = fields.random_elements(C.scalar_field, n, 6)
    off, cnt = shard_range(n, rank, world)
    part = C.msm_naive(pts[off:off + cnt], sc[off:off + cnt])
    R = pow(2, 32 * cl, C.base_field.p)
    if part is None:
        limbs = [0] * (4 * cl)
    else:
        limbs = (fields.to_limbs32(part[0] * R % C.base_field.p, cl) + fields.to_limbs32(part[1] * R % C.base_field.p, cl)
                 + fields.to_limbs32(R % C.base_field.p, cl) * 2)
    got = allgather_combine(C.name, np.array(limbs, dtype=np.uint32))
    q.put((rank, got == C.msm_naive(pts, sc)))
    dist.barrier()
    dist.destroy_process_group()


def test_point_sharded_msm_combine_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
