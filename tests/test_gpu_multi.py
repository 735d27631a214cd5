"""Multi-GPU paths on the GPU box (skipped when fewer than 2 GPUs): point-sharded MSM over NCCL."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from crypto3_zk_b200 import Context
    from crypto3_zk_b200.sharding import msm_sharded, shard_range
    from oracle import curves, fields
    C = curves.BLS12_381_G1
    n = 600
    pts = C.random_points(n, 21)
    sc = fields.random_elements(C.scalar_field, n, 22)
    off, cnt = shard_range(n, rank, world)
    ctx = Context(rank)
    pa = fields.ints_to_u32_array([c for P in pts[off:off + cnt] for c in P], 12).reshape(cnt, 2, 12)
    bases = ctx.msm_bases(C.name, pa)
    got = msm_sharded(ctx, bases, fields.ints_to_u32_array(sc[off:off + cnt], 8), device=torch.device("cuda", rank))
    q.put((rank, got == C.msm_bdlo12(pts, sc)))
    dist.barrier()
    dist.destroy_process_group()


def test_point_sharded_msm_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
