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


def _lpc_worker(rank, world, port, q):
    """Polynomial-sharded LPC commit on CPU tensors: the product's regroup (sharding.lpc_regroup_send) + a gloo
    all-to-all must hand every rank the evaluations of a sub-domain whose tree is the rank's subtree of the full
    tree (leaf pattern basic_fri.hpp:466-492); hashing is done by the oracle here (no GPU in this container)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from crypto3_zk_b200.sharding import lpc_regroup_send
    from oracle import fri, hashes, ntt
    F = fields.PALLAS_FQ
    ok = True
    for fri_step, n_in, n_out, per_rank in ((1, 8, 32, 2), (2, 4, 64, 1), (3, 8, 64, 3)):
        polys = [fields.random_elements(F, n_in, 100 + b) for b in range(per_rank * world)]
        mine = polys[rank * per_rank:(rank + 1) * per_rank]
        ext = np.stack([fields.ints_to_u32_array(ntt.dfs_resize(p, F, n_out), 8) for p in mine])
        send = torch.from_numpy(lpc_regroup_send(ext, world, fri_step).view(np.int32).copy())
        # gloo has no all_to_all: emulate it with one all_gather (same data movement semantics for the test)
        gathered = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(gathered, send)
        recv = torch.stack([g[rank] for g in gathered])            # [src, per_rank, n_out / world, 8]
        sub = recv.numpy().view(np.uint32).reshape(world * per_rank, n_out // world, 8)
        sub_polys = [fields.u32_array_to_ints(sub[b]) for b in range(world * per_rank)]
        levels, _ = fri.precommit(sub_polys, F, n_out // world, fri_step, hashes.keccak256)
        root = torch.frombuffer(bytearray(levels[-1][0]), dtype=torch.uint8)
        roots = [torch.empty_like(root) for _ in range(world)]
        dist.all_gather(roots, root)
        level = [bytes(r.numpy().tobytes()) for r in roots]
        while len(level) > 1:
            level = [hashes.keccak256(level[2 * i] + level[2 * i + 1]) for i in range(len(level) // 2)]
        want, _ = fri.precommit(polys, F, n_out, fri_step, hashes.keccak256)
        ok = ok and level[0] == want[-1][0]
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_lpc_commit_regroup_by_leaf_range(world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_lpc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
