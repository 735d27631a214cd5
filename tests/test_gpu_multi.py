"""Multi-GPU paths on the GPU box (each world size is skipped when the box has fewer GPUs): point-sharded MSM and
polynomial-sharded LPC commit over NCCL at 2, 4 and 8 ranks.  `gpurun --gpus N -- python -m pytest tests/test_gpu_multi.py`
runs them; the logs of those runs are committed under profiles/ (r2_multi_*.log)."""
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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_point_sharded_msm_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


def _lpc_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from crypto3_zk_b200 import Context
    from crypto3_zk_b200.sharding import lpc_commit_sharded
    ctx = Context(rank)
    ok = True
    for hid, fri_step, log_in, log_out, per_rank in ((0, 1, 10, 13, 3), (1, 2, 8, 10, 1), (0, 1, 14, 17, 4)):
        g = torch.Generator(device="cpu").manual_seed(5 + log_in)
        allp = torch.randint(-2**31, 2**31 - 1, (per_rank * world, 1 << log_in, 8), dtype=torch.int32, generator=g)
        allp[..., 7] &= 0x0FFFFFFF
        mine = allp[rank * per_rank:(rank + 1) * per_rank].cuda()
        got = lpc_commit_sharded(ctx, "pallas_fq", hid, mine, log_in, log_out, fri_step)
        want = ctx.lpc_commit("pallas_fq", hid, allp.cuda(), log_in, log_out, fri_step)   # whole batch on one GPU
        ok = ok and got == want
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_polynomial_sharded_lpc_commit_nccl(world):
    """LDE sharded by polynomial, one all-to-all regroup by leaf range, per-rank subtrees, top levels
    from the all-gathered subtree roots - same root as the single-GPU commit of the whole batch."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_lpc_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_single_process_multi_gpu_abi(world):
    """zkb_msm_multi / zkb_lpc_commit_multi (include/zkb200.h, multi-GPU section): one process, a device
    list, host threads and peer copies inside the library - same MSM result as the oracle and the same LPC root as the
    single-GPU commit (and as the CPU port) on identical inputs."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from crypto3_zk_b200 import Context
    from crypto3_zk_b200.api import MultiContext
    from oracle import cref, curves, fields
    mc = MultiContext(list(range(world)))
    C = curves.BLS12_381_G1
    n = 1000
    pts = C.random_points(n, 31)
    sc = fields.random_elements(C.scalar_field, n, 32)
    sc[3], sc[4] = 0, 1
    pa = fields.ints_to_u32_array([c for P in pts for c in P], 12).reshape(n, 2, 12)
    sa = fields.ints_to_u32_array(sc, 8)
    want = C.msm_bdlo12(pts, sc)
    for pre in (False, True):
        b = mc.msm_bases(C.name, pa, precompute=pre)
        assert mc.multiexp(b, sa) == want
        assert mc.multiexp(b, sa[:777]) == C.msm_bdlo12(pts[:777], sc[:777])     # a prefix: later devices get short or empty slices
        b.free()
    ctx = Context(0)
    for hid, fri_step, log_in, log_out, batch in ((0, 1, 10, 13, 8), (1, 2, 9, 12, 8), (2, 1, 12, 15, 16)):
        rng = np.random.Generator(np.random.PCG64(100 + log_in))
        a = rng.integers(0, 1 << 32, size=(batch, 1 << log_in, 8), dtype=np.uint64).astype(np.uint32)
        a[..., 7] &= 0x0FFFFFFF
        got = mc.lpc_commit("pallas_fq", hid, a, log_in, log_out, fri_step)
        assert got == ctx.lpc_commit("pallas_fq", hid, a, log_in, log_out, fri_step)
        assert got == cref.lpc_commit(3, hid, a, log_in, log_out, fri_step, threads=cref.host_cores())[0]
    ctx.close()
    mc.close()
