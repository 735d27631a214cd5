"""Multi-GPU partitioning of the two hot paths (one process per GPU, torch.distributed for plumbing).

* batched NTT / LDE: polynomials are independent -> contiguous slices per rank, no collective;
* MSM: contiguous point ranges per rank; the only exchange is an all-gather of one XYZZ partial sum per
  rank (4 * coord_limbs uint32 words, <= 192 bytes), added on the host by zkb_msm_combine.
"""
import numpy as np

from .api import msm_combine
from .fields import CURVE_BY_NAME, coord_limbs


def shard_range(n, rank, world):
    """Contiguous [offset, offset+count) of n items owned by `rank` (sizes differ by at most one)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def allgather_combine(curve, partial, group=None, device=None):
    """All-gathers the per-rank XYZZ partial sums and returns the combined affine point on every rank.
    Works with any backend: NCCL needs `device` (the rank's GPU), gloo uses CPU tensors."""
    import torch
    import torch.distributed as dist
    c = CURVE_BY_NAME[curve] if isinstance(curve, str) else curve
    words = 4 * coord_limbs(c)
    p = np.ascontiguousarray(partial, dtype=np.uint32).reshape(words)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return msm_combine(c, p.reshape(1, words))
    t = torch.from_numpy(p.view(np.int32).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    parts = np.stack([o.cpu().numpy().view(np.uint32) for o in out])
    return msm_combine(c, parts)


def msm_sharded(ctx, bases_local, scalars_local, group=None, device=None, stream=None):
    """Point-sharded MSM: `bases_local` / `scalars_local` are this rank's slice.  Returns the full result
    (affine ints or None) on every rank."""
    partial = ctx.multiexp_partial(bases_local, scalars_local, stream=stream)
    return allgather_combine(bases_local.curve, partial, group=group, device=device)
