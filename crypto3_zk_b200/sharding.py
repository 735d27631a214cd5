"""Multi-GPU partitioning of the two hot paths (one process per GPU, torch.distributed for plumbing).

* batched NTT / LDE: polynomials are independent -> contiguous slices per rank, no collective;
* MSM: contiguous point ranges per rank; the only exchange is an all-gather of one XYZZ partial sum per
  rank (4 * coord_limbs uint32 words, <= 192 bytes), added on the host by zkb_msm_combine;
* LPC commit: the LDE shards by polynomial, but a Merkle leaf holds ALL polynomials at one index coset, so the
  extended evaluations are regrouped by leaf range with one all-to-all; every rank then commits the subtree of
  its leaf range and the top log2(world) levels are hashed from the all-gathered subtree roots.
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


# ---------------------------------------------------------------------------------------- LPC commit
def lpc_regroup_send(ext, world, fri_step):
    """ext: [polys_local, N, 8] extended evaluations of this rank's polynomials.  Leaf x of the tree over a domain
    of size N holds, per polynomial, the 2^fri_step elements x + t * (N >> fri_step) (basic_fri.hpp:466-492), so
    rank g, which owns leaves [g L/G, (g+1) L/G) with L = N >> fri_step, needs 2^fri_step runs of L/G elements
    of every polynomial.  Returns the send buffer [world, polys_local, N / world, 8] (block g goes to rank g)."""
    pl, n, limbs = ext.shape
    t = 1 << fri_step
    lg = (n >> fri_step) // world
    if lg * world * t != n or lg == 0:
        raise ValueError("the leaf count must be a positive multiple of the world size")
    v = ext.reshape(pl, t, world, lg, limbs)
    perm = v.permute(2, 0, 1, 3, 4) if hasattr(v, "permute") else v.transpose(2, 0, 1, 3, 4)
    out = perm.contiguous() if hasattr(perm, "contiguous") else np.ascontiguousarray(perm)
    return out.reshape(world, pl, n // world, limbs)


def lpc_commit_sharded(ctx, field, hash_id, polys_local, log_n_in, log_n_out, fri_step=1, group=None, stream=None):
    """lpc::commit (precommit, basic_fri.hpp:445-496) of a batch whose polynomials are sharded over the ranks
    (equal count per rank, rank order = batch order).  polys_local: torch CUDA tensor [polys_local, 2^log_n_in, 8].
    Returns the commitment (root bytes), identical on every rank and identical to the single-GPU
    Context.lpc_commit of the concatenated batch."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ctx.lpc_commit(field, hash_id, polys_local, log_n_in, log_n_out, fri_step, stream=stream)
    if world & (world - 1) or (log_n_out - fri_step) < (world.bit_length() - 1):
        raise ValueError("world size must be a power of two not larger than the leaf count")
    ext = ctx.lde(field, polys_local, log_n_in, log_n_out, stream=stream)
    send = lpc_regroup_send(ext, world, fri_step)
    del ext
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    del send
    # recv[src] = the polynomials of rank src restricted to my leaf range: [all polys][N / world] - exactly the
    # evaluations of a domain of size N / world with the same leaf pattern
    log_sub = log_n_out - (world.bit_length() - 1)
    sub_root = ctx.merkle_commit(field, hash_id, recv.reshape(-1, 1 << log_sub, 8), log_sub, fri_step, stream=stream)
    db = len(sub_root)
    t = torch.frombuffer(bytearray(sub_root), dtype=torch.uint8).to(recv.device)
    roots = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(roots, t, group=group)
    return ctx.merkle_root_of_digests(hash_id, [bytes(r.cpu().numpy().tobytes()) for r in roots])
