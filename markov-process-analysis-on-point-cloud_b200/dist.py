"""Data-parallel plumbing for the one exchange step of the path: batches of clouds are sharded across ranks
(one process per GPU), nothing is exchanged in forward, and after backward the gradients are averaged with one
coalesced all-reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).  BatchNorm statistics stay per replica --
the reference is single-GPU, so a replica reproduces the reference on its shard (plain DDP semantics).
"""
import torch
import torch.distributed as dist


def shard_batch(batch_size, rank, world):
    """Contiguous batch shard [lo, hi) of rank `rank`; sizes differ by at most one cloud."""
    base, rem = divmod(batch_size, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_grads(params, world=None, group=None):
    """Average .grad over the ranks in one coalesced collective.  Parameters that received no gradient (the
    reference's constructed-but-unused sub-modules: normal_Trans, norm1, ...) are skipped -- on every rank alike,
    because which parameters are used does not depend on the data."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world <= 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    torch._foreach_div_(grads, float(world))
    if grads[0].is_cuda:
        with dist._coalescing_manager(group=group, device=grads[0].device, async_ops=False):
            for g in grads:
                dist.all_reduce(g, group=group)
    else:  # gloo: one flat buffer (coalescing is an NCCL feature)
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, group=group)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
    return len(grads)
