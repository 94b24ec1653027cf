"""Multi-GPU plumbing (torch.distributed; NCCL over NVLink on the box, gloo in CPU tests).

Inference: independent measurements are split into contiguous index ranges, one per rank, with
no data-path collective — only scalars (timings, PSNRs) are reduced.  Training: the one exchange
step of the path is the gradient average between backward() and the optimizer step; the whole
gradient (486,080 floats for FFDNet, 1.9 MB) goes through ONE flat all-reduce, which is
latency-bound on NVLink and needs no bucketing."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total, rank, world_size):
    """Contiguous [lo, hi) of `n_total` independent measurements for `rank`; sizes differ by <= 1."""
    base, extra = divmod(n_total, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_mean_gradients(params):
    """Averages .grad of `params` over all ranks with a single flat all-reduce (no-op at world 1)."""
    rank, ws = world()
    grads = [p.grad for p in params if p.grad is not None]
    if ws == 1 or not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= ws
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel()


def max_over_ranks(value, device="cpu"):
    """max of a python float over ranks (timings are reported as the slowest rank's)."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_floats(values, device="cpu"):
    """All ranks' lists of floats concatenated in rank order (e.g. per-measurement PSNRs)."""
    rank, ws = world()
    if ws == 1:
        return list(values)
    out = [None] * ws
    dist.all_gather_object(out, list(values))
    return [v for part in out for v in part]
