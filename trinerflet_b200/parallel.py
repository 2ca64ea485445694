"""Ray-sharded data parallelism over the GPUs of one box.  The reference has no working
equivalent (its DDP hooks are dead code, nerf/utils.py:412-414): every rank holds the full wavelet coefficients
and MLP weights, takes a contiguous shard of the step's rays, and the coefficient + MLP gradients are summed
with NCCL over NVLink before the (replicated, identical) optimizer step.  Full-frame rendering shards the
pixels into contiguous ray tiles per rank; the only collective is the final gather of image/depth/weights_sum.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world_size):
    """Contiguous shard [lo, hi) of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dense_view(t):
    """A contiguous view over the same memory (channels-last parameters are permuted-contiguous)."""
    if t.is_contiguous():
        return t
    if t.dim() == 4 and t.permute(0, 2, 3, 1).is_contiguous():
        return t.permute(0, 2, 3, 1)
    if t.dim() == 5 and t.permute(0, 2, 3, 4, 1).is_contiguous():
        return t.permute(0, 2, 3, 4, 1)
    raise RuntimeError("gradient tensor is neither contiguous nor channels-last")


def allreduce_gradients(model, world_size, average=True):
    """Sum (or average) every parameter gradient across ranks, in place.  Each rank's loss is the mean over its
    own shard, so the average over ranks is the gradient of the global-batch mean."""
    if world_size <= 1 or not dist.is_initialized():
        return
    small = []
    for p in model.parameters():
        if p.grad is None:
            continue
        g = _dense_view(p.grad)
        if g.numel() >= (1 << 20):
            dist.all_reduce(g, op=dist.ReduceOp.SUM)       # large coefficient grads: one NCCL call each, in place
            if average:
                g.div_(world_size)
        else:
            small.append(g)
    if small:                                               # MLP weights + coarse levels: one fused bucket
        flat = torch.cat([g.reshape(-1) for g in small])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat.div_(world_size)
        off = 0
        for g in small:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n


def gather_frame(local, n_total, rank, world_size):
    """Final gather of a ray-tile-sharded render: `local` [n_local, ...] -> [n_total, ...] on every rank."""
    if world_size <= 1 or not dist.is_initialized():
        return local
    sizes = [shard_range(n_total, r, world_size) for r in range(world_size)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(out, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def render_frame_sharded(model, rays_o, rays_d, rank, world_size, **render_kwargs):
    """Full-frame inference with contiguous ray tiles per rank (renderer.py:549-576 path, no collective until the end)."""
    n = rays_o.shape[0]
    lo, hi = shard_range(n, rank, world_size)
    out = model.render(rays_o[lo:hi].unsqueeze(0), rays_d[lo:hi].unsqueeze(0), staged=True, perturb=False, **render_kwargs)
    image = gather_frame(out['image'].reshape(-1, 3), n, rank, world_size)
    depth = gather_frame(out['depth'].reshape(-1), n, rank, world_size)
    ws = gather_frame(out['weights_sum'].reshape(-1), n, rank, world_size)
    return {'image': image, 'depth': depth, 'weights_sum': ws}
