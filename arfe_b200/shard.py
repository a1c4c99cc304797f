"""Data-parallel sharding of the region-aware feature path.

The path partitions by image: every op is per image or per RoI with an image
index in ``rois[:, 0]`` (mmdet/core/bbox/transforms.py:41-60) and nothing crosses
images, so ranks own disjoint image subsets and the RoIs of those images; there
is no collective on the data path.  torch.distributed (NCCL on GPUs, gloo in the
CPU tests) is used only for timing reductions and the final result gather -- the
counterpart of the reference's ``collect_results_gpu`` (mmdet/apis/test.py:179-197).
"""
import torch
import torch.distributed as dist


def shard_images(num_images, rank, world):
    """Contiguous block of image indices owned by `rank` (sizes differ by <= 1)."""
    base, extra = divmod(num_images, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def shard_rois(rois, images):
    """RoIs of the given (global) images with the batch index re-based to the
    local position of the image; order within an image is preserved."""
    if rois.numel() == 0 or not images:
        return rois.new_zeros((0, 5)), torch.zeros(0, dtype=torch.long)
    b = rois[:, 0].long()
    lut = torch.full((int(b.max().item()) + 1 if b.numel() else 1,), -1, dtype=torch.long)
    for local, g in enumerate(images):
        if g < lut.numel():
            lut[g] = local
    local_idx = lut[b]
    keep = (local_idx >= 0).nonzero().flatten()
    out = rois[keep].clone()
    out[:, 0] = local_idx[keep].to(out.dtype)
    return out, keep


def max_over_ranks(value, device=None):
    """Whole-job time of a step = the slowest rank's."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_variable(t, device=None):
    """all_gather of per-rank tensors with different leading sizes (padded to
    the maximum, like mmdet/apis/test.py:183-197). Returns a list per rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [t]
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.long, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s.item()) for s in sizes))
    pad = t.new_zeros((m,) + tuple(t.shape[1:]))
    pad[:t.shape[0]] = t
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return [b[:int(s.item())] for b, s in zip(bufs, sizes)]
