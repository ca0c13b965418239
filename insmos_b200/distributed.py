"""Multi-GPU partition of the path: samples of a batch are independent, so they are
sharded one per rank (weights replicated) and the only exchange step is ONE gather of the per-point MOS logits
(variable length, ~1.44 MB per 120k-point scan) at the end.  No collective on the data path."""
import torch


def shard_samples(n_samples, rank, world):
    """indices of the samples rank `rank` processes (round robin, sample i -> GPU i mod world)."""
    return list(range(rank, n_samples, world))


def gather_logits(logits, world, group=None):
    """all ranks receive every rank's [Nc_r, C] logits: one tiny size exchange + one padded all_gather.
    Returns (list of per-rank tensors with padding removed)."""
    import torch.distributed as dist
    if world == 1:
        return [logits]
    n = torch.tensor([logits.shape[0]], device=logits.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = logits.new_zeros((mx, logits.shape[1]))
    pad[:logits.shape[0]] = logits
    out = logits.new_empty((world * mx, logits.shape[1]))
    dist.all_gather_into_tensor(out, pad, group=group)
    return [out[r * mx: r * mx + sizes[r]] for r in range(world)]
