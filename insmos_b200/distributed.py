"""Multi-GPU partition of the path: samples of a batch are independent, so they are
sharded one per rank (weights replicated) and the only exchange step is ONE gather of the per-point MOS logits
(variable length, ~1.44 MB per 120k-point scan) at the end.  No collective on the data path.

The exchange is a SINGLE fixed-size collective with no host synchronisation: every rank contributes a padded block
[pad_rows + 1, C] whose last row carries its row count, so neither a size exchange nor a `.item()` sits between the
forward and the collective (round 1 did both inside every step, which drained the stream and destroyed the two-in-flight
overlap of ScanPipeline)."""
import torch


def shard_samples(n_samples, rank, world):
    """indices of the samples rank `rank` processes (round robin, sample i -> GPU i mod world)."""
    return list(range(rank, n_samples, world))


class GatheredLogits:
    """result of gather_logits_padded: `buf` [world, pad_rows + 1, C] on every rank; counts stay on the device.
    rows(r) / counts() read the counts back (host sync) -- consumers on the device use `buf` and `count_tensor()`."""

    def __init__(self, buf, pad_rows):
        self.buf, self.pad_rows = buf, pad_rows

    def count_tensor(self):
        return self.buf[:, self.pad_rows, 0].to(torch.int64)

    def counts(self):
        return [int(v) for v in self.count_tensor().cpu().tolist()]

    def rows(self, r, n=None):
        n = self.counts()[r] if n is None else n
        return self.buf[r, :n]


def gather_logits_padded(logits, world, pad_rows, group=None, out=None):
    """all ranks receive every rank's [Nc_r, C] logits with ONE all_gather of fixed-size blocks; the row counts travel in
    the payload (exact in fp32 up to 2^24 rows).  No host synchronisation.  `out` may be a preallocated
    [world, pad_rows + 1, C] buffer."""
    import torch.distributed as dist
    n, C = logits.shape
    if n > pad_rows:
        raise ValueError("gather_logits_padded: %d rows exceed pad_rows=%d" % (n, pad_rows))
    if pad_rows >= (1 << 24):
        raise ValueError("gather_logits_padded: pad_rows must stay below 2^24 (count travels as fp32)")
    block = logits.new_empty((pad_rows + 1, C))
    block[:n] = logits
    block[n:] = 0
    block[pad_rows, 0] = float(n)
    if out is None:
        out = logits.new_empty((world, pad_rows + 1, C))
    if world == 1:
        out[0] = block
    else:
        dist.all_gather_into_tensor(out.view(world * (pad_rows + 1), C), block, group=group)
    return GatheredLogits(out, pad_rows)


def gather_logits(logits, world, group=None, pad_rows=None):
    """list of per-rank logits with the padding removed (reads the counts back: one host sync AFTER the collective).
    pad_rows defaults to the next multiple of 4096 above this rank's row count, which must then be equal on all ranks --
    pass an explicit bound when rank sizes differ by more than that."""
    if world == 1:
        return [logits]
    if pad_rows is None:
        pad_rows = -(-max(int(logits.shape[0]), 1) // 4096) * 4096
    g = gather_logits_padded(logits, world, pad_rows, group=group)
    counts = g.counts()
    return [g.rows(r, counts[r]) for r in range(world)]
