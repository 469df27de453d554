"""Data-parallel partition rule shared by the CUDA kernels (update_kernel's ownership test) and the host.

Replicated tables, global batch of B = world * B_local triples, every rank holds the same plan: the triple with in-batch
arrival index p belongs to rank p // (B // world).  Each rank accumulates gradient partial sums only for its own
triples into a dense per-slot buffer; the buffers are all-reduced (NCCL) and every replica applies the same update.
The reference has no multi-device path: equivalence is defined against the 1-GPU / oracle step on the
global batch."""
import numpy as np


def dp_owner(p, batch_size, world):
    """Rank that owns in-batch arrival index ``p`` (mirrors ``rec.w / dp_block == dp_rank`` in update_kernel)."""
    if batch_size % world:
        raise ValueError("global batch %d not divisible by world size %d" % (batch_size, world))
    return np.asarray(p) // (batch_size // world)


def dp_shard(batch_idx, rank, world):
    """The slice of a global batch (index list into the COO) that ``rank`` accumulates."""
    batch_idx = np.asarray(batch_idx)
    block = len(batch_idx) // world
    if len(batch_idx) % world:
        raise ValueError("global batch %d not divisible by world size %d" % (len(batch_idx), world))
    return batch_idx[rank * block:(rank + 1) * block]
