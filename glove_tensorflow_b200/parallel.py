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


# ---- row-sharded owner-computes scheme (glove_prepare_batches_sharded / glove_shard_* in include/glove_b200.h) ---------
# Row id -> (owner = id % world, local row = id // world).  Every rank owns V_rows = ceil(V / world) rows of BOTH tables
# with their optimizer state; per step it (1) brings its own rows up to date, (2) receives the opposite-side snapshot
# rows its segments reference from their owners (request-only all-to-all), (3) updates the segments it owns and
# (4) all-reduces three floats (loss, d loss / d global bias, regulariser) so every rank applies the same global-bias
# update.  No gradient leaves its owner.

def shard_owner(ids, world):
    """Rank that owns row ``id`` (mirrors ``id % n_shards`` in prepare_sharded)."""
    return np.asarray(ids) % world


def shard_local(ids, world):
    """Row index inside the owner's table."""
    return np.asarray(ids) // world


def shard_rows(num_rows, world):
    """Rows every rank allocates per table (the last ranks may hold one zero pad row)."""
    return (int(num_rows) + world - 1) // world


def shard_requests(ids, opposite_ids, rank, world):
    """For the segments ``rank`` owns on one side of a batch: the sorted unique opposite-side ids it has to fetch,
    split by owner -- ``requests[k]`` = ids owned by rank k (``requests[rank]`` are local reads).  This is the content of
    the plan's need_pos / need_off lists, expressed in global ids."""
    ids, opposite_ids = np.asarray(ids), np.asarray(opposite_ids)
    need = np.unique(opposite_ids[shard_owner(ids, world) == rank])
    return [need[shard_owner(need, world) == k] for k in range(world)]
