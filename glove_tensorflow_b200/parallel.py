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


# ---- frequency-balanced owner map (GloveEngine.balance_owners) ---------------------------------------------------------
def balanced_labels(freq, world, batch_size, nnz, hot=65536):
    """Relabelling ``label[id]`` (owner = label % world, local row = label // world) that spreads the work of a Zipf
    vocabulary evenly: the ``hot`` costliest ids are dealt greedily to the least loaded owner (cost of an id = expected
    triples per step and side + ~3 triples' worth of fixed work when its row is touched), the tail round-robin; every
    owner keeps exactly the rows it has under ``id % world``, and inside an owner ids keep their order (top-k ties still
    break towards the lower original id).  Returns (label, unlabel, relative load of the hot part per owner)."""
    freq = np.asarray(freq, np.float64)
    V, N = len(freq), int(world)
    t = freq * (float(batch_size) / (2.0 * max(int(nnz), 1)))
    cost = t + 3.0 * (1.0 - np.exp(-t))
    order = np.argsort(-cost, kind="stable")
    cap = np.array([(V - o + N - 1) // N for o in range(N)], np.int64)
    owner = np.full(V, -1, np.int64)
    load = np.zeros(N)
    n_hot = int(min(hot, V))
    for i in order[:n_hot]:
        o = int(np.argmin(np.where(cap > 0, load, np.inf)))
        owner[i] = o
        load[o] += cost[i]
        cap[o] -= 1
    rest = order[n_hot:]
    if rest.size:
        slots = np.concatenate([np.stack([np.arange(cap[o]), np.full(cap[o], o)], 1) for o in range(N)])
        slots = slots[np.lexsort((slots[:, 1], slots[:, 0]))]
        owner[rest] = slots[:, 1]
    label = np.empty(V, np.int64)
    for o in range(N):
        ids = np.flatnonzero(owner == o)                     # ascending original id
        label[ids] = np.arange(ids.size, dtype=np.int64) * N + o
    unlabel = np.empty(V, np.int64)
    unlabel[label] = np.arange(V, dtype=np.int64)
    return label, unlabel, load / max(load.mean(), 1e-30)


# ---- host-fed chunks, one share per rank (GloveEngine.train_chunks_from_host(sliced=True)) -----------------------------
def assemble_chunk(gathered, world, n_steps, b_local):
    """``gathered`` = all-gather over ranks of [4 arrays][n_steps][b_local] (every rank's share of each batch of a chunk) ->
    view [4][n_steps][world][b_local]: array j of the chunk in batch order (global batch = the ranks' shares in rank
    order).  Works on torch tensors and numpy arrays alike."""
    g = gathered.reshape(world, 4, n_steps, b_local)
    return g.permute(1, 2, 0, 3) if hasattr(g, "permute") else g.transpose(1, 2, 0, 3)


# ---- shared plan construction (GloveEngine.enable_plan_sharing) ----------------------------------------------------------
def shared_plan_slot(chunk, world):
    """(round, builder rank, build buffer) of plan chunk ``chunk``: the chunks are dealt round-robin, rank ``chunk % world``
    builds the plan of chunk ``chunk`` into its build buffer ``round & 1`` while the previous round is being trained on."""
    r = int(chunk) // int(world)
    return r, int(chunk) % int(world), r & 1


def round_send_buffer(shares, world):
    """Host-fed rounds: ``shares[q]`` = this rank's share of chunk q of the round, 4 arrays of ``m`` 32-bit words each.
    Returns the all-to-all send buffer [dest rank q][array][m]: after ``all_to_all_single`` rank q holds
    [src rank][array][m] -- every rank's share of ITS chunk, the layout ``assemble_chunk`` takes.  numpy in, numpy out
    (the engine writes the same layout straight into a device buffer)."""
    assert len(shares) == world and all(len(s) == 4 for s in shares)
    return np.stack([np.stack([np.asarray(a).view(np.int32).reshape(-1) for a in s]) for s in shares])
