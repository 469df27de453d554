"""PREDICT path: cosine top-k against the oracle (ids exact, ties broken by lower id)."""
import numpy as np
import pytest

from oracle import glove_oracle as o

pytestmark = pytest.mark.gpu


def _table(V, d, seed, dup=True):
    rng = np.random.default_rng(seed)
    T = rng.uniform(-0.05, 0.05, (V, d)).astype(np.float32)
    if dup:
        T[V // 2] = T[7]                 # exact duplicate rows: identical similarity to everything -> tie by id
        T[V - 1] = T[7]
        T[11] = 0.0                       # zero row: l2_normalize's 1e-12 floor
    return T


def _check(sim, idx, T, q, k):
    want_sim, want_idx = o.cosine_topk(T, q, k)
    assert np.all(np.diff(sim, axis=1) <= 0)
    np.testing.assert_allclose(sim, want_sim, atol=2e-6)
    bad = idx != want_idx
    if bad.any():   # a swap is legitimate only between entries whose similarities agree to fp32 summation noise
        r, c = np.nonzero(bad)
        assert np.all(np.abs(want_sim[r, c] - sim[r, c]) < 5e-7), (idx[r[0]], want_idx[r[0]])
        for rr in np.unique(r):
            assert set(idx[rr]) == set(want_idx[rr]) or np.abs(want_sim[rr, -1] - sim[rr, -1]) < 5e-7


@pytest.mark.parametrize("V,d,k,exact", [(3000, 64, 20, True), (3000, 64, 20, False), (5000, 300, 10, True),
                                          (5000, 300, 10, False), (700, 30, 1, True), (40, 8, 32, True), (40, 8, 32, False)])
def test_topk_matches_oracle(V, d, k, exact):
    from glove_tensorflow_b200.engine import GloveEngine
    T = _table(V, d, 1)
    eng = GloveEngine(V, d, batch_size=64, plan_steps=1, max_steps=4)
    eng.load_state(T, T[::-1].copy(), np.zeros(V, np.float32) + 0.3, np.zeros(V, np.float32))
    q = np.array([7, V // 2, V - 1, 0, 11, 3, V - 2] + list(range(20, 20 + 57)), np.int32) % V
    sim, idx = eng.topk(q, k, exact_fp32=exact)
    _check(sim, idx, T, q, k)
    if not exact and eng.tc_path_covers(k):
        # the tensor-core candidates must carry the answer themselves: at most a few queries may need the exact fallback
        assert eng.last_topk_fallbacks is not None and eng.last_topk_fallbacks <= len(q) // 8, eng.last_topk_fallbacks
    # the three duplicates of row 7 lead every one of their own lists in id order
    if k >= 3:
        for r in range(3):
            assert list(idx[r, :3]) == sorted([7, V // 2, V - 1])


def test_topk_every_vocab_row_like_the_exporter():
    """export_embeddings runs PREDICT over every vocab line (ref src/models/estimator.py:59-69)."""
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, k = 2000, 64, 20
    T = _table(V, d, 2, dup=False)
    eng = GloveEngine(V, d, batch_size=64, plan_steps=1, max_steps=4)
    eng.load_state(T, T, np.zeros(V, np.float32), np.zeros(V, np.float32))
    q = np.arange(V, dtype=np.int32)
    sim, idx = eng.topk(q, k)
    assert np.array_equal(idx[:, 0], q)                      # every row is its own nearest neighbour
    _check(sim, idx, T, q, k)
    assert eng.last_topk_fallbacks <= V // 8, eng.last_topk_fallbacks


def test_tensor_core_candidates_alone_are_correct():
    """Large-ish table, no fallback allowed to hide a broken MMA: compare the tensor-core result with the exact scan and
    require (almost) no guarantee-check fallbacks on well separated data."""
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, k = 40000, 300, 10
    T = _table(V, d, 5, dup=False)
    eng = GloveEngine(V, d, batch_size=64, plan_steps=1, max_steps=4)
    eng.load_state(T, T, np.zeros(V, np.float32), np.zeros(V, np.float32))
    q = np.random.default_rng(6).integers(0, V, 777).astype(np.int32)
    s1, i1 = eng.topk(q, k)
    fallbacks = eng.last_topk_fallbacks
    s2, i2 = eng.topk(q, k, exact_fp32=True)
    assert np.array_equal(i1, i2) and np.array_equal(s1, s2)     # same fp32 routine => bit-identical similarities
    assert fallbacks <= len(q) // 20, fallbacks


def test_topk_rejects_unsupported_k():
    from glove_tensorflow_b200.engine import GloveEngine
    from glove_tensorflow_b200._lib import GloveError
    eng = GloveEngine(100, 8, batch_size=8, plan_steps=1, max_steps=2)
    eng.init_uniform(0)
    with pytest.raises(GloveError):
        eng.topk(np.array([1], np.int32), 33, exact_fp32=True)


@pytest.mark.parametrize("world,V,d,k,exact", [(2, 3001, 64, 10, False), (2, 3001, 64, 10, True), (4, 2601, 300, 20, False),
                                               (3, 4000, 40, 31, True), (8, 9001, 300, 10, False)])
def test_row_sharded_topk_matches_oracle(world, V, d, k, exact):
    """Row-sharded tables (owner = id % world), emulated as `world` engines on one GPU: the three phases of
    GloveEngine.topk with the collectives done by hand (sum of the partial query rows; stacking of the per-shard lists).
    Ties across shards must still go to the lower GLOBAL id; the pad row of the short shards must never surface."""
    import torch
    from glove_tensorflow_b200.engine import GloveEngine
    T = _table(V, d, 4)
    T[V - 2] = T[3]                                            # a tie that straddles two shards
    engs = []
    for r in range(world):
        e = GloveEngine(V, d, batch_size=64, plan_steps=1, max_steps=4, dp_rank=r, dp_world=world, dp_mode="sharded")
        if world >= 4:                                         # frequency-balanced owner map: ids are relabelled inside the engine
            rng = np.random.default_rng(5)
            e.balance_owners(rng.integers(0, V, 4000) // 3, rng.integers(0, V, 4000) // 2, hot=50)
        e.load_state(T, T[::-1].copy(), np.zeros(V, np.float32) + 0.3, np.zeros(V, np.float32))
        engs.append(e)
    q = np.array([7, V // 2, V - 1, 0, 11, 3, V - 2] + list(range(20, 20 + 90)), np.int32) % V
    qrows = sum(e.topk_shard_query_rows(q) for e in engs)     # == dist.all_reduce
    parts = [e.topk_shard_local(qrows, k, exact) for e in engs]
    all_sim = torch.stack([p[0] for p in parts])               # == dist.all_gather_into_tensor
    all_idx = torch.stack([p[1] for p in parts])
    assert int(all_idx.max()) < V
    sim, idx = engs[0].topk_shard_merge(all_sim, all_idx, k)
    _check(sim, idx, T, q, k)
    # and it is the same answer as the unsharded engine gives
    one = GloveEngine(V, d, batch_size=64, plan_steps=1, max_steps=4)
    one.load_state(T, T[::-1].copy(), np.zeros(V, np.float32) + 0.3, np.zeros(V, np.float32))
    sim1, idx1 = one.topk(q, k, exact_fp32=True)
    assert np.array_equal(idx, idx1) or np.allclose(sim, sim1, atol=5e-7)
