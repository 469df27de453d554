"""Parity on the configurations BASELINE.json itself names (configs[1], [2], [4]) at their full shapes, through the C ABI.

cfg2  text8 shape, V=10,001, d=64, Adam, B=65,536 -- 100 injected steps vs the C oracle (fp32) and its fp64 shadow
cfg3  wiki6b shape, V=400,000, d=300, Adam, B=65,536 -- 24 injected steps; the first two batches are drawn uniformly over
      the vocabulary so that ~100k rows per side sit idle for up to 22 steps before Zipf batches touch them again
cfg5  cosine top-k over a 2.2M x 300 table, k=10, 4,096 queries: tensor-core path vs the exact fp32 scan (bit-equal ids
      and similarities), planted exact duplicates (ties -> lower id), and the NumPy oracle on a subset of the queries

The reference CPU side of these tests is the C port of the oracle (oracle/glove_oracle.c, OpenMP); the NumPy restatement
it is pinned on (tests/test_oracle.py) would take minutes at these sizes.
"""
import os

import numpy as np
import pytest

from conftest import make_coo
from oracle import glove_oracle as o
from test_train_gpu import RTOL, SHADOW_C, _rel, _record, _shadow, _dump_maxima  # noqa: F401  (fixture re-export)

pytestmark = pytest.mark.gpu


def _zipf_ids(rng, V, n):
    p = 1.0 / np.arange(1, V + 1, dtype=np.float64)
    cdf = np.cumsum(p / p.sum())
    return np.minimum(np.searchsorted(cdf, rng.random(n)), V - 1).astype(np.int32)


def _run(V, d, B, coo, batches, lr, K, name):
    from glove_tensorflow_b200.engine import GloveEngine
    steps = len(batches)
    st = o.init_state(V, d, 77)
    (c32, l32), (c64, l64) = _shadow(st, coo, batches, optimizer="Adam", head="glove", lr=lr, reg_scale=2.0,
                                     adam_mode="replay", neg_factor=1.0)
    eng = GloveEngine(V, d, learning_rate=lr, batch_size=B, plan_steps=K, max_steps=steps + K)
    eng.load_state(st.R, st.C, st.rb, st.cb, st.g)
    eng.set_coo(coo["row"], coo["col"], coo["target"], coo["weight"])
    eng.set_batches(batches)
    losses = eng.train(steps)
    got = eng.get_state()
    e_loss = float(np.max(np.abs(losses - l32) / np.abs(l32)))
    _record(name, loss_vs_oracle32=e_loss, loss_vs_shadow64=float(np.max(np.abs(losses - l64) / np.abs(l64))),
            final_loss=float(losses[-1]), final_loss_oracle=float(l32[-1]))
    assert e_loss < RTOL, (losses[:3], l32[:3])
    assert abs(losses[-1] - l32[-1]) <= 0.01 * abs(l32[-1])                    # north_star: final loss within 1 %
    for k in ("R", "C", "rb", "cb"):
        e_gpu, e_o32 = _rel(got[k], getattr(c64, k)), _rel(getattr(c32, k), getattr(c64, k))
        _record(name, **{k + "_gpu_vs_shadow64": e_gpu, k + "_oracle32_vs_shadow64": e_o32,
                         k + "_gpu_vs_oracle32": _rel(got[k], getattr(c32, k))})
        assert e_gpu <= max(RTOL, SHADOW_C * e_o32), (k, e_gpu, e_o32)
    assert abs(float(got["g"]) - float(c64.g)) <= RTOL * max(abs(float(c64.g)), 1e-3)
    assert got["step"] == steps
    return eng


def test_cfg2_text8_shape_batch_64k_100_steps():
    """BASELINE configs[1]: V=10,001, d=64, Adam lr 1e-3, l2 0.01, batch 65,536, 100 injected steps."""
    V, d, B, steps, n = 10_001, 64, 65_536, 100, 873_186
    coo = make_coo(V, n, 21)
    batches = np.random.default_rng(22).integers(0, n, (steps, B))
    _run(V, d, B, coo, batches, 0.001, 16, "cfg2_text8_B65536_100steps")


def test_cfg3_wiki6b_shape_with_idle_gaps():
    """BASELINE configs[2] shape: V=400,000, d=300, Adam, batch 65,536.  Steps 0-1 touch ~2 x 60k rows per side uniformly;
    the Zipf batches of steps 2..23 come back to them after idle gaps of up to 22 steps, every other row they touch for
    the first time."""
    V, d, B, steps, n = 400_000, 300, 65_536, 24, 1 << 21
    rng = np.random.default_rng(31)
    row, col = _zipf_ids(rng, V, n), _zipf_ids(rng, V, n)
    nu = 2 * B                                                         # the first 2B triples: uniform ids
    row[:nu], col[:nu] = rng.integers(0, V, nu), rng.integers(0, V, nu)
    col = np.where(col == row, (col + 1) % V, col).astype(np.int32)
    count = 10 + np.floor(np.minimum(rng.pareto(1.0, n), 1e6))
    value = count * rng.uniform(0.3, 0.6, n)
    coo = {"row": row, "col": col, "target": np.log(value).astype(np.float32),
           "weight": np.minimum(1.0, (count / 100.0) ** 0.75).astype(np.float32)}
    batches = np.concatenate([np.arange(nu).reshape(2, B), rng.integers(nu, n, (steps - 2, B))])
    eng = _run(V, d, B, coo, batches, 0.001, 8, "cfg3_wiki6b_shape_24steps_idle_gaps")
    # the idle rows really were there: rows touched in steps 0-1 and again later
    early = np.unique(row[:nu])
    late = np.unique(row[batches[2:].reshape(-1)])
    assert len(np.intersect1d(early, late)) > 10_000
    del eng


def test_cfg5_topk_2p2m_vocab():
    """BASELINE configs[4]: 2.2M x 300 table, k=10.  4,096 queries through the tcgen05 candidate pass + exact fp32 re-score
    vs (a) the exact fp32 scan of the whole table (ids AND similarities bit-equal), (b) planted exact duplicates: every
    copy of a planted row must list all copies first, in id order, (c) the NumPy oracle on 48 of the queries."""
    import torch
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, k, Q = 2_200_000, 300, 10, 4096
    rng = np.random.default_rng(41)
    T = rng.uniform(-0.05, 0.05, (V, d)).astype(np.float32)
    groups = []                                                        # planted exact ties: 3 copies of 64 source rows
    src = rng.choice(V // 2, 64, replace=False)
    for g, s in enumerate(src):
        a, b = V // 2 + 1000 + 17 * g, V - 1 - 13 * g
        T[a], T[b] = T[s], T[s]
        groups.append(sorted((int(s), a, b)))
    T[12345] = 0.0                                                     # zero row: the 1e-12 floor of l2_normalize
    eng = GloveEngine(V, d, optimizer="SGD", batch_size=64, plan_steps=1, max_steps=4)
    zb = np.zeros(V, np.float32)
    eng.load_state(T, T, zb, zb)
    q = np.concatenate([np.array(groups).reshape(-1), [12345, 0, V - 1],
                        rng.integers(0, V, Q - 3 * len(groups) - 3)]).astype(np.int32)
    assert eng.tc_path_covers(k)
    s_tc, i_tc = eng.topk(q, k)
    fallbacks = eng.last_topk_fallbacks
    s_ex, i_ex = eng.topk(q, k, exact_fp32=True)
    assert np.array_equal(i_tc, i_ex)
    assert np.array_equal(s_tc, s_ex)
    assert fallbacks <= Q // 20, fallbacks
    for g, grp in enumerate(groups):                                   # ties broken by id, at 2.2M scale
        for c in range(3):
            assert list(i_tc[3 * g + c, :3]) == grp, (grp, i_tc[3 * g + c])
            assert s_tc[3 * g + c, 0] == s_tc[3 * g + c, 1] == s_tc[3 * g + c, 2]
    sub = np.concatenate([np.arange(0, 24), np.arange(200, 224)])      # NumPy oracle on planted + random queries
    want_sim, want_idx = o.cosine_topk(T, q[sub], k)
    np.testing.assert_allclose(s_tc[sub], want_sim, atol=2e-6)
    bad = i_tc[sub] != want_idx
    if bad.any():        # fp32 summation order differs from NumPy's: a swap is legitimate only inside that noise
        r, c = np.nonzero(bad)
        assert np.all(np.abs(want_sim[r, c] - s_tc[sub][r, c]) < 5e-7)
    _record("cfg5_topk_V2200000_Q4096", fallbacks=fallbacks, id_mismatch_vs_exact_scan=int((i_tc != i_ex).sum()),
            id_swaps_vs_numpy_inside_noise=int(bad.sum()))
    del eng
    torch.cuda.empty_cache()
