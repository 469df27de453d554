"""GPU tests of the input pipeline (keyed shuffle), EVAL pass and the host-buffer entry point, against the oracle."""
import os

import numpy as np
import pytest

from conftest import make_coo
from oracle import glove_oracle as o

pytestmark = pytest.mark.gpu


def _shuffle(key, nnz, first, count):
    import ctypes
    import torch
    from glove_tensorflow_b200._lib import lib, check
    out = torch.empty(count, dtype=torch.int64, device="cuda:0")
    check(lib.glove_shuffle_indices(key, nnz, first, count, ctypes.c_void_p(out.data_ptr()),
                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("nnz", [1, 2, 3, 17, 1000, 873186])
def test_shuffle_matches_oracle_bijection(nnz):
    key = 0xC0FFEE
    got = _shuffle(key, nnz, 0, 2 * nnz + 5).cpu().numpy()
    pos = np.arange(2 * nnz + 5)
    want = np.concatenate([o.feistel_permute(pos[pos // nnz == e] % nnz, nnz, (key + e) & 0xFFFFFFFF) for e in range(int(pos[-1] // nnz) + 1)])
    assert np.array_equal(got, want)
    assert np.array_equal(np.sort(got[:nnz]), np.arange(nnz))          # an epoch is a permutation
    if nnz > 100:
        assert not np.array_equal(got[:nnz], got[nnz:2 * nnz])          # epochs differ
        assert np.mean(got[:nnz] == np.arange(nnz)) < 0.01


def test_shuffle_full_size_is_a_permutation():
    """Size-independent property at the cfg3 nnz (2^28; 1e9 follows the same code path with a wider Feistel)."""
    import torch
    nnz = 1 << 28
    idx = _shuffle(7, nnz, 0, nnz)
    assert int(idx.min()) == 0 and int(idx.max()) == nnz - 1
    s = torch.sort(idx).values
    assert bool((s[1:] - s[:-1] == 1).all())
    del idx, s
    nnz = 1_000_000_007  # odd size above 2^29: cycle-walking path
    idx = _shuffle(9, nnz, nnz - 5, 1 << 20)  # straddles the epoch boundary
    assert int(idx.min()) >= 0 and int(idx.max()) < nnz
    assert len(torch.unique(idx[5:])) == (1 << 20) - 5


@pytest.mark.parametrize("head", ["glove", "logistic"])
def test_eval_metrics_match_oracle(head):
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, n, B = 300, 20, 5000, 512   # last batch is short (5000 = 9*512 + 392)
    coo = make_coo(V, n, 3)
    st = o.init_state(V, d, 4)
    st.g = np.float32(0.3)
    eng = GloveEngine(V, d, head=head, batch_size=B, plan_steps=2, max_steps=8)
    eng.load_state(st.R, st.C, st.rb, st.cb, st.g)
    a, b = ("target", "weight") if head == "glove" else ("pos", "neg")
    eng.set_coo(coo["row"], coo["col"], coo[a], coo[b])
    got = eng.eval_metrics(B)
    if head == "glove":
        want = o.eval_metrics(st, coo, B)
        for k, v in want.items():
            assert abs(got[k] - v) <= 1e-5 * max(abs(v), 1e-3), (k, got[k], v)
    else:
        losses = []
        for s in range(0, n, B):
            bt = {k: v[s:s + B] for k, v in coo.items()}
            losses.append(float(o.forward_loss(st, bt, head="logistic", neg_factor=1.0)[0]))
        assert abs(got["loss"] - np.mean(losses)) <= 1e-5 * abs(np.mean(losses))


def test_train_then_eval_final_loss_within_1pct():
    """north_star: final loss within 1 % of the reference on the same inputs."""
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, n, B, steps = 2000, 32, 40000, 1024, 150
    coo = make_coo(V, n, 8)
    batches = np.random.default_rng(9).integers(0, n, (steps, B))
    st = o.init_state(V, d, 10)
    ref = st.copy()
    o.train(ref, coo, batches, learning_rate=0.01)
    want = o.eval_metrics(ref, coo, B)
    eng = GloveEngine(V, d, learning_rate=0.01, batch_size=B, plan_steps=16, max_steps=steps + 16)
    eng.load_state(st.R, st.C, st.rb, st.cb, st.g)
    eng.set_coo(coo["row"], coo["col"], coo["target"], coo["weight"])
    eng.set_batches(batches)
    eng.train(steps)
    got = eng.eval_metrics(B)
    assert abs(got["loss"] - want["loss"]) <= 0.01 * want["loss"]
    assert abs(got["average_loss"] - want["average_loss"]) <= 1e-4 * want["average_loss"]


def test_keyed_shuffle_training_runs_and_learns():
    """No injected order: batches come from the on-GPU keyed shuffle; the loss must fall and the device step counter
    must track the host."""
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, n, B = 1000, 16, 30000, 2048
    coo = make_coo(V, n, 11)
    eng = GloveEngine(V, d, learning_rate=0.01, batch_size=B, plan_steps=8, max_steps=128)
    eng.init_uniform(0)
    eng.set_coo(coo["row"], coo["col"], coo["target"], coo["weight"], shuffle_key=5)
    l = eng.train(100)
    assert np.all(np.isfinite(l)) and l[-10:].mean() < 0.5 * l[:10].mean()


def test_host_entry_point_matches_device_path():
    import torch
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, n, B, K = 500, 24, 8192, 256, 4
    coo = make_coo(V, n, 12)
    st = o.init_state(V, d, 13)
    batches = np.arange(2 * K * B).reshape(2 * K, B)          # file order, two calls
    ref = st.copy()
    want = np.array(o.train(ref, coo, batches, learning_rate=0.01))
    eng = GloveEngine(V, d, learning_rate=0.01, batch_size=B, plan_steps=K, max_steps=64)
    eng.load_state(st.R, st.C, st.rb, st.cb, st.g)
    got = []
    for c in range(2):
        sl = slice(c * K * B, (c + 1) * K * B)
        hl = torch.empty(K, dtype=torch.float32).pin_memory()
        eng.train_steps_host(torch.from_numpy(coo["row"][sl].copy()).pin_memory(), torch.from_numpy(coo["col"][sl].copy()).pin_memory(),
                             torch.from_numpy(coo["target"][sl].copy()).pin_memory(), torch.from_numpy(coo["weight"][sl].copy()).pin_memory(), hl)
        got.append(hl.numpy().copy())
    got = np.concatenate(got)
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-5
    state = eng.get_state()
    assert np.max(np.abs(state["R"] - ref.R)) / np.max(np.abs(ref.R)) < 1e-5


@pytest.mark.parametrize("optimizer,adam_mode", [("Adam", "replay"), ("Adam", "lazy"), ("Adagrad", "replay")])
def test_host_entry_pipelined_call_is_bit_identical_to_chunked_calls(optimizer, adam_mode):
    """One glove_train_steps_host call over 6 plan chunks (H2D + plan of chunk c+1 and the Adam catch-up overlap the
    steps in flight) == 6 calls of one chunk each == the device-resident path, bit for bit; and == the oracle to 1e-5."""
    import torch
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, n, B, K, chunks = 3000, 40, 6 * 5 * 512, 512, 5, 6
    coo = make_coo(V, n, 31)
    st = o.init_state(V, d, 32)
    pin = {k: torch.from_numpy(coo[k].copy()).pin_memory() for k in ("row", "col", "target", "weight")}

    def run(calls):
        eng = GloveEngine(V, d, optimizer=optimizer, adam_mode=adam_mode, learning_rate=0.01, batch_size=B, plan_steps=K, max_steps=64)
        eng.load_state(st.R, st.C, st.rb, st.cb, st.g)
        per = chunks // calls * K * B
        losses = []
        for c in range(calls):
            hl = torch.empty(per // B, dtype=torch.float32).pin_memory()
            eng.train_steps_host(*[pin[k][c * per:(c + 1) * per] for k in ("row", "col", "target", "weight")], hl)
            losses.append(hl.numpy().copy())
        return np.concatenate(losses), eng.get_state()
    l1, s1 = run(1)
    l6, s6 = run(6)
    l2, s2 = run(2)
    assert np.array_equal(l1, l6) and np.array_equal(l1, l2)
    for k in ("R", "C", "rb", "cb"):
        assert np.array_equal(s1[k], s6[k]) and np.array_equal(s1[k], s2[k]), k
    dev = GloveEngine(V, d, optimizer=optimizer, adam_mode=adam_mode, learning_rate=0.01, batch_size=B, plan_steps=K, max_steps=64)
    dev.load_state(st.R, st.C, st.rb, st.cb, st.g)
    dev.set_coo(coo["row"], coo["col"], coo["target"], coo["weight"])
    dev.set_batches(np.arange(chunks * K * B).reshape(chunks * K, B))
    ld = dev.train(chunks * K)
    assert np.array_equal(ld, l1) and np.array_equal(dev.get_state()["R"], s1["R"])
    if adam_mode == "lazy":
        return                                        # LazyAdam is not the reference's optimizer: no oracle for it
    ref = st.copy()
    want = np.array(o.train(ref, coo, np.arange(chunks * K * B).reshape(chunks * K, B), optimizer=optimizer, learning_rate=0.01))
    assert np.max(np.abs(l1 - want) / np.abs(want)) < 1e-5
    assert np.max(np.abs(s1["R"] - ref.R)) / np.max(np.abs(ref.R)) < 1e-5


def test_train_chunk_from_host_at_any_step_alignment():
    """The data-parallel-aware host entry (bench e2e at N > 1; here N = 1): a chunk planned from HOST buffers must be the
    one the steps use even when the current step is not a multiple of plan_steps, and the device-resident plan prefetch
    must not overwrite it."""
    import torch
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, n, B, K = 800, 24, 16384, 256, 4
    coo = make_coo(V, n, 41)
    st = o.init_state(V, d, 42)
    pre = np.random.default_rng(43).integers(0, n, (3, B))              # 3 device-resident steps first: host_step = 3
    host_idx = np.random.default_rng(44).integers(0, n, (2 * K, B))     # then two host-fed chunks
    ref = st.copy()
    want = np.array(o.train(ref, coo, np.concatenate([pre, host_idx]), learning_rate=0.01))
    eng = GloveEngine(V, d, learning_rate=0.01, batch_size=B, plan_steps=K, max_steps=64)
    eng.load_state(st.R, st.C, st.rb, st.cb, st.g)
    eng.set_coo(coo["row"], coo["col"], coo["target"], coo["weight"])
    eng.set_batches(np.concatenate([pre, np.zeros((2 * K + 8, B), np.int64)]))   # resident batches differ from the host-fed ones
    got = list(eng.train(3))
    pins = []
    for c in range(2):
        sel = host_idx[c * K:(c + 1) * K].reshape(-1)
        pins.append(tuple(torch.from_numpy(coo[k][sel].copy()).pin_memory() for k in ("row", "col", "target", "weight")))
    got += list(eng.train_chunks_from_host(pins))                      # both chunks in one pipelined call
    assert eng.host_step == 3 + 2 * K
    assert np.max(np.abs(np.array(got) - want) / np.abs(want)) < 1e-5
    assert np.max(np.abs(eng.get_state()["R"] - ref.R)) / np.max(np.abs(ref.R)) < 1e-5
    # chunk by chunk gives the same bits, and the engine goes back to its resident COO afterwards
    eng2 = GloveEngine(V, d, learning_rate=0.01, batch_size=B, plan_steps=K, max_steps=64)
    eng2.load_state(st.R, st.C, st.rb, st.cb, st.g)
    eng2.set_coo(coo["row"], coo["col"], coo["target"], coo["weight"])
    eng2.set_batches(np.concatenate([pre, np.zeros((2 * K + 8, B), np.int64)]))
    got2 = list(eng2.train(3)) + list(eng2.train_chunk_from_host(*pins[0])) + list(eng2.train_chunk_from_host(*pins[1]))
    assert np.array_equal(np.array(got2), np.array(got)) and np.array_equal(eng2.get_state()["R"], eng.get_state()["R"])
    assert np.all(np.isfinite(eng2.train(2)))


def test_resume_after_long_idle_gaps_is_bit_identical(tmp_path):
    """Checkpoint / resume with rows that have sat idle for > 1,000 steps (their Adam m has underflowed to exactly 0 while
    v has not): the resumed run must continue bit for bit like the run that wrote the checkpoint and went on.  Rows
    V-8 .. V-1 are touched in step 0 only, then again after the checkpoint; the checkpoint stores the per-row
    "ever updated" mask, so their v keeps decaying across the restart."""
    from glove_tensorflow_b200 import train_utils
    from glove_tensorflow_b200.engine import GloveEngine
    V, d, B, T_ckpt, T_end = 64, 8, 16, 1100, 1130
    rng = np.random.default_rng(9)
    n = T_end * B
    row = rng.integers(0, V - 8, n).astype(np.int32)
    col = rng.integers(0, V - 8, n).astype(np.int32)
    row[:8] = np.arange(V - 8, V); col[8:16] = np.arange(V - 8, V)          # step 0 touches the rare rows ...
    late = (T_ckpt + 20) * B
    row[late:late + 8] = np.arange(V - 8, V); col[late + 8:late + 16] = np.arange(V - 8, V)   # ... and step T_ckpt+20 again
    coo = dict(row=row, col=col, target=rng.normal(1.0, 0.5, n).astype(np.float32), weight=rng.uniform(0.1, 1, n).astype(np.float32))
    batches = np.arange(n).reshape(T_end, B)

    def engine():
        e = GloveEngine(V, d, learning_rate=0.05, batch_size=B, plan_steps=8, max_steps=T_end + 8)
        e.init_uniform(3)
        e.set_coo(coo["row"], coo["col"], coo["target"], coo["weight"])
        e.set_batches(batches)
        return e

    a = engine()
    a.train(T_ckpt)
    job = str(tmp_path)
    path = train_utils.save_checkpoint(a, job)
    m_rare = a.get_state(slots=True)["R/s0"][V - 8:]
    assert not m_rare.any(), "the rare rows' first moment should have underflowed to 0 (the case the mask exists for)"
    assert a.get_state(slots=True)["R/s1"][V - 8:].any()
    la = a.train(T_end - T_ckpt)
    sa = a.get_state(slots=True)

    b = engine()
    assert train_utils.latest_checkpoint(job) == path
    assert train_utils.load_checkpoint(b, path) == T_ckpt
    lb = b.train(T_end - T_ckpt)
    sb = b.get_state(slots=True)
    assert np.array_equal(la, lb)
    for k in sa:
        assert np.array_equal(np.asarray(sa[k]), np.asarray(sb[k])), k
    # a truncated newer file must not be picked up, and a checkpoint beyond max_steps is refused with a clear message
    open(os.path.join(job, "model.ckpt-999999.npz"), "wb").write(b"PK\x03\x04 truncated")
    assert train_utils.latest_checkpoint(job) == path
    os.remove(os.path.join(job, "checkpoint"))
    assert train_utils.latest_checkpoint(job) == path
    small = GloveEngine(V, d, batch_size=B, plan_steps=8, max_steps=100)
    with pytest.raises(ValueError, match="beyond this run"):
        train_utils.load_checkpoint(small, path)
