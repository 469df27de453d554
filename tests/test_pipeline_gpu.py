"""GPU tests of the input pipeline (keyed shuffle), EVAL pass and the host-buffer entry point, against the oracle."""
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
