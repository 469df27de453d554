"""Pins the CPU oracle (no GPU): reference README known-answer rows, golden output of the reference's own preprocessor,
autograd cross-checks of every closed-form gradient, the dense-vs-replay Adam identity, C port vs NumPy."""
import json
import os

import numpy as np
import pytest

from conftest import make_coo
from oracle import glove_oracle as o

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_readme_known_answer_rows():
    kat = json.load(open(os.path.join(GOLD, "readme_kat.json")))
    c = {n: i for i, n in enumerate(kat["columns"])}
    for r in kat["rows"]:
        # ref README.md:48-59 prints 4 decimals, truncated (0.34289 -> 0.3428); src/data/text8.py:132-139
        assert 0 <= o.glove_weight(r[c["count"]]) - r[c["glove_weight"]] < 1e-4
        assert abs(o.glove_value(r[c["value"]]) - r[c["glove_value"]]) < 2e-5 * max(1, r[c["glove_value"]]) + 1e-4


def test_golden_reference_preprocessor_output():
    """tests/golden/text8_small was written by the reference's src/data/text8.py (imported unmodified)."""
    import pandas as pd
    df = pd.read_csv(os.path.join(GOLD, "text8_small", "interaction.csv"), keep_default_na=False)
    assert list(df.columns) == ["row_token_id", "col_token_id", "count", "value", "row_token", "col_token",
                                "neg_weight", "glove_weight", "glove_value"]
    assert (df["count"] >= 10).all() and (df.row_token_id != df.col_token_id).all()
    np.testing.assert_allclose(o.glove_weight(df["count"]), df["glove_weight"], rtol=1e-12)
    np.testing.assert_allclose(o.glove_value(df["value"]), df["glove_value"], rtol=1e-12)
    # symmetric matrix: (i, j) and (j, i) carry the same count and value (text8.py:103-108)
    a = df.set_index(["row_token_id", "col_token_id"])[["count", "value"]]
    b = df.set_index(["col_token_id", "row_token_id"])[["count", "value"]]
    b.index.names = a.index.names
    assert a.sort_index().equals(b.sort_index().astype(a.dtypes.to_dict()))
    vocab = open(os.path.join(GOLD, "text8_small", "vocab.txt"), encoding="utf8").read().split("\n")
    vc = pd.read_csv(os.path.join(GOLD, "text8_small", "vocab.csv"), keep_default_na=False)
    assert vocab == list(vc["token"]) and not vocab[-1] == ""          # no trailing newline (text8.py:149-150)
    assert list(df["row_token"]) == [vocab[i] for i in df["row_token_id"]]
    # neg_weight = count_row * proportion_col of the vocab-level unigram stats (text8.py:113-116)
    np.testing.assert_allclose(vc["count"].to_numpy()[df.row_token_id] * vc["proportion"].to_numpy()[df.col_token_id],
                               df["neg_weight"], rtol=1e-9)
    assert {"na", "null", "nan"} <= set(vocab)                          # NA-like tokens survive ingest


def _torch_loss(R, C, rb, cb, g, b, head, l2, s, nu):
    """Literal restatement of the Keras layer + estimator heads in torch (fp64) for autograd."""
    import torch
    i, j = torch.as_tensor(b["row"], dtype=torch.long), torch.as_tensor(b["col"], dtype=torch.long)
    B, d = len(i), R.shape[1]
    re, ce_, rbi, cbj = R[i], C[j], rb[i], cb[j]
    z = (re * ce_).sum(-1) + rbi + cbj + g
    reg = s * ((l2 / d) * (re ** 2).sum() / B + (l2 / d) * (ce_ ** 2).sum() / B + l2 * (rbi ** 2).sum() / B
               + l2 * (cbj ** 2).sum() / B + l2 * g ** 2)
    if head == "glove":
        y, w = torch.as_tensor(b["target"], dtype=torch.float64), torch.as_tensor(b["weight"], dtype=torch.float64)
        data = (w * (z - y) ** 2).sum() / B
    else:
        p, n = torch.as_tensor(b["pos"], dtype=torch.float64), torch.as_tensor(b["neg"], dtype=torch.float64)
        sp = torch.nn.functional.softplus
        data = (p * sp(-z)).sum() / B + nu * (n * sp(z)).sum() / B
    return data + reg


@pytest.mark.parametrize("head", ["glove", "logistic"])
@pytest.mark.parametrize("reg_scale", [1.0, 2.0])
def test_gradients_match_autograd(head, reg_scale):
    import torch
    V, d, B = 30, 7, 40
    coo = make_coo(V, B, 1, zipf=False)
    st = o.init_state(V, d, 2)
    st.g = np.float32(0.2)
    loss, _, e = o.forward_loss(st, coo, head, 0.01, reg_scale, 0.7)
    grads, dg = o.sparse_grads(st, coo, e, 0.01, reg_scale)
    t = [torch.tensor(np.asarray(x, np.float64), requires_grad=True) for x in (st.R, st.C, st.rb, st.cb, st.g)]
    L = _torch_loss(*t, coo, head, 0.01, reg_scale, 0.7)
    L.backward()
    assert abs(float(L.detach()) - float(loss)) < 1e-5 * abs(float(L.detach()))
    for name, tt in zip(("R", "C", "rb", "cb"), t[:4]):
        uniq, G = grads[name]
        dense = np.zeros_like(getattr(st, name))
        dense[uniq] = G
        np.testing.assert_allclose(dense, tt.grad.numpy(), rtol=2e-4, atol=1e-7)
    assert abs(float(dg) - float(t[4].grad)) < 1e-5 * max(1e-3, abs(float(t[4].grad)))


def test_dense_adam_is_plain_adam_on_the_dense_gradient():
    """Legacy Keras sparse Adam == textbook Adam on the densified gradient (zeros on untouched rows)."""
    V, d, B, steps = 25, 5, 8, 12
    coo = make_coo(V, 200, 3, zipf=False)
    batches = np.random.default_rng(4).integers(0, 200, (steps, B))
    st = o.init_state(V, d, 5)
    ref = st.copy()
    o.train(ref, coo, batches, learning_rate=0.01)
    x = st.copy()
    m, v = np.zeros_like(x.R), np.zeros_like(x.R)
    alpha = o.alpha_table(0.01, steps)
    f = np.float32
    for s in range(steps):
        b = {k: a[batches[s]] for k, a in coo.items()}
        _, _, e = o.forward_loss(x, b)
        grads, dg = o.sparse_grads(x, b, e, 0.01, 2.0)
        g = np.zeros_like(x.R)
        g[grads["R"][0]] = grads["R"][1]
        m = (m * f(0.9)).astype(f) + (g * f(1 - f(0.9))).astype(f)
        v = (v * f(0.999)).astype(f) + ((g * g).astype(f) * (f(1) - f(0.999))).astype(f)
        newR = x.R - ((alpha[s] * m).astype(f) / (np.sqrt(v).astype(f) + f(1e-7))).astype(f)
        o.apply_adam(x, grads, dg, alpha[s])   # moves C, rb, cb, g (and R the oracle's way)
        np.testing.assert_array_equal(x.R, newR)
        x.step += 1
    np.testing.assert_array_equal(x.R, ref.R)


def test_lazy_replay_is_bit_identical_to_dense():
    """SURVEY §7 hard part 1: replaying the missed idle steps when a row is next touched == the dense sweep."""
    rng = np.random.default_rng(0)
    f = np.float32
    n, steps = 64, 300
    alpha = o.alpha_table(0.001, steps)
    x0, m0, v0 = rng.normal(0, .05, n).astype(f), rng.normal(0, 1e-3, n).astype(f), rng.uniform(0, 1e-6, n).astype(f)
    xd, md, vd = x0.copy(), m0.copy(), v0.copy()
    xl, ml, vl, last = x0.copy(), m0.copy(), v0.copy(), np.zeros(n, int)
    for s in range(steps):
        touched = rng.random(n) < 0.1
        g = np.where(touched, rng.normal(0, 1e-4, n), 0).astype(f)
        # dense
        md = (md * f(.9)).astype(f); md[touched] += (g[touched] * f(1 - f(.9))).astype(f)
        vd = (vd * f(.999)).astype(f); vd[touched] += ((g[touched] ** 2).astype(f) * (f(1) - f(.999))).astype(f)
        xd = xd - ((alpha[s] * md).astype(f) / (np.sqrt(vd).astype(f) + f(1e-7))).astype(f)
        # lazy with replay
        for k in np.nonzero(touched)[0]:
            xs, ms, vs = xl[k:k + 1], ml[k:k + 1], vl[k:k + 1]
            for t in range(last[k], s):
                o._adam_untouched_step(xs, ms, vs, alpha[t], f(.9), f(.999), f(1e-7))
            ms[:] = (ms * f(.9)).astype(f) + (g[k] * f(1 - f(.9))).astype(f)
            vs[:] = (vs * f(.999)).astype(f) + ((g[k] ** 2).astype(f) * (f(1) - f(.999))).astype(f)
            xs[:] = xs - ((alpha[s] * ms).astype(f) / (np.sqrt(vs).astype(f) + f(1e-7))).astype(f)
            last[k] = s + 1
    for k in range(n):
        for t in range(last[k], steps):
            o._adam_untouched_step(xl[k:k + 1], ml[k:k + 1], vl[k:k + 1], alpha[t], f(.9), f(.999), f(1e-7))
    np.testing.assert_array_equal(xl, xd)
    np.testing.assert_array_equal(ml, md)
    np.testing.assert_array_equal(vl, vd)


@pytest.mark.parametrize("optimizer,head,mode", [("Adam", "glove", "keras_dense"), ("Adam", "logistic", "lazy"),
                                                 ("Adagrad", "glove", "keras_dense"), ("SGD", "logistic", "keras_dense")])
def test_c_port_matches_numpy_oracle(optimizer, head, mode):
    from oracle import c_oracle
    V, d, B, n = 60, 9, 24, 400
    coo = make_coo(V, n, 6)
    batches = np.random.default_rng(7).integers(0, n, (25, B))
    st = o.init_state(V, d, 8)
    ref = st.copy()
    L = np.array(o.train(ref, coo, batches, optimizer=optimizer, head=head, adam_mode=mode, learning_rate=0.01))
    c = c_oracle.COracle(st.R, st.C, st.rb, st.cb, optimizer=optimizer)
    Lc = c.train(coo, batches, head=head, adam_mode=mode, learning_rate=0.01)
    assert np.max(np.abs(L - Lc) / np.abs(L)) < 2e-6
    assert np.max(np.abs(ref.R - c.R)) < 1e-6 * np.max(np.abs(ref.R)) + 1e-7
    assert abs(float(ref.g) - float(c.g)) < 1e-6


def test_oracle_regression_trajectory():
    z = np.load(os.path.join(GOLD, "oracle_train.npz"))
    coo = {k[4:]: z[k] for k in z.files if k.startswith("coo_")}
    for opt in ("Adam", "Adagrad", "SGD"):
        for head in ("glove", "logistic"):
            st = o.State(z["R0"].copy(), z["C0"].copy(), z["rb0"].copy(), z["cb0"].copy(), np.float32(0))
            L = o.train(st, coo, z["batches"], optimizer=opt, head=head, learning_rate=0.05)
            key = "%s_%s" % (opt, head)
            np.testing.assert_allclose(np.array(L), z[key + "_losses"], rtol=1e-6)
            np.testing.assert_allclose(st.R, z[key + "_R"], rtol=1e-5, atol=1e-7)
            np.testing.assert_allclose(st.cb, z[key + "_cb"], rtol=1e-5, atol=1e-7)


def test_alpha_table_and_engine_copy_agree():
    a = o.alpha_table(0.001, 5000)
    assert abs(a[0] - 0.001 * np.sqrt(1 - 0.999) / (1 - 0.9)) < 1e-7 and abs(a[-1] - 0.001) < 1e-5
    import importlib.util
    spec = importlib.util.find_spec("glove_tensorflow_b200")
    src = open(os.path.join(os.path.dirname(spec.origin), "engine.py")).read()
    ns = {}
    start = src.index("def adam_alpha_table")
    exec("import numpy as np\nADAM_BETA1, ADAM_BETA2 = 0.9, 0.999\n" + src[start:src.index("def _ptr")], ns)
    np.testing.assert_array_equal(ns["adam_alpha_table"](0.001, 5000), a)


@pytest.mark.parametrize("n", [1, 2, 5, 64, 1000, 65537])
def test_feistel_is_a_bijection(n):
    p = o.feistel_permute(np.arange(n), n, 12345)
    assert np.array_equal(np.sort(p), np.arange(n))
    if n > 100:
        q = o.feistel_permute(np.arange(n), n, 12346)
        assert not np.array_equal(p, q)


def test_cosine_topk_ties_break_to_lower_id():
    rng = np.random.default_rng(0)
    T = rng.normal(size=(50, 8)).astype(np.float32)
    T[17] = T[3]            # exact duplicates: identical similarity to everything
    T[40] = 2.0 * T[3]      # same direction, different norm: cosine equal up to rounding
    sim, idx = o.cosine_topk(T, np.array([3, 17]), 5)
    assert idx[0, 0] == 3 and idx[0, 1] == 17 and idx[1, 0] == 3 and idx[1, 1] == 17
    assert np.all(np.diff(sim, axis=1) <= 0)
    emb = o.format_embeddings(T[:3], ["<UNK>", "a", "b"])
    assert list(emb) == ["a", "b"] and emb["a"]["item_id"] == "a" and len(emb["a"]["item_embedding"]) == 8


def test_eval_metrics_definition():
    coo = make_coo(20, 100, 9, zipf=False)
    st = o.init_state(20, 4, 1)
    m = o.eval_metrics(st, coo, 32)
    z = o.logits(st, coo["row"], coo["col"]).astype(np.float64)
    w, y = coo["weight"].astype(np.float64), coo["target"].astype(np.float64)
    assert abs(m["average_loss"] - np.sum(w * (z - y) ** 2) / np.sum(w)) < 1e-9
    assert abs(m["label/mean"] - np.sum(w * y) / np.sum(w)) < 1e-9


def test_cooc_oracle_is_pinned_on_the_reference_preprocessor_output():
    """oracle/cooc_oracle.py against tests/golden/text8_small (written by the reference's src/data/text8.py, unmodified):
    ids and counts exact, float64 columns to 1e-15 (the reference evaluates them through pandas/numexpr)."""
    import os
    import pandas as pd
    from oracle import cooc_oracle
    gold = os.path.join(os.path.dirname(__file__), "golden", "text8_small")
    voc = pd.read_csv(os.path.join(gold, "vocab.csv"), keep_default_na=False)
    tokens = open(os.path.join(gold, "corpus.txt")).read().split()
    ids = cooc_oracle.token_ids(tokens, list(voc["token"]))
    got = cooc_oracle.interaction_table(ids, voc["count"].to_numpy(), 5, 10)
    ref = pd.read_csv(os.path.join(gold, "interaction.csv"), keep_default_na=False)
    ref = ref.sort_values(["row_token_id", "col_token_id"]).reset_index(drop=True)
    assert len(got) == len(ref) == 2746
    for k in ("row_token_id", "col_token_id", "count"):
        assert np.array_equal(got[k].to_numpy(), ref[k].to_numpy()), k
    for k in ("value", "neg_weight", "glove_weight", "glove_value"):
        np.testing.assert_allclose(got[k].to_numpy(), ref[k].to_numpy(), rtol=1e-15, atol=0, err_msg=k)


def test_host_vocabulary_matches_the_reference_preprocessor_output():
    import os
    import pandas as pd
    from glove_tensorflow_b200 import text8
    gold = os.path.join(os.path.dirname(__file__), "golden", "text8_small")
    ref = pd.read_csv(os.path.join(gold, "vocab.csv"), keep_default_na=False)
    tokens = open(os.path.join(gold, "corpus.txt")).read().split()
    got = text8.create_vocabulary(tokens, 60, 0.9)
    assert list(got["token"]) == list(ref["token"]) and list(got["count"]) == list(ref["count"])
    np.testing.assert_allclose(got["proportion"], ref["proportion"], rtol=1e-15)
    ids = text8.token_ids(tokens, got["token"].to_numpy())
    assert ids.dtype == np.int32 and ids.max() == len(got) - 1 and np.bincount(ids)[0] > got["count"][0]   # unknown -> row 0


def test_host_vocabulary_tie_order_matches_the_oracle_restatement():
    """Count ties are where a vocabulary can silently diverge (most_common order, then pandas' sort): random corpora with
    many ties, the product's host function against the oracle's restatement of the reference's lines (which
    tests/golden/pin_cooc_oracle.py pins on the reference itself)."""
    from glove_tensorflow_b200 import text8
    from oracle import cooc_oracle
    rng = np.random.default_rng(3)
    for _ in range(25):
        types, n = int(rng.integers(5, 300)), int(rng.integers(50, 8000))
        p = 1.0 / np.arange(1, types + 1) ** rng.uniform(0.3, 1.5)
        toks = ["t%d" % i for i in rng.choice(types, size=n, p=p / p.sum())]
        vs, cov = int(rng.integers(1, types + 5)), float(rng.uniform(0.3, 0.999))
        a, b = cooc_oracle.vocabulary_frame(toks, vs, cov), text8.create_vocabulary(toks, vs, cov)
        assert list(a["token"]) == list(b["token"]) and list(a["count"]) == list(b["count"])
        np.testing.assert_allclose(a["proportion"], b["proportion"], rtol=1e-15)


# ---- closed-form replay of idle Adam steps (the CUDA path's default adam_mode) vs the sequential recurrence -----------
@pytest.mark.parametrize("scale_v", [1e-12, 1e-8, 1e-4, 1.0])
@pytest.mark.parametrize("ls,gap", [(1, 1), (1, 5), (3, 40), (10, 200), (2000, 1), (2000, 17), (2000, 130), (2000, 1000),
                                    (5, 3000)])
def test_closed_form_replay_is_closer_to_fp64_than_the_fp32_recurrence(scale_v, ls, gap):
    """x after a run of idle steps: |closed form (fp32) - sequential fp64| must not exceed the error of the sequential
    fp32 recurrence (+ 2 ulp of x), for moments from the eps-dominated to the v-dominated regime."""
    rng = np.random.default_rng(ls * 7919 + gap)
    n = 2048
    alpha = o.alpha_table(1e-3, ls + gap + 1)
    x = rng.uniform(-.05, .05, n).astype(np.float32)
    v = (rng.uniform(0.01, 1, n) * scale_v).astype(np.float32)
    m = (rng.normal(0, 1, n) * np.sqrt(scale_v) * rng.uniform(0.05, 1, n)).astype(np.float32)
    x64, m64, v64 = o.idle_run_sequential(x, m, v, alpha, ls, ls + gap, np.float64)
    x32, m32, v32 = o.idle_run_sequential(x, m, v, alpha, ls, ls + gap, np.float32)
    xc, mc, vc = o.idle_run_closed_form(x, m, v, alpha, ls, ls + gap)
    err_c, err_s = np.max(np.abs(xc - x64)), np.max(np.abs(x32 - x64))
    ulp = float(np.spacing(np.float32(np.max(np.abs(x64)))))
    assert err_c <= err_s + 2 * ulp, (err_c, err_s, ulp)
    assert err_c <= 1e-5 * np.max(np.abs(x64))            # far inside the north-star tolerance on its own
    big = np.abs(m64) > 1e-30
    if big.any():
        assert np.max(np.abs(mc[big] - m64[big]) / np.abs(m64[big])) < 2e-6
    assert np.max(np.abs(vc - v64) / np.abs(v64)) < 2e-6


def test_closed_form_replay_zero_state_and_padding():
    """never-updated elements (m = v = 0) and a zero gap are exact no-ops"""
    alpha = o.alpha_table(1e-3, 64)
    x = np.linspace(-0.05, 0.05, 16).astype(np.float32)
    z = np.zeros_like(x)
    xc, mc, vc = o.idle_run_closed_form(x, z, z, alpha, 3, 50)
    assert np.array_equal(xc, x) and not mc.any() and not vc.any()
    xc, mc, vc = o.idle_run_closed_form(x, x, np.abs(x), alpha, 7, 7)
    assert np.array_equal(xc, x) and np.array_equal(mc, x)


def test_fp64_shadow_of_the_c_oracle():
    """libglove_oracle64.so is the same source in double: it must agree with the fp32 build to fp32 noise and with a
    float64 NumPy evaluation of the first step's loss."""
    from oracle import c_oracle
    V, d, B, steps = 300, 16, 64, 25
    coo = make_coo(V, 4 * B, 3)
    batches = np.random.default_rng(4).integers(0, 4 * B, (steps, B))
    st = o.init_state(V, d, 5)
    c32 = c_oracle.COracle(st.R, st.C, st.rb, st.cb)
    c64 = c_oracle.COracle(st.R, st.C, st.rb, st.cb, dtype=np.float64)
    l32 = c32.train(coo, batches, learning_rate=0.01)
    l64 = c64.train(coo, batches, learning_rate=0.01)
    assert c64.R.dtype == np.float64 and l64.dtype == np.float64
    assert np.max(np.abs(l32 - l64) / np.abs(l64)) < 2e-6
    assert np.max(np.abs(c32.R - c64.R)) / np.max(np.abs(c64.R)) < 5e-5
    b = {k: v[batches[0]] for k, v in coo.items()}
    z = (np.sum(st.R[b["row"]].astype(np.float64) * st.C[b["col"]], -1) + st.rb[b["row"]] + st.cb[b["col"]])
    data = np.sum(b["weight"].astype(np.float64) * (z - b["target"]) ** 2) / B
    reg = 2.0 * (0.01 / d * (np.sum(st.R[b["row"]].astype(np.float64) ** 2) + np.sum(st.C[b["col"]].astype(np.float64) ** 2)) / B
                 + 0.01 * (np.sum(st.rb[b["row"]].astype(np.float64) ** 2) + np.sum(st.cb[b["col"]].astype(np.float64) ** 2)) / B)
    assert abs(l64[0] - (data + reg)) < 1e-6 * abs(l64[0])      # l2 = 0.01f, not 0.01: fp32-valued hyper-parameters


def test_oracle_matches_tf_crosscheck():
    """Consumes tests/golden/tf_crosscheck.npz -- per-step losses and variables dumped from the UNMODIFIED reference
    model_fn under TensorFlow 2.11 by tests/golden/tf_crosscheck.py (which can only run where TensorFlow imports).
    While the file is absent the train-path oracle stays "parity unpinned" and this test skips; once it is committed
    the oracle (keras_dense Adam, whichever reg_scale TF turns out to use) must reproduce it to 1e-5."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_crosscheck.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/tf_crosscheck.npz not generated yet (needs tensorflow==2.11.0; see tests/golden/tf_crosscheck.py)")
    z = np.load(path, allow_pickle=False)
    V, d, steps = int(z["in/V"]), int(z["in/d"]), int(z["in/steps"])
    coo = {"row": z["in/row"].reshape(-1).astype(np.int32), "col": z["in/col"].reshape(-1).astype(np.int32),
           "target": z["in/target"].reshape(-1), "weight": z["in/weight"].reshape(-1)}
    B = int(z["in/B"])
    batches = np.arange(steps * B).reshape(steps, B)
    names = [str(n) for n in z["variable_names"]]
    pick = lambda s, *parts: z["step%02d/%s" % (s, [n for n in names if all(q in n for q in parts)
                                                    and not any(t in n for t in ("Adam", "/m", "/v", "accumulator"))][0])]
    ok = {}
    for reg_scale in (1.0, 2.0):
        st = o.State(z["in/R"].copy(), z["in/C"].copy(), z["in/rb"].copy(), z["in/cb"].copy(), np.float32(z["in/g"]))
        losses = np.array(o.train(st, coo, batches, optimizer=str(z["in/optimizer"]), learning_rate=float(z["in/learning_rate"]),
                                  reg_scale=reg_scale, adam_mode="keras_dense"))
        e_loss = float(np.max(np.abs(losses - z["losses"]) / np.abs(z["losses"])))
        e_R = float(np.max(np.abs(st.R - pick(steps - 1, "row_embedding", "embeddings"))) / np.max(np.abs(st.R)))
        e_C = float(np.max(np.abs(st.C - pick(steps - 1, "col_embedding", "embeddings"))) / np.max(np.abs(st.C)))
        e_b = float(np.max(np.abs(st.rb - pick(steps - 1, "row_bias", "embeddings").reshape(-1))) / np.max(np.abs(st.rb)))
        ok[reg_scale] = max(e_loss, e_R, e_C, e_b)
    assert min(ok.values()) < 1e-5, ok
    assert ok[2.0] < 1e-5, ("TensorFlow counts the activity losses ONCE: switch the default reg_scale to 1", ok)
