"""End-to-end drop-in path on the GPU: the trainer CLI on a csv written by the reference's own preprocessor
(tests/golden/text8_small), checkpoint/resume, embeddings.json export, predictions -- checked against the oracle fed with the
same initial tables and the same (keyed-shuffle) batch order."""
import json
import os

import numpy as np
import pytest

from oracle import glove_oracle as o

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "text8_small")


def _oracle_run(params, steps, head="glove"):
    """Replays what the CLI does: same uniform init (read back from the engine), same Feistel batch order."""
    from glove_tensorflow_b200 import data_utils
    names = (params["target_name"], params["weight_name"]) if head == "glove" else (params["pos_name"], params["neg_name"])
    coo = data_utils.load_interaction_csv(params["train_csv"], params["vocab_txt"], params["row_name"], params["col_name"],
                                          names, cache=False)
    n, B = len(coo["row"]), params["batch_size"]
    pos = np.arange(steps * B)
    idx = np.empty(steps * B, np.int64)
    for e in range(int(pos[-1] // n) + 1):
        m = pos // n == e
        idx[m] = o.feistel_permute(pos[m] % n, n, (params["seed"] + e) & 0xFFFFFFFF)
    ocoo = {"row": coo["row"], "col": coo["col"]}
    if head == "glove":
        ocoo.update(target=coo[names[0]], weight=coo[names[1]])
    else:
        ocoo.update(pos=coo[names[0]], neg=coo[names[1]])
    return ocoo, idx.reshape(steps, B)


def test_cli_train_resume_export_matches_oracle(tmp_path):
    from glove_tensorflow_b200 import config_utils, estimator, export_embeddings, train_utils
    job = str(tmp_path / "job")
    argv = ["--train-csv", os.path.join(GOLD, "interaction.csv"), "--vocab-txt", os.path.join(GOLD, "vocab.txt"),
            "--job-dir", job, "--disable-datetime-path", "--embedding-size", "16", "--batch-size", "256",
            "--learning-rate", "0.01", "--top-k", "5", "--seed", "3", "--plan-steps", "4"]
    # 1) train 10 steps, 2) resume from the checkpoint and continue to 25 steps (max_steps semantics)
    params = config_utils.parse_args(argv + ["--train-steps", "10"])
    eng, hist = estimator.train(params)
    init = None
    assert hist and hist[-1][0] == 10 and os.path.exists(train_utils.checkpoint_path(job, 10))
    params = config_utils.parse_args(argv + ["--train-steps", "25"])
    eng2, hist2 = estimator.train(params)
    assert eng2.host_step == 25 and hist2[-1][0] == 25
    got = eng2.get_state()

    # oracle with the same initial tables (a fresh engine with the same seed reproduces them) and batch order
    from glove_tensorflow_b200.engine import GloveEngine
    fresh = GloveEngine(61, 16, batch_size=256, plan_steps=4, max_steps=8)
    fresh.init_uniform(3)
    s0 = fresh.get_state()
    st = o.State(s0["R"].copy(), s0["C"].copy(), s0["rb"].copy(), s0["cb"].copy(), np.float32(0))
    ocoo, batches = _oracle_run(params, 25)
    o.train(st, ocoo, batches, learning_rate=0.01)
    for k in ("R", "C", "rb", "cb"):
        err = np.max(np.abs(got[k] - getattr(st, k))) / np.max(np.abs(getattr(st, k)))
        assert err < 1e-5, (k, err)
    want = o.eval_metrics(st, ocoo, 256)
    assert abs(hist2[-1][1]["loss"] - want["loss"]) <= 1e-4 * want["loss"]

    # TensorBoard summaries: the reference's tags [ref src/models/model_utils.py:113-118] in <job_dir>, EVAL metrics in <job_dir>/eval
    import glob
    from glove_tensorflow_b200 import summary
    ev = [e for f in sorted(glob.glob(os.path.join(job, "events.out.tfevents.*"))) for e in summary.read_events(f)]
    last = [e for e in ev if e["step"] == 25][-1]
    assert set(last["scalars"]) == {"loss", "global_step/sec", "mf/global_bias"}
    assert abs(last["scalars"]["mf/global_bias"] - float(got["g"])) < 1e-7
    h = last["histograms"]["mf/row_biases"]
    assert h["num"] == 61 and abs(h["sum"] - float(got["rb"].astype(np.float64).sum())) < 1e-5 and "mf/col_biases" in last["histograms"]
    ee = [e for f in glob.glob(os.path.join(job, "eval", "events.out.tfevents.*")) for e in summary.read_events(f)]
    assert any(e["step"] == 25 and abs(e["scalars"]["loss"] - hist2[-1][1]["loss"]) < 1e-6 * abs(hist2[-1][1]["loss"]) for e in ee)

    # export: same json the reference writes (row table, <UNK> skipped, indent 2)
    out = export_embeddings.main(job, str(tmp_path / "embeddings.json"))
    emb = json.load(open(out))
    vocab = open(os.path.join(GOLD, "vocab.txt")).read().split("\n")
    assert "<UNK>" not in emb and set(emb) == set(vocab) - {"<UNK>"}
    i = vocab.index("nan")
    np.testing.assert_allclose(emb["nan"]["item_embedding"], st.R[i], rtol=0, atol=1e-5 * np.max(np.abs(st.R)))
    assert json.load(open(os.path.join(job, "params.json")))["vocab_txt"] == os.path.join(job, "vocab.txt")

    # predictions (PREDICT mode keys of the reference)
    pred = estimator.estimator_predict(params, np.array([i, 1, 2], np.int32))
    assert set(pred) == {"input_string", "input_embedding", "top_k_similarity", "top_k_string"}
    assert pred["input_string"][0] == "nan" and pred["top_k_string"][0][0] == "nan"      # nearest neighbour of a row is itself
    sim, idx = o.cosine_topk(got["R"], np.array([i, 1, 2]), 5)
    assert [vocab[j] for j in idx[0]] == pred["top_k_string"][0]


def test_logistic_cli_runs_and_matches_oracle(tmp_path):
    from glove_tensorflow_b200 import config_utils, estimator, logistic_matrix_factorisation
    job = str(tmp_path / "job")
    argv = ["--train-csv", os.path.join(GOLD, "interaction.csv"), "--vocab-txt", os.path.join(GOLD, "vocab.txt"),
            "--job-dir", job, "--disable-datetime-path", "--embedding-size", "8", "--batch-size", "128",
            "--learning-rate", "0.01", "--seed", "5", "--train-steps", "12", "--neg-factor", "0.5", "--optimizer", "Adagrad"]
    logistic_matrix_factorisation.main(argv)
    saved = json.load(open(os.path.join(job, "params.json")))
    assert saved["input_fn_args"]["select_columns"] == ["row_token", "col_token", "value", "neg_weight"]
    assert saved["input_fn_args"]["target_names"] == []
    from glove_tensorflow_b200 import train_utils
    z = np.load(train_utils.latest_checkpoint(job))
    from glove_tensorflow_b200.engine import GloveEngine
    fresh = GloveEngine(61, 8, optimizer="Adagrad", batch_size=128, plan_steps=4, max_steps=8)
    fresh.init_uniform(5)
    s0 = fresh.get_state()
    st = o.State(s0["R"].copy(), s0["C"].copy(), s0["rb"].copy(), s0["cb"].copy(), np.float32(0))
    ocoo, batches = _oracle_run(saved, 12, head="logistic")
    o.train(st, ocoo, batches, optimizer="Adagrad", head="logistic", learning_rate=0.01, neg_factor=0.5)
    assert int(z["step"]) == 12
    assert np.max(np.abs(z["R"] - st.R)) / np.max(np.abs(st.R)) < 1e-5
