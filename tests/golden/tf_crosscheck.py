"""TF cross-check of the train-path oracle (SURVEY 8c last row / Appendix B items 1-3).  RUNS ONLY WHERE TENSORFLOW IMPORTS.

The arithmetic of the reference's TRAIN path lives in tensorflow==2.11.0 / keras==2.11.0 / tensorflow-estimator==2.11.0
(requirements.txt of the reference), which cannot be installed in the build container (no index; Python 3.12 vs the
pins' < 3.11).  Until this script has run somewhere, oracle/glove_oracle.py is "parity unpinned" for A3-A6.  It drives
the UNMODIFIED ``src/models/estimator.py:model_fn`` [ref src/models/estimator.py:13-56] in a TF1 graph (what
tf.estimator does under the hood) with

* a deterministic feed instead of make_csv_dataset (explicit batches, token strings looked up through the reference's own
  StaticHashTable),
* the four Embedding variables + global_bias overwritten with known values after initialisation,

runs N TRAIN steps and writes per-step loss, every variable and every optimizer slot after every step to an .npz that
``tests/test_oracle.py::test_oracle_matches_tf_crosscheck`` consumes (it skips while the file is absent).  What this
settles: reg_scale in {1, 2} (does get_losses_for(None) + get_losses_for(features) count the activity losses twice under
TF 2.11?), the legacy-Keras Adam sparse apply (dense decay of m, v on untouched rows), the Adam epsilon (1e-7) and the
RegressionHead loss reduction (SUM_OVER_BATCH_SIZE over the batch, weights applied per example).

usage (on a machine with the reference's pinned environment):
    python tests/golden/tf_crosscheck.py --reference /path/to/glove-tensorflow --out tests/golden/tf_crosscheck.npz
"""
import argparse
import os
import sys
import tempfile

import numpy as np


def make_problem(V=48, d=8, B=32, steps=10, seed=1234):
    """Small Zipf-ish problem with duplicate ids inside a batch and rows that sit idle for several steps."""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, V + 1)
    p /= p.sum()
    row = rng.choice(V, (steps, B), p=p).astype(np.int64)
    col = rng.choice(V, (steps, B), p=p).astype(np.int64)
    # rows V-4 .. V-1 are touched in step 0 only and then again in the last step: 8 idle steps of dense Adam decay
    row[0, :4] = np.arange(V - 4, V)
    row[-1, :4] = np.arange(V - 4, V)
    row[1:-1][row[1:-1] >= V - 4] = 0
    target = rng.normal(1.5, 1.0, (steps, B)).astype(np.float32)
    weight = rng.uniform(0.05, 1.0, (steps, B)).astype(np.float32)
    init = {
        "R": rng.uniform(-0.05, 0.05, (V, d)).astype(np.float32),
        "C": rng.uniform(-0.05, 0.05, (V, d)).astype(np.float32),
        "rb": rng.uniform(-0.05, 0.05, (V,)).astype(np.float32),
        "cb": rng.uniform(-0.05, 0.05, (V,)).astype(np.float32),
        "g": np.float32(0.1),
    }
    return dict(V=V, d=d, B=B, steps=steps, row=row, col=col, target=target, weight=weight, **init)


def run_reference(reference, prob, optimizer="Adam", learning_rate=0.01, l2_reg=0.01):
    sys.path.insert(0, reference)
    import tensorflow as tf
    from src.models.estimator import model_fn                      # the UNMODIFIED reference model_fn

    tf1 = tf.compat.v1
    V, d, steps = prob["V"], prob["d"], prob["steps"]
    vocab = ["tok%03d" % i for i in range(V)]
    tmp = tempfile.mkdtemp()
    vocab_txt = os.path.join(tmp, "vocab.txt")
    with open(vocab_txt, "w") as f:
        f.write("\n".join(vocab))                                  # V lines, no trailing newline: file_lines() == V
    params = {"row_name": "row_token", "col_name": "col_token", "target_name": "glove_value",
              "weight_name": "glove_weight", "vocab_txt": vocab_txt, "embedding_size": d, "l2_reg": l2_reg,
              "optimizer": optimizer, "learning_rate": learning_rate, "top_k": 5}
    out = {"tf_version": tf.__version__, "vocab": np.array(vocab)}
    with tf.Graph().as_default():
        tf1.set_random_seed(0)
        feats = {"row_token": tf1.placeholder(tf.string, [None]), "col_token": tf1.placeholder(tf.string, [None]),
                 "glove_weight": tf1.placeholder(tf.float32, [None])}
        labels = {"glove_value": tf1.placeholder(tf.float32, [None])}
        spec = model_fn(feats, labels, tf.estimator.ModeKeys.TRAIN, params)
        gvars = tf1.global_variables()

        def find(*parts):
            hits = [v for v in gvars if all(q in v.name for q in parts) and "Adam" not in v.name and "/m" not in v.name
                    and "/v" not in v.name and "accumulator" not in v.name]
            assert len(hits) == 1, (parts, [v.name for v in gvars])
            return hits[0]

        var = {"R": find("row_embedding", "embeddings"), "C": find("col_embedding", "embeddings"),
               "rb": find("row_bias", "embeddings"), "cb": find("col_bias", "embeddings"), "g": find("global_bias")}
        with tf1.Session() as sess:
            sess.run([tf1.global_variables_initializer(), tf1.tables_initializer()])
            var["R"].load(prob["R"], sess)
            var["C"].load(prob["C"], sess)
            var["rb"].load(prob["rb"].reshape(V, 1), sess)
            var["cb"].load(prob["cb"].reshape(V, 1), sess)
            var["g"].load(prob["g"], sess)
            losses = []
            for s in range(steps):
                feed = {feats["row_token"]: [vocab[i] for i in prob["row"][s]],
                        feats["col_token"]: [vocab[i] for i in prob["col"][s]],
                        feats["glove_weight"]: prob["weight"][s], labels["glove_value"]: prob["target"][s]}
                loss, _ = sess.run([spec.loss, spec.train_op], feed)
                losses.append(np.float32(loss))
                vals = sess.run(gvars)
                for v, a in zip(gvars, vals):
                    out["step%02d/%s" % (s, v.name)] = np.asarray(a)
            out["losses"] = np.array(losses, np.float32)
            out["variable_names"] = np.array([v.name for v in gvars])
    return out


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "tf_crosscheck.npz"))
    ap.add_argument("--optimizer", default="Adam")
    ap.add_argument("--learning-rate", type=float, default=0.01)
    args = ap.parse_args()
    try:
        import tensorflow  # noqa: F401
    except ImportError as e:
        sys.exit("tf_crosscheck: tensorflow is not importable here (%s); run this where the reference's pinned "
                 "environment (tensorflow==2.11.0) is installed" % e)
    prob = make_problem()
    out = run_reference(args.reference, prob, args.optimizer, args.learning_rate)
    out.update({"in/" + k: np.asarray(v) for k, v in prob.items()})
    out["in/optimizer"] = np.array(args.optimizer)
    out["in/learning_rate"] = np.float32(args.learning_rate)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, "losses", out["losses"])


if __name__ == "__main__":
    main()
