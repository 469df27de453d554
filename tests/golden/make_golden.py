"""Generates the committed golden fixtures.  Run ONCE in the build container (needs /root/reference, which does not
exist on the GPU box):   PYTHONHASHSEED=0 python tests/golden/make_golden.py

1. text8_small/: output of the reference's OWN preprocessor (src/data/text8.py:process_data + save_data, imported
   unmodified from /root/reference) on a small seeded synthetic corpus.  Pins the on-disk schema (column order,
   vocab.txt without trailing newline), glove_weight / glove_value / neg_weight and the symmetric count>=10 filter.
2. readme_kat.json: the sample rows printed in the reference README.md:48-59 (the only known-answer vectors the
   reference publishes).
3. oracle_train.npz: a short seeded trajectory of the oracle itself (regression pin for the checker; it is NOT
   reference output -- the trainer arithmetic lives in un-vendored TensorFlow 2.11: parity unpinned).
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def reference_preprocess():
    if os.environ.get("PYTHONHASHSEED") != "0":
        raise SystemExit("run with PYTHONHASHSEED=0 (row order comes from Python hash(), src/data/text8.py:118-123)")
    out = os.path.join(HERE, "text8_small")
    os.makedirs(out, exist_ok=True)
    scratch = tempfile.mkdtemp()
    for name in ("src", "configs"):
        os.symlink(os.path.join(REF, name), os.path.join(scratch, name))
    cwd = os.getcwd()
    os.chdir(scratch)  # the reference logger opens main.log in cwd at import (src/utils/logger.py:103)
    sys.path.insert(0, scratch)
    from src.data import text8
    rng = np.random.default_rng(0)
    types = 120
    p = 1.0 / np.arange(1, types + 1)
    p /= p.sum()
    tokens = rng.choice(types, size=40000, p=p)
    words = ["na", "null", "nan", "none"] + ["w%d" % i for i in range(4, types)]  # pandas NA-like tokens on purpose
    text = " ".join(words[t] for t in tokens)
    data = text8.process_data(text, vocab_size=60, coverage=0.9, context_size=5)
    text8.save_data(data, out)
    os.chdir(cwd)
    with open(os.path.join(out, "corpus.txt"), "w") as f:
        f.write(text)
    print("text8_small:", data["vocabulary"].shape, data["interaction"].shape)


def readme_kat():
    rows = [  # row_token_id, col_token_id, count, value, row_token, col_token, neg_weight, glove_weight, glove_value
        [6125, 38, 24, 16.9500, "altogether", "not", 0.6421, 0.3428, 2.83027],
        [18, 1571, 176, 74.1000, "was", "prominent", 7.5889, 1.0000, 4.30542],
        [91, 372, 19, 5.4500, "th", "society", 3.1999, 0.2877, 1.69562],
        [432, 541, 12, 5.9000, "numbers", "note", 0.6461, 0.2038, 1.77495],
        [1304, 285, 25, 11.1667, "na", "europe", 0.4112, 0.3535, 2.41293],
        [32, 18, 2312, 723.2000, "be", "was", 406.5180, 1.0000, 6.58369],
        [2247, 1154, 136, 46.5833, "html", "www", 0.0740, 1.0000, 3.84124],
        [710, 229, 18, 9.0500, "cannot", "point", 0.8569, 0.2763, 2.20276],
        [467, 3756, 12, 5.2000, "style", "width", 0.0911, 0.2038, 1.64866],
        [80, 543, 35, 20.6333, "over", "lost", 2.6989, 0.4550, 3.02691],
    ]
    with open(os.path.join(HERE, "readme_kat.json"), "w") as f:
        json.dump({"source": "reference README.md:48-59", "columns": ["row_token_id", "col_token_id", "count", "value",
                   "row_token", "col_token", "neg_weight", "glove_weight", "glove_value"], "rows": rows}, f, indent=1)


def oracle_trajectory():
    sys.path.insert(0, ROOT)
    from oracle import glove_oracle as o
    V, d, B, n, steps = 40, 6, 12, 300, 25
    rng = np.random.default_rng(7)
    coo = {"row": rng.integers(0, V, n).astype(np.int32), "col": rng.integers(0, V, n).astype(np.int32),
           "target": rng.normal(2, 1, n).astype(np.float32), "weight": rng.uniform(0.1, 1, n).astype(np.float32),
           "pos": rng.uniform(1, 30, n).astype(np.float32), "neg": rng.uniform(0, 5, n).astype(np.float32)}
    batches = rng.integers(0, n, (steps, B))
    st0 = o.init_state(V, d, 11)
    out = {"V": V, "d": d, "B": B, "batches": batches, "R0": st0.R, "C0": st0.C, "rb0": st0.rb, "cb0": st0.cb}
    out.update({"coo_" + k: v for k, v in coo.items()})
    for opt in ("Adam", "Adagrad", "SGD"):
        for head in ("glove", "logistic"):
            st = st0.copy()
            losses = o.train(st, coo, batches, optimizer=opt, head=head, learning_rate=0.05)
            key = "%s_%s" % (opt, head)
            out[key + "_losses"] = np.array(losses, np.float32)
            out[key + "_R"] = st.R
            out[key + "_cb"] = st.cb
            out[key + "_g"] = np.float32(st.g)
    np.savez_compressed(os.path.join(HERE, "oracle_train.npz"), **out)


if __name__ == "__main__":
    readme_kat()
    oracle_trajectory()
    if os.path.isdir(REF):
        reference_preprocess()
