"""One-off pin of the preprocessing oracle against the reference ITSELF (needs /root/reference, which exists only in the
build container -- so this is a script, not a collected test):   python tests/golden/pin_cooc_oracle.py

Imports src/data/text8.py unmodified and checks, on random corpora with many count ties and context sizes 1..6, that
  * oracle/cooc_oracle.vocabulary_frame and glove_tensorflow_b200.text8.create_vocabulary give the reference's vocabulary
    (tokens, counts, order -- the pandas sort_values tie order included);
  * oracle/cooc_oracle.interaction_table gives the reference's interaction frame: ids and counts exact, float64 columns
    to 1e-14.
Last run (build container, round 1): 30 / 30 vocabularies equal; 6 / 6 interaction tables equal with max error 0.0."""
import logging
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def main():
    sys.path.insert(0, ROOT)
    from oracle import cooc_oracle
    from glove_tensorflow_b200 import text8 as mine
    scratch = tempfile.mkdtemp()
    for name in ("src", "configs"):
        os.symlink(os.path.join(REF, name), os.path.join(scratch, name))
    os.chdir(scratch)                      # the reference logger opens main.log in cwd at import
    sys.path.insert(0, scratch)
    from src.data import text8 as ref
    logging.disable(logging.CRITICAL)
    rng = np.random.default_rng(0)
    ok_v = 0
    for _ in range(30):
        types, n = int(rng.integers(5, 400)), int(rng.integers(50, 20000))
        p = 1.0 / np.arange(1, types + 1) ** rng.uniform(0.3, 1.5)
        toks = ["t%d" % i for i in rng.choice(types, size=n, p=p / p.sum())]
        vs, cov = int(rng.integers(1, types + 5)), float(rng.uniform(0.3, 0.999))
        a = ref.create_vocabulary(toks, vs, cov)
        ok_v += all(list(a["token"]) == list(x["token"]) and list(a["count"]) == list(x["count"])
                    for x in (cooc_oracle.vocabulary_frame(toks, vs, cov), mine.create_vocabulary(toks, vs, cov)))
    print("vocabularies equal to the reference: %d / 30" % ok_v)
    rng = np.random.default_rng(1)
    ok_t, worst = 0, 0.0
    for _ in range(6):
        types, n, ctx = int(rng.integers(20, 200)), int(rng.integers(5000, 40000)), int(rng.integers(1, 7))
        p = 1.0 / np.arange(1, types + 1)
        toks = ["t%d" % i for i in rng.choice(types, size=n, p=p / p.sum())]
        data = ref.process_data(" ".join(toks), vocab_size=int(types * 0.6), coverage=0.95, context_size=ctx)
        voc = data["vocabulary"]
        r = data["interaction"].sort_values(["row_token_id", "col_token_id"]).reset_index(drop=True)
        m = cooc_oracle.interaction_table(cooc_oracle.token_ids(toks, list(voc["token"])), voc["count"].to_numpy(), ctx, 10)
        same = len(r) == len(m) and all(np.array_equal(r[k].to_numpy(), m[k].to_numpy()) for k in ("row_token_id", "col_token_id", "count"))
        if same:
            worst = max([worst] + [float(np.max(np.abs(r[k].to_numpy() - m[k].to_numpy()))) for k in ("value", "neg_weight", "glove_weight", "glove_value")])
        ok_t += same
    print("interaction tables equal to the reference: %d / 6, max abs float error %.3g" % (ok_t, worst))
    assert ok_v == 30 and ok_t == 6 and worst < 1e-12


if __name__ == "__main__":
    main()
