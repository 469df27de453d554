"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference preprocessor's pair counting, the checker for the CUDA
co-occurrence kernels (glove_tensorflow_b200/csrc/glove_cooc.cu).  Nothing in the product path may import this file.

PINNED: tests/test_oracle.py checks this restatement against tests/golden/text8_small, which is the output of the
reference's own preprocessor (src/data/text8.py imported unmodified, tests/golden/make_golden.py), and
tests/golden/pin_cooc_oracle.py ran it against the reference itself on random corpora (30 vocabularies with count ties, 6
interaction tables with context sizes 1..6: ids, counts and every float64 column equal, max error 0.0).

Follows create_interaction_dataframe [ref src/data/text8.py:84-126] and create_glove_dataframe / glove_weight
[ref src/data/text8.py:129-139], with the position cross-join replaced by shifted slices of the id array (the same pairs
in the same order: distance 1 first, then 2, ...; within a distance by position)."""
import numpy as np
import pandas as pd


def token_ids(tokens, vocab_tokens):
    """token -> row of the vocabulary frame, unknown -> 0 [ref text8.py:85-86]."""
    lut = {t: i for i, t in enumerate(vocab_tokens)}
    return np.fromiter((lut.get(t, 0) for t in tokens), np.int32, len(tokens))


def interaction_table(ids, vocab_count, context_size=5, count_minimum=10):
    """ids int[T], vocab_count int[V] (count column of the vocabulary frame, total = its sum) -> DataFrame with the
    reference's numeric columns, sorted by (row_token_id, col_token_id) (the reference's own order is a salted hash)."""
    ids = np.asarray(ids, np.int64)
    rows, cols, vals = [], [], []
    for k in range(1, context_size + 1):                       # right context only, row != col [ref text8.py:90-95]
        a, b = ids[:-k], ids[k:]
        m = a != b
        rows.append(a[m])
        cols.append(b[m])
        vals.append(np.full(int(m.sum()), 1 / k))
    co = pd.DataFrame({"row_token_id": np.concatenate(rows), "col_token_id": np.concatenate(cols), "value": np.concatenate(vals)})
    agg = co.groupby(["row_token_id", "col_token_id"])["value"].agg(["count", "sum"]).reset_index().rename(columns={"sum": "value"})
    agg = agg[(agg["count"] != 0) & (agg["value"] != 0)]       # [ref text8.py:98-102]
    swapped = agg.rename(columns={"row_token_id": "col_token_id", "col_token_id": "row_token_id"})
    sym = pd.concat([agg, swapped], sort=False).groupby(["row_token_id", "col_token_id"]).sum().reset_index()   # [ref :105-110]
    vc = np.asarray(vocab_count, np.int64)
    total = vc.sum()
    sym["neg_weight"] = vc[sym["row_token_id"]] * (vc[sym["col_token_id"]] / total)    # count_row * proportion_col [ref :113-117]
    sym = sym[sym["count"] >= count_minimum].copy()                                     # [ref :129]
    sym["glove_weight"] = np.clip(np.power(sym["count"] / 100, 0.75), 0, 1)             # [ref :130, :137-139]
    sym["glove_value"] = np.log(sym["value"])                                           # [ref :131]
    return sym.sort_values(["row_token_id", "col_token_id"]).reset_index(drop=True)


def vocabulary_frame(tokens, vocab_size, coverage):
    """create_vocabulary restated [ref src/data/text8.py:61-81]: Counter, coverage cut-off on the count-sorted cumulative
    share, most_common(vocab_size) filtered by the cut-off, '<UNK>' with the remaining mass, ordered by count."""
    from collections import Counter
    freq = Counter(tokens)
    counts = np.sort(list(freq.values()))[::-1]
    total = np.sum(counts)
    cutoff = counts[np.searchsorted(np.cumsum(counts) / total, coverage)]
    vocab = [t for t, n in freq.most_common(vocab_size) if n >= cutoff]
    vc = [freq[t] for t in vocab]
    df = pd.DataFrame({"token": ["<UNK>"] + vocab, "count": [total - np.sum(vc)] + vc})
    df["proportion"] = df["count"] / total
    return df.sort_values("count", ascending=False).reset_index(drop=True)
