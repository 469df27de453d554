"""Input side of the drop-in: interaction.csv + vocab.txt -> COO arrays.

Replaces ``get_csv_input_fn`` (tf.data ``make_csv_dataset`` with select_columns, ref src/models/data_utils.py:4-26) and
the per-step ``StaticHashTable`` string->id lookup (ref src/models/model_utils.py:121-127, src/models/estimator.py:26-28):
the csv is parsed ONCE, token strings are resolved to vocab line numbers once (missing -> 0, the table's default), and
the result is uploaded as the device-resident COO triple buffer.  A binary sidecar (<csv>.coo.npz) caches the parse."""
import os

import numpy as np


def read_vocab(vocab_txt):
    """One token per line, line number = id, no trailing newline (ref src/data/text8.py:149-150)."""
    with open(vocab_txt, encoding="utf8") as f:
        return f.read().split("\n")


def file_lines(fname):
    """ref src/models/utils.py:4-9 (counts iterated lines; a missing trailing newline still counts the last line)."""
    i = -1
    with open(fname, encoding="utf8") as f:
        for i, _ in enumerate(f):
            pass
    return i + 1


def lookup_ids(tokens, vocab):
    """string -> id with default 0, like tf.lookup.StaticHashTable(..., default_value=0)."""
    table = {}
    for i, tok in enumerate(vocab):
        table.setdefault(tok, i)  # TextFileInitializer rejects duplicate keys; first occurrence is the safe reading
    return np.fromiter((table.get(t, 0) for t in tokens), dtype=np.int32, count=len(tokens))


def load_interaction_csv(train_csv, vocab_txt, row_name="row_token", col_name="col_token",
                         value_names=("glove_value", "glove_weight"), cache=True):
    """Returns {'row': i32[n], 'col': i32[n], <value_name>: f32[n] ...} in file order.

    ``row_name`` / ``col_name`` may name string columns (resolved through vocab.txt) or integer id columns
    (``row_token_id``: equal by construction).  keep_default_na=False: tokens like 'na' / 'null' / 'nan' are
    ordinary vocabulary words (ref README.md:54)."""
    import pandas as pd

    key = "|".join([row_name, col_name] + list(value_names))
    side = train_csv + ".coo.npz"
    if cache and os.path.exists(side) and os.path.getmtime(side) >= os.path.getmtime(train_csv):
        z = np.load(side, allow_pickle=False)
        if str(z["key"]) == key:
            return {k: z[k] for k in z.files if k != "key"}
    cols = [row_name, col_name] + list(value_names)
    df = pd.read_csv(train_csv, usecols=cols, keep_default_na=False,
                     dtype={row_name: str, col_name: str, **{v: np.float32 for v in value_names}})
    vocab = read_vocab(vocab_txt)
    out = {}
    for name, src in (("row", row_name), ("col", col_name)):
        vals = df[src].to_numpy()
        if src.endswith("_id"):
            ids = vals.astype(np.int64)
            if ids.min() < 0 or ids.max() >= len(vocab):
                raise ValueError("%s out of range [0, %d)" % (src, len(vocab)))
            out[name] = ids.astype(np.int32)
        else:
            out[name] = lookup_ids(vals, vocab)
    for v in value_names:
        out[v] = df[v].to_numpy(np.float32)
    if cache:
        try:
            np.savez(side, key=np.array(key), **out)
        except OSError:
            pass
    return out
