"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's input side, the checker for the CUDA ingest kernels
(glove_tensorflow_b200/csrc/glove_ingest.cu).  Nothing in the product path may import this file.

PARITY UNPINNED: the reference holds no fixture for its input pipeline; this restates
  * tf.data.experimental.make_csv_dataset(header=True, select_columns=[row, col, weight, target]) as called by
    get_csv_input_fn [ref src/models/data_utils.py:4-26]: RFC-4180 records (quoted fields, "" escapes), blank lines
    skipped, float columns inferred as float32 and converted from the field text with ONE rounding (DecodeCSV ->
    strtof semantics), empty field -> the inferred default 0;
  * tf.lookup.StaticHashTable(TextFileInitializer(vocab.txt, key = whole line, value = line number), default 0)
    [ref src/models/model_utils.py:121-127] applied to the row / col token columns [ref src/models/estimator.py:26-28].
It is pinned against the reference's own preprocessor output (tests/golden/text8_small, written by
[ref src/data/text8.py:97-150]): the *_token_id columns the preprocessor wrote must equal the looked-up ids."""
import csv
import io
import re
from fractions import Fraction

import numpy as np

_NUM = re.compile(rb"^([+-]?)(\d*)(?:\.(\d*))?(?:[eE]([+-]?\d+))?$")


def f32_exact(text: bytes) -> np.float32:
    """Correctly rounded (nearest, ties to even) float32 of a decimal field, via exact rational arithmetic."""
    if text == b"":
        return np.float32(0.0)
    low = text.lower()
    sign = -1.0 if low[:1] == b"-" else 1.0
    body = low[1:] if low[:1] in (b"-", b"+") else low
    if body in (b"inf", b"infinity"):
        return np.float32(sign * np.inf)
    if body == b"nan":
        return np.float32(np.nan)
    m = _NUM.match(text)
    if not m or (m.group(2) == b"" and not m.group(3)):
        raise ValueError("not a number: %r" % text)
    ip, fp, ex = m.group(2) or b"", m.group(3) or b"", int(m.group(4) or 0)
    exact = Fraction(int(ip + fp or b"0"), 1) * Fraction(10) ** (ex - len(fp))
    neg = m.group(1) == b"-"
    if exact == 0:
        return np.float32(-0.0 if neg else 0.0)
    # candidate from the double rounding, then settle against its float32 neighbours exactly
    with np.errstate(over="ignore"):
        try:
            c = np.float32(float(exact))
        except OverflowError:
            c = np.float32(np.inf)
    if np.isinf(c):
        c = np.float32(np.finfo(np.float32).max)
    best = None
    with np.errstate(over="ignore"):
        cands = [np.nextafter(c, np.float32(-np.inf), dtype=np.float32), c, np.nextafter(c, np.float32(np.inf), dtype=np.float32)]
    for v in cands:
        if np.isinf(v):
            continue
        err = abs(Fraction(float(v)) - exact)
        even = (int(np.float32(v).view(np.uint32)) & 1) == 0
        if best is None or err < best[0] or (err == best[0] and even):
            best = (err, even, v)
    v = best[2]
    fmax = Fraction(float(np.finfo(np.float32).max))
    if exact >= fmax + Fraction(2) ** 103:     # beyond the rounding boundary of FLT_MAX (half an ulp = 2^103)
        v = np.float32(np.inf)
    return np.float32(-v if neg else v)


def read_vocab(vocab_txt):
    with open(vocab_txt, "rb") as f:
        return f.read().split(b"\n")


def parse_interaction_csv(data: bytes, vocab, row_name, col_name, value_names):
    """bytes of the whole file (header included) -> {'row','col',value_names...}; ids int32, values float32."""
    table = {}
    for i, tok in enumerate(vocab):
        table.setdefault(tok, i)
    text = data.decode("latin-1")                      # byte-transparent: tokens are compared as bytes
    rows = [r for r in csv.reader(io.StringIO(text, newline="")) if r]
    header = rows[0]
    cols = [header.index(n) for n in (row_name, col_name) + tuple(value_names)]
    out_ids = [[], []]
    out_val = [[], []]
    # a key column holds ids when its first 100 values are integers (make_csv_dataset's num_rows_for_inference=100
    # [ref src/models/data_utils.py:19]), tokens otherwise -- whatever the column is called
    import re
    is_id = [all(len(r) > cols[s] and re.fullmatch(r"[+-]?[0-9]+", r[cols[s]].strip()) for r in rows[1:101]) for s in range(2)]
    for r in rows[1:]:
        if len(r) != len(header):
            raise ValueError("expected %d fields, got %d" % (len(header), len(r)))
        for s, name in enumerate((row_name, col_name)):
            field = r[cols[s]].encode("latin-1")
            if is_id[s]:
                v = int(field)
                if not 0 <= v < len(vocab):
                    raise ValueError("id out of range")
                out_ids[s].append(v)
            else:
                out_ids[s].append(table.get(field, 0))
        for s in range(2):
            out_val[s].append(f32_exact(r[cols[2 + s]].encode("latin-1")))
    out = {"row": np.asarray(out_ids[0], np.int32), "col": np.asarray(out_ids[1], np.int32)}
    for s, n in enumerate(value_names):
        out[n] = np.asarray(out_val[s], np.float32)
    return out
