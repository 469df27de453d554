"""CUDA csv ingest (csrc/glove_ingest.cu through data_utils.ingest_csv) against the CPU oracle (oracle/csv_oracle.py):
bit-exact ids and float32 values, on the csv written by the reference's own preprocessor (tests/golden/text8_small), on
hand-made records that exercise the RFC-4180 corners, across chunk boundaries, and -- at a size the oracle cannot
follow -- through the write -> ingest round trip."""
import os

import numpy as np
import pytest

from oracle import csv_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "text8_small")
VALS = ("glove_value", "glove_weight")


def _ingest(csv, voc, row="row_token", col="col_token", vals=VALS, **kw):
    from glove_tensorflow_b200 import data_utils
    return {k: v.cpu().numpy() for k, v in data_utils.ingest_csv(str(csv), str(voc), row, col, vals, **kw).items()}


def _same(got, exp):
    assert set(got) == set(exp)
    for k in exp:
        assert got[k].dtype == exp[k].dtype and got[k].shape == exp[k].shape, k
        assert np.array_equal(got[k].view(np.uint32), exp[k].view(np.uint32)), k     # bit-exact, -0.0 / NaN included


def test_golden_preprocessor_csv_bit_exact():
    import pandas as pd
    csv, voc = os.path.join(GOLD, "interaction.csv"), os.path.join(GOLD, "vocab.txt")
    data, vocab = open(csv, "rb").read(), csv_oracle.read_vocab(voc)
    got = _ingest(csv, voc)
    _same(got, csv_oracle.parse_interaction_csv(data, vocab, "row_token", "col_token", VALS))
    df = pd.read_csv(csv, keep_default_na=False)
    assert np.array_equal(got["row"], df["row_token_id"]) and np.array_equal(got["col"], df["col_token_id"])
    by_id = _ingest(csv, voc, "row_token_id", "col_token_id", ("value", "neg_weight"))
    _same(by_id, csv_oracle.parse_interaction_csv(data, vocab, "row_token_id", "col_token_id", ("value", "neg_weight")))
    assert np.array_equal(by_id["row"], got["row"])


def test_rfc4180_corners(tmp_path):
    vocab = [b"<UNK>", b"the", b"na", b"w,x", b'say "hi"', b"two\nlines", "café".encode(), b"the", b"", b"nan"]
    voc = tmp_path / "vocab.txt"
    voc.write_bytes(b"\n".join(vocab))
    vocab = csv_oracle.read_vocab(str(voc))   # 11 lines: "two\nlines" cannot be ONE vocab line (keys are whole lines)
    body = (b"id,row_token,col_token,glove_value,glove_weight,extra\r\n"
            b'1,the,na,1.5,0.25,x\r\n'
            b'\r\n'
            b'2,"w,x","say ""hi""",-2e-3,,"q"\n'
            b'\n\n'
            b'3,"two\nlines",caf\xc3\xa9,+.5,1E2,\n'
            b'4,,unknown-token,-0.0,3.4028236e38,"a""b"\n'
            b'5,nan,"",1e-46,16777217,z\r\n'
            b'6,"the",the,0.1,0.30000000000000004,no-trailing-newline')
    csv = tmp_path / "i.csv"
    csv.write_bytes(body)
    exp = csv_oracle.parse_interaction_csv(body, vocab, "row_token", "col_token", VALS)
    assert list(exp["row"]) == [1, 3, 0, 9, 10, 1] and list(exp["col"]) == [2, 4, 7, 0, 9, 1]   # duplicate line: first wins
    got = _ingest(csv, voc)
    _same(got, exp)
    assert np.isinf(got["glove_weight"][3]) and np.signbit(got["glove_value"][3])
    # same records, final newline present / CRLF only
    csv.write_bytes(body + b"\n")
    _same(_ingest(csv, voc), exp)
    csv.write_bytes(body.replace(b"\r\n", b"\n").replace(b"\n", b"\r\n"))
    exp_crlf = csv_oracle.parse_interaction_csv(body.replace(b"\r\n", b"\n").replace(b"\n", b"\r\n"), vocab, "row_token", "col_token", VALS)
    _same(_ingest(csv, voc), exp_crlf)
    # header only -> empty COO
    csv.write_bytes(b"id,row_token,col_token,glove_value,glove_weight,extra\n")
    assert all(len(v) == 0 for v in _ingest(csv, voc).values())


def test_chunk_boundaries_and_unknown_tokens(tmp_path):
    rng = np.random.default_rng(5)
    V, n = 5000, 40000
    vocab = [b"<UNK>"] + [b"w%d" % i for i in range(1, V)]
    voc = tmp_path / "vocab.txt"
    voc.write_bytes(b"\n".join(vocab))
    rows = [b"row_token_id,col_token_id,count,value,row_token,col_token,neg_weight,glove_weight,glove_value"]
    for i in range(n):
        r, c = int(rng.integers(0, V + 50)), int(rng.integers(0, V))          # ids >= V: tokens missing from the vocab
        v = float(np.exp(rng.normal(0, 3)))
        tok = b"w%d" % r if r else b"<UNK>"
        if i % 97 == 0:
            tok = b'"' + tok + b'"'
        rows.append(b"%d,%d,%d,%s,%s,%s,%s,%s,%s" % (min(r, V - 1), c, i, repr(v).encode(), tok, vocab[c], repr(v / 7).encode(),
                                                       repr(min(1.0, v ** 0.75)).encode(), repr(float(np.log(v))).encode()))
    body = b"\n".join(rows) + b"\n"
    csv = tmp_path / "i.csv"
    csv.write_bytes(body)
    exp = csv_oracle.parse_interaction_csv(body, vocab, "row_token", "col_token", VALS)
    assert (exp["row"] == 0).sum() > 100
    for chunk in (1 << 16, 100003 // 16 * 16 + 65536, 256 << 20):            # many chunks / odd size / one chunk
        _same(_ingest(csv, voc, chunk_bytes=chunk), exp)


def test_malformed_records_fail_loudly(tmp_path):
    from glove_tensorflow_b200._lib import GloveError
    voc = tmp_path / "vocab.txt"
    voc.write_bytes(b"<UNK>\na\nb")
    head = b"row_token,col_token,glove_value,glove_weight\n"
    csv = tmp_path / "i.csv"
    for body, what in ((b"a,b,1,1\na,b,1\n", "record 1: wrong number of fields"),
                       (b"a,b,1,1\na,b,1,1\na,b,x1,1\n", "record 2: not a number"),
                       (b'a,"b,1,1\n', "quot"),
                       (b'a,"b"x,1,1\n', "record 0: malformed quoting"),
                       (b"a,b,1.00000005960464477539062500000,1\n", "19 significant digits")):
        csv.write_bytes(head + body)
        with pytest.raises(GloveError, match=what):
            _ingest(csv, voc)
    csv.write_bytes(b"row_token_id,col_token_id,glove_value,glove_weight\n0,1,1,1\n2,3,1,1\n")
    with pytest.raises(GloveError, match="record 1: id outside"):
        _ingest(csv, voc, "row_token_id", "col_token_id")
    # a key column is an id column when its first 100 values are integers (the rule make_csv_dataset infers dtypes by); a
    # value that is not an integer further down is an error, never a silent 0
    csv.write_bytes(b"row_token_id,col_token_id,glove_value,glove_weight\n" + b"0,1,1,1\n" * 100 + b"2,1x,1,1\n")
    with pytest.raises(GloveError, match="record 100: not an integer"):
        _ingest(csv, voc, "row_token_id", "col_token_id")
    # ... and a column with a non-integer among its first values is a token column whatever it is called: resolved through
    # vocab.txt, unknown tokens -> id 0 (the reference's string_id_table default)
    csv.write_bytes(b"row_token_id,col_token_id,glove_value,glove_weight\nb,a,1,1\na,1x,1,1\n")
    got = _ingest(csv, voc, "row_token_id", "col_token_id")
    assert got["row"].tolist() == [2, 1] and got["col"].tolist() == [1, 0]
    csv.write_bytes(head + b"a,b,1,1\n")
    with pytest.raises(ValueError, match="not in the csv header"):
        _ingest(csv, voc, "row_token", "nope")


def test_round_trip_at_scale(tmp_path):
    """2M records (about 150 MB of text, several chunks): float32 -> shortest repr of its double -> ingest gives the same
    bits back; ids survive the token round trip."""
    import pandas as pd
    rng = np.random.default_rng(11)
    V, n = 400000, 2000000
    voc = tmp_path / "vocab.txt"
    voc.write_bytes(b"\n".join([b"<UNK>"] + [b"w%d" % i for i in range(1, V)]))
    row, col = rng.integers(0, V, n).astype(np.int32), rng.integers(0, V, n).astype(np.int32)
    gv = rng.normal(0, 3, n).astype(np.float32)
    gw = rng.random(n, dtype=np.float32)
    tok = np.array(["<UNK>"] + ["w%d" % i for i in range(1, V)], dtype=object)
    df = pd.DataFrame({"row_token_id": row, "col_token_id": col, "row_token": tok[row], "col_token": tok[col],
                       "glove_weight": gw.astype(np.float64), "glove_value": gv.astype(np.float64)})
    csv = tmp_path / "big.csv"
    df.to_csv(csv, index=False)
    got = _ingest(csv, voc, chunk_bytes=32 << 20)
    assert np.array_equal(got["row"], row) and np.array_equal(got["col"], col)
    assert np.array_equal(got["glove_value"].view(np.uint32), gv.view(np.uint32))
    assert np.array_equal(got["glove_weight"].view(np.uint32), gw.view(np.uint32))
