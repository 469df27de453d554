"""Host-side drop-in boundary (no GPU): CLI flags/defaults, params.json, vocab copy, csv ingest, embeddings.json."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "text8_small")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_flags_and_defaults_match_the_reference():
    from glove_tensorflow_b200 import config_utils
    p = config_utils.build_parser()
    d = vars(p.parse_args([]))
    # ref src/models/config_utils.py:76-180 + configs/app.ini:15-53 ($ENVIRONMENT=dev -> TRAIN_STEPS 1024)
    want = {"train_csv": "data/interaction.csv", "vocab_txt": "data/vocab.txt", "row_name": "row_token",
            "col_name": "col_token", "target_name": "glove_value", "weight_name": "glove_weight", "pos_name": "value",
            "neg_name": "neg_weight", "job_dir": "checkpoints/estimator", "disable_datetime_path": False,
            "embedding_size": 64, "l2_reg": 0.01, "neg_factor": 1.0, "optimizer": "Adam", "learning_rate": 0.001,
            "batch_size": 1024, "steps_per_epoch": 16384, "top_k": 20}
    for k, v in want.items():
        assert d[k] == v, (k, d[k], v)
    assert d["train_steps"] in (1024, 16384, 65536)
    assert isinstance(d["learning_rate"], float)
    a = vars(p.parse_args("--embedding-size 300 --l2-reg 0.1 --neg-factor 2 --optimizer Adagrad --learning-rate 0.05 "
                          "--batch-size 65536 --train-steps 7 --top-k 10".split()))
    assert (a["embedding_size"], a["optimizer"], a["batch_size"], a["train_steps"], a["top_k"]) == (300, "Adagrad", 65536, 7, 10)


def test_init_params_writes_params_json_and_copies_vocab(tmp_path):
    from glove_tensorflow_b200 import config_utils
    job = tmp_path / "job"
    params = config_utils.parse_args(["--train-csv", os.path.join(GOLD, "interaction.csv"), "--vocab-txt",
                                      os.path.join(GOLD, "vocab.txt"), "--job-dir", str(job), "--disable-datetime-path"])
    assert params["vocab_txt"] == str(job / "vocab.txt") and (job / "vocab.txt").read_text() == open(os.path.join(GOLD, "vocab.txt")).read()
    saved = json.load(open(job / "params.json"))
    assert saved["input_fn_args"]["select_columns"] == ["row_token", "col_token", "glove_weight", "glove_value"]
    assert saved["input_fn_args"]["target_names"] == ["glove_value"]
    assert saved["dataset_args"]["weight_names"] == ["glove_weight"]
    assert saved["serving_input_fn_args"] == {"string_features": ["row_token", "col_token"]}
    p2 = config_utils.parse_args(["--vocab-txt", os.path.join(GOLD, "vocab.txt"), "--job-dir", str(tmp_path / "j2")])
    assert p2["job_dir"].startswith(str(tmp_path / "j2") + "-20")    # datetime suffix %Y%m%d-%H%M%S


def test_csv_oracle_resolves_tokens_like_the_reference_preprocessor():
    """Pins the ingest oracle on the reference's own preprocessor output: the *_token_id columns it wrote must be what
    the vocab lookup of the token columns gives (tokens 'na' / 'null' / 'nan' included)."""
    import pandas as pd
    from oracle import csv_oracle
    csv, voc = os.path.join(GOLD, "interaction.csv"), os.path.join(GOLD, "vocab.txt")
    df = pd.read_csv(csv, keep_default_na=False)
    data, vocab = open(csv, "rb").read(), csv_oracle.read_vocab(voc)
    coo = csv_oracle.parse_interaction_csv(data, vocab, "row_token", "col_token", ("glove_value", "glove_weight"))
    assert np.array_equal(coo["row"], df["row_token_id"]) and np.array_equal(coo["col"], df["col_token_id"])
    np.testing.assert_allclose(coo["glove_value"], df["glove_value"].astype(np.float32), rtol=1e-7)
    by_id = csv_oracle.parse_interaction_csv(data, vocab, "row_token_id", "col_token_id", ("value", "neg_weight"))
    assert np.array_equal(by_id["row"], coo["row"]) and by_id["value"].dtype == np.float32
    # out-of-vocabulary -> 0, the StaticHashTable default (ref model_utils.py:121-127); quoting; blank lines; CRLF
    tricky = b'a,row_token,col_token,glove_value,glove_weight\r\n1,w5,not-a-token,1.5,0.25\r\n\r\n2,"nan","w,x",-2e-3,\n'
    t = csv_oracle.parse_interaction_csv(tricky, vocab + [b"w,x"], "row_token", "col_token", ("glove_value", "glove_weight"))
    assert list(t["row"]) == [vocab.index(b"w5"), vocab.index(b"nan")] and list(t["col"]) == [0, len(vocab)]
    assert list(t["glove_value"]) == [np.float32(1.5), np.float32(-2e-3)] and list(t["glove_weight"]) == [0.25, 0.0]


def test_decimal_to_float32_is_correctly_rounded():
    """The host entry of the kernels' decimal -> float32 routine against exact rational rounding (one rounding, like
    DecodeCSV; NOT text -> double -> float)."""
    import ctypes
    import random
    from fractions import Fraction
    from glove_tensorflow_b200._lib import lib
    from oracle.csv_oracle import f32_exact

    def parse(s):
        out = ctypes.c_float()
        rc = lib.glove_parse_float32(s, len(s), ctypes.byref(out))
        return rc, np.float32(out.value)
    rng = random.Random(7)
    cases = [b"0", b"-0", b"1", b"-1.5", b"3.5596246182566738", b"1e38", b"3.4028235e38", b"3.4028236e38", b"1e39", b"1e-45",
             b"7e-46", b"7.1e-46", b"1.17549435e-38", b"1.1754942e-38", b"16777217", b"16777219", b"9007199791611905",
             b"1E+5", b"+2.5e-3", b".5", b"5.", b"1e-64", b"1e-65", b"12345678901234567890", b"8388609.5", b"8388610.5",
             b"inf", b"-Infinity", b"", b"0.000000000000000000000000000000000000000000001"]
    for _ in range(20000):
        cases.append(repr(rng.lognormvariate(0, 8) * rng.choice((1, -1))).encode())        # what pandas writes
        f = np.array([rng.getrandbits(31)], np.uint32).view(np.float32)[0]                   # float32 midpoints
        with np.errstate(invalid="ignore", over="ignore"):
            g = np.nextafter(f, np.float32(np.inf))
        if np.isfinite(f) and np.isfinite(g) and f > 0:
            mid = (Fraction(float(f)) + Fraction(float(g))) / 2
            digits = rng.choice((9, 12, 17, 19))
            e = len(str(mid.numerator // mid.denominator)) - 1 if mid >= 1 else -len(str(mid.denominator // mid.numerator))
            w = mid / Fraction(10) ** (e - digits + 1)
            cases.append(b"%de%d" % (w.numerator // w.denominator + rng.choice((0, 0, 1)), e - digits + 1))
        cases.append(b"%de%d" % (rng.randint(0, 10 ** rng.randint(1, 19) - 1), rng.randint(-70, 45)))
    for s in cases:
        rc, got = parse(s)
        assert rc == 0, s
        assert got.view(np.uint32) == f32_exact(s).view(np.uint32), s
    assert np.isnan(parse(b"nan")[1])
    for s in (b"abc", b"1e", b"--1", b"1.2.3", b"1 ", b" 1", b"e5", b".", b"+", b"infx", b"1e+"):
        assert parse(s)[0] != 0, s
    assert parse(b"1.00000005960464477539062500000")[0] != 0   # > 19 digits AND the tail decides: refused, never guessed


def test_ingest_host_helpers(tmp_path):
    from glove_tensorflow_b200 import _lib, data_utils
    csv, voc = os.path.join(GOLD, "interaction.csv"), os.path.join(GOLD, "vocab.txt")
    names, off = data_utils.read_header(csv)
    assert names[:2] == ["row_token_id", "col_token_id"] and open(csv, "rb").read()[off - 1:off] == b"\n"
    sc = data_utils.make_schema(names, "row_token", "col_token_id", ("glove_value", "glove_weight"))
    assert sc.n_cols == 9 and list(sc.column) == [4, 1, 8, 7] and list(sc.kind) == [_lib.CSV_TOKEN, _lib.CSV_INT, _lib.CSV_FLOAT, _lib.CSV_FLOAT]
    with pytest.raises(ValueError):
        data_utils.make_schema(names, "row_token", "nope", ("glove_value", "glove_weight"))
    # the kind of a key column comes from the data, as in make_csv_dataset: the golden file's id columns hold integers, its
    # token columns strings; a STRING column that happens to be called '*_id' is a token column
    sample = data_utils.read_sample_records(csv, off)
    assert len(sample) == 100 and len(sample[0]) == 9
    sc = data_utils.make_schema(names, "row_token", "col_token_id", ("glove_value", "glove_weight"), sample=sample)
    assert list(sc.kind) == [_lib.CSV_TOKEN, _lib.CSV_INT, _lib.CSV_FLOAT, _lib.CSV_FLOAT]
    # fewer records than the sample size, quoted fields with embedded newlines / commas, blank lines, \r\n
    r = tmp_path / "r.csv"
    r.write_bytes(b'a,b,c\r\n"x,1","y\nz",3\r\n\r\n7,8,9\r\n')
    rn, roff = data_utils.read_header(str(r))
    assert data_utils.read_sample_records(str(r), roff) == [["x,1", "y\nz", "3"], ["7", "8", "9"]]
    # a record cut by the read limit is not part of the sample
    big = tmp_path / "big.csv"
    big.write_bytes(b"a,b\n" + b"12345,67890\n" * 50)
    bn, boff = data_utils.read_header(str(big))
    cut = data_utils.read_sample_records(str(big), boff, n=100, limit=12 * 10 + 5)
    assert cut == [["12345", "67890"]] * 10
    assert [data_utils._int_like(x) for x in ("7", "-3", "+12", " 5 ", "", "1.0", "1e3", "a1", "--1")] == \
        [True, True, True, True, False, False, False, False, False]
    u = tmp_path / "u.csv"
    u.write_text("user_id,item_id,v,w\nalice,7,1.0,1.0\nbob,-3,2.0,1.0\n")
    un, uoff = data_utils.read_header(str(u))
    usc = data_utils.make_schema(un, "user_id", "item_id", ("v", "w"), sample=data_utils.read_sample_records(str(u), uoff))
    assert list(usc.kind)[:2] == [_lib.CSV_TOKEN, _lib.CSV_INT]
    blob, offs = data_utils.vocab_blob(voc)
    vocab = data_utils.read_vocab(voc)
    assert data_utils.file_lines(voc) == len(vocab) == len(offs) - 1 == 61
    assert [blob[offs[i]:offs[i + 1] - 1].decode() for i in range(61)] == vocab
    q = tmp_path / "q.csv"
    q.write_bytes(b'"a ""x""","b\nc",d\r\n1,2,3\n')
    assert data_utils.read_header(str(q)) == (['a "x"', "b\nc", "d"], 19)
    # sidecar cache: a current <csv>.coo.npz with the same column key is returned without touching the GPU
    c2 = tmp_path / "i.csv"
    c2.write_text("x\n")
    # (the key names the columns AND the vocabulary the ids were resolved through: line count, size, mtime)
    vst = os.stat(voc)
    key = "r|c|a|b|vocab:%d:%d:%d" % (61, vst.st_size, vst.st_mtime_ns)
    side = dict(row=np.arange(3, dtype=np.int32), col=np.arange(3, dtype=np.int32), a=np.ones(3, np.float32), b=np.ones(3, np.float32))
    np.savez(str(c2) + ".coo.npz", key=np.array(key), **side)
    got = data_utils.load_interaction_csv(str(c2), voc, "r", "c", ("a", "b"))
    assert list(got["row"]) == [0, 1, 2]
    # cached ids beyond the vocabulary are refused (nothing downstream bounds-checks them) ...
    np.savez(str(c2) + ".coo.npz", key=np.array(key), **dict(side, col=np.array([0, 61, 2], np.int32)))
    with pytest.raises(ValueError, match="out of range"):
        data_utils.load_interaction_csv(str(c2), voc, "r", "c", ("a", "b"))
    # ... and a sidecar written for ANOTHER vocab.txt is not reused (here the re-ingest then fails on the dummy csv)
    np.savez(str(c2) + ".coo.npz", key=np.array("r|c|a|b|vocab:60:1:1"), **side)
    with pytest.raises(ValueError, match="not in the csv header"):
        data_utils.load_interaction_csv(str(c2), voc, "r", "c", ("a", "b"))
    # a vocab.txt that ends with a newline has the same number of lines
    v2 = tmp_path / "vocab_nl.txt"
    v2.write_bytes(open(voc, "rb").read() + b"\n")
    assert data_utils.file_lines(str(v2)) == len(data_utils.read_vocab(str(v2))) == len(data_utils.vocab_blob(str(v2))[1]) - 1 == 61


def test_export_embeddings_format(tmp_path):
    from glove_tensorflow_b200 import export_embeddings, train_utils
    job = tmp_path / "job"
    job.mkdir()
    vocab = ["<UNK>", "the", "na", "of"]
    (job / "vocab.txt").write_text("\n".join(vocab))
    json.dump({"vocab_txt": str(job / "vocab.txt")}, open(job / "params.json", "w"))
    R = np.arange(12, dtype=np.float32).reshape(4, 3) / 7
    np.savez(train_utils.checkpoint_path(str(job), 5), R=R, step=5)
    np.savez(train_utils.checkpoint_path(str(job), 12), R=R + 1, step=12)
    out = export_embeddings.main(str(job), str(tmp_path / "out" / "embeddings.json"))
    emb = json.load(open(out))
    assert list(emb) == ["the", "na", "of"]                       # '<UNK>' skipped (ref export_embeddings.py:22-24)
    assert emb["na"] == {"item_id": "na", "item_embedding": [float(x) for x in (R + 1)[2]]}   # latest checkpoint wins
    assert open(out).read().startswith('{\n  "the": {\n    "item_id"')                        # indent=2


def test_dp_partition_rule():
    from glove_tensorflow_b200 import parallel
    assert list(parallel.dp_owner(np.arange(8), 8, 4)) == [0, 0, 1, 1, 2, 2, 3, 3]
    idx = np.arange(100, 112)
    parts = [parallel.dp_shard(idx, r, 3) for r in range(3)]
    assert np.array_equal(np.concatenate(parts), idx)
    with pytest.raises(ValueError):
        parallel.dp_shard(idx, 0, 5)


def test_decimal_to_float32_round_trips_shortest_representations():
    """Every finite float32 printed in its shortest round-trip form (and in scientific form with 9 significant digits) must
    parse back to the same bits: subnormals, powers of two, the largest finite value, both signs."""
    import ctypes
    from glove_tensorflow_b200._lib import lib
    rng = np.random.default_rng(5)
    bits = np.concatenate([rng.integers(0, 0x7f800000, 60000, dtype=np.uint32),               # all exponents, incl. subnormal
                           rng.integers(0, 0x00800000, 5000, dtype=np.uint32),                 # subnormals
                           np.arange(1, 255, dtype=np.uint32) << 23,                           # powers of two
                           np.array([1, 2, 0x007fffff, 0x00800000, 0x7f7fffff, 0x3f800000, 0x3f7fffff], np.uint32)])
    vals = bits.view(np.float32)
    out = ctypes.c_float()
    for i, v in enumerate(vals):
        forms = [np.format_float_positional(v, unique=True, trim="-") if 1e-5 < v < 1e9 else np.format_float_scientific(v, unique=True),
                 "%.8e" % float(v)]
        if i % 2:
            forms = ["-" + f for f in forms]
        for f in forms:
            s = f.encode()
            assert lib.glove_parse_float32(s, len(s), ctypes.byref(out)) == 0, f
            want = (-v if i % 2 else v)
            assert np.float32(out.value).view(np.uint32) == np.float32(want).view(np.uint32), (f, out.value, want)


def test_preprocess_defaults_are_the_references():
    """ref src/config.py:28-33 + configs/app.ini:25-28 (compared once against the imported reference module: all 27
    upper-case constants of src.config equal ours).  Note VOCAB_SIZE = None: by default only the coverage cut-off limits
    the vocabulary."""
    from glove_tensorflow_b200 import config, text8
    assert config.VOCAB_SIZE is None and config.COVERAGE == 0.9 and config.CONTEXT_SIZE == 5 and config.DATA_DIR == "data"
    assert config.TEXT8_URL == "http://mattmahoney.net/dc/text8.zip" and config.STRING_IDX is None and config.NAME_IDX is None
    assert text8.VOCAB_SIZE is None and text8.CONTEXT_SIZE == 5
    v = text8.create_vocabulary(["a", "b", "a", "c", "b", "a", "d"], None, 0.9)     # no cap: every token above the cut-off
    assert list(v["token"]) == ["a", "b", "c", "d", "<UNK>"] and list(v["count"]) == [3, 2, 1, 1, 0]


def test_tensorboard_event_files_round_trip(tmp_path):
    """summary.EventWriter writes what tf.summary.FileWriter would (TFRecord framing, masked CRC-32C, Event / Summary /
    HistogramProto) with the reference's tags [ref src/models/model_utils.py:113-118]; read_events checks the CRCs and
    decodes it again.  CRC-32C known answers: RFC 3720 B.4."""
    from glove_tensorflow_b200 import summary
    assert summary.crc32c(b"123456789") == 0xE3069283
    assert summary.crc32c(bytes(32)) == 0x8A9136AA and summary.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    w = summary.EventWriter(str(tmp_path))
    rng = np.random.default_rng(0)
    rb = rng.normal(0, 0.05, 1000).astype(np.float32)
    w.add(100, scalars={"loss": 1.25, "global_step/sec": 4321.0, "mf/global_bias": -0.5},
          histograms={"mf/row_biases": rb, "mf/col_biases": np.zeros(7, np.float32)})
    w.add(200, scalars={"loss": 1.0})
    w.close()
    assert os.path.basename(w.path).startswith("events.out.tfevents.")
    ev = summary.read_events(w.path)
    assert ev[0]["file_version"] == "brain.Event:2" and [e["step"] for e in ev] == [0, 100, 200]
    assert ev[1]["scalars"] == {"loss": 1.25, "global_step/sec": 4321.0, "mf/global_bias": -0.5}
    h = ev[1]["histograms"]["mf/row_biases"]
    assert h["num"] == 1000 and abs(h["sum"] - float(rb.astype(np.float64).sum())) < 1e-9
    assert h["min"] == float(rb.min()) and h["max"] == float(rb.max()) and h["bucket"].sum() == 1000
    assert np.all(np.diff(h["bucket_limit"]) > 0) and h["bucket_limit"][-1] >= h["max"]
    inside = (rb[:, None] < h["bucket_limit"][None, :]).argmax(1)           # first limit above each value = its bucket
    assert np.array_equal(np.bincount(inside, minlength=len(h["bucket"])), h["bucket"].astype(np.int64))
    z = ev[1]["histograms"]["mf/col_biases"]
    assert z["num"] == 7 and z["bucket"].sum() == 7 and z["min"] == z["max"] == 0.0
    assert ev[2]["scalars"] == {"loss": 1.0}


def test_bench_reference_arm_line_and_rank_gating():
    """bench.py --impl reference (the arm the driver runs beside the GPU arm): one JSON line with the contract's keys on the
    host cores; under torchrun only rank 0 works and prints, the other ranks exit 0 without output."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "text8", "--steps", "4", "--warmup", "3"]
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "co-occurrence updates/sec" and j["unit"] == "updates/s"
    assert j["higher_is_better"] is True and j["scaling"] == "weak" and j["vs_baseline"] is None and j["dtype"] == "f32"
    assert j["steps"] == 4 and j["warmup"] == 3 and j["value"] > 0 and abs(j["value"] - 65536 / (j["ms_per_step"] * 1e-3)) < 1e-3 * j["value"]
    assert j["config"]["workload"].startswith("text8:") and j["config"]["global_batch"] == 65536
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "TRAIN steps" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # rank 1 of a 2-rank launch: exits 0 without work or output
    out = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"),
                         timeout=60)
    assert out.returncode == 0 and out.stdout.strip() == ""
