"""Host-side drop-in boundary (no GPU): CLI flags/defaults, params.json, vocab copy, csv ingest, embeddings.json."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "text8_small")


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


def test_csv_ingest_resolves_tokens_like_the_hash_table(tmp_path):
    import pandas as pd
    from glove_tensorflow_b200 import data_utils
    csv, voc = os.path.join(GOLD, "interaction.csv"), os.path.join(GOLD, "vocab.txt")
    df = pd.read_csv(csv, keep_default_na=False)
    coo = data_utils.load_interaction_csv(csv, voc, cache=False)
    assert np.array_equal(coo["row"], df["row_token_id"]) and np.array_equal(coo["col"], df["col_token_id"])
    np.testing.assert_allclose(coo["glove_value"], df["glove_value"].astype(np.float32))
    by_id = data_utils.load_interaction_csv(csv, voc, "row_token_id", "col_token_id", ("value", "neg_weight"), cache=False)
    assert np.array_equal(by_id["row"], coo["row"]) and by_id["value"].dtype == np.float32
    assert data_utils.file_lines(voc) == len(data_utils.read_vocab(voc)) == 61
    # out-of-vocabulary -> 0, the StaticHashTable default (ref model_utils.py:121-127)
    assert list(data_utils.lookup_ids(["w5", "not-a-token", "nan"], data_utils.read_vocab(voc))) == \
        [data_utils.read_vocab(voc).index("w5"), 0, data_utils.read_vocab(voc).index("nan")]
    # sidecar cache round trip
    c2 = tmp_path / "i.csv"
    c2.write_text(open(csv).read())
    first = data_utils.load_interaction_csv(str(c2), voc)
    again = data_utils.load_interaction_csv(str(c2), voc)
    assert os.path.exists(str(c2) + ".coo.npz") and np.array_equal(first["row"], again["row"])


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
