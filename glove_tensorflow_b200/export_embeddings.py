"""``python -m glove_tensorflow_b200.export_embeddings`` -- drop-in for src/models/export_embeddings.py:13-64:
reload params.json from the job dir, read the ROW embedding of every vocab token from the latest checkpoint and write
{token: {"item_id": token, "item_embedding": [d floats]}} (indent 2), skipping '<UNK>'."""
import json
import logging
import os
import sys
from argparse import ArgumentParser

import numpy as np

from .config import EMBEDDINGS_JSON, JOB_DIR
from . import data_utils, train_utils

logger = logging.getLogger(__name__)


def format_predictions(tokens, embeddings):
    out = {}
    for tok, emb in zip(tokens, embeddings):
        if tok != "<UNK>":
            out[tok] = {"item_id": tok, "item_embedding": [float(x) for x in emb]}
    logger.info("embedding dict size: %s.", len(out))
    return out


def main(job_dir=JOB_DIR, embeddings_json=EMBEDDINGS_JSON, **kwargs):
    with open(os.path.join(job_dir, "params.json")) as f:
        params = json.load(f)
    ckpt = train_utils.latest_checkpoint(job_dir)
    if ckpt is None:
        raise FileNotFoundError("no checkpoint in %s" % job_dir)
    # the checkpoint holds the flushed (dense-Keras-exact) tables in the reference's layout: the export needs no GPU
    table = np.load(ckpt, allow_pickle=False)["R"]
    vocab = data_utils.read_vocab(params["vocab_txt"])
    embeddings = format_predictions(vocab, table)
    os.makedirs(os.path.dirname(os.path.abspath(embeddings_json)), exist_ok=True)
    with open(embeddings_json, "w") as f:
        json.dump(embeddings, f, indent=2)
    logger.info("json saved: %s.", embeddings_json)
    return embeddings_json


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    ap = ArgumentParser()
    ap.add_argument("--job-dir", default=JOB_DIR, help="job directory (default: %(default)s)")
    ap.add_argument("--embeddings-json", default=EMBEDDINGS_JSON, help="path to the embeddings json (default: %(default)s)")
    args = ap.parse_args()
    logger.info("call: %s.", " ".join(sys.argv))
    try:
        main(**args.__dict__)
    except KeyboardInterrupt:
        pass
    except Exception as e:
        logger.exception(e)
        raise e
