"""ini -> module constants, mirroring the reference's src/config.py:8-54: section chosen by $ENVIRONMENT (default
'dev'), file ``configs/app.ini`` relative to the working directory (falling back to the copy shipped in this repo).
``python -m glove_tensorflow_b200.config KEY`` prints a value, like the reference's Makefile helper (config.py:56-63)."""
import os
import sys
from argparse import ArgumentParser
from configparser import ConfigParser
from pathlib import Path

_REPO_INI = Path(__file__).resolve().parent.parent / "configs" / "app.ini"


def read_config(ini_file="app.ini", environment=None):
    environment = environment or os.environ.get("ENVIRONMENT", "dev")
    parser = ConfigParser()
    found = parser.read([str(Path("configs", ini_file))])
    if not found:
        parser.read([str(_REPO_INI)])
    return parser[environment]


CONFIG = read_config()

JOB_DIR = CONFIG["JOB_DIR"]

# preprocess (ref src/config.py:28-33: no cap on the vocabulary size by default, only the coverage cut-off)
TEXT8_URL = CONFIG.get("TEXT8_URL", fallback="http://mattmahoney.net/dc/text8.zip")
DATA_DIR = CONFIG.get("DATA_DIR", fallback="data")
VOCAB_SIZE = None
COVERAGE = CONFIG.getfloat("COVERAGE", fallback=0.9)
CONTEXT_SIZE = CONFIG.getint("CONTEXT_SIZE", fallback=5)
TRAIN_CSV = CONFIG["TRAIN_CSV"]
VOCAB_TXT = CONFIG["VOCAB_TXT"]
EMBEDDINGS_JSON = CONFIG["EMBEDDINGS_JSON"]

ROW_NAME = CONFIG["ROW_NAME"]
COL_NAME = CONFIG["COL_NAME"]
TARGET_NAME = CONFIG["TARGET_NAME"]
WEIGHT_NAME = CONFIG["WEIGHT_NAME"]
POS_NAME = CONFIG["POS_NAME"]
NEG_NAME = CONFIG["NEG_NAME"]
STRING_IDX = CONFIG.getint("STRING_IDX", fallback=None)   # absent from app.ini in the reference too (src/config.py:42-43)
NAME_IDX = CONFIG.getint("NAME_IDX", fallback=None)

EMBEDDING_SIZE = CONFIG.getint("EMBEDDING_SIZE")
L2_REG = CONFIG.getfloat("L2_REG")
NEG_FACTOR = CONFIG.getfloat("NEG_FACTOR")
OPTIMIZER = CONFIG["OPTIMIZER"]
LEARNING_RATE = CONFIG["LEARNING_RATE"]  # a string, exactly like the reference (config.py:50); argparse makes it float
BATCH_SIZE = CONFIG.getint("BATCH_SIZE")
TRAIN_STEPS = CONFIG.getint("TRAIN_STEPS")
STEPS_PER_EPOCH = CONFIG.getint("STEPS_PER_EPOCH")
TOP_K = CONFIG.getint("TOP_K")

if __name__ == "__main__":
    ap = ArgumentParser()
    ap.add_argument("key", help="key name to get value")
    sys.stdout.write(CONFIG[ap.parse_args().key])
    sys.stdout.flush()
