import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def make_coo(V, n, seed=0, zipf=True, hot=None):
    """Synthetic co-occurrence COO with the reference schema columns (ids, value, glove_value, glove_weight,
    neg_weight).  ``hot`` forces a fraction of the rows onto id 0 (heavy-hitter segments)."""
    rng = np.random.default_rng(seed)
    if zipf:
        p = 1.0 / np.arange(1, V + 1)
        p /= p.sum()
        row = rng.choice(V, size=n, p=p)
        col = rng.choice(V, size=n, p=p)
    else:
        row = rng.integers(0, V, n)
        col = rng.integers(0, V, n)
    if hot:
        m = rng.random(n) < hot
        row[m] = 0
        m = rng.random(n) < hot
        col[m] = 1 % V
    count = 10 + np.floor(rng.pareto(1.0, n)).clip(0, 1e6)
    value = count * rng.uniform(0.3, 0.6, n)
    return {
        "row": row.astype(np.int32), "col": col.astype(np.int32),
        "target": np.log(value).astype(np.float32),
        "weight": np.clip((count / 100.0) ** 0.75, 0, 1).astype(np.float32),
        "pos": value.astype(np.float32),
        "neg": (count * rng.uniform(0.0, 0.5, n)).astype(np.float32),
    }
