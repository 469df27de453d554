"""Measures the co-occurrence preprocessing on one B200 at text8 scale (17,005,207 tokens, 10,001
vocabulary rows, context 5 = 85 M pairs) beside the CPU restatement of the reference's pandas algorithm on a bounded
sample.

    python tools/bench_cooc.py [--tokens 17005207] [--vocab 10001] > gpurun_out/cooc.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=17_005_207)
    ap.add_argument("--vocab", type=int, default=10_001)
    ap.add_argument("--context", type=int, default=5)
    ap.add_argument("--cpu-tokens", type=int, default=1_000_000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from glove_tensorflow_b200 import text8
    from oracle import cooc_oracle
    rng = np.random.default_rng(0)
    T, V, K = args.tokens, args.vocab, args.context
    p = 1.0 / np.arange(1, V + 1)
    cdf = np.cumsum(p / p.sum())
    ids = np.searchsorted(cdf, rng.random(T)).clip(0, V - 1).astype(np.int32)
    vc = np.bincount(ids, minlength=V).astype(np.int64)
    text8.cooccurrence_table(ids[:100000], vc, K, 10)            # warm-up (context, allocator)
    times = []
    for _ in range(args.reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = text8.cooccurrence_table(ids, vc, K, 10, as_numpy=False)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    gpu_s = float(np.median(times))
    n_out = int(out["count"].numel())
    pairs = sum(int((ids[:-k] != ids[k:]).sum()) for k in range(1, K + 1))
    # CPU: the reference's groupby algorithm restated (oracle), bounded sample of the same stream
    n_cpu = min(args.cpu_tokens, T)
    t0 = time.perf_counter()
    ref = cooc_oracle.interaction_table(ids[:n_cpu], vc, K, 10)
    cpu_s = time.perf_counter() - t0
    small = text8.cooccurrence_table(ids[:n_cpu], vc, K, 10)
    order = np.lexsort((small["col_token_id"], small["row_token_id"]))
    ok = (np.array_equal(small["count"][order], ref["count"].to_numpy())
          and np.allclose(small["value"][order], ref["value"].to_numpy(), rtol=1e-14, atol=0))
    # algorithmic bytes: read the ids once, write the table once
    alg = 4 * T + n_out * (4 + 4 + 8 + 4 * 8)
    print(json.dumps({
        "workload": "co-occurrence table, %d tokens, V=%d, context %d (%d pairs emitted, %d records with count >= 10)" % (T, V, K, pairs, n_out),
        "parity_on_cpu_sample": bool(ok),
        "gpu": {"s": gpu_s, "tokens_per_s": T / gpu_s, "pairs_per_s": pairs / gpu_s, "includes": "H2D of the ids, chunk/merge/finish kernels, all syncs"},
        "cpu_port": {"s": cpu_s, "tokens": n_cpu, "tokens_per_s": n_cpu / cpu_s, "kind": "port (pandas groupby restatement of ref src/data/text8.py:84-139)"},
        "algorithmic_bytes": alg, "sorted_bytes_per_pair": 9}))


if __name__ == "__main__":
    main()
