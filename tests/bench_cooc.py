"""Measures the co-occurrence preprocessing on one B200 at text8 scale (17,005,207 tokens, 10,001
vocabulary rows, context 5 = 85 M pairs) beside the CPU restatement of the reference's pandas algorithm on a bounded
sample.

    python tests/bench_cooc.py [--tokens 17005207] [--vocab 10001] > gpurun_out/cooc.json

Lives under tests/ (not collected by pytest) because it times the oracle as the CPU baseline, and only tests/, smoke() and
bench.py's cpu_baseline leg may touch oracle/.
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
    # ---- whole preprocessor from corpus TEXT (SURVEY cfg1 shape: Zipf over 253,854 types "w<id>"), stage by stage
    types = 253_854
    pt = 1.0 / np.arange(1, types + 1)
    tid = np.searchsorted(np.cumsum(pt / pt.sum()), rng.random(T)).clip(0, types - 1)
    names = np.array(["w%d" % i for i in range(types)], dtype=object)
    text = " ".join(names[tid].tolist())
    del tid
    stages = {}
    text8.process_data(text[:2_000_000], 10000, 0.9, K)         # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    corpus = text8.DeviceCorpus(text)
    torch.cuda.synchronize(); stages["encode + H2D + tokenise"] = time.perf_counter() - t0; t1 = time.perf_counter()
    toks, cnts = corpus.distinct()
    torch.cuda.synchronize(); stages["distinct tokens (sort by hash, verify, D2H, decode %d strings)" % len(toks)] = time.perf_counter() - t1; t1 = time.perf_counter()
    dfv = text8._frame_from_counts(toks, cnts, 10000, 0.9)
    stages["vocabulary frame (host, pandas)"] = time.perf_counter() - t1; t1 = time.perf_counter()
    dids = corpus.ids(dfv["token"].to_numpy())
    torch.cuda.synchronize(); stages["token -> id"] = time.perf_counter() - t1; t1 = time.perf_counter()
    tab = text8.cooccurrence_table(dids, dfv["count"].to_numpy(), K, 10, as_numpy=False)
    torch.cuda.synchronize(); stages["co-occurrence table"] = time.perf_counter() - t1
    text_s = time.perf_counter() - t0
    # host equivalents of the front half on a sample (what the reference does in Python: split, Counter, dict lookups)
    sample = text[:len(text) // 16]
    t1 = time.perf_counter()
    stoks = sample.split()
    from collections import Counter
    Counter(stoks)
    text8.token_ids(stoks, dfv["token"].to_numpy())
    host_front_s = time.perf_counter() - t1
    # algorithmic bytes: read the ids once, write the table once
    alg = 4 * T + n_out * (4 + 4 + 8 + 4 * 8)
    print(json.dumps({
        "workload": "co-occurrence table, %d tokens, V=%d, context %d (%d pairs emitted, %d records with count >= 10)" % (T, V, K, pairs, n_out),
        "parity_on_cpu_sample": bool(ok),
        "gpu": {"s": gpu_s, "tokens_per_s": T / gpu_s, "pairs_per_s": pairs / gpu_s, "includes": "H2D of the ids, chunk/merge/finish kernels, all syncs"},
        "cpu_port": {"s": cpu_s, "tokens": n_cpu, "tokens_per_s": n_cpu / cpu_s, "kind": "port (pandas groupby restatement of ref src/data/text8.py:84-139)"},
        "from_text": {"s": text_s, "bytes": len(text), "tokens_per_s": T / text_s, "records": int(tab["count"].numel()),
                      "vocab_rows": int(len(dfv)), "stages_s": stages},
        "host_front_half_sample": {"s": host_front_s, "tokens": len(stoks), "tokens_per_s": len(stoks) / host_front_s,
                                   "what": "str.split + Counter + dict lookups (ref text8.py:47,62,85-86)"},
        "algorithmic_bytes": alg, "sorted_bytes_per_pair": 9}))


if __name__ == "__main__":
    main()
