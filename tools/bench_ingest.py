"""Measures the csv ingest path on one B200: kernel-only (text resident in HBM, CUDA events) and end to end
from the file (page cache -> pinned -> HBM -> COO), beside a CPU parse of the same file (pandas C parser + dict lookup).

    python tools/bench_ingest.py [--rows 10000000] [--vocab 400000] > gpurun_out/ingest.json
"""
import argparse
import ctypes
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--vocab", type=int, default=400_000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import pandas as pd
    import torch
    from glove_tensorflow_b200 import _lib, data_utils
    lib, check = _lib.lib, _lib.check
    rng = np.random.default_rng(3)
    V, n = args.vocab, args.rows
    tmp = tempfile.mkdtemp(prefix="ingest_bench_")
    voc, csv = os.path.join(tmp, "vocab.txt"), os.path.join(tmp, "interaction.csv")
    tok = np.array(["<UNK>"] + ["w%d" % i for i in range(1, V)], dtype=object)
    open(voc, "w").write("\n".join(tok))
    p = 1.0 / np.arange(1, V + 1)
    cdf = np.cumsum(p / p.sum())
    row = np.searchsorted(cdf, rng.random(n)).clip(0, V - 1).astype(np.int32)
    col = np.searchsorted(cdf, rng.random(n)).clip(0, V - 1).astype(np.int32)
    count = 10 + np.floor(1.0 / rng.random(n)).clip(0, 1e6)
    value = count * rng.uniform(0.3, 0.6, n)
    df = pd.DataFrame({"row_token_id": row, "col_token_id": col, "count": count.astype(np.int64), "value": value,
                       "row_token": tok[row], "col_token": tok[col], "neg_weight": value / 7.0,
                       "glove_weight": np.minimum(1.0, (count / 100.0) ** 0.75), "glove_value": np.log(value)})
    df.to_csv(csv, index=False)                      # the schema and float formatting of ref src/data/text8.py:97-139
    size = os.path.getsize(csv)
    gv32, gw32 = df["glove_value"].to_numpy().astype(np.float32), df["glove_weight"].to_numpy().astype(np.float32)
    del df

    # ---- end to end from the file (warm page cache), default chunking
    data_utils.ingest_csv(csv, voc)                  # warm-up: page cache, CUDA context, pinned allocation
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = data_utils.ingest_csv(csv, voc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    ok = (np.array_equal(out["row"].cpu().numpy(), row) and np.array_equal(out["col"].cpu().numpy(), col))
    # text -> float32 in one rounding vs float64 -> float32 of the parsed double: equal except in double-rounding cases
    dbl = int((out["glove_value"].cpu().numpy() != gv32).sum() + (out["glove_weight"].cpu().numpy() != gw32).sum())

    # ---- kernel only: whole text resident in HBM
    dev = torch.device("cuda:0")
    names, off = data_utils.read_header(csv)
    schema = data_utils.make_schema(names, "row_token", "col_token", ("glove_value", "glove_weight"))
    raw = np.fromfile(csv, dtype=np.uint8)[off:]
    text = torch.from_numpy(raw).to(dev)
    blob, voff_h = data_utils.vocab_blob(voc)
    vb, voff = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev), torch.from_numpy(voff_h).to(dev)
    slots = lib.glove_vocab_slots(V)
    table = torch.empty(slots, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ws = torch.empty(lib.glove_csv_workspace_bytes(text.numel()), dtype=torch.uint8, device=dev)
    cap = n + 1
    ends = torch.empty(cap, dtype=torch.int64, device=dev)
    o = [torch.empty(cap, dtype=dt, device=dev) for dt in (torch.int32, torch.int32, torch.float32, torch.float32)]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_build, t_index, t_parse = [], [], []
    for _ in range(args.reps + 1):
        ev[0].record()
        check(lib.glove_vocab_build(table.data_ptr(), slots, vb.data_ptr(), voff.data_ptr(), V, st))
        ev[1].record()
        nrec = ctypes.c_int64()
        check(lib.glove_csv_index(text.data_ptr(), text.numel(), ws.data_ptr(), ws.numel(), ctypes.byref(nrec), st))
        ev[2].record()
        nrows, cons = ctypes.c_int64(), ctypes.c_int64()
        check(lib.glove_csv_parse(text.data_ptr(), text.numel(), 1, ws.data_ptr(), ws.numel(), ctypes.byref(schema),
                                  table.data_ptr(), slots, vb.data_ptr(), voff.data_ptr(), V, ends.data_ptr(),
                                  o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), cap, 0,
                                  ctypes.byref(nrows), ctypes.byref(cons), st))
        ev[3].record()
        torch.cuda.synchronize()
        t_build.append(ev[0].elapsed_time(ev[1])); t_index.append(ev[1].elapsed_time(ev[2])); t_parse.append(ev[2].elapsed_time(ev[3]))
    assert nrows.value == n
    k_ms = float(np.median(t_index[1:]) + np.median(t_parse[1:]))

    # ---- CPU: pandas C parser + python dict lookup on the same file (what a host-side ingest costs)
    t0 = time.perf_counter()
    d2 = pd.read_csv(csv, usecols=["row_token", "col_token", "glove_value", "glove_weight"], keep_default_na=False,
                     dtype={"row_token": str, "col_token": str, "glove_value": np.float32, "glove_weight": np.float32})
    lut = {t: i for i, t in enumerate(tok)}
    r2 = np.fromiter((lut.get(t, 0) for t in d2["row_token"]), np.int32, n)
    c2 = np.fromiter((lut.get(t, 0) for t in d2["col_token"]), np.int32, n)
    cpu_s = time.perf_counter() - t0
    ok = ok and np.array_equal(r2, row) and np.array_equal(c2, col)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except OSError:
        pass
    alg = text.numel() + 16 * n                      # read the text once, write the COO once
    print(json.dumps({
        "workload": "csv ingest, %d records, %d bytes (%.1f B/record), V=%d, reference preprocessor schema" % (n, size, size / n, V),
        "parity": {"ids_equal": bool(ok), "float_bits_differing_from_f64_cast": dbl},
        "kernel_only": {"ms": k_ms, "index_ms": float(np.median(t_index[1:])), "parse_ms": float(np.median(t_parse[1:])),
                        "vocab_build_ms": float(np.median(t_build[1:])), "records_per_s": n / k_ms * 1e3,
                        "text_GB_per_s": text.numel() / k_ms / 1e6, "algorithmic_GB_per_s": alg / k_ms / 1e6},
        "e2e_from_file": {"s": e2e_s, "records_per_s": n / e2e_s, "text_GB_per_s": size / e2e_s / 1e9},
        "cpu_pandas": {"s": cpu_s, "records_per_s": n / cpu_s, "text_GB_per_s": size / cpu_s / 1e9},
        "peaks": peaks}))
    for f in (csv, voc):
        os.remove(f)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
