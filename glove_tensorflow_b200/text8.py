"""``python -m glove_tensorflow_b200.text8`` -- drop-in for ``python -m src.data.text8`` (ref src/data/text8.py:154-196):
corpus text -> vocab.csv, vocab.txt, interaction.csv with the reference's schema.

The vocabulary (a few thousand strings) is counted on the host; the pair counting -- the position cross-join, groupby
and symmetrisation that take minutes and tens of GB in pandas -- runs on the GPU (csrc/glove_cooc.cu): token ids are
uploaded once, pairs are emitted / radix-sorted / reduced per chunk of positions, chunks are merged, and the final
kernel writes the reference's numeric columns.  ``cooccurrence_table`` also hands the result over as device tensors, so a
trainer can skip the csv altogether."""
import ctypes
import logging
import os
import sys
from argparse import ArgumentParser
from collections import Counter

import numpy as np

logger = logging.getLogger(__name__)

TEXT8_URL = "http://mattmahoney.net/dc/text8.zip"      # ref src/config.py
DATA_DIR, VOCAB_SIZE, COVERAGE, CONTEXT_SIZE = "data", 10000, 0.9, 5


def download_data(url=TEXT8_URL, dest_dir=DATA_DIR):
    """Fetch and unzip the corpus unless ``<dest_dir>/text8`` is already there (ref text8.py:18-35)."""
    import shutil
    import urllib.request
    from zipfile import ZipFile
    os.makedirs(dest_dir, exist_ok=True)
    if os.path.exists(os.path.join(dest_dir, "text8")):
        return
    dest = os.path.join(dest_dir, os.path.basename(url))
    if not os.path.exists(dest):
        logger.info("downloading file: %s.", url)
        with urllib.request.urlopen(url) as r, open(dest, "wb") as f:
            shutil.copyfileobj(r, f)
    with ZipFile(dest) as zf:
        zf.extractall(dest_dir)


def load_data(src_dir=DATA_DIR):
    with open(os.path.join(src_dir, "text8")) as f:
        return f.read()


def create_vocabulary(text_tokens, vocab_size=VOCAB_SIZE, coverage=COVERAGE):
    """Frame (token, count, proportion): '<UNK>' plus at most ``vocab_size`` tokens whose count reaches the cut-off at
    which the cumulative token share passes ``coverage``; ordered by count (ref text8.py:61-81)."""
    import pandas as pd
    freq = Counter(text_tokens)
    by_count = np.sort(np.fromiter(freq.values(), np.int64, len(freq)))[::-1]
    total = int(by_count.sum())
    cutoff = by_count[np.searchsorted(np.cumsum(by_count) / total, coverage)]
    logger.info("count cufoff: %s; token coverage: %s.", cutoff, coverage)
    kept = [(tok, n) for tok, n in freq.most_common(vocab_size) if n >= cutoff]
    counts = [n for _, n in kept]
    frame = pd.DataFrame({"token": ["<UNK>"] + [tok for tok, _ in kept], "count": [total - int(np.sum(counts))] + counts})
    frame["proportion"] = frame["count"] / total
    return frame.sort_values("count", ascending=False).reset_index(drop=True)


def token_ids(text_tokens, vocab_tokens):
    """token -> row of the vocabulary frame, unknown -> 0 (ref text8.py:85-86)."""
    lut = {t: i for i, t in enumerate(vocab_tokens)}
    get = lut.get
    return np.fromiter((get(t, 0) for t in text_tokens), np.int32, len(text_tokens))


def cooccurrence_table(ids, vocab_count, context_size=CONTEXT_SIZE, count_minimum=10, device="cuda:0",
                       chunk_positions=32 << 20, order_key=0, as_numpy=True):
    """Symmetric co-occurrence table of an id stream on the GPU.  Returns the numeric columns of interaction.csv
    (row_token_id, col_token_id int32; count int64; value, neg_weight, glove_weight, glove_value float64)."""
    import torch
    from . import _lib
    lib, check = _lib.lib, _lib.check
    dev = torch.device(device)
    V, T = len(vocab_count), len(ids)
    if T < 2:
        raise ValueError("need at least two tokens")
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream().cuda_stream
        d_ids = torch.as_tensor(np.array(ids, np.int32)).to(dev)
        d_vc = torch.as_tensor(np.array(vocab_count, np.int64)).to(dev)
        total = int(np.asarray(vocab_count, np.int64).sum())
        chunk = int(min(chunk_positions, T, ((1 << 31) - 1) // context_size))
        n_items = chunk * context_size
        ws = torch.empty(lib.glove_cooc_workspace_bytes(n_items), dtype=torch.uint8, device=dev)
        acc_k = acc_a = None
        n_acc = 0
        got = ctypes.c_int64()
        for p0 in range(0, T, chunk):
            n_pos = min(chunk, T - p0)
            n_tok = min(n_pos + context_size, T - p0)
            cap = n_pos * context_size
            k = torch.empty(cap, dtype=torch.int64, device=dev)
            a = torch.empty(2 * cap, dtype=torch.int64, device=dev)
            check(lib.glove_cooc_chunk(d_ids[p0:].data_ptr(), n_pos, n_tok, V, context_size, ws.data_ptr(), ws.numel(),
                                       k.data_ptr(), a.data_ptr(), cap, ctypes.byref(got), st), "glove_cooc_chunk")
            n = got.value
            k, a = k[:n].clone(), a[:2 * n].clone()
            if acc_k is None:
                acc_k, acc_a, n_acc = k, a, n
                continue
            if n == 0:
                continue
            m = n_acc + n
            if lib.glove_cooc_workspace_bytes(m) > ws.numel():
                ws = torch.empty(lib.glove_cooc_workspace_bytes(m), dtype=torch.uint8, device=dev)
            ok_, oa = torch.empty(m, dtype=torch.int64, device=dev), torch.empty(2 * m, dtype=torch.int64, device=dev)
            check(lib.glove_cooc_merge(acc_k.data_ptr(), acc_a.data_ptr(), n_acc, k.data_ptr(), a.data_ptr(), n, V,
                                       ws.data_ptr(), ws.numel(), ok_.data_ptr(), oa.data_ptr(), m, ctypes.byref(got), st),
                  "glove_cooc_merge")
            n_acc = got.value
            acc_k, acc_a = ok_[:n_acc].clone(), oa[:2 * n_acc].clone()
        names = ("row_token_id", "col_token_id", "count", "value", "neg_weight", "glove_weight", "glove_value")
        dts = (torch.int32, torch.int32, torch.int64, torch.float64, torch.float64, torch.float64, torch.float64)
        if n_acc == 0:
            out = [torch.empty(0, dtype=dt, device=dev) for dt in dts]
        else:
            if lib.glove_cooc_workspace_bytes(2 * n_acc) > ws.numel():
                ws = torch.empty(lib.glove_cooc_workspace_bytes(2 * n_acc), dtype=torch.uint8, device=dev)
            cap = 2 * n_acc
            out = [torch.empty(cap, dtype=dt, device=dev) for dt in dts]
            check(lib.glove_cooc_finish(acc_k.data_ptr(), acc_a.data_ptr(), n_acc, V, context_size, count_minimum,
                                        d_vc.data_ptr(), total, order_key, ws.data_ptr(), ws.numel(),
                                        *[t.data_ptr() for t in out], cap, ctypes.byref(got), st), "glove_cooc_finish")
            out = [t[:got.value] for t in out]
    if as_numpy:
        return {n: t.cpu().numpy() for n, t in zip(names, out)}
    return dict(zip(names, out))


def create_interaction_dataframe(text_tokens, df_vocab, context_size=CONTEXT_SIZE, count_minimum=10, device="cuda:0"):
    """interaction frame in the reference's column order, GloVe columns included (ref text8.py:84-139: the reference builds
    it in two steps, create_interaction_dataframe + create_glove_dataframe; the GPU kernel writes all columns at once)."""
    import pandas as pd
    vocab_tokens = df_vocab["token"].to_numpy()
    ids = token_ids(text_tokens, vocab_tokens)
    t = cooccurrence_table(ids, df_vocab["count"].to_numpy(), context_size, count_minimum, device)
    frame = pd.DataFrame({"row_token_id": t["row_token_id"], "col_token_id": t["col_token_id"], "count": t["count"],
                          "value": t["value"], "row_token": vocab_tokens[t["row_token_id"]],
                          "col_token": vocab_tokens[t["col_token_id"]], "neg_weight": t["neg_weight"],
                          "glove_weight": t["glove_weight"], "glove_value": t["glove_value"]})
    logger.info("dataframe shape: %s.", frame.shape)
    return frame


def process_data(text8, vocab_size=VOCAB_SIZE, coverage=COVERAGE, context_size=CONTEXT_SIZE, device="cuda:0"):
    tokens = text8.split()
    df_vocab = create_vocabulary(tokens, int(vocab_size), coverage)
    logger.info("vocab created, size: %s.", df_vocab.shape[0])
    return {"vocabulary": df_vocab, "interaction": create_interaction_dataframe(tokens, df_vocab, context_size, device=device)}


def save_data(data, save_dir=DATA_DIR):
    """vocab.csv, vocab.txt (one token per line, no trailing newline), interaction.csv (ref text8.py:142-156)."""
    os.makedirs(save_dir, exist_ok=True)
    data["vocabulary"].to_csv(os.path.join(save_dir, "vocab.csv"), index=False)
    with open(os.path.join(save_dir, "vocab.txt"), "w") as f:
        f.write("\n".join(data["vocabulary"]["token"]))
    data["interaction"].to_csv(os.path.join(save_dir, "interaction.csv"), index=False)
    return data


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(name)s - %(message)s")
    ap = ArgumentParser(description="Download, extract and prepare text8 data.")
    ap.add_argument("--url", default=TEXT8_URL, help="url of text8 data (default: %(default)s)")
    ap.add_argument("--dest", default=DATA_DIR, help="destination directory for downloaded and extracted files (default: %(default)s)")
    ap.add_argument("--vocab-size", type=int, default=VOCAB_SIZE, help="maximum size of vocab (default: %(default)s)")
    ap.add_argument("--coverage", type=float, default=COVERAGE, help="token coverage to set token count cutoff (default: %(default)s)")
    ap.add_argument("--context-size", type=int, default=CONTEXT_SIZE, help="size of context window (default: %(default)s)")
    ap.add_argument("--device", default="cuda:0")
    args = ap.parse_args(argv)
    logger.info("call: %s.", " ".join(sys.argv))
    download_data(args.url, args.dest)
    save_data(process_data(load_data(args.dest), args.vocab_size, args.coverage, args.context_size, args.device), args.dest)


if __name__ == "__main__":
    main()
