"""world_size-2 CPU (gloo) test of the data-parallel scheme the GPU path uses: every rank accumulates gradient partial
sums only for the triples it owns (parallel.dp_owner), dense per-slot buffers are all-reduced, every replica applies the
same update -- and the result equals the single-process oracle step on the global batch."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import make_coo
    from glove_tensorflow_b200 import parallel
    from oracle import glove_oracle as o
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, d, B, steps, n = 80, 6, 32, 6, 600
    coo = make_coo(V, n, 21)
    batches = np.random.default_rng(22).integers(0, n, (steps, B))
    st = o.init_state(V, d, 23)
    alpha = o.alpha_table(0.01, steps)
    for s in range(steps):
        gb = {k: v[batches[s]] for k, v in coo.items()}
        # global plan: slots = sorted unique ids of the GLOBAL batch (same on every rank)
        ur, uc = np.unique(gb["row"]), np.unique(gb["col"])
        loss, z, e = o.forward_loss(st, gb)            # forward over the global batch (each rank needs only its own e)
        mine = parallel.dp_owner(np.arange(B), B, world) == rank
        Bf, df = np.float32(B), np.float32(d)
        ce, cb_ = np.float32(0.04) / (df * Bf), np.float32(0.04) / Bf
        GR = np.zeros((len(ur), d + 1), np.float32)
        GC = np.zeros((len(uc), d + 1), np.float32)
        for b in np.nonzero(mine)[0]:
            i, j = gb["row"][b], gb["col"][b]
            sr, sc = np.searchsorted(ur, i), np.searchsorted(uc, j)
            GR[sr, :d] += e[b] * st.C[j] + ce * st.R[i]
            GR[sr, d] += e[b] + cb_ * st.rb[i]
            GC[sc, :d] += e[b] * st.R[i] + ce * st.C[j]
            GC[sc, d] += e[b] + cb_ * st.cb[j]
        sums = np.array([np.sum(e[mine], dtype=np.float32)], np.float32)
        for buf in (GR, GC, sums):
            t = torch.from_numpy(buf)
            dist.all_reduce(t)
        grads = {"R": (ur, GR[:, :d].copy()), "C": (uc, GC[:, :d].copy()), "rb": (ur, GR[:, d].copy()), "cb": (uc, GC[:, d].copy())}
        dg = np.float32(sums[0] + np.float32(0.04) * st.g)
        o.apply_adam(st, grads, dg, alpha[s])
        st.step += 1
    np.savez(os.path.join(tmp, "rank%d.npz" % rank), R=st.R, C=st.C, rb=st.rb, cb=st.cb, g=st.g)
    dist.destroy_process_group()


def test_dp_allreduce_equals_single_process_oracle(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import make_coo
    from oracle import glove_oracle as o
    world, port = 2, 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    V, d, B, steps, n = 80, 6, 32, 6, 600
    coo = make_coo(V, n, 21)
    batches = np.random.default_rng(22).integers(0, n, (steps, B))
    ref = o.init_state(V, d, 23)
    o.train(ref, coo, batches, learning_rate=0.01)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in ("R", "C", "rb", "cb"):
        assert np.array_equal(r0[k], r1[k])                               # replicas stay bit-identical
        assert np.max(np.abs(r0[k] - getattr(ref, k))) <= 1e-5 * np.max(np.abs(getattr(ref, k)))
    assert abs(float(r0["g"]) - float(ref.g)) < 1e-6
