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


# ---- row-sharded owner-computes scheme (the default N>1 path: bench.py --dp-mode sharded) ----------------------------
def _shard_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import make_coo
    from glove_tensorflow_b200 import parallel as par
    from oracle import glove_oracle as o
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, d, B, steps, n = 81, 6, 32, 6, 600                  # V not divisible by world: the last rank holds a pad row
    coo = make_coo(V, n, 31)
    batches = np.random.default_rng(32).integers(0, n, (steps, B))
    full = o.init_state(V, d, 33)
    alpha = o.alpha_table(0.01, steps)
    Vl = par.shard_rows(V, world)
    own = np.nonzero(par.shard_owner(np.arange(V), world) == rank)[0]

    def a2a(send):                                         # gloo has no all-to-all: gather every rank's send lists
        box = [None] * world
        dist.all_gather_object(box, send)
        return [box[k][rank] for k in range(world)]

    def local(a):                                          # this rank's rows only, zero pad row at the end
        out = np.zeros((Vl,) + a.shape[1:], np.float32)
        out[:len(own)] = a[own]
        return out
    tabs = {k: local(getattr(full, k)) for k in ("R", "C", "rb", "cb")}
    mom = {k: (np.zeros_like(v), np.zeros_like(v)) for k, v in tabs.items()}
    g, gm, gv = np.float32(full.g), np.float32(0), np.float32(0)
    b1, b2, eps, l2 = np.float32(0.9), np.float32(0.999), np.float32(1e-7), np.float32(0.04)
    Bf, df = np.float32(B), np.float32(d)
    sent_rows = 0
    for s in range(steps):
        gb = {k: v[batches[s]] for k, v in coo.items()}
        snap = {k: v.copy() for k, v in tabs.items()}     # step-start snapshot of the rows held here
        # exchange: for each side, fetch the opposite-side snapshot rows my segments reference from their owners
        fetched = {}
        for ids, opp, emb, bias in ((gb["row"], gb["col"], "C", "cb"), (gb["col"], gb["row"], "R", "rb")):
            req = par.shard_requests(ids, opp, rank, world)
            want = [torch.from_numpy(r.astype(np.int64)) for r in req]
            asked = a2a(want)
            reply = []
            for k, a in enumerate(asked):                  # rows rank k asked me for: they must all be mine
                a = a.numpy()
                assert np.all(par.shard_owner(a, world) == rank)
                la = par.shard_local(a, world)
                reply.append(torch.from_numpy(np.concatenate([snap[emb][la], snap[bias][la, None]], 1)))
                if k != rank:
                    sent_rows += len(a)
            rows = a2a(reply)
            fetched[emb] = {int(i): r.numpy() for w, rr in zip(want, rows) for i, r in zip(w, rr)}
        # forward + gradients of the segments owned here (each triple is visited once per side, on two ranks at most)
        partial = np.zeros(3, np.float32)                  # loss sum, sum e, regulariser -- from the ROW side only
        grads = {}
        for side, (ids, opp, emb, bias, oemb) in enumerate(((gb["row"], gb["col"], "R", "rb", "C"), (gb["col"], gb["row"], "C", "cb", "R"))):
            for b in np.nonzero(par.shard_owner(ids, world) == rank)[0]:
                i, li = int(ids[b]), int(par.shard_local(ids[b], world))
                other = fetched[oemb][int(opp[b])]
                z = np.float32(np.dot(snap[emb][li], other[:d]) + snap[bias][li] + other[d] + g)
                err = np.float32(z - gb["target"][b])
                e = np.float32(2) * gb["weight"][b] * err / Bf
                ge, gbias = grads.setdefault((emb, li), [np.zeros(d, np.float32), np.float32(0)])
                ge += e * other[:d] + (l2 / (df * Bf)) * snap[emb][li]
                grads[(emb, li)][1] = np.float32(gbias + e + (l2 / Bf) * snap[bias][li])
                if side == 0:
                    partial[0] += gb["weight"][b] * err * err / Bf
                    partial[1] += e
        t = torch.from_numpy(partial)
        dist.all_reduce(t)
        a = alpha[s]
        for (emb, li), (ge, gbias) in grads.items():
            bias = "rb" if emb == "R" else "cb"
            for name, grad in ((emb, ge), (bias, gbias)):
                m, v = mom[name]
                m[li] = b1 * m[li] + (np.float32(1) - b1) * grad
                v[li] = b2 * v[li] + (np.float32(1) - b2) * grad * grad
                tabs[name][li] = snap[name][li] - a * m[li] / (np.sqrt(v[li]) + eps)
        # rows not in the batch still decay under dense Adam (legacy OptimizerV2): m, v shrink, x moves
        for name in tabs:
            touched = np.zeros(Vl, bool)
            touched[[li for (e_, li) in grads if e_ == ("R" if name in ("R", "rb") else "C")]] = True
            m, v = mom[name]
            idle = ~touched
            m[idle] = b1 * m[idle]
            v[idle] = b2 * v[idle]
            tabs[name][idle] = tabs[name][idle] - a * m[idle] / (np.sqrt(v[idle]) + eps)
        dg = np.float32(partial[1] + l2 * g)
        gm = b1 * gm + (np.float32(1) - b1) * dg
        gv = b2 * gv + (np.float32(1) - b2) * dg * dg
        g = np.float32(g - a * gm / (np.sqrt(gv) + eps))
    np.savez(os.path.join(tmp, "shard%d.npz" % rank), own=own, g=g, sent=sent_rows, **tabs)
    dist.destroy_process_group()


def test_sharded_owner_computes_equals_single_process_oracle(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import make_coo
    from oracle import glove_oracle as o
    world, port = 2, 30100 + os.getpid() % 500
    mp.spawn(_shard_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    V, d, B, steps, n = 81, 6, 32, 6, 600
    coo = make_coo(V, n, 31)
    batches = np.random.default_rng(32).integers(0, n, (steps, B))
    ref = o.init_state(V, d, 33)
    o.train(ref, coo, batches, learning_rate=0.01)
    parts = [np.load(tmp_path / ("shard%d.npz" % r)) for r in range(world)]
    for k in ("R", "C", "rb", "cb"):
        got = np.zeros_like(getattr(ref, k))
        for p in parts:
            got[p["own"]] = p[k][:len(p["own"])]
        assert np.max(np.abs(got - getattr(ref, k))) <= 1e-5 * np.max(np.abs(getattr(ref, k))), k
    assert abs(float(parts[0]["g"]) - float(ref.g)) < 1e-6 and float(parts[0]["g"]) == float(parts[1]["g"])
    assert all(int(p["sent"]) > 0 for p in parts)          # the exchange really moved remote rows


def _gather_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from glove_tensorflow_b200 import parallel
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, Bl = 3, 5
    B = Bl * world
    # the global chunk every rank must end up with: array j, step k, in-batch position p -> 1000 j + 100 k + p
    full = torch.tensor([[[1000 * j + 100 * k + p for p in range(B)] for k in range(K)] for j in range(4)], dtype=torch.int32)
    mine = full.view(4, K, world, Bl)[:, :, rank, :].contiguous().view(-1)          # this rank's share, [4][K][Bl]
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    got = parallel.assemble_chunk(torch.cat(out), world, K, Bl).reshape(4, K, B)
    assert torch.equal(got, full)
    np.save(os.path.join(tmp, "gather%d.npy" % rank), got.numpy())
    dist.destroy_process_group()


def test_sliced_host_chunks_assemble_to_the_global_batch(tmp_path):
    """train_chunks_from_host(sliced=True): every rank feeds 1/world of each batch; the all-gathered shares must assemble
    to the global chunk in batch order (world_size 2, gloo)."""
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + 7
    mp.spawn(_gather_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "gather0.npy"), np.load(tmp_path / "gather1.npy")
    assert np.array_equal(a, b)


def _round_worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from glove_tensorflow_b200 import parallel
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K, Bl = 3, 5
    B, m = Bl * world, K * Bl
    # chunk q of the round, array j, step k, in-batch position p -> 10000 q + 1000 j + 100 k + p
    full = np.array([[[[10000 * q + 1000 * j + 100 * k + p for p in range(B)] for k in range(K)] for j in range(4)]
                     for q in range(world)], np.int32)
    shares = [[full[q, j].reshape(K, world, Bl)[:, rank, :].reshape(-1) for j in range(4)] for q in range(world)]
    send = torch.from_numpy(parallel.round_send_buffer(shares, world)).reshape(-1)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)                       # rank q receives every rank's share of chunk q
    got = parallel.assemble_chunk(recv, world, K, Bl).reshape(4, K, B)
    assert np.array_equal(got.numpy(), full[rank]), rank
    np.save(os.path.join(tmp, "round%d.npy" % rank), got.numpy())
    dist.destroy_process_group()


def test_shared_plan_round_exchange_hands_every_rank_its_chunk(tmp_path):
    """Shared plan construction, host-fed path: of a round of `world` chunks every rank holds 1/world of each batch; ONE
    all-to-all must leave rank q with the whole of chunk q in batch order (world_size 2, gloo).  Plus the round-robin
    deal itself: consecutive rounds alternate build buffers, a round's chunks have distinct builders."""
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from glove_tensorflow_b200 import parallel
    port = 29500 + (os.getpid() % 2000) + 11
    mp.spawn(_round_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "round0.npy"), np.load(tmp_path / "round1.npy")
    assert a[0, 0, 0] == 0 and b[0, 0, 0] == 10000           # rank 0 got chunk 0, rank 1 chunk 1
    for world in (2, 3, 4, 8):
        slots = [parallel.shared_plan_slot(c, world) for c in range(4 * world)]
        for R in range(4):
            rnd = [s for s in slots if s[0] == R]
            assert sorted(s[1] for s in rnd) == list(range(world)) and {s[2] for s in rnd} == {R & 1}


def _hostfed_worker(rank, world, port, tmp):
    """The engine's own host-fed path with shared plans (train_chunks_from_host(sliced=True) ->
    _train_chunks_shared_plans) on CPU tensors over gloo: streams / events / C entry points are stubs, the tensors, the
    all-to-all and the assembly are real.  What reaches glove_prepare_batches_sharded on rank q in round R must be the
    whole of chunk R * world + q in batch order."""
    sys.path.insert(0, ROOT)
    import contextlib
    import ctypes
    import types
    import torch
    import torch.distributed as dist
    from glove_tensorflow_b200 import engine as E
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Stream:
        cuda_stream = 1
        def wait_event(self, ev): pass
        def wait_stream(self, other): pass

    class Event:
        def __init__(self, *a, **k): pass
        def record(self, stream=None): pass

    E.torch.cuda.Event = Event
    E.torch.cuda.stream = lambda s: contextlib.nullcontext()
    E.torch.cuda.current_stream = lambda *a, **k: Stream()
    K, Bl, rounds = 3, 5, 2
    B, m = Bl * world, K * Bl
    seen = {}

    def prepare(dst, ws, wsb, row, col, ca, cb, nnz, sidx, first_sample, key, first, K_, B_, V, n, stream):
        coo = next(c for c in eng._ring_coo if c[0].data_ptr() == row.value)
        assert [t.data_ptr() for t in coo] == [row.value, col.value, ca.value, cb.value] and nnz == K * B
        seen[first] = [t.clone() for t in coo]
        return 0

    E.lib = types.SimpleNamespace(glove_prepare_batches_sharded=prepare, glove_plan_pull_slice=lambda *a: 0,
                                  glove_shard_train_step=lambda *a: 0)
    E.check = lambda rc, what="": None
    eng = object.__new__(E.GloveEngine)
    eng.dp_rank, eng.dp_world, eng.sharded, eng.K, eng.B, eng.V_global = rank, world, True, K, B, 1000
    eng.max_steps, eng.host_step, eng.device = 10 ** 6, 7, torch.device("cpu")
    eng.optimizer, eng.adam_mode, eng.shard_exchange, eng.overlap = "Adam", "replay", "peer-push", True
    eng._side, eng._prep_stream = Stream(), Stream()
    eng.plans = [torch.zeros(8, dtype=torch.uint8), torch.zeros(8, dtype=torch.uint8)]
    eng.plan_bytes = 64
    eng.plan_first, eng._plan_counts, eng._plan_shards, eng._plan_need = [None, None], [None, None], [None, None], [None, None]
    eng._ev_plan, eng._keep, eng._plan_override, eng._ev_catchup = [None, None], [None, None], None, None
    eng._ev_step_done = [Event(), Event()]
    eng.prep_ws, eng._ev_coo, eng._label = torch.zeros(8, dtype=torch.uint8), None, None
    eng.loss_cap, eng.loss_out = 64, torch.zeros(64)
    eng._args = [ctypes.c_int(0), ctypes.c_int(1)]
    eng._ring = dict(buf=torch.zeros(128, dtype=torch.uint8), hdl=object(), ptrs=[0] * world, barrier=lambda: None, built={},
                     opened=set(), ev_barrier=None, keep={}, build_stream=Stream(), pull_stream=Stream(),
                     builder=eng._ring_build_resident)
    # chunk c, array j, step k, in-batch position p -> an int pattern for the id arrays, a float pattern for the value arrays
    def full(c, j):
        v = np.array([[100000 * c + 10000 * j + 100 * k + p for p in range(B)] for k in range(K)])
        return v.astype(np.int32) if j < 2 else (v / 7.0).astype(np.float32)
    chunks = [tuple(torch.from_numpy(np.ascontiguousarray(full(c, j).reshape(K, world, Bl)[:, rank, :]).reshape(-1)) for j in range(4))
              for c in range(rounds * world)]
    assert all(t.numel() == m for t in chunks[0])
    eng.train_chunks_from_host(chunks, sliced=True)
    assert eng.host_step == 7 + rounds * world * K
    mine = sorted(seen)                                   # the plans this rank built: chunk R * world + rank of every round
    assert mine == [7 + (R * world + rank) * K for R in range(rounds)], mine
    for R in range(rounds):
        got = seen[7 + (R * world + rank) * K]
        for j in range(4):
            exp = full(R * world + rank, j).reshape(-1)
            assert got[j].dtype == (torch.int32 if j < 2 else torch.float32)
            assert np.array_equal(got[j].numpy().view(np.int32), exp.view(np.int32)), (R, j)
    np.save(os.path.join(tmp, "hostfed%d.npy" % rank), np.array(mine))
    dist.destroy_process_group()


def test_shared_plan_host_fed_path_plans_the_whole_chunk(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + 13
    mp.spawn(_hostfed_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "hostfed0.npy").tolist() == [7, 13] and np.load(tmp_path / "hostfed1.npy").tolist() == [10, 16]


def test_balanced_owner_labels():
    """balance_owners: a permutation of the vocabulary that keeps every owner's row count, keeps id order inside an owner
    and evens out a Zipf head that puts 1.5x the mean work on owner 0 under id % world."""
    sys.path.insert(0, ROOT)
    from glove_tensorflow_b200 import parallel
    V, world, B = 50_003, 8, 8 * 65_536
    rng = np.random.default_rng(3)
    p = 1.0 / np.arange(1, V + 1)
    ids = rng.choice(V, 1_000_000, p=p / p.sum())
    freq = np.bincount(ids, minlength=V).astype(np.float64)
    label, unlabel, rel = parallel.balanced_labels(freq, world, B, len(ids) // 2, hot=4096)
    assert sorted(label.tolist()) == list(range(V)) and np.array_equal(unlabel[label], np.arange(V))
    for o in range(world):
        mine = np.flatnonzero(label % world == o)
        assert len(mine) == len(range(o, V, world))                       # same rows per owner as id % world
        assert np.all(np.diff(label[mine]) > 0)                            # id order kept inside the owner
    naive = np.bincount(np.arange(V) % world, weights=freq, minlength=world)
    bal = np.bincount(label % world, weights=freq, minlength=world)
    assert naive.max() / naive.mean() > 1.3 and bal.max() / bal.mean() < 1.03, (naive / naive.mean(), bal / bal.mean())
    assert rel.max() < 1.02
