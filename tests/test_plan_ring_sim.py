"""Shared plan construction (GloveEngine.enable_plan_sharing): the REAL host logic of the engine -- _plan_for,
_prefetch_plan, _ring_build / _ring_open / _ring_pull / _ring_reset, batch_counts, the row-sharded step -- driven on the
CPU against a simulated set of GPUs (test infrastructure; no CUDA, no kernels).

torch.cuda streams / events and the C entry points are replaced by recorders: every rank's calls are enqueued on fake
streams exactly as the engine issues them, and a scheduler then executes the queued operations of all ranks in random
(seeded) and in deliberately lopsided orders, honouring only what the real system honours: stream order, event waits,
the symmetric-memory barrier (a collective of the ranks' k-th calls) and the per-step device-side synchronisation of the
row-sharded step.  Checked while executing:
  * a slice pull finds, in the builder's buffer, the complete plan of exactly the chunk it was issued for (never a plan
    that a later build has started to overwrite, never one that is not built yet);
  * a step (and a host read of a plan) finds the plan of ITS chunk in the local plan buffer it uses;
  * nothing deadlocks: every queue drains, i.e. all ranks issued matching barrier sequences.
"""
import contextlib
import ctypes
import itertools
import os
import random
import sys
import types
from collections import deque

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Sim:
    def __init__(self, world):
        self.world = world
        self.streams = []
        self.by_id = {}
        self.records = itertools.count(1)
        self.done = set()
        self.mem = {}              # device address -> first step of the complete plan it holds
        self.current = None        # engine whose host code is running
        self.stack = []            # torch.cuda.stream(...) contexts
        self.executed = {"build": 0, "pull": 0, "step": 0, "read": 0, "barrier": 0, "a2a": 0}

    def current_stream(self):
        return self.stack[-1] if self.stack else self.current._main


class FakeStream:
    ids = itertools.count(7000)

    def __init__(self, sim, rank, kind):
        self.sim, self.rank, self.kind, self.ops = sim, rank, kind, deque()
        self.cuda_stream = next(FakeStream.ids)
        sim.streams.append(self)
        sim.by_id[self.cuda_stream] = self

    def wait_event(self, ev):
        self.ops.append(("wait", ev.last))          # CUDA semantics: the most recent record at the time of the call

    def wait_stream(self, other):
        rid = next(self.sim.records)
        other.ops.append(("record", rid))
        self.ops.append(("wait", rid))


class FakeBuf:
    def __init__(self, ptr, n=1 << 20):
        self.ptr, self.n = ptr, n

    def data_ptr(self):
        return self.ptr

    def numel(self):
        return self.n


def make_world(monkeypatch, world, K, first_step, max_steps):
    from glove_tensorflow_b200 import engine as E
    sim = Sim(world)

    class FakeEvent:
        def __init__(self, *a, **k):
            self.last = None

        def record(self, stream=None):
            stream = stream if stream is not None else sim.current_stream()
            self.last = next(sim.records)
            stream.ops.append(("record", self.last))

    @contextlib.contextmanager
    def stream_ctx(s):
        sim.stack.append(s)
        try:
            yield
        finally:
            sim.stack.pop()

    monkeypatch.setattr(E.torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(E.torch.cuda, "stream", stream_ctx)
    monkeypatch.setattr(E.torch.cuda, "current_stream", lambda *a, **k: sim.current_stream())

    PLAN = 1000

    def local_ptr(rank, which):
        return (rank + 1) * 10 ** 9 + 10 ** 6 + which

    def on(stream_arg):
        return sim.by_id[stream_arg.value]

    def prepare(dst, ws, wsb, row, col, ca, cb, nnz, sidx, first_sample, key, first, K_, B, V, n, stream):
        on(stream).ops.append(("build", dst.value, first))
        return 0

    def pull(dst, src, K_, B, n, rank, stream):
        on(stream).ops.append(("pull", dst.value, src.value, sim.current._expect))
        return 0

    def train_step(args, stream):
        eng = sim.current
        which = args._obj.value
        on(stream).ops.append(("step", local_ptr(eng.dp_rank, which), eng.host_step))
        return 0

    def batch_counts(plan, K_, B, k, out, stream):
        eng = sim.current
        which = plan.value - local_ptr(eng.dp_rank, 0)
        on(stream).ops.append(("read", plan.value, eng.plan_first[which]))
        return 0

    monkeypatch.setattr(E, "lib", types.SimpleNamespace(glove_prepare_batches_sharded=prepare, glove_plan_pull_slice=pull,
                                                        glove_shard_train_step=train_step, glove_plan_batch_counts=batch_counts))
    monkeypatch.setattr(E, "check", lambda rc, what="": None)
    real_pull = E.GloveEngine._ring_pull

    def ring_pull(self, R, j, which, first, after=None):
        self._expect = first                            # first step of the plan the pull about to be issued must find
        return real_pull(self, R, j, which, first, after)

    monkeypatch.setattr(E.GloveEngine, "_ring_pull", ring_pull)

    engs = []
    for r in range(world):
        e = object.__new__(E.GloveEngine)
        e.dp_rank, e.dp_world, e.sharded, e.K, e.B, e.V_global = r, world, True, K, 64 * world, 1000
        e.max_steps, e.host_step, e.device = max_steps, first_step, "sim"
        e.optimizer, e.adam_mode, e.shard_exchange, e.overlap = "Adam", "replay", "peer-push", True
        e._main = FakeStream(sim, r, "main")
        e._side, e._prep_stream = FakeStream(sim, r, "side"), FakeStream(sim, r, "prep")
        e.plans = [FakeBuf(local_ptr(r, 0)), FakeBuf(local_ptr(r, 1))]
        e.plan_bytes = PLAN
        e.plan_first, e._plan_counts, e._plan_shards, e._plan_need = [None, None], [None, None], [None, None], [None, None]
        e._ev_plan, e._keep, e._plan_override, e._ev_catchup = [None, None], [None, None], None, None
        e._ev_step_done = [FakeEvent(), FakeEvent()]
        e.coo, e.nnz, e.shuffle_key, e.sample_idx = tuple(FakeBuf(50 + i) for i in range(4)), 1 << 20, 7, None
        e.prep_ws, e._ev_coo = FakeBuf(99), None
        e._args = [ctypes.c_int(0), ctypes.c_int(1)]
        e._barriers = 0

        def barrier(e=e):
            sim.current_stream().ops.append(("barrier", e._barriers))
            e._barriers += 1

        e._ring = dict(buf=FakeBuf((r + 1) * 10 ** 9), hdl=None, ptrs=[(q + 1) * 10 ** 9 for q in range(world)], barrier=barrier,
                       built={}, opened=set(), ev_barrier=None, keep={}, build_stream=FakeStream(sim, r, "build"),
                       pull_stream=FakeStream(sim, r, "pull"), builder=e._ring_build_resident)
        engs.append(e)
    return sim, engs


def host(sim, eng, fn, *a):
    sim.current = eng
    try:
        return fn(*a)
    finally:
        sim.current = None


def execute(sim, rng, bias, limit=None):
    """Run queued device operations (all of them, or at most `limit`); `bias` = weight per stream kind: a lopsided
    scheduler lets e.g. the build streams race ahead or holds one rank's pull stream back."""
    n = 0
    while limit is None or n < limit:
        heads = [s for s in sim.streams if s.ops]
        if not heads:
            return True
        runnable = []
        for s in heads:
            op = s.ops[0]
            if op[0] == "wait":
                if op[1] is None or op[1] in sim.done:
                    runnable.append((s, None))
            elif op[0] in ("barrier", "step", "a2a"):   # collectives: the ranks' matching calls, all at the head of their queues
                kind = {"barrier": "pull", "step": "main", "a2a": "build"}[op[0]]
                peers = [t for t in sim.streams if t.kind == kind and t.ops and t.ops[0][0] == op[0] and t.ops[0][-1] == op[-1]]
                if len(peers) == sim.world:
                    runnable.append((s, peers))
            else:
                runnable.append((s, None))
        if not runnable:
            return False                            # deadlock
        w = [bias.get(s.kind, 1.0) * bias.get(("rank", s.rank), 1.0) for s, _ in runnable]
        s, peers = rng.choices(runnable, weights=w)[0]
        for t in (peers or [s]):
            op = t.ops.popleft()
            if op[0] == "record":
                sim.done.add(op[1])
            elif op[0] == "build":
                sim.mem[op[1]] = op[2]
            elif op[0] == "pull":
                assert sim.mem.get(op[2]) == op[3], "rank %d pulled the plan of steps %r.. where %r.. was expected" % (t.rank, sim.mem.get(op[2]), op[3])
                sim.mem[op[1]] = op[3]
            elif op[0] == "step":
                first = sim.mem.get(op[1])
                assert first is not None and first <= op[2] < first + sim.K, "rank %d ran step %d on the plan of steps %r.." % (t.rank, op[2], first)
            elif op[0] == "read":
                assert sim.mem.get(op[1]) == op[2], "rank %d read the plan of steps %r.. for %r.." % (t.rank, sim.mem.get(op[1]), op[2])
            if op[0] in sim.executed:
                sim.executed[op[0]] += 1
        n += 1
    return True


BIASES = [{}, {"build": 50.0}, {"build": 0.02}, {"pull": 50.0}, {"pull": 0.02}, {"main": 50.0}, {"main": 0.02},
          {("rank", 0): 0.02}, {("rank", 1): 50.0, "build": 20.0}]


@pytest.mark.parametrize("world,K,first_step", [(2, 4, 0), (2, 4, 2048), (3, 4, 20), (4, 2, 6), (8, 2, 32), (8, 4, 0)])
def test_shared_plan_ring_protocol(monkeypatch, world, K, first_step):
    n_steps = 6 * world * K + 3                         # several rounds
    for seed, bias in enumerate(BIASES * 3):
        rng = random.Random(1000 * world + seed)
        sim, engs = make_world(monkeypatch, world, K, first_step, first_step + 4 * n_steps)
        sim.K = K
        lag = rng.choice([0, 3, 40, None])              # device operations executed per host step (None: GPU fully lazy)
        for s in range(n_steps):
            for e in (engs if rng.random() < 0.5 else engs[::-1]):     # the ranks' hosts are not synchronised
                host(sim, e, e._step_sharded)
            if lag:
                assert execute(sim, rng, bias, lag)
            if s == n_steps // 2 and seed % 3 == 1:
                # the batch order changes in mid-run (set_coo / set_batches / set_step): every rank resets its ring
                for e in engs:
                    e.plan_first = [None, None]
                    host(sim, e, e._ring_reset)
            if s == n_steps // 2 and seed % 3 == 2:
                # plans of EARLIER steps asked for again (bench.py reads the segment counts of the timed steps): rounds out of
                # order -- must start over behind a barrier, not overwrite a buffer a peer still copies from
                for e in engs:
                    for t in range(first_step + 1, first_step + s, max(1, K - 1)):
                        host(sim, e, e.batch_counts, t)
        assert execute(sim, rng, bias), "deadlock: the ranks' barrier sequences do not match"
        assert len({e._barriers for e in engs}) == 1
        assert sim.executed["step"] == world * n_steps and sim.executed["pull"] >= world * (n_steps // K)
        # per rank the construction cost is that of 1 / world of the chunks: one build per round (+ the look-ahead round, the
        # partial first round, and the rounds rebuilt after a reset / asked for again out of order)
        rounds = n_steps / (K * world)
        per_rank = sim.executed["build"] / world
        assert per_rank <= rounds + 4 + (3 if seed % 3 == 1 else 0) + (rounds + 6 if seed % 3 == 2 else 0), (per_rank, rounds, sim.executed)


class Anything:
    """Stand-in for a tensor whose contents do not matter here: every method returns the object itself."""
    ptrs = itertools.count(500)

    def __init__(self, n=0, is_cuda=True):
        self.n, self.is_cuda, self.ptr = n, is_cuda, next(Anything.ptrs)

    def numel(self):
        return self.n

    def data_ptr(self):
        return self.ptr

    def __getitem__(self, k):
        return self

    def __mod__(self, k):
        return self

    def __getattr__(self, name):
        return lambda *a, **k: self


@pytest.mark.parametrize("world,K", [(2, 4), (4, 2), (8, 2)])
def test_shared_plan_rounds_of_host_fed_chunks(monkeypatch, world, K):
    """train_chunks_from_host(sliced=True) with shared plans: rounds of `world` chunks -- H2D of the shares, one all-to-all,
    rank q plans chunk q, slices pulled as the steps reach them -- between stretches of steps on the resident COO (the ring
    is reset on the way in and out).  Same checks as above, plus matching all-to-all sequences."""
    import torch.distributed as dist
    from glove_tensorflow_b200 import engine as E
    for seed, bias in enumerate(BIASES * 2):
        rng = random.Random(77 * world + seed)
        first_step = K * world * rng.randrange(0, 3)
        sim, engs = make_world(monkeypatch, world, K, first_step, 10 ** 6)
        sim.K = K
        monkeypatch.setattr(E.torch, "empty", lambda n, *a, **k: Anything(n if isinstance(n, int) else 0))
        monkeypatch.setattr(E.torch, "arange", lambda *a, **k: Anything())
        monkeypatch.setattr(dist, "get_backend", lambda *a, **k: "gloo")

        def a2a(recv, send, group=None):
            e = sim.current
            sim.current_stream().ops.append(("a2a", e._a2a))
            e._a2a += 1

        monkeypatch.setattr(dist, "all_to_all_single", a2a)
        m = K * 64                                         # a rank's share of a chunk: K steps x B / world triples
        for e in engs:
            e._ring["hdl"] = object()
            e._a2a, e._label, e.loss_cap, e.loss_out = 0, None, 4096, Anything()
        n_resident = 0
        for phase in range(3):
            for _ in range(rng.randrange(0, 2 * K * world)):   # steps on the resident COO
                for e in engs:
                    host(sim, e, e._step_sharded)
                n_resident += 1
                assert execute(sim, rng, bias, rng.choice([0, 5, 50]))
            n_chunks = world * rng.randrange(1, 4)
            chunks = [tuple(Anything(m, is_cuda=False) for _ in range(4)) for _ in range(n_chunks)]
            for e in engs:                                  # every rank enqueues its whole call; the final loss read drains
                host(sim, e, e.train_chunks_from_host, chunks, True)
            assert execute(sim, rng, bias), "deadlock"
            assert all(e.host_step == engs[0].host_step for e in engs)
        assert len({e._barriers for e in engs}) == 1 and len({e._a2a for e in engs}) == 1 and engs[0]._a2a > 0
        assert sim.executed["step"] == world * (engs[0].host_step - first_step)


def test_plan_reads_with_host_synchronisation_do_not_deadlock(monkeypatch):
    """bench.py at N > 1 reads the segment counts of the timed steps back (GloveEngine.batch_counts): plans of EARLIER
    chunks, asked for out of order, and every call synchronises the caller's step stream on the host.  A rank blocked in
    that synchronisation depends on barriers its peers have not even enqueued yet; the ranks' hosts run the same loop
    independently.  Modelled here with one coroutine per rank that may only proceed past a call once its step stream has
    drained: the loop must complete on every rank, for every world size and schedule."""
    for world, K, first_step, n_steps in [(2, 16, 2048, 25), (2, 4, 0, 30), (4, 4, 8, 40), (8, 2, 32, 40), (3, 4, 20, 30)]:
        for seed, bias in enumerate(BIASES):
            rng = random.Random(31 * world + seed)
            sim, engs = make_world(monkeypatch, world, K, first_step, 10 ** 6)
            sim.K = K
            for s in range(n_steps):
                for e in engs:
                    host(sim, e, e._step_sharded)
                assert execute(sim, rng, bias, rng.choice([0, 10, 100]))
            assert execute(sim, rng, bias)                       # torch.cuda.synchronize() + barrier after the timed region

            def reader(e):
                for t in range(first_step, first_step + min(n_steps, 32)):
                    host(sim, e, e.batch_counts, t)
                    while e._main.ops:                           # cudaStreamSynchronize inside glove_plan_batch_counts
                        before = sum(sim.executed.values())
                        execute(sim, rng, bias, 50)
                        if e._main.ops and sum(sim.executed.values()) == before:
                            yield                                # blocked on something a peer's host has yet to enqueue
                for _ in range(2 * K):                           # ... and training goes on afterwards
                    host(sim, e, e._step_sharded)

            runs = [reader(e) for e in engs]
            idle = 0
            while runs and idle < 4 * world:
                for g in list(runs):
                    before = sum(sim.executed.values()) + sum(len(st.ops) for st in sim.streams)
                    try:
                        next(g)
                    except StopIteration:
                        runs.remove(g)
                        idle = 0
                        continue
                    idle = 0 if sum(sim.executed.values()) + sum(len(st.ops) for st in sim.streams) != before else idle + 1
            assert not runs, "world %d: the ranks' hosts block one another" % world
            assert execute(sim, rng, bias), "deadlock"
            assert len({e._barriers for e in engs}) == 1 and sim.executed["read"] == world * min(n_steps, 32)
