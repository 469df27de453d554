"""Per-phase device timings of one data-parallel step (run under torchrun): python -m torch.distributed.run ... tools/dp_profile.py [sharded|replicated]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from glove_tensorflow_b200.engine import GloveEngine
mode = sys.argv[1] if len(sys.argv) > 1 else "sharded"
rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev); dist.init_process_group("nccl", device_id=dev)
V, d, Bl = 400000, 300, 65536; B = Bl * world
eng = GloveEngine(V, d, batch_size=B, plan_steps=16, max_steps=4096 + 2048, device=dev, dp_rank=rank, dp_world=world, dp_mode=mode)
row, col, t, w = bench.gen_coo_device(V, 1 << 24, 1234, dev)
if len(sys.argv) > 2 and sys.argv[2].startswith("peer"): eng.enable_peer_gather(direct=sys.argv[2] == "peer-direct", sync=sys.argv[2] == "peer-sync", push=sys.argv[2] == "peer-push")
if mode == "sharded": eng.balance_owners(row, col)
eng.init_uniform(1); eng.set_coo(row, col, t, w, shuffle_key=1)
for _ in range(400): eng.step()
torch.cuda.synchronize(); dist.barrier()
n = 64
ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
acc = np.zeros(6)
host = 0.0
for _ in range(n):
    if mode == "sharded":
        th = time.perf_counter()
        ev[0].record(); upad = eng.shard_stage(); ev[1].record()
        N, r = world, rank
        if len(sys.argv) > 2 and sys.argv[2] == "peer-push":
            eng.shard_signal_staged(); eng.shard_wait_staged()
        elif len(sys.argv) > 2 and sys.argv[2] == "peer-sync":
            eng.shard_signal_staged(); eng.shard_pull()
        elif len(sys.argv) > 2 and sys.argv[2].startswith("peer"):
            eng._symm.barrier()
            if sys.argv[2] == "peer": eng.shard_pull()
        elif len(sys.argv) > 2 and sys.argv[2] == "allgather":
            for side in (0, 1):
                snap, u = eng.snapshot_view(side), upad[side]
                dist.all_gather_into_tensor(snap[: N * u], snap[r * u:(r + 1) * u])
        else:
            send, recv = eng.shard_pack()
            dist.all_to_all_single(eng._xbuf[1][: sum(recv)], eng._xbuf[0][: sum(send)], recv, send)
            eng.shard_unpack()
        ev[2].record(); eng.shard_update(); ev[3].record()
        if len(sys.argv) > 2 and sys.argv[2] in ("peer-sync", "peer-push"):
            ev[4].record(); eng.shard_finish_sync(); ev[5].record()
        else:
            dist.all_reduce(eng._shard_scalars())
            ev[4].record(); eng.shard_finish(); ev[5].record()
        host += time.perf_counter() - th
        torch.cuda.synchronize()
        acc[:5] += [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
    else:
        which = eng._plan_for(eng.host_step); eng._before_step(which)
        ev[0].record(); gr, gc, gs = eng.grad_step(); ev[1].record()
        n_r, n_c = eng._counts_for(eng.host_step)
        dist.all_reduce(eng._grad_flat[:8 + n_r * eng.S]); dist.all_reduce(gc[:n_c * eng.S]); ev[2].record()
        eng.apply_step(); ev[3].record()
        torch.cuda.synchronize()
        acc[:3] += [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
if rank == 0:
    names = ["stage", "exchange", "update(own)", "all_reduce(3 floats)", "finish"] if mode == "sharded" else ["stage+grad", "all_reduce", "apply"]
    print(mode, "per-phase ms:", {k: float(v) for k, v in zip(names, np.round(acc / n, 4))}, "host enqueue ms/step", round(host / n * 1e3, 4))
while eng.host_step % eng.K: eng.step()
for use_graph in (False, True):
    eng.use_graph = use_graph and eng.shard_exchange in ("peer-sync", "peer-push")
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    eng.train(n)
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0: print("pipelined ms/step (graph=%s)" % eng.use_graph, (time.perf_counter() - t0) / n * 1e3)
dist.destroy_process_group()
