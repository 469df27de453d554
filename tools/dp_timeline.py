"""Device-side timeline of the row-sharded step WITHOUT per-step synchronisation (run under torchrun on N GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_timeline.py [shared] [priority]

The five phases of a peer-push step (stage + push | announce + wait for every owner | update of the owned segments |
announce + wait + finish) are enqueued as separate launches with CUDA events between them, 64 steps back to back; the
medians say where a pipelined step spends its time on this rank (waits include waiting for the slowest peer)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from glove_tensorflow_b200.engine import GloveEngine, _stream
from glove_tensorflow_b200._lib import lib, check
opts = sys.argv[1:]
rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev); dist.init_process_group("nccl", device_id=dev)
if "priority" in opts:
    torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
V, d, Bl = (2_200_000 if "cc" in opts else 400_000), 300, 65536; B = Bl * world
eng = GloveEngine(V, d, batch_size=B, plan_steps=16, max_steps=8192, device=dev, dp_rank=rank, dp_world=world, dp_mode="sharded")
row, col, t, w = bench.gen_coo_device(V, 1 << 26, 1234, dev)
eng.enable_peer_gather(push=True)
if "shared" in opts: eng.enable_plan_sharing()
eng.balance_owners(row, col)
eng.init_uniform(1); eng.set_coo(row, col, t, w, shuffle_key=1)
for _ in range(1024): eng.step()
torch.cuda.synchronize(); dist.barrier()
M = 96
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(M)]
scal = eng._shard_scalars()
t0 = time.perf_counter()
for i in range(M):
    which = eng._plan_for(eng.host_step); eng._before_step(which)
    a, st = ctypes.byref(eng._args[which]), _stream()
    e = ev[i]
    e[0].record(); check(lib.glove_shard_stage_step(a, st))
    e[1].record(); check(lib.glove_shard_signal_staged(a, st)); check(lib.glove_shard_wait_staged(a, st))
    e[2].record(); check(lib.glove_shard_update_step(a, ctypes.c_void_p(scal.data_ptr()), st))
    e[3].record(); check(lib.glove_shard_finish_sync(a, ctypes.c_void_p(scal.data_ptr()), st))
    e[4].record(); eng._after_step(); eng.host_step += 1
host = (time.perf_counter() - t0) / M
torch.cuda.synchronize()
ph = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(4)] for e in ev])
gap = np.array([ev[i][4].elapsed_time(ev[i + 1][0]) for i in range(M - 1)])
tot = ev[0][0].elapsed_time(ev[-1][4]) / M
res = torch.tensor(list(np.median(ph, 0)) + [float(np.median(gap)), tot, host * 1e3] + list(ph.mean(0)), device=dev)
allr = [torch.empty_like(res) for _ in range(world)]
dist.all_gather(allr, res)
if rank == 0:
    print("opts", opts, "world", world)
    print("rank | median ms: stage+push, signal+wait, update, finish(sync) | gap between steps | ms/step | host enqueue ms/step | mean ms of the 4 phases")
    for r, x in enumerate(allr):
        print(r, [round(float(v), 4) for v in x])
# and the plain pipelined loop (one C call per step), for reference
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(128): eng.step()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print("one call per step: %.4f ms/step" % (e0.elapsed_time(e1) / 128))
dist.destroy_process_group()
