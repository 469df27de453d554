"""Pipelined ms/step of the row-sharded step under torchrun for the synchronisation variants (measurement aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from glove_tensorflow_b200.engine import GloveEngine
rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev); dist.init_process_group("nccl", device_id=dev)
V, d, Bl = 400000, 300, 65536; B = Bl * world
row, col, t, w = bench.gen_coo_device(V, 1 << 24, 1234, dev)
eng = GloveEngine(V, d, batch_size=B, plan_steps=16, max_steps=8192, device=dev, dp_rank=rank, dp_world=world, dp_mode="sharded")
eng.enable_peer_gather(sync=True, push=True)
eng.balance_owners(row, col); eng.init_uniform(1); eng.set_coo(row, col, t, w, shuffle_key=1)
for _ in range(400): eng.step()
n = 128
for name, ex, graph in (("peer", "peer", False), ("peer-sync", "peer-sync", False), ("peer-push", "peer-push", False), ("peer-push graph", "peer-push", True)):
    eng.shard_exchange = ex
    for a in eng._args: a.peer_gather = {"peer": 2, "peer-sync": 3, "peer-push": 4}[ex]
    eng.use_graph = graph
    eng._graphs = [None, None] if not graph else eng._graphs
    while eng.host_step % eng.K: eng.step()
    eng.train(32)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    left = n
    while left:
        k = eng.step_chunk_graph() if graph else 0
        if not k: eng.step(); k = 1
        left -= k
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    if rank == 0: print("%-28s %.4f ms/step" % (name, e0.elapsed_time(e1) / n), flush=True)
dist.destroy_process_group()
