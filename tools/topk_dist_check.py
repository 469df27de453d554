"""torchrun check of the row-sharded GloveEngine.topk (real NCCL collectives) against the single-GPU answer.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/topk_dist_check.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from glove_tensorflow_b200.engine import GloveEngine
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    V, d, k, nq = 400_001, 300, 10, 8192
    rng = np.random.default_rng(0)
    T = rng.uniform(-0.05, 0.05, (V, d)).astype(np.float32)
    T[V - 2] = T[3]
    z = np.zeros(V, np.float32)
    q = rng.integers(0, V, nq).astype(np.int32)
    q[:3] = (3, V - 2, V - 1)
    eng = GloveEngine(V, d, batch_size=64, plan_steps=1, max_steps=4, device=dev, dp_rank=rank, dp_world=world, dp_mode="sharded")
    eng.load_state(T, T, z, z)
    eng.topk(q[:256], k)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    sim, idx = eng.topk(q, k)
    dt = time.perf_counter() - t0
    if rank == 0:
        one = GloveEngine(V, d, batch_size=64, plan_steps=1, max_steps=4, device=dev)
        one.load_state(T, T, z, z)
        one.topk(q[:256], k)
        t0 = time.perf_counter()
        sim1, idx1 = one.topk(q, k)
        dt1 = time.perf_counter() - t0
        same = np.array_equal(idx, idx1)
        print({"world": world, "ids_equal": bool(same), "max_sim_diff": float(np.abs(sim - sim1).max()),
               "sharded_s": dt, "single_gpu_s": dt1, "fallbacks": eng.last_topk_fallbacks})
        assert same or np.allclose(sim, sim1, atol=5e-7)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
