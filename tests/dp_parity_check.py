"""Real multi-process parity of the row-sharded path (test infrastructure; run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/dp_parity_check.py [exchange[+shared-plans]] [graph]

Every rank owns 1/N of the rows (frequency-balanced owner map), steps run through the real exchange -- by default the
peer-memory pull with device-side synchronisation ('peer-sync', as CUDA graphs), or 'peer' / 'alltoall' / 'allgather'
(NCCL) -- on the GLOBAL batch, and rank 0 compares the union of the shards and the per-step losses with the
single-process oracle (C port, fp32 + fp64 shadow) on the same injected batches.  Exit code 0 = parity within the
north-star tolerance (1e-5 relative on losses and tables); the observed maxima are printed as one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import make_coo
    from glove_tensorflow_b200.engine import GloveEngine
    from oracle import c_oracle, glove_oracle as o
    exchange = sys.argv[1] if len(sys.argv) > 1 else "peer-push"
    shared = exchange.endswith("+shared-plans")      # rank c % N builds the plan of chunk c, every rank pulls its slice
    exchange = exchange.split("+")[0]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    V, d, B_local, steps, K = 20_000, 300, 4096, 40, 8
    B = B_local * world
    n = 40 * B
    coo = make_coo(V, n, 71, hot=0.05)
    batches = np.random.default_rng(72).integers(0, n, (steps, B))
    st = o.init_state(V, d, 73)
    eng = GloveEngine(V, d, learning_rate=0.01, batch_size=B, plan_steps=K, max_steps=steps + K, device=dev,
                      dp_rank=rank, dp_world=world, dp_mode="sharded")
    eng.shard_exchange = exchange
    if exchange.startswith("peer"):
        eng.enable_peer_gather(direct=exchange == "peer-direct", sync=exchange == "peer-sync", push=exchange == "peer-push")
    if shared:
        eng.enable_plan_sharing()
    eng.balance_owners(coo["row"], coo["col"], hot=2048)
    eng.load_state(st.R, st.C, st.rb, st.cb, st.g)
    eng.set_coo(coo["row"], coo["col"], coo["target"], coo["weight"])
    eng.set_batches(batches)
    eng.use_graph = exchange in ("peer-sync", "peer-push") and len(sys.argv) > 2 and sys.argv[2] == "graph"
    losses = eng.train(steps)
    got = eng.get_state()
    ids = torch.from_numpy(eng.owned_ids()).to(dev)
    out = {}
    for k in ("R", "C", "rb", "cb"):
        full = torch.zeros((V,) + got[k].shape[1:], dtype=torch.float32, device=dev)
        full[ids] = torch.from_numpy(got[k][: ids.numel()]).to(dev)
        dist.all_reduce(full)                      # the shards are disjoint: the sum is their union
        out[k] = full.cpu().numpy()
    all_losses = [torch.empty(steps, dtype=torch.float32, device=dev) for _ in range(world)]
    dist.all_gather(all_losses, torch.from_numpy(np.asarray(losses, np.float32)).to(dev))
    rc = 0
    if rank == 0:
        c32 = c_oracle.COracle(st.R, st.C, st.rb, st.cb)
        c64 = c_oracle.COracle(st.R, st.C, st.rb, st.cb, dtype=np.float64)
        l32 = c32.train(coo, batches, learning_rate=0.01)
        l64 = c64.train(coo, batches, learning_rate=0.01)
        rel = lambda a, b: float(np.max(np.abs(np.asarray(a, np.float64) - b)) / np.max(np.abs(b)))
        res = {"world": world, "exchange": exchange, "shared_plans": shared, "graph": bool(eng.use_graph), "steps": steps, "global_batch": B,
               "loss_vs_oracle32": float(np.max(np.abs(losses - l32) / np.abs(l32))),
               "loss_vs_shadow64": float(np.max(np.abs(losses - l64) / np.abs(l64))),
               "ranks_agree_on_losses": all(bool(torch.equal(all_losses[0], t)) for t in all_losses)}
        for k in ("R", "C", "rb", "cb"):
            res[k + "_vs_shadow64"] = rel(out[k], getattr(c64, k))
            res[k + "_oracle32_vs_shadow64"] = rel(getattr(c32, k), getattr(c64, k))
        ok = (res["loss_vs_oracle32"] < 1e-5 and res["ranks_agree_on_losses"]
              and all(res[k + "_vs_shadow64"] <= max(1e-5, 3.0 * res[k + "_oracle32_vs_shadow64"]) for k in ("R", "C", "rb", "cb")))
        res["parity"] = "ok" if ok else "FAILED"
        print(json.dumps(res))
        rc = 0 if ok else 1
    flag = torch.tensor([rc], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(int(flag.item()))


if __name__ == "__main__":
    main()
