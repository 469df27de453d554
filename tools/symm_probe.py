"""Probe: torch symmetric memory across the ranks of one node (peer pointers + barrier) and a timed remote read."""
import os
import time

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    n = 64 << 20
    t = symm_mem.empty(n, dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(rank, "ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal", hasattr(hdl, "signal_pad_ptrs"), flush=True)
    hdl.barrier()
    peer = (rank + 1) % world
    remote = hdl.get_buffer(peer, (n,), torch.float32)
    torch.cuda.synchronize()
    out = torch.empty_like(t)
    for _ in range(3):
        out.copy_(remote)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out.copy_(remote)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    # barrier latency
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        hdl.barrier()
    torch.cuda.synchronize()
    bt = (time.perf_counter() - t0) / 100
    print(rank, "remote value", float(out[0]), "expected", float(peer + 1), "GB/s", n * 4 / dt / 1e9, "barrier us", bt * 1e6, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
