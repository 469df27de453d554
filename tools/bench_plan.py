"""Device time of one plan construction (glove_prepare_batches_sharded) for a global batch of B triples x K steps, as every
rank of an N-way row-sharded job builds it: python tools/bench_plan.py B K n_shards"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from glove_tensorflow_b200._lib import lib, check
B, K, NS = (int(x) for x in sys.argv[1:4])
V = 400000
dev = torch.device("cuda:0")
row, col, t, w = bench.gen_coo_device(V, 1 << 26, 1234, dev)
u8 = dict(dtype=torch.uint8, device=dev)
plan = torch.empty(lib.glove_plan_bytes(K, B), **u8)
ws = torch.empty(lib.glove_prepare_workspace_bytes(K, B), **u8)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda x: ctypes.c_void_p(x.data_ptr())
def build(first):
    check(lib.glove_prepare_batches_sharded(p(plan), p(ws), ws.numel(), p(row), p(col), p(t), p(w), row.numel(), None, first * B, 7,
                                            first, K, B, V, NS, st), "prepare")
for i in range(3): build(i * K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): build((3 + i) * K)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("B=%d K=%d n_shards=%d: %.3f ms per plan = %.1f us per step (plan %.2f GB)" % (B, K, NS, ms, 1e3 * ms / K, plan.numel() / 1e9))
