import os, sys, time, json
sys.path.insert(0, '/root/repo')
import torch, torch.distributed as dist, numpy as np
import bench
from glove_tensorflow_b200.engine import GloveEngine
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); dev=torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev); dist.init_process_group("nccl", device_id=dev)
V,d,Bl=400000,300,65536; B=Bl*world
eng=GloveEngine(V,d,batch_size=B,plan_steps=16,max_steps=4096+2048,device=dev,dp_rank=rank,dp_world=world)
eng.init_uniform(1); row,col,t,w=bench.gen_coo_device(V,1<<24,1234,dev); eng.set_coo(row,col,t,w,shuffle_key=1)
bench.steady_state(eng,V,B,99)
for _ in range(20): eng.step()
torch.cuda.synchronize(); dist.barrier()
ev=[torch.cuda.Event(enable_timing=True) for _ in range(5)]
acc=np.zeros(4); n=64
for _ in range(n):
    which=eng._plan_for(eng.host_step); eng._before_step(which)
    ev[0].record(); gr,gc,gs=eng.grad_step(); ev[1].record()
    n_r,n_c=eng._counts_for(eng.host_step)
    dist.all_reduce(eng._grad_flat[:8+n_r*eng.S]); dist.all_reduce(gc[:n_c*eng.S]); ev[2].record()
    eng.apply_step(); ev[3].record()
    torch.cuda.synchronize()
    acc+= [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), 0]
if rank==0: print("DP breakdown ms (grad, allreduce, apply):", acc[:3]/n, "n_r,n_c", n_r, n_c)
t0=time.perf_counter()
for _ in range(n): eng.step()
torch.cuda.synchronize(); dist.barrier()
if rank==0: print("pipelined ms/step", (time.perf_counter()-t0)/n*1e3)
dist.destroy_process_group()
