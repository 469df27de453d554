import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch, bench
from glove_tensorflow_b200.engine import GloveEngine
V,d,N=400000,300,8; B=65536*N
eng=GloveEngine(V,d,batch_size=B,plan_steps=16,max_steps=4096,dp_rank=7,dp_world=N,dp_mode="sharded")
row,col,t,w=bench.gen_coo_device(V,1<<24,1234,eng.device); eng.set_coo(row,col,t,w,shuffle_key=1)
send,recv=eng._need_counts(0)
need=eng._plan_need[0]
print("shape",need.shape,"send",send,"recv",recv)
flat=need[:, :, :, :].reshape(2,-1)
for s in (0,1):
    a=need[s]            # [K][8][9]
    d_=np.diff(a,axis=2)
    print("side",s,"min diff within lists",d_.min(),"max",d_.max(),"first",a[0,0,:], "last list", a[15,7,:])
    # list ends should equal next list starts
    ends=a[:,:,8].reshape(-1); starts=a[:,:,0].reshape(-1)
    print("  continuity ok:", np.array_equal(ends[:-1], starts[1:]), "total", ends[-1])
assert all(np.all(np.diff(need[s].reshape(-1, 9), axis=1) >= 0) and np.array_equal(need[s][:, :, 8].reshape(-1)[:-1], need[s][:, :, 0].reshape(-1)[1:]) for s in (0, 1)), "request lists are not contiguous"
own,upad=eng._shard_info(0); print("own",own,"upad",upad)
