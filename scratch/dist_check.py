"""2-rank check of the sharded path against the unsharded engine (run under torchrun)."""
import os, sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, torch.distributed as dist
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE']); lr=int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
print('rank',rank,'init pg',flush=True)
dist.init_process_group('nccl', device_id=torch.device('cuda',lr))
import workloads
from oracle import mrgp_oracle as O
from cimrgp_b200.distributed import ShardedEngine
from cimrgp_b200.engine import Engine
n,res,M=200000,7,30
x,y=workloads.workload1(n); xs=(x-x.mean(0))/x.std(0); offs=O.uniform_offsets(n,res,2)
print('rank',rank,'building',flush=True)
e=ShardedEngine(xs,y,offs,M,rank,world,device=lr)
print('rank',rank,'built',flush=True)
e.sweep(2,use_graph=False); e.synchronize(); print('rank',rank,'2 plain sweeps ok',flush=True)
t=time.time(); e.sweep(3,use_graph=True); e.synchronize(); print('rank',rank,'graph sweeps ok %.2fs'%(time.time()-t),flush=True)
ref=Engine(xs,y,offs,M,device=lr); ref.sweep(5); ref.synchronize()
a,b=e.state(),ref.state(latent=False)
worst=max(float(np.max(np.abs(a[k]-b[k])/(np.abs(b[k])+1e-12*np.abs(b[k]).max()+1e-300))) for k in b)
print('rank',rank,'max rel diff vs unsharded %.2e'%worst,flush=True)
dist.barrier(); dist.destroy_process_group()
