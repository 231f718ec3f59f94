"""In-graph timeline of the sharded sweep (rank 0 prints). torchrun."""
import os, sys, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, torch.distributed as dist
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE']); lr=int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device('cuda',lr))
import workloads
from oracle import mrgp_oracle as O
from cimrgp_b200.distributed import ShardedEngine
n,res,M=1000000,9,30
x,y=workloads.workload1(n); xs=(x-x.mean(0))/x.std(0); offs=O.uniform_offsets(n,res,2)
e=ShardedEngine(xs,y,offs,M,rank,world,device=lr)
e.sweep(12); e.synchronize(); dist.barrier()
ev0,ev1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
ev0.record(e.stream); e.sweep(20); ev1.record(e.stream); e.synchronize()
if rank==0: print('graph sweep %.1f us'%(ev0.elapsed_time(ev1)*1000/20),flush=True)
dist.barrier()
e.lib.mrgp_timeline_enable(e.handle, 1)
e.sweep(2); e.synchronize(); dist.barrier()
tags=(C.c_int32*128)(); ms=(C.c_float*256)()
nn=e.lib.mrgp_timeline_read(e.handle, tags, ms, 128)
e.synchronize(); dist.barrier()
if rank==0:
    J=10; names=['A','mid','B','omega']
    for j in range(J):
        line='L%d:'%j
        for k in range(4):
            b,en=ms[2*(4*j+k)]*1000, ms[2*(4*j+k)+1]*1000
            line+='  %s [%.1f -> %.1f] %.1f'%(names[k], b, en, en-b)
        print(line,flush=True)
    print('iters', e.get(-1, 51, (10,)))
dist.barrier(); sys.stdout.flush(); os._exit(0)
