import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden'); sys.path.insert(0,'/root/repo/tests')
import workloads
from oracle import mrgp_oracle as O
from cimrgp_b200.distributed import ShardedEngine
from cimrgp_b200.engine import Engine
class Comm:
    def all_reduce(self, t, op): pass
n,res,M=20000,5,30
x,y=workloads.workload1(n); xs=(x-x.mean(0))/x.std(0); offs=O.uniform_offsets(n,res,2)
e=ShardedEngine(xs,y,offs,M,0,1,comm=Comm()); e.sweep(3,use_graph=False); e.synchronize()
r=Engine(xs,y,offs,M); r.sweep(3); r.synchronize()
print(e.elbo()); print(r.elbo())
from cimrgp_b200 import _lib
for j in range(3):
    R=e.R[j]
    for f,name,shape in ((30,'sumsB',(R,5)),(29,'yvar',(R,)),(23,'noise_mean',(R,)),(10,'prec',(R,30)),(21,'m2',(R,30))):
        a=e.get(j,f,shape); b=r.get(j,f,shape)
        print(j,name,'nan' if np.isnan(a).any() else 'ok', float(np.max(np.abs(a-b)/(np.abs(b)+1e-300))))
sh=e.shared_state(); print({k: bool(np.isnan(v).any()) for k,v in sh.items()})
