import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden'); sys.path.insert(0,'/root/repo/tests')
import workloads
from oracle import mrgp_oracle as O
from cimrgp_b200.distributed import ShardedEngine
from cimrgp_b200 import _lib
class Comm:
    def all_reduce(self, t, op): pass
n,res,M=20000,5,30
x,y=workloads.workload1(n); xs=(x-x.mean(0))/x.std(0); offs=O.uniform_offsets(n,res,2)
e=ShardedEngine(xs,y,offs,M,0,1,comm=Comm())
lib,h=e.lib,e.handle
def bad(tag):
    out=[]
    for j in range(e.J):
        for k,v in e.layer_state(j).items():
            if not np.all(np.isfinite(v)): out.append('L%d.%s'%(j,k))
    for k,v in e.shared_state().items():
        if not np.all(np.isfinite(v)): out.append('S.'+k)
    if out: print(tag, out[:10]); sys.exit(0)
for s in range(3):
    for j in range(e.J):
        e._ck(lib.mrgp_phase_a(h,j)); e._ck(lib.mrgp_region_sums(h,j,0)); e.synchronize()
        xa=e._xchg[j].cpu().numpy().reshape(e.R[j],-1)
        if not np.isfinite(xa).all(): print('sweep',s,'layer',j,'xchg A nonfinite', np.argwhere(~np.isfinite(xa))[:5]); sys.exit(0)
        e._ck(lib.mrgp_axis_update(h,j)); e.synchronize(); bad('sweep %d layer %d after mid'%(s,j))
        e._ck(lib.mrgp_phase_b(h,j)); e._ck(lib.mrgp_region_sums(h,j,1)); e.synchronize()
        xb=e._xchg[j].cpu().numpy().reshape(e.R[j],-1)
        if not np.isfinite(xb).all(): print('sweep',s,'layer',j,'xchg B nonfinite', np.argwhere(~np.isfinite(xb))[:5]); sys.exit(0)
        e._ck(lib.mrgp_bias_noise(h,j)); e.synchronize(); bad('sweep %d layer %d after post'%(s,j))
print('clean')
