import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden'); sys.path.insert(0,'/root/repo/tests')
import workloads
from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel, _lib
from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
x,y = workloads.workload1(4096)
m = MultiResolutionGaussianProcess([x,y],30,IndexSetUniform(4096,5,2),LaplacianEigenpairs(),MaternKernel(1,1,1))
e = m._engine
def bad():
    out=[]
    for j in range(e.J):
        for k,v in e.layer_state(j).items():
            if not np.all(np.isfinite(v)): out.append('L%d.%s'%(j,k))
    for k,v in e.shared_state().items():
        if not np.all(np.isfinite(v)): out.append('S.'+k)
    lw = e.get(-1,_lib.F_LOG_OMEGA_HAT,(30,30))
    if not np.all(np.isfinite(lw)): out.append('S.lw')
    return out, lw
for sweep in range(4):
    for j in range(e.J):
        for name,fn in (('A',e.phase_a),('mid',e.axis_update),('B',e.phase_b),('post',e.bias_noise)):
            fn(j); e.synchronize()
            b,lw = bad()
            if name=='mid': print('sweep',sweep,'layer',j,'lw range %.4g %.4g'%(lw.min(), lw.max()), flush=True)
            if b:
                print('sweep',sweep,'layer',j,'after',name,'non-finite:',b[:12]); 
                sh=e.shared_state(); print('omega', sh['omega'][:2,:6]); print('kappa', sh['kappa'][:4]); print('ardshape', sh['ard_shape'][:4], sh['ard_scale'][:4])
                sys.exit(0)
print('no NaN')
