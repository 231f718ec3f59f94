import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import workloads
from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel
from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
N=int(sys.argv[1]); J=int(sys.argv[2]); pre=int(sys.argv[3])
x,y = workloads.workload1(N)
m = MultiResolutionGaussianProcess([x,y],30,IndexSetUniform(N,J-1,2),LaplacianEigenpairs(),MaternKernel(1,1,1))
e=m._engine
if pre: st=e.state(); print('state ok', flush=True)
e.sweep(1); e.synchronize(); print('graph sweep ok', flush=True)
