import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import bench
from cimrgp_b200 import _lib
m = bench.make_model(1000000, 0)
e = m._engine
for s in range(30):
    e.sweep(1); e.synchronize()
    if s in (0,1,2,3,5,8,12,16,20,25,29): print('sweep', s, e.get(-1, _lib.F_OMEGA_ITERS, (10,)).astype(int), 'noise0 %.6f'%e.get(0,_lib.F_NOISE_MEAN,(1,))[0])
