import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import bench
from cimrgp_b200 import _lib
m = bench.make_model(1000000, 0)
e = m._engine
for s in range(6):
    e.sweep(1); e.synchronize()
    print('sweep', s, 'omega iters', e.get(-1, _lib.F_OMEGA_ITERS, (10,)).astype(int))
