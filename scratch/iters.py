import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import bench
from cimrgp_b200 import _lib
m = bench.make_model(1000000, 0)
e = m._engine
for s in range(6):
    e.sweep(1); e.synchronize()
it=e.get(-1, _lib.F_OMEGA_ITERS, (10,)).astype(int)
tr=e.get(-1, 52, (10,64*64))
for j in (1,5,9): print('L%d iters %d  cycles: prologue %d eval %d S %d chol %d solve %d total %d'%((j,it[j])+tuple(int(v) for v in tr[j][16:22])))
