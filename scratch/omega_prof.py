import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import bench
m = bench.make_model(1000000, 0)
e = m._engine
for s in range(12): e.sweep(1)
e.synchronize()
print('Newton cycles H, chol+fwd, back+exp:', e.get(-1, 52, (3,)))
print('k_ard cycles stage/partials, ard stats, table, rowmax, colmax+exp:', e.get(-1, 53, (5,)))
print('iters', e.get(-1, 51, (10,)))
