import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import bench
m = bench.make_model(1000000, 0)
e = m._engine
for s in range(16):
    e.sweep(1); e.synchronize()
    print(s+1, e.get(-1, 51, (10,)).astype(int).tolist(), flush=True)
