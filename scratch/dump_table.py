import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import bench
from cimrgp_b200 import _lib
m = bench.make_model(1000000, 0)
e = m._engine
out = {}
for s in range(1, 8):
    for j in range(10):
        e.phase_a(j); e.axis_update(j); e.synchronize()
        if j <= 2:
            out['lw_s%d_j%d' % (s, j)] = e.get(-1, _lib.F_LOG_OMEGA_HAT, (30, 30)).copy()
            out['it_s%d_j%d' % (s, j)] = e.get(-1, 51, (10,))[j]
        e.phase_b(j); e.bias_noise(j)
    e.synchronize()
np.savez('/root/repo/gpurun_out/tables.npz', **out)
print({k: float(v) for k, v in out.items() if k.startswith('it_')})
