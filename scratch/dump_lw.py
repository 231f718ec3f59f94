import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/golden')
import bench
from cimrgp_b200 import _lib
m = bench.make_model(1000000, 0)
e = m._engine
out = {}
for s in range(30):
    for j in range(e.J):
        e.phase_a(j); e.axis_update(j)
        if s in (1, 5, 12, 20, 29):
            out['s%d_L%d' % (s, j)] = e.get(-1, _lib.F_LOG_OMEGA_HAT, (30, 30))
            out['om_s%d_L%d' % (s, j)] = e.get(-1, _lib.F_OMEGA, (30, 30))
        e.phase_b(j); e.bias_noise(j)
    if s in (1,5,12,20,29): print(s, e.get(-1, _lib.F_OMEGA_ITERS, (10,)).astype(int), flush=True)
np.savez_compressed('/root/repo/gpurun_out/lw_dump.npz', **out)
