"""Config 5 throughput: S independent series of N=2048, 6 resolutions, ci vs fi (SeriesBatch)."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import numpy as np, torch
import workloads
from cimrgp_b200 import LaplacianEigenpairs, MaternKernel, SeriesBatch
S = int(sys.argv[1]); streams = int(sys.argv[2]); group = int(sys.argv[3]) if len(sys.argv) > 3 else 128
xs, ys = zip(*[workloads.workload1(2048, seed=10 + s) for s in range(S)])
for fi in (False, True):
    t0 = time.time()
    b = SeriesBatch(xs, ys, 30, 5, LaplacianEigenpairs(), MaternKernel(nu=1, l=1, sf=1), forced_independence=fi, n_streams=streams, group_size=group)
    t_build = time.time() - t0
    b.fit(3)
    torch.cuda.synchronize(); t0 = time.time()
    b.fit(5)
    torch.cuda.synchronize(); dt = (time.time() - t0) / 5
    print('%s S=%d streams=%d group=%d: build %.1fs, %.2f ms per batch iteration = %.0f series-sweeps/s' % ('fi' if fi else 'ci', S, streams, group, t_build, dt * 1e3, S / dt), flush=True)
    b.close(); del b
