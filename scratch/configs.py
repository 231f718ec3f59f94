"""Throughput of the other BASELINE configs (one GPU): config 3 (N=1e5, 8 resolutions, ci) and config 4 in fi mode."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import numpy as np, torch
import workloads
from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel
from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
for name, n, res, fi in (('config 3 ci', 100000, 7, False), ('config 3 fi', 100000, 7, True), ('config 4 fi', 1000000, 9, True)):
    x, y = workloads.workload1(n)
    m = MultiResolutionGaussianProcess([x, y], 30, IndexSetUniform(n, res, 2), LaplacianEigenpairs(), MaternKernel(nu=1, l=1, sf=1), forced_independence=fi)
    e = m._engine
    e.sweep(5); e.synchronize()
    ms = []
    for k in range(20):
        flush.zero_(); e.stream.wait_stream(torch.cuda.current_stream())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(e.stream); e.sweep(1); b.record(e.stream); b.synchronize()
        ms.append(a.elapsed_time(b))
    print('%s: N=%d, %d layers: %.3f ms per sweep (median of 20, L2 flushed), %.0f it/s' % (name, n, res + 1, float(np.median(ms)), 1e3 / float(np.median(ms))), flush=True)
