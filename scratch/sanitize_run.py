"""Small ci runs for compute-sanitizer: a cluster of 4 CTAs (N = 20000, 8 layers), a cluster of 16 (N = 120000, 10 layers), one CTA."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
import workloads
from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel
from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
for n, res in ((20000, 7), (120000, 9), (2048, 5)):
    x, y = workloads.workload1(n)
    m = MultiResolutionGaussianProcess([x, y], 30, IndexSetUniform(n, res, 2), LaplacianEigenpairs(), MaternKernel(1, 1, 1))
    m.fit(3, None)
    e = m._engine
    e.prefetch_observations(y * 1.01)
    e.refresh_statistics()
    e.sweep(2)
    print(n, res, float(np.sum(e.elbo())))
