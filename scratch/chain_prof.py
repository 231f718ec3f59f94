"""Cycle breakdown of the fused ci sweep (MRGP_CHAIN_PROF=1): per layer, SM-clock stamps inside CTA 0."""
import os, sys
os.environ['MRGP_CHAIN_PROF'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
import workloads
from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel
from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
n, res = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000, int(sys.argv[2]) if len(sys.argv) > 2 else 9
x, y = workloads.workload1(n)
m = MultiResolutionGaussianProcess([x, y], 30, IndexSetUniform(n, res, 2), LaplacianEigenpairs(), MaternKernel(1, 1, 1))
m.fit(12, None)
if os.environ.get('FLUSH') == '1':
    import torch
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    m._engine.synchronize(); flush.zero_(); torch.cuda.synchronize()
    m.fit(1, None)
p = m._engine.get(-1, 54, (res + 1, 16))
names = ['gather', 'bingham|ard', 'mean', 'table', 'exp', '->sync2', 'solve:load', 'solve:iter', 'solve:fb', None,
         'w:pubcopy', 'w:mid1', 'w:mid2stats', 'w:bar', None]
print('iters', m._engine.get(-1, 51, (res + 1,)))
for j in range(res + 1):
    r = p[j]
    sh = ' '.join('%s=%d' % (names[k], r[k + 1] - r[k]) for k in range(5))
    so = ' '.join('%s=%d' % (names[k], r[k + 1] - r[k]) for k in (6, 7, 8))
    wo = ' '.join('%s=%d' % (names[k], r[k + 1] - r[k]) for k in (10, 11, 12, 13))
    nxt = (p[j + 1][0] - r[9]) if j < res else 0
    print('L%d shared[%s] sync2=%d | solve[%s] ->next=%d | worker(other SM clock)[%s] | layer=%d' % (
        j, sh, r[6] - r[5], so, nxt, wo, (p[j + 1][0] - r[0]) if j < res else r[9] - r[0]))
k = p[:4, 15]
print('kernel (CTA 0 clock): prologue=%d loop=%d last finish + barrier=%d total=%d cycles' % (k[1] - k[0], k[2] - k[1], k[3] - k[2], k[3] - k[0]))
print('background finish of layers 4.. (worker warp 0, cycles):', [int(v) for v in p[4:res, 15]])
