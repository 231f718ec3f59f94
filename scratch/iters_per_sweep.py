"""Solver iterations and cycles per layer for the first sweeps of the headline configuration (MRGP_CHAIN_PROF=1)."""
import os, sys
os.environ['MRGP_CHAIN_PROF'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
import bench
m = bench.make_model(1000000, 10, 0)
res = 9
for s in range(int(sys.argv[1]) if len(sys.argv) > 1 else 26):
    m.fit(1, None)
    it = m._engine.get(-1, 51, (res + 1,))
    p = m._engine.get(-1, 54, (res + 1, 16))
    solve = p[:, 8] - p[:, 7]
    k = p[:4, 15]
    print('sweep %2d iters %s sum %3d | solve:iter cycles %s sum %6d | kernel(CTA0) %d' % (
        s, ' '.join('%2d' % v for v in it), it.sum(), ' '.join('%5d' % v for v in solve), solve.sum(), k[3] - k[0]))
