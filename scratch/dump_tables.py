"""Dump log omega_hat of every layer for the first sweeps of the headline configuration (offline study of the solver)."""
import os, sys
os.environ['MRGP_CHAIN_PROF'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
import bench
m = bench.make_model(1000000, 10, 0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 30
out = []
its = []
for s in range(S):
    m.fit(1, None)
    out.append(m._engine.get(-1, 56, (10, 30, 30)).copy())
    its.append(m._engine.get(-1, 51, (10,)).copy())
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'chain_tables.npz'), tables=np.array(out), iters=np.array(its))
print('ok', np.array(out).shape)
