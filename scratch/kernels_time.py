"""Launches of the streaming kernels for ncu: the layer-0 statistics pass (ci) and two fi sweeps (phase A / phase B of every layer)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import bench
m = bench.make_model(1000000, 10, 0)
m.fit(2, None)
m._engine.refresh_statistics()
m._engine.synchronize()
f = bench.make_model(1000000, 10, 0, fi=True)
f.fit(3, None)
print('ok')
