import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import bench
m = bench.make_model(1000000, 10, 0)
for _ in range(3):
    m._engine.refresh_statistics()
m._engine.synchronize()
print('ok')
