"""Back-to-back fused sweeps of the headline configuration (for ncu: kernel durations with warm caches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import bench
m = bench.make_model(1000000, 10, 0)
m.fit(int(sys.argv[1]) if len(sys.argv) > 1 else 30, None)
m._engine.synchronize()
print('ok')
