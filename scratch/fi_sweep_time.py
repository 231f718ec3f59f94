"""Config-4 fi sweep time (back-to-back sweeps) and the streaming kernels behind it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import torch
import bench
m = bench.make_model(1000000, 10, 0, fi=True)
m.fit(5, None)
torch.cuda.synchronize()
t = time.perf_counter()
m._engine.sweep(50)
torch.cuda.synchronize()
print('fi sweep ms', (time.perf_counter() - t) / 50 * 1e3)
