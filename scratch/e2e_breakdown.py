"""Where a pipelined end-to-end step goes: device time of each stage (events on the engine's stream) and host wall time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np, torch
import bench
m = bench.make_model(1000000, 10, 0)
eng = m._engine
m.fit(5, None)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
ev = lambda: torch.cuda.Event(enable_timing=True)
K = 20
rows = []
eng.prefetch_observations()
torch.cuda.synchronize()
t_wall0 = time.perf_counter()
for k in range(K):
    e = [ev() for _ in range(5)]
    h0 = time.perf_counter()
    e[0].record(eng.stream)
    bench.flush_l2(eng, flush, torch)
    e[1].record(eng.stream)
    eng.refresh_statistics()
    e[2].record(eng.stream)
    if k + 1 < K: eng.prefetch_observations()
    eng.sweep(1)
    e[3].record(eng.stream)
    h1 = time.perf_counter()
    lb = eng.elbo()
    e[4].record(eng.stream)
    e[4].synchronize()
    h2 = time.perf_counter()
    rows.append([e[i].elapsed_time(e[i + 1]) * 1e3 for i in range(4)] + [(h1 - h0) * 1e6, (h2 - h1) * 1e6])
wall = (time.perf_counter() - t_wall0) / K * 1e6
r = np.median(np.array(rows), axis=0)
print('device us: flush %.0f | wait copy + statistics %.0f | sweep %.0f | elbo + read-back %.0f ; host us: enqueue %.0f, elbo call %.0f ; wall per step %.0f us' % (*r, wall))
