"""world_size-2 gloo test of the host-side sharding logic (CPU): contiguous sample chunks, per-region statistics
computed per rank with the oracle and summed with an all-reduce equal the unsharded statistics."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, 'tests', 'golden'), os.path.join(ROOT, 'tests')):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from cimrgp_b200.distributed import chunk_bounds
    from oracle import mrgp_oracle as O
    import workloads
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    n, res, M = 5000, 4, 20
    x, y = workloads.workload1(n)
    offsets = O.uniform_offsets(n, res, 2)
    full = O.OracleMRGP(x, y, M, offsets, mode='fi')
    lo, hi = chunk_bounds(n, world, rank)
    ok = True
    for j, off in enumerate(offsets):
        ly = full.layers[j]
        # local contribution to Phi^T y and sum phi^2 per region: rows of the chunk only
        mask = np.zeros(n, dtype=bool)
        mask[lo:hi] = True
        T_loc = np.zeros((ly.R, 2, M))
        d_loc = np.zeros((ly.R, M))
        for r in range(ly.R):
            rows = np.arange(off[r], off[r + 1])
            rows = rows[mask[rows]]
            T_loc[r] = (ly.Phi[rows].T @ y[rows]).T
            d_loc[r] = np.sum(ly.Phi[rows] ** 2, axis=0)
        t = torch.from_numpy(np.concatenate([T_loc.ravel(), d_loc.ravel()]))
        dist.all_reduce(t)
        got = t.numpy()
        T_all = np.stack([(ly.Phi[off[r]:off[r + 1]].T @ y[off[r]:off[r + 1]]).T for r in range(ly.R)])
        want = np.concatenate([T_all.ravel(), ly.d.ravel()])
        ok = ok and np.allclose(got, want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    # the one exchange of the fused ci sweep: layer-0 sufficient statistics of the observations, summed over the ranks
    ly = full.layers[0]
    rows = np.arange(lo, hi)
    loc = np.concatenate([(ly.Phi[rows].T @ y[rows]).ravel(), y[rows].sum(0), [np.sum(y[rows] ** 2)]])
    t = torch.from_numpy(loc)
    dist.all_reduce(t)
    want = np.concatenate([(ly.Phi.T @ y).ravel(), y.sum(0), [np.sum(y ** 2)]])
    ok = ok and np.allclose(t.numpy(), want, rtol=1e-12, atol=1e-12 * np.abs(want).max())
    # series of a batch split by rank: a partition, no collective
    from cimrgp_b200.batch import series_range
    mine = torch.zeros(4097, dtype=torch.int64)
    a, b = series_range(4097, rank, world)
    mine[a:b] = 1
    dist.all_reduce(mine)
    ok = ok and bool(torch.all(mine == 1))
    bounds = [chunk_bounds(n, world, r) for r in range(world)]
    ok = ok and bounds[0][0] == 0 and bounds[-1][1] == n and all(b[0] % 32 == 0 for b in bounds) \
        and all(bounds[k][1] == bounds[k + 1][0] for k in range(world - 1))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_statistics_sum_to_the_unsharded_ones():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_chunk_bounds_cover_everything():
    from cimrgp_b200.distributed import chunk_bounds
    for n in (1, 31, 32, 33, 1000, 1000000, 999999):
        for w in (1, 2, 3, 4, 8):
            b = [chunk_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and max(x[1] for x in b) == n
            assert all(b[k][1] == b[k + 1][0] or b[k + 1][0] == n for k in range(w - 1))
            assert sum(x[1] - x[0] for x in b) == n
