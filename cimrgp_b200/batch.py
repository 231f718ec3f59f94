"""Batch of independent series (BASELINE config 5): one MultiResolutionGaussianProcess per series,
all of them in ONE device allocation and one pinned staging area.  The series of a group are swept together
(mrgp_group_*, include/cimrgp.h): ci models take the batched fused sweep - ONE kernel launch per iteration for the
whole group, one thread-block cluster per series; fi models are captured as parallel branches of one CUDA graph.
(group_size=0: one graph launch per series on a pool of streams.)

Over several GPUs the series are split by rank in contiguous blocks (`rank`, `world`): replicas only, no collective
(SURVEY.md §8e).

The reference has no such API - every series is a separate object fitted one after the other
(scripts/tests/*.py loop over models) - so this is a front end over the drop-in class, not a new model: series s
behaves exactly like `MultiResolutionGaussianProcess([xs[s], ys[s]], ...)` and `batch[s]` IS that object.
Series are independent: over several GPUs they are split by rank with no collective ("replicas only").
"""
import ctypes as C

import numpy as np

from . import _lib
from .engine import Engine
from .IndexSetGenerator import IndexSetUniform, offsets_of
from .MRGP import MultiResolutionGaussianProcess


def series_range(n_series, rank, world):
    """Contiguous block of series kept by `rank` of `world` processes (replicas only, no collective)."""
    per = -(-int(n_series) // int(world))
    return min(n_series, int(rank) * per), min(n_series, (int(rank) + 1) * per)


class SeriesBatch(object):
    def __init__(self, xs, ys, n_basis, resolution, basis_function_obj, spectral_density_obj=None, divider=2,
                 forced_independence=False, n_streams=16, device=0, n_ctas=None, group_size=None, rank=0, world=1,
                 **model_kw):
        """xs[s]: (N_s, 1) inputs, ys[s]: (N_s, dy) observations of series s; the other arguments as for
        MultiResolutionGaussianProcess (one index set IndexSetUniform(N_s, resolution, divider) per series).
        n_ctas: streaming grid per model (default: one CTA per 1024 samples, so that many models fit on the GPU).
        group_size: models per group; 0 = one graph launch per model on the stream pool.  Default: the whole batch in
        ci mode (one fused launch per iteration), 64 in fi mode (branch graphs).
        rank, world: this process keeps the series [rank * per, (rank + 1) * per), per = ceil(S / world)."""
        import torch
        self.torch = torch
        self.series_range = series_range(len(xs), rank, world)
        if world > 1:
            xs, ys = xs[self.series_range[0]:self.series_range[1]], ys[self.series_range[0]:self.series_range[1]]
        self.n_series = len(xs)
        if len(ys) != self.n_series or self.n_series == 0:
            raise ValueError('xs and ys must list the same, non-zero number of series')
        dev = torch.device('cuda', int(device))
        mode = 'fi' if forced_independence else 'ci'
        sizes = [int(np.asarray(x).shape[0]) for x in xs]
        dy = int(np.asarray(ys[0]).shape[1])
        index_sets, need, ctas = [], [], []
        probe = {}
        for n in sizes:
            idx = IndexSetUniform(n, resolution, divider)
            index_sets.append(idx)
            c = int(n_ctas) if n_ctas else max(1, min(64, n // 1024))
            ctas.append(c)
            if (n, c) not in probe:
                probe[(n, c)] = Engine.probe_workspace_bytes(offsets_of(idx), n_basis, dy=dy, mode=mode, n_ctas=c,
                                                             device=device)
            need.append((probe[(n, c)] + 255) & ~255)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, min(int(n_streams), self.n_series)))]
        self.workspace = torch.empty(int(sum(need)), dtype=torch.uint8, device=dev)
        total = int(sum(sizes))
        self.pinned_x = torch.empty((total, 1), dtype=torch.float64).pin_memory()
        self.pinned_y = torch.empty((total, dy), dtype=torch.float64).pin_memory()
        self.models = []
        w_off = s_off = 0
        for s in range(self.n_series):
            opts = dict(stream=self.streams[s % len(self.streams)], workspace=self.workspace[w_off:w_off + need[s]],
                        pinned=(self.pinned_x[s_off:s_off + sizes[s]], self.pinned_y[s_off:s_off + sizes[s]]))
            self.models.append(MultiResolutionGaussianProcess(
                [np.asarray(xs[s], dtype=np.float64), np.asarray(ys[s], dtype=np.float64)], n_basis, index_sets[s],
                basis_function_obj, spectral_density_obj, forced_independence=forced_independence, device=device,
                n_ctas=ctas[s], _engine_opts=opts, **model_kw))
            w_off += need[s]
            s_off += sizes[s]
        self.lib = _lib.load()
        self.groups = []
        if group_size is None:
            group_size = 64 if forced_independence else self.n_series
        if group_size and group_size > 0:
            for g0 in range(0, self.n_series, int(group_size)):
                members = self.models[g0:g0 + int(group_size)]
                arr = (C.c_void_p * len(members))(*[m._engine.handle for m in members])
                out = C.c_void_p()
                rc = self.lib.mrgp_group_create(arr, len(members), None, C.byref(out))
                if rc != 0:
                    raise _lib.MrgpError(rc, 'mrgp_group_create failed: ' + self.lib.mrgp_last_error(members[0]._engine.handle).decode())
                self.groups.append(out)

    def __len__(self):
        return self.n_series

    def __getitem__(self, s):
        return self.models[s]

    def fit(self, n_iter=1):
        """n_iter sweeps of every series (fit(n_iter, None) of each model), interleaved over the stream pool."""
        if self.groups:
            for g in self.groups:
                rc = self.lib.mrgp_group_sweep(g, int(n_iter))
                if rc != 0:
                    raise _lib.MrgpError(rc, 'mrgp_group_sweep failed')
        else:
            for _ in range(int(n_iter)):
                for m in self.models:
                    m._engine.sweep(1)
        self.synchronize()

    def set_observations(self, ys):
        """New observations for every series at unchanged inputs: staged in the pinned area, uploaded on the models'
        streams; the groups refresh the layer-0 statistics of all series with one launch at the next fit()."""
        for m, y in zip(self.models, ys):
            m._engine.set_observations(np.asarray(y, dtype=np.float64))

    def launch_count(self):
        n = sum(int(self.lib.mrgp_group_launch_count(g)) for g in self.groups)
        return n if self.groups else sum(m._engine.launch_count() for m in self.models)

    def synchronize(self):
        for g in self.groups:
            rc = self.lib.mrgp_group_synchronize(g)
            if rc != 0:
                raise _lib.MrgpError(rc, 'a group sweep failed on the device')
        for st in self.streams:
            st.synchronize()

    def close(self):
        for g in self.groups:
            self.lib.mrgp_group_destroy(g)
        self.groups = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
