"""Parity of the CUDA path (through the drop-in API and the C ABI) with the golden fixtures produced by
the unmodified reference and with the CPU oracle.  Tolerance (BASELINE.json north_star): relative 1e-6 in
fp64, plus 1e-12 * max|ref| per array for exact zeros; index sets bit-exact."""
import os

import numpy as np
import pytest

from oracle import mrgp_oracle as O
from parity import ATOL_ABS, ATOL_ABS_SHARED, assert_state_close, compare, expand_shared, mask_degenerate, mismatch
import workloads

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RTOL = 1e-6


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def split(g, prefix):
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


def build(x, y, m, resolution, fi, basis_interval_obj=None, **kw):
    from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel
    from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
    return MultiResolutionGaussianProcess(
        train_xy=[x, y], n_basis=m, index_set_obj=IndexSetUniform(x.shape[0], resolution, 2),
        basis_function_obj=LaplacianEigenpairs(), spectral_density_obj=MaternKernel(nu=1, l=1, sf=1),
        adaptive_inputs=False, standard_normalized_inputs=True, basis_interval_obj=basis_interval_obj, interval_factor=1,
        forced_independence=fi, **kw)


CASES = [('c1_ci', False, {}), ('c1_fi', True, {}), ('n2000_ci', False, {}), ('n2000_fi', True, {}),
         ('c2_ci', False, {}), ('c2_fi', True, {}), ('n600_ci_snr', False, dict(snr_ratio=10.))]


@pytest.mark.parametrize('name,fi,kw', CASES)
def test_sweeps_match_reference_goldens(name, fi, kw):
    g = load(name)
    x, y = g['x'], g['y']
    m = build(x, y, int(g['meta.M']), int(g['meta.resolution']), fi, **kw)
    compare(m._engine.state(), split(g, 'k0.'))
    done = 0
    for k in g['meta.checkpoints']:
        m.fit(int(k) - done, None)
        done = int(k)
        compare(m._engine.state(), split(g, 'k%d.' % k))
    if 'pred.x' in g.files:
        from cimrgp_b200 import IndexSetUniform
        xt = g['pred.x']
        assert mismatch(m.get_predicted_mean(xt), g['pred.mean_global'], RTOL) is None
        idx_t = IndexSetUniform(xt.shape[0], int(g['meta.resolution']), 2)
        assert mismatch(m.get_predicted_mean(xt, index_set_obj=idx_t), g['pred.mean_indexed'], RTOL) is None
        assert mismatch(m.get_central_moment2(xt), g['pred.var_global'], RTOL) is None


def test_ci_upper_layers_are_exactly_inert():
    """SURVEY.md §8c behavioural KAT: exact zeros, and indexed prediction == layer-0 prediction."""
    from cimrgp_b200 import IndexSetUniform
    x, y = workloads.workload1(4096)
    m = build(x, y, 30, 5, False)
    m.fit(4, None)
    for j in range(1, 6):
        st = m._engine.layer_state(j)
        assert not np.any(st['A']) and not np.any(st['bias_mean']) and not np.any(st['ytil'])
    xt = np.atleast_2d(np.linspace(1, 3, 5000)).T
    a = m.get_predicted_mean(xt)
    b = m.get_predicted_mean(xt, index_set_obj=IndexSetUniform(5000, 5, 2))
    assert np.array_equal(a, b)


@pytest.mark.parametrize('fi', [False, True])
def test_elbo_and_oracle_mid_size(fi):
    """N = 20000, 8 layers: three sweeps against the CPU oracle (fsolve vs the exact scaling: 1e-6)."""
    x, y = workloads.workload1(20000)
    offsets = O.uniform_offsets(20000, 7, 2)
    ora = O.OracleMRGP(x, y, 30, offsets, mode='fi' if fi else 'ci')
    m = build(x, y, 30, 7, fi)
    for _ in range(3):
        ora.sweep()
    m.fit(3, None)
    compare(m._engine.state(), ora.state())
    if not fi:
        total, per_layer, terms = ora.elbo()
        got = m._engine.elbo()
        assert mismatch(got, terms, RTOL) is None


def test_elbo_matches_reference_golden():
    g = load('c1_ci_elbo')
    x, y = g['x'], g['y']
    n_iter = g['lower_bound'].shape[0]
    m = build(x, y, int(g['meta.M']), int(g['meta.resolution']), False)
    m.fit(n_iter=n_iter, tol=1e-300, min_iter=n_iter)
    assert mismatch(np.array(m.lower_bound_terms), g['terms'], RTOL) is None
    assert mismatch(np.array(m.lower_bound_layer), g['lower_bound_layer'], RTOL) is None
    assert mismatch(np.array(m.lower_bound), g['lower_bound'], RTOL) is None


def test_adaptive_intervals_match_reference_golden():
    """B1 (BasisInterval.learn): the reference's own run with BasisInterval(opt_interval_factor=(1, 1.2))."""
    from cimrgp_b200 import BasisInterval
    g = load('c1_ci_adaptive')
    x, y = g['x'], g['y']
    m = build(x, y, int(g['meta.M']), int(g['meta.resolution']), False,
              basis_interval_obj=BasisInterval(opt_interval_factor=(1, 1.2)))
    assert m.adaptive_basis_intervals is True
    done = 0
    for k in g['meta.checkpoints']:
        m.fit(int(k) - done, None)
        done = int(k)
        compare(m._engine.state(), split(g, 'k%d.' % k))
    assert m._engine.interval_failures() == 0


@pytest.mark.parametrize('use_prior', [True, False])
def test_adaptive_intervals_against_oracle(use_prior):
    """Ragged regions over several CTAs, with and without the spectral prior in the objective."""
    from cimrgp_b200 import BasisInterval
    n = 6000
    x, y = workloads.workload1(n)
    offsets = O.uniform_offsets(n, 4, 2)
    ora = O.OracleMRGP(x, y, 20, offsets, mode='ci', adaptive=dict(use_prior=use_prior, opt_interval_factor=(1., 1.3)))
    m = build(x, y, 20, 4, False, basis_interval_obj=BasisInterval(use_prior=use_prior, opt_interval_factor=(1., 1.3)))
    for _ in range(2):
        ora.sweep()
    m.fit(2, None)
    ref = ora.state()
    got = m._engine.state()
    # the optimum is located to xatol = 1e-5 by both sides with the same steps; the intervals do move
    static = np.max(np.abs(x - np.mean(x)) / np.std(x))
    assert abs(float(ref['L0.L'].ravel()[0]) / static - 1.0) > 1e-7
    compare(got, ref)
    assert m._engine.interval_failures() == 0


def test_adaptive_intervals_are_switched_off_in_fi_mode():
    """MRGP.py:108-109."""
    from cimrgp_b200 import BasisInterval
    x, y = workloads.workload1(512)
    m = build(x, y, 20, 3, True, basis_interval_obj=BasisInterval())
    assert m.adaptive_basis_intervals is False


def test_graph_replay_equals_stepwise_phases_and_is_deterministic():
    x, y = workloads.workload1(50000)
    a = build(x, y, 30, 6, False)
    b = build(x, y, 30, 6, False)
    c = build(x, y, 30, 6, False)
    a.fit(3, None)
    c.fit(3, None)
    for _ in range(3):
        b._engine.sweep_stepwise()
    b._engine.synchronize()
    sa, sb, sc = a._engine.state(), b._engine.state(), c._engine.state()
    for k in sa:
        assert np.array_equal(sa[k], sc[k]), k          # same launch geometry -> bitwise reproducible
    # sweep == per-phase ABI calls, up to the summation order of the layer statistics: the per-phase calls stream every
    # layer, the fused sweep takes layer 0 from the sufficient statistics of y and the layers above in closed form
    compare(sa, sb, rtol=1e-9)


def test_skipped_statistics_pass_of_inferred_layers_is_an_identity():
    """The sweep does not launch the P1 statistics pass on ci layers above the first, whose targets are inferred
    from the layer's own posterior (Phi^T r == 0, y_tilde == d a).  The per-phase ABI still streams it.  On a state
    whose upper layers are NOT inert (coefficients and biases set by hand) both forms must agree to roundoff."""
    from cimrgp_b200 import _lib
    x, y = workloads.workload1(40000)
    models = [build(x, y, 30, 5, False) for _ in range(2)]
    rng = np.random.RandomState(3)
    pert = [(0.05 * rng.standard_normal((2 ** j, 30, 2)), 0.3 * rng.standard_normal((2 ** j, 2))) for j in range(6)]
    for m in models:
        m.fit(2, None)
        for j in range(1, 6):
            m._engine.put(j, _lib.F_A, pert[j][0])
            m._engine.put(j, _lib.F_BIAS_MEAN, pert[j][1])
    models[0]._engine.sweep(1)              # captured sweep: no statistics pass on layers 1..5
    models[1]._engine.sweep_stepwise()      # per-phase calls: every layer streams both passes
    sa, sb = models[0]._engine.state(), models[1]._engine.state()
    assert np.max(np.abs(sb['L3.ytil'])) > 1e-3          # the upper layers really carry signal here
    compare(sa, sb, rtol=1e-9)


def test_results_do_not_depend_on_grid_size():
    x, y = workloads.workload1(30000)
    a = build(x, y, 30, 6, True, n_ctas=148)
    b = build(x, y, 30, 6, True, n_ctas=7)
    a.fit(2, None)
    b.fit(2, None)
    compare(b._engine.state(), a._engine.state(), rtol=1e-9)


def test_ragged_random_regions_against_oracle():
    """IndexSetUniform(n_regions=...) random boundaries (IndexSetGenerator.py:67-92): ragged regions whose
    boundaries do not nest across layers."""
    from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel
    from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
    x, y = workloads.workload1(3000)
    np.random.seed(5)
    idx = IndexSetUniform(3000, 3, 2, n_regions=[1, 3, 4, 7], min_percentage_of_samples_per_region=0.1)
    for fi in (False, True):
        m = MultiResolutionGaussianProcess([x, y], 20, idx, LaplacianEigenpairs(), MaternKernel(2.5, 0.7, 1.3),
                                           forced_independence=fi, interval_factor=[1.0, 1.1, 1.2, 1.3])
        ora = O.OracleMRGP(x, y, 20, idx.offsets, mode='fi' if fi else 'ci', spectral=(2.5, 0.7, 1.3),
                           interval_factor=[1.0, 1.1, 1.2, 1.3])
        m.fit(2, None)
        ora.sweep()
        ora.sweep()
        compare(m._engine.state(), ora.state())


def test_full_size_properties():
    """BASELINE config 4 (N = 1e6, 10 layers, 1023 regions): size-independent properties."""
    from cimrgp_b200 import IndexSetUniform
    n = 1000000
    x, y = workloads.workload1(n)
    m = build(x, y, 30, 9, False)
    m.fit(3, None)
    eng = m._engine
    sh = eng.shared_state()
    assert np.max(np.abs(sh['omega'].sum(0) - 1)) < 1e-11 and np.max(np.abs(sh['omega'].sum(1) - 1)) < 1e-11
    tot = 0
    for j in range(10):
        st = eng.layer_state(j)
        assert all(np.all(np.isfinite(v)) for v in st.values())
        # noise shape counts the samples of the region: c = c0 + dy/2 * n  (Posteriors.py:136)
        assert np.array_equal(st['noise_shape'], 1e-45 + 0.5 * 2 * np.diff(eng.offsets[j]).astype(np.float64))
        assert np.allclose(st['bias_prec'], np.diff(eng.offsets[j]))
        tot += st['A'].shape[0]
        if j > 0:
            assert not np.any(st['A']) and not np.any(st['bias_mean'])
    assert tot == 1023
    # layer 0 explains the signal: residual variance ~ noise level of the generator (0.1 * U(1,2))^2
    st0 = eng.layer_state(0)
    assert 20. < st0["noise_mean"][0] < 60.
    xt = np.atleast_2d(np.linspace(1, 3, 100000)).T
    pm = m.get_predicted_mean(xt)
    truth = workloads.signal1(xt)[:, :, 0].T
    assert np.sqrt(np.mean((pm - truth) ** 2)) < 0.2
    assert np.array_equal(pm, m.get_predicted_mean(xt, index_set_obj=IndexSetUniform(100000, 9, 2)))
    # fi at full size: every layer is active; phase-B sums are consistent with the bias posterior
    f = build(x, y, 30, 9, True)
    f.fit(2, None)
    for j in (0, 5, 9):
        st = f._engine.layer_state(j)
        assert all(np.all(np.isfinite(v)) for v in st.values())
        assert np.any(st['A'])
    assert f._engine.cholesky_count() >= 2 * 1023 * 30


def test_batched_cholesky_matches_lapack():
    import ctypes as C
    import torch
    from cimrgp_b200 import _lib
    lib = _lib.load()
    rng = np.random.RandomState(0)
    for n in (2, 5, 16, 32):
        batch = 1000
        a = rng.randn(batch, n, n + 3)
        a = a @ np.swapaxes(a, 1, 2)
        a[7] = -a[7]                                  # not PD: potrf info = 1
        ref = np.linalg.cholesky(np.delete(a, 7, axis=0))
        t = torch.as_tensor(a, device='cuda').contiguous()
        info = torch.zeros(batch, dtype=torch.int32, device='cuda')
        torch.cuda.synchronize()
        assert lib.mrgp_batched_cholesky(None, C.c_void_p(t.data_ptr()), n, batch, C.c_void_p(info.data_ptr())) == 0
        torch.cuda.synchronize()
        got = np.tril(np.delete(t.cpu().numpy(), 7, axis=0))
        info = info.cpu().numpy()
        assert info[7] == 1 and not np.any(np.delete(info, 7))
        assert np.max(np.abs(got - ref)) < 1e-11 * np.max(np.abs(ref))


def test_constructor_errors_match_reference():
    from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel, BasisInterval
    from cimrgp_b200.MRGP import MultiResolutionGaussianProcess as M
    x, y = workloads.workload1(64)
    idx = IndexSetUniform(64, 2, 2)
    le, mk = LaplacianEigenpairs(), MaternKernel()
    with pytest.raises(TypeError):
        M([x, y], 30, idx, le, mk, axis_resolution_specific=True)                       # MRGP.py:50
    with pytest.raises(ValueError):
        M([x, y[:, :1]], 30, idx, le, mk)                                               # MRGP.py:66
    with pytest.raises(ValueError):
        M([x, y], 30, idx, le, [mk, mk])                                                # MRGP.py:75
    with pytest.raises(ValueError):
        M([x, y], 30, idx, le, mk, interval_factor=[1, 1])                              # MRGP.py:102
    with pytest.raises(ValueError):
        M([x, y], 30, idx, le, mk, basis_interval_obj=[BasisInterval()])                # MRGP.py:123
    m = M([x, y], 30, idx, le, mk)
    with pytest.raises(ValueError):
        m.get_predicted_mean(x, index_set_obj=IndexSetUniform(64, 3, 2))               # MRGP.py:758-760
    with pytest.raises(ValueError):
        m.get_predicted_mean(x, index_set_obj=IndexSetUniform(64, 2, 3))               # MRGP.py:762-764
    assert len(m.posterior_obj) == 3 and m.stats_obj[2].scale_axis_mean[3].shape == (2, 30)
    assert m.phi_x[1][1].shape == (32, 30) and m.stats_obj[1].latent_f_var[0].shape == (32, 1)


class _ThreadComm(object):
    """In-process stand-in for the NCCL all-reduce: the ranks are Python threads driving their own handle (all on
    cuda:0); tensors are summed through host-synchronised device ops."""

    def __init__(self, world):
        import threading
        self.world, self.slots, self.barrier = world, [None] * world, threading.Barrier(world)
        self.local = threading.local()

    def all_reduce(self, tensor, op):
        import torch
        torch.cuda.synchronize()
        self.slots[self.local.rank] = tensor
        self.barrier.wait()
        stack = torch.stack([t for t in self.slots])
        res = stack.max(0).values if op == 'max' else stack.sum(0)
        torch.cuda.synchronize()
        self.barrier.wait()
        tensor.copy_(res)
        torch.cuda.synchronize()
        self.barrier.wait()

    def all_gather_bytes(self, payload):
        self.slots[self.local.rank] = bytes(payload)
        self.barrier.wait()
        out = list(self.slots)
        self.barrier.wait()
        return out

    def sync(self):
        self.barrier.wait()


@pytest.mark.parametrize('fi,world,exchange', [(False, 2, 'nccl'), (True, 3, 'nccl'), (False, 3, 'peer'), (True, 2, 'peer')])
def test_sample_sharding_matches_single_handle(fi, world, exchange):
    """The multi-GPU decomposition (sample chunks + summed region statistics) on one device: every rank must
    end with the state of the unsharded run (up to the summation order of the exchanged statistics).  'peer':
    the library's own exchange through mapped arenas (here: handles of one process, plain pointers), driven phase
    by phase (the captured-graph form runs between processes, next test); 'nccl': the caller-side all-reduce arm."""
    import threading
    from cimrgp_b200.distributed import ShardedEngine, chunk_bounds
    from cimrgp_b200.engine import Engine
    n, res, M = 30000, 6, 30
    x, y = workloads.workload1(n)
    xs = (x - x.mean(0)) / x.std(0)
    offsets = O.uniform_offsets(n, res, 2)
    mode = 'fi' if fi else 'ci'
    ref = Engine(xs, y, offsets, M, mode=mode)
    ref.sweep(3)
    ref.synchronize()
    comm = _ThreadComm(world)
    engines, errors = [None] * world, []

    def work(rank):
        try:
            comm.local.rank = rank
            e = ShardedEngine(xs, y, offsets, M, rank, world, comm=comm, exchange=exchange, mode=mode)
            for _ in range(3):
                e.sweep(1, use_graph=False)
            e.synchronize()
            engines[rank] = e
        except Exception as ex:   # pragma: no cover
            errors.append(ex)
            comm.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    assert sum(hi - lo for lo, hi in (chunk_bounds(n, world, r) for r in range(world))) == n
    want = ref.state(latent=False)
    for e in engines:
        compare(e.state(), want, rtol=1e-9)
    a, b = engines[0].state(), engines[-1].state()
    for k in a:
        assert np.array_equal(a[k], b[k]), k     # replicated small-matrix steps: identical on every rank


@pytest.mark.parametrize('mode,world', [('ci', 2), ('fi', 3)])
def test_peer_exchange_between_processes(mode, world):
    """The product form of the multi-GPU path: one process per rank, arenas mapped through CUDA IPC, the sweep one
    CUDA graph with the exchanges inside.  The ranks share cuda:0 here (the parity box has one GPU)."""
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sharded_worker.py')
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(port), worker, mode]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    import re
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    oks = re.findall(r'rank \d+ ok ', res.stdout)      # the ranks' lines may interleave
    assert len(oks) == world and 'MISMATCH' not in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.parametrize('fi,group_size', [(False, 4), (True, 4), (False, 0)])
def test_series_batch_equals_separate_models(fi, group_size):
    """Config 5 front end: models carved out of one device allocation, sweeps interleaved over a stream pool, must
    give exactly the states of models built and fitted one by one (same launch geometry -> bit for bit)."""
    from cimrgp_b200 import LaplacianEigenpairs, MaternKernel, SeriesBatch, IndexSetUniform
    from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
    n, res, M, S = 2048, 5, 30, 6
    xs, ys = [], []
    for s in range(S):
        x, y = workloads.workload1(n, seed=10 + s)
        xs.append(x)
        ys.append(y)
    batch = SeriesBatch(xs, ys, M, res, LaplacianEigenpairs(), MaternKernel(nu=1, l=1, sf=1), forced_independence=fi,
                        n_streams=3, group_size=group_size)
    batch.fit(3)
    batch.fit(1)
    for s in (0, 3, 5):
        m = MultiResolutionGaussianProcess([xs[s], ys[s]], M, IndexSetUniform(n, res, 2), LaplacianEigenpairs(),
                                           MaternKernel(nu=1, l=1, sf=1), forced_independence=fi, n_ctas=2)
        m.fit(4, None)
        a, b = batch[s]._engine.state(), m._engine.state()
        for k in b:
            assert np.array_equal(a[k], b[k]), (s, k)
    xt = np.atleast_2d(np.linspace(1, 3, 500)).T
    assert batch[1].get_predicted_mean(xt).shape == (500, 2)


@pytest.mark.parametrize('name,fi', [('predvar_ci', False), ('predvar_fi', True)])
def test_indexed_second_moment_and_test_likelihood_match_reference(name, fi):
    """O1, index-set form (MRGP.py:863-932, 825-831) against the reference's own output."""
    from cimrgp_b200 import IndexSetUniform
    g = load(name)
    m = build(g['x'], g['y'], int(g['meta.M']), int(g['meta.resolution']), fi)
    m.fit(int(g['meta.sweeps']), None)
    xt, yt = g['pred.x'], g['pred.y']
    idx_t = IndexSetUniform(xt.shape[0], int(g['meta.resolution']), 2)
    before = m._engine.state()
    assert mismatch(m.get_predicted_mean(xt, index_set_obj=idx_t), g['pred.mean_indexed'], RTOL) is None
    assert mismatch(m.get_central_moment2(xt), g['pred.var_global'], RTOL) is None
    assert mismatch(m.get_central_moment2(xt, index_set_obj=idx_t), g['pred.var_indexed'], RTOL) is None
    got = m.get_test_likelihood([xt, yt], index_set_obj=idx_t)
    assert abs(got - float(g['pred.test_likelihood_indexed'])) <= RTOL * abs(float(g['pred.test_likelihood_indexed']))
    after = m._engine.state()
    for k in before:
        assert np.array_equal(before[k], after[k]), k       # prediction leaves the model untouched


SHARED = [('shared_ci_nb', False, True, False), ('shared_ci_sn', False, False, True), ('shared_ci_ss', False, False, False),
          ('shared_fi_nb', True, True, False), ('shared_fi_sn', True, False, True), ('shared_fi_ss', True, False, False)]


@pytest.mark.parametrize('name,fi,noise_rs,bias_rs', SHARED)
def test_shared_noise_and_bias_variants_match_reference(name, fi, noise_rs, bias_rs):
    """The three other variants of the bias / noise update (Posteriors.py:94-110, 150-211, 358-374, 414-475)."""
    g = load(name)
    m = build(g['x'], g['y'], int(g['meta.M']), int(g['meta.resolution']), fi, noise_region_specific=noise_rs,
              bias_region_specific=bias_rs)
    done = 0
    for k in g['meta.checkpoints']:
        m.fit(int(k) - done, None)
        done = int(k)
        st = m._engine.state()
        compare(st, expand_shared(split(g, 'k%d.' % k), st), atol_abs=ATOL_ABS_SHARED)
    # public shapes (Posteriors.py:17-25, Stats.py:29-49)
    R1 = m.n_regions[1]
    assert np.shape(m.posterior_obj[1].noise_gamma_shape) == ((R1,) if noise_rs else ())
    assert np.shape(m.stats_obj[1].noise_mean) == ((R1,) if noise_rs else ())
    assert np.shape(m.posterior_obj[1].bias_normal_mean) == ((R1, 2) if bias_rs else (2,))
    assert np.shape(m.stats_obj[1].bias_var) == ((R1,) if bias_rs else ())
    if 'pred.x' in g.files:
        from cimrgp_b200 import IndexSetUniform
        xt = g['pred.x']
        assert mismatch(m.get_predicted_mean(xt), g['pred.mean_global'], RTOL) is None
        idx_t = IndexSetUniform(xt.shape[0], int(g['meta.resolution']), 2)
        assert mismatch(m.get_predicted_mean(xt, index_set_obj=idx_t), g['pred.mean_indexed'], RTOL) is None
        assert mismatch(m.get_central_moment2(xt), g['pred.var_global'], RTOL) is None


@pytest.mark.parametrize('name,noise_rs,bias_rs', [('shared_ci_ss_elbo', False, False), ('shared_ci_nb_elbo', True, False)])
def test_elbo_of_shared_variants_matches_reference(name, noise_rs, bias_rs):
    g = load(name)
    n_iter = g['lower_bound'].shape[0]
    m = build(g['x'], g['y'], int(g['meta.M']), int(g['meta.resolution']), False, noise_region_specific=noise_rs,
              bias_region_specific=bias_rs)
    m.fit(n_iter=n_iter, tol=1e-300, min_iter=n_iter)
    assert mismatch(np.array(m.lower_bound_terms), g['terms'], RTOL) is None
    assert mismatch(np.array(m.lower_bound_layer), g['lower_bound_layer'], RTOL) is None


def test_adaptive_inputs_with_an_injected_input_model_match_reference():
    """adaptive_inputs=True trains on the warped coordinate z (Inputs.py:12, 50-51) and
    routes test inputs through input_model.predict (MRGP.py:733-739); the GPy warp model itself is out of scope, a
    deterministic interpolating model is injected on both sides."""
    from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel
    from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
    g = load('c2_ci_warp')
    x, y = g['x'], g['y']
    m = MultiResolutionGaussianProcess(
        train_xy=[x, y], n_basis=int(g['meta.M']), index_set_obj=IndexSetUniform(x.shape[0], int(g['meta.resolution']), 2),
        basis_function_obj=LaplacianEigenpairs(), spectral_density_obj=MaternKernel(nu=1, l=1, sf=1),
        adaptive_inputs=True, standard_normalized_inputs=True, input_model=workloads.InterpInputModel(x))
    compare(m._engine.state(), split(g, 'k0.'))
    done = 0
    for k in g['meta.checkpoints']:
        m.fit(int(k) - done, None)
        done = int(k)
        compare(m._engine.state(), split(g, 'k%d.' % k))
    xt = g['pred.x']
    assert mismatch(m.get_predicted_mean(xt), g['pred.mean_global'], RTOL) is None
    idx_t = IndexSetUniform(xt.shape[0], int(g['meta.resolution']), 2)
    assert mismatch(m.get_predicted_mean(xt, index_set_obj=idx_t), g['pred.mean_indexed'], RTOL) is None
    assert mismatch(m.get_central_moment2(xt), g['pred.var_global'], RTOL) is None


@pytest.mark.parametrize('name', ['c1_ci_adaptive_elbo', 'n600_ci_adaptive_elbo'])
def test_elbo_under_adaptive_intervals_matches_reference(name):
    """fit(n_iter, tol) with BasisInterval: the data term of the bound uses the re-learnt basis of every layer while
    its targets were inferred with the basis of the beginning of the step (MRGP.py:535-569 after :632-641)."""
    from cimrgp_b200 import BasisInterval
    g = load(name)
    n_iter = g['lower_bound'].shape[0]
    m = build(g['x'], g['y'], int(g['meta.M']), int(g['meta.resolution']), False,
              basis_interval_obj=BasisInterval(opt_interval_factor=(1, 1.2)))
    m.fit(n_iter=n_iter, tol=1e-300, min_iter=n_iter)
    assert mismatch(np.array(m.lower_bound_terms), g['terms'], RTOL) is None
    assert mismatch(np.array(m.lower_bound_layer), g['lower_bound_layer'], RTOL) is None


@pytest.mark.parametrize('fi', [False, True])
def test_smallest_compiled_basis_against_oracle(fi):
    """n_basis = 8: the smallest instantiation of the streaming kernels, the one-warp omega solve and the invariant
    builders (30 and 20 are covered by the goldens and the tests above, 40 by config 2 with the block solver)."""
    x, y = workloads.workload1(2500)
    offsets = O.uniform_offsets(2500, 4, 2)
    ora = O.OracleMRGP(x, y, 8, offsets, mode='fi' if fi else 'ci')
    m = build(x, y, 8, 4, fi)
    for _ in range(3):
        ora.sweep()
    m.fit(3, None)
    compare(m._engine.state(), ora.state())
    if not fi:
        assert mismatch(m._engine.elbo(), ora.elbo()[2], RTOL) is None


@pytest.mark.parametrize('fi', [False, True])
def test_single_layer_model_against_oracle(fi):
    """resolution 0: one layer, one region - no closed-form layers, no third stream, nothing to propagate."""
    x, y = workloads.workload1(777)
    offsets = O.uniform_offsets(777, 0, 2)
    ora = O.OracleMRGP(x, y, 20, offsets, mode='fi' if fi else 'ci')
    m = build(x, y, 20, 0, fi)
    for _ in range(3):
        ora.sweep()
    m.fit(3, None)
    compare(m._engine.state(), ora.state())
    xt = np.atleast_2d(np.linspace(1, 3, 100)).T
    assert mismatch(m.get_predicted_mean(xt), ora.predict_mean(xt), RTOL) is None


def test_divider_three_and_block_omega_solver_against_oracle():
    """divider 3 (27 regions on the finest layer, none nested in thirds of 5000) with n_basis = 40: the block version
    of the omega solve, the table built by k_ard, closed-form statistics over three-way splits; with the ELBO."""
    from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel
    from cimrgp_b200.MRGP import MultiResolutionGaussianProcess
    x, y = workloads.workload1(5000)
    idx = IndexSetUniform(5000, 3, 3)
    m = MultiResolutionGaussianProcess([x, y], 40, idx, LaplacianEigenpairs(), MaternKernel(1, 1, 1))
    ora = O.OracleMRGP(x, y, 40, idx.offsets, mode='ci')
    for _ in range(3):
        ora.sweep()
    m.fit(3, None)
    compare(m._engine.state(), ora.state())
    assert mismatch(m._engine.elbo(), ora.elbo()[2], RTOL) is None


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs 3, 4 and 5 at their full sizes ("pin configs 3 and 4 at full size")
# ------------------------------------------------------------------------------------------------------------
# Fields downstream of the chain of permutation weights omega(0) ... omega(J-1).  The reference solves for omega with
# MINPACK hybrd at xtol 1.5e-8 (Stats.py:413); the device computes the exact doubly-stochastic scaling.  Through 10
# layers and 3 sweeps the reference's own solver slack accumulates to 2.4e-6 on S.B / S.logC at N = 1e6 (measured: the
# oracle with the exact scaling agrees with the device to 1e-13, with the reference-faithful fsolve oracle to 2.4e-6).
OMEGA_CHAIN = ('S.B', 'S.kappa', 'S.logC', 'S.rho', 'S.omega')


def _oracle_vs_device_at_size(n, resolution, fi, checkpoints, elbo, omega_chain_rtol=None):
    x, y = workloads.workload1(n)
    ora = O.OracleMRGP(x, y, 30, O.uniform_offsets(n, resolution, 2), mode='fi' if fi else 'ci')
    exact = None
    if omega_chain_rtol is not None:
        exact = O.OracleMRGP(x, y, 30, O.uniform_offsets(n, resolution, 2), mode='ci', omega_solver='sinkhorn')
    m = build(x, y, 30, resolution, fi)
    done = 0
    for k in checkpoints:
        for _ in range(k - done):
            ora.sweep()
            if exact is not None:
                exact.sweep()
        m.fit(k - done, None)
        done = k
        got, ref = m._engine.state(), ora.state()
        if exact is None:
            compare(got, ref)
        else:
            # reference-faithful oracle: 1e-6 everywhere except the omega chain (its own solver tolerance, see above)
            compare({k_: v for k_, v in got.items() if k_ not in OMEGA_CHAIN}, {k_: v for k_, v in ref.items() if k_ not in OMEGA_CHAIN})
            for k_ in OMEGA_CHAIN:
                assert mismatch(got[k_], ref[k_], omega_chain_rtol, atol_scale=omega_chain_rtol) is None, k_
            # oracle with the exact scaling: the shared state to 1e-9
            ex = exact.state()
            for k_ in [k_ for k_ in ex if k_.startswith('S.')]:
                assert mismatch(got[k_], ex[k_], 1e-9, atol_scale=1e-9 if k_ == 'S.omega' else 1e-12) is None, k_
    if elbo:
        assert mismatch(m._engine.elbo(), ora.elbo()[2], RTOL) is None     # per layer and per term
    return m, ora


@pytest.mark.parametrize('fi', [False, True])
def test_config3_full_size_against_oracle(fi):
    """BASELINE config 3: N = 1e5, 8 resolutions (255 regions), M = 30: every state array after 1 and 3 sweeps and
    (ci) the six ELBO terms per layer against the CPU oracle at the SURVEY §8c rule."""
    _oracle_vs_device_at_size(100000, 7, fi, (1, 3), elbo=not fi)


def test_config4_full_size_against_oracle():
    """BASELINE config 4 (the headline): N = 1e6, 10 resolutions (1023 regions), M = 30, ci: state after 1 and 3 sweeps
    and the ELBO per term against the CPU oracle (the oracle needs ~45 s per sweep at this size)."""
    _oracle_vs_device_at_size(1000000, 9, False, (1, 3), elbo=True, omega_chain_rtol=1e-5)


def test_config4_closed_form_statistics_at_full_size_on_a_non_inert_state(monkeypatch):
    """The cancellation test of the Gram-based closed forms at N = 1e6: on a state whose upper layers carry signal
    (coefficients and biases set by hand) the captured sweep (sufficient statistics / closed-form layer statistics)
    must agree with the sweep that streams every layer over the samples (MRGP_STREAM_ALL=1) to 1e-9."""
    from cimrgp_b200 import _lib
    n, res = 1000000, 9
    x, y = workloads.workload1(n)
    rng = np.random.RandomState(7)
    pert = [(0.05 * rng.standard_normal((2 ** j, 30, 2)), 0.3 * rng.standard_normal((2 ** j, 2))) for j in range(res + 1)]
    states = []
    for stream_all in ('0', '1'):
        monkeypatch.setenv('MRGP_STREAM_ALL', stream_all)
        m = build(x, y, 30, res, False)
        m.fit(2, None)
        for j in range(1, res + 1):
            m._engine.put(j, _lib.F_A, pert[j][0])
            m._engine.put(j, _lib.F_BIAS_MEAN, pert[j][1])
        m.fit(1, None)       # (a second sweep would find the upper layers inert again: their noise precision is ~1e-45)
        states.append(m._engine.state(latent=False))
        del m
    assert np.max(np.abs(states[1]['L5.ytil'])) > 1e-3          # the upper layers really carry signal here
    compare(states[0], states[1], rtol=1e-9)


@pytest.mark.parametrize('name,fi', [('c5_ci', False), ('c5_fi', True)])
def test_config5_series_matches_reference_golden(name, fi):
    """One series of BASELINE config 5 (N = 2048, 6 resolutions, M = 30) against the unmodified reference."""
    g = load(name)
    m = build(g['x'], g['y'], int(g['meta.M']), int(g['meta.resolution']), fi)
    compare(m._engine.state(), split(g, 'k0.'))
    done = 0
    for k in g['meta.checkpoints']:
        m.fit(int(k) - done, None)
        done = int(k)
        compare(m._engine.state(), split(g, 'k%d.' % k))


# ------------------------------------------------------------------------------------------------------------
# boundary generality: any number of basis functions, public state of SURVEY.md §8b
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('m_basis,fi', [(25, False), (25, True), (13, False), (45, False), (45, True), (3, False)])
def test_any_number_of_basis_functions_against_oracle(m_basis, fi):
    """The reference accepts any n_basis (MRGP.py:16-66).  The streaming kernels are instantiated for 8 / 20 / 30 / 40 /
    48 functions and pad; the fused sweep pads its omega solve with an identity block (n_basis <= 32), n_basis > 32
    takes the multi-kernel sweep with the block solver."""
    x, y = workloads.workload1(3000)
    offsets = O.uniform_offsets(3000, 4, 2)
    ora = O.OracleMRGP(x, y, m_basis, offsets, mode='fi' if fi else 'ci')
    m = build(x, y, m_basis, 4, fi)
    for _ in range(3):
        ora.sweep()
    m.fit(3, None)
    compare(m._engine.state(), ora.state())
    if not fi:
        assert mismatch(m._engine.elbo(), ora.elbo()[2], RTOL) is None
    xt = np.atleast_2d(np.linspace(1, 3, 200)).T
    assert mismatch(m.get_predicted_mean(xt), ora.predict_mean(xt), RTOL) is None


@pytest.mark.parametrize('fi', [False, True])
def test_public_state_priors_and_targets(fi):
    """prior_obj, shared_prior, y_mean, y_var (MRGP.py:181-226, 262-271, 650-652; Priors.py)."""
    x, y = workloads.workload1(1500)
    offsets = O.uniform_offsets(1500, 3, 2)
    m = build(x, y, 20, 3, fi, snr_ratio=5.)
    assert m.y_mean[2][1] == [] and m.y_var[0][0] == []                       # MRGP.py:262-271
    pr = m.prior_obj
    assert len(pr) == 4 and len(pr[2].scale_precision) == 4 and pr[2].scale_precision[3].shape == (20,)
    assert pr[1].noise_gamma_shape == [1e-45] * 2 and pr[1].noise_gamma_scale == [1e-45 + 1] * 2
    y_var0 = (np.linalg.norm(y) ** 2) / 1500 - np.dot(np.mean(y, axis=0), np.mean(y, axis=0))
    assert abs(pr[0].noise_gamma_scale[0] - (1e-45 + 1) * y_var0 / 5.) < 1e-12    # MRGP.py:195-199, 966-971
    assert pr[3].bias_normal_precision == [1e-45] * 8 and np.array_equal(pr[3].bias_normal_mean[7], np.zeros(2))
    if fi:
        assert pr[2].axis_bingham_rho[1].shape == (20, 2) and abs(pr[2].axis_bingham_log_const[0][0] - 1.9189385332046727) < 1e-12
        with pytest.raises(AttributeError):
            m.shared_prior
    else:
        sp = m.shared_prior
        assert np.array_equal(sp.axis_bingham_b, np.zeros((20, 2, 2))) and np.allclose(sp.axis_bingham_rho, 0.5)
        assert abs(sp.axis_bingham_log_const[3] - 1.9189385332046727) < 1e-12            # SURVEY.md §8c KAT
        assert np.array_equal(sp.ard_gamma_shape, 1e-45 * np.ones(20)) and np.array_equal(sp.ard_gamma_scale, 1e-45 * np.ones(20) / 1.0)
    ora = O.OracleMRGP(x, y, 20, offsets, mode='fi' if fi else 'ci', snr_ratio=5.)
    ora.sweep()
    ora.sweep()
    m.fit(2, None)
    ym, yv = m.y_mean, m.y_var
    for j in range(4):
        ly = ora.layers[j]
        if j == 0 and not fi:
            assert np.array_equal(ym[0][0], y)                              # LatentOutputs.py:6-9: one entry with all of Y
        else:
            for l in range(ly.R):
                ref = ly.y_target[ly.off[l]:ly.off[l + 1]]
                assert mismatch(ym[j][l], ref, RTOL) is None
        assert mismatch(np.array(yv[j], dtype=np.float64), ly.y_var, RTOL) is None


@pytest.mark.parametrize('n', [6000, 150000])
def test_fused_sweep_guard_falls_back_to_streamed_statistics(n, monkeypatch):
    """Layer 0 of the fused sweep forms sum |r|^2 from the sufficient statistics of y as a difference of large terms;
    below sum |r|^2 / sum |y|^2 = 1e-5 (high-SNR data that layer 0 explains almost exactly) a guard hands the statistics of
    layer 0 to a streamed pass (small layer-0 regions: k_l0_fix_small; large ones: the gated phase-B kernel).  Variational
    inference approaches such a fit only over hundreds of sweeps, so the test raises the threshold instead
    (MRGP_CHAIN_GUARD) and checks the fallback itself: guard tripped, results equal to the multi-kernel sweep, which streams
    layer 0 anyway."""
    from cimrgp_b200 import _lib
    x, y = workloads.workload1(n)
    states = []
    for fused, thr in (('1', '0.9'), ('1', '1e-5'), ('0', '1e-5')):
        monkeypatch.setenv('MRGP_FUSED', fused)
        monkeypatch.setenv('MRGP_CHAIN_GUARD', thr)
        m = build(x, y, 12, 2, False)
        m.fit(4, None)
        states.append(m._engine.state(latent=False))
        if fused == '1':
            guard = m._engine.get(-1, _lib.F_FUSED_GUARD, (2,))
            assert guard[1] == (1.0 if thr == '0.9' else 0.0) and 1e-3 < guard[0] < 0.9, guard
        del m
    compare(states[0], states[2], rtol=1e-9)       # streamed fallback == multi-kernel sweep
    compare(states[1], states[2], rtol=1e-9)       # closed form == multi-kernel sweep
    assert any(not np.array_equal(states[0][k], states[1][k]) for k in states[0])   # the fallback did run (other summation order)


def test_prefetched_observations_match_synchronous_uploads():
    """mrgp_prefetch_observations_host: a stream of data sets through the double-buffered upload (copy of the next set beside
    the sweep of the current one) gives the state of the synchronous uploads, bit for bit; one set can be pending."""
    from cimrgp_b200 import _lib
    n = 60000                                   # layer 0 is one region of 60000 samples: the streaming statistics pass
    x, y = workloads.workload1(n)
    rng = np.random.default_rng(5)
    ys = [y + 0.05 * rng.standard_normal(y.shape) for _ in range(3)]
    a, b = build(x, y, 30, 5, False), build(x, y, 30, 5, False)
    a.fit(2, None)
    b.fit(2, None)
    ea, eb = a._engine, b._engine
    eb.prefetch_observations(ys[0])
    with pytest.raises(_lib.MrgpError):
        eb.prefetch_observations(ys[1])         # the first set has not been taken over yet
    for k in range(3):
        ea.set_observations(ys[k])
        ea.sweep(2)
        eb.refresh_statistics()                 # takes over set k (waits for its copy) ...
        if k + 1 < 3:
            eb.prefetch_observations(ys[k + 1])  # ... and the copy of set k + 1 runs beside the sweeps of set k
        eb.sweep(2)
        sa, sb = ea.state(), eb.state()
        for key in sa:
            assert np.array_equal(sa[key], sb[key]), (k, key)
    assert np.array_equal(ea.elbo(), eb.elbo())


@pytest.mark.parametrize('n,res', [(20000, 7), (150000, 9)])
def test_fused_sweep_is_bit_reproducible(n, res):
    """Two identical models swept 25 times (cluster of 4 / of 16 CTAs: DSMEM pushes, mbarrier hand-offs, background worker
    steps): every state array and the ELBO agree bit for bit - a race in the hand-offs would show as a difference."""
    x, y = workloads.workload1(n)
    states = []
    for rep in range(3):
        m = build(x, y, 30, res, False)
        m.fit(25, None)
        st = m._engine.state()
        st['elbo'] = m._engine.elbo()
        states.append(st)
        del m
    for st in states[1:]:
        for key in states[0]:
            assert np.array_equal(states[0][key], st[key], equal_nan=True), key


def test_elbo_async_matches_blocking_call():
    x, y = workloads.workload1(20000)
    m = build(x, y, 30, 6, False)
    m.fit(3, None)
    e = m._engine
    ref = e.elbo()
    e.elbo_async(0)
    e.sweep(1)
    e.elbo_async(1)
    assert np.array_equal(e.elbo_result(0), ref)
    assert np.array_equal(e.elbo_result(1), e.elbo())


def test_switch_from_the_fused_to_the_multi_kernel_sweep(monkeypatch):
    """Engine.set_fused(False) in the middle of a fit (what the host does when the accelerated omega solve of the fused sweep
    runs out of budget): same state arrays, results equal to a model that took the multi-kernel sweep throughout."""
    x, y = workloads.workload1(20000)
    a = build(x, y, 30, 6, False)
    a.fit(2, None)
    assert a.omega_solve_report().max() < 100
    a._engine.set_fused(False)
    a.fit(2, None)
    monkeypatch.setenv('MRGP_FUSED', '0')
    b = build(x, y, 30, 6, False)
    b.fit(4, None)
    compare(a._engine.state(latent=False), b._engine.state(latent=False), rtol=1e-9)
