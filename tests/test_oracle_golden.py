"""Pin the CPU oracle (oracle/mrgp_oracle.py) to fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import mrgp_oracle as O
from parity import ATOL_ABS_SHARED, assert_state_close, expand_shared, mismatch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# the reference's own fsolve is converged to ~1e-8; the oracle calls the same
# solver, so the two agree far below that.
RTOL = 1e-9


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def split(g, prefix):
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}


def test_kat_saddle_point():
    k = load('kats')
    for p in ('saddle2', 'saddle3'):
        lc, rho = O.log_partition_saddle(k[p + '.kappa'])
        assert mismatch(lc, k[p + '.logC'], 1e-14) is None
        assert mismatch(rho, k[p + '.rho'], 1e-13) is None
    # SURVEY.md §8c known answers
    lc, rho = O.log_partition_saddle(np.zeros(2))
    assert abs(lc - 1.9189385332046727) < 1e-15 and np.allclose(rho, 0.5, atol=1e-15)
    lc, rho = O.log_partition_saddle(np.array([3., 1., 0.]))
    assert abs(lc - 4.200667217497404) < 1e-14
    assert np.allclose(rho, [0.5904957658136983, 0.24103618159199133, 0.16846805259431055], atol=1e-14)


def test_kat_bingham_guard():
    k = load('kats')
    B, kap, ax, rho, lc, n = O.bingham_batch(k['bingham.B_in'])
    assert n == k['bingham.B_in'].shape[0]
    assert mismatch(B, k['bingham.B'], 1e-14) is None
    assert mismatch(kap, k['bingham.kappa'], 1e-13) is None
    assert mismatch(rho, k['bingham.rho'], 1e-12) is None
    assert mismatch(lc, k['bingham.logC'], 1e-13) is None
    assert mismatch(O.axis_cov_from(rho, ax), k['bingham.axis_cov'], 1e-12) is None


def test_kat_basis_and_spectral():
    k = load('kats')
    phi = O.eigenfunctions(k['basis.x'], np.full((k['basis.x'].shape[0], 1), 2.0), 40)
    assert mismatch(phi, k['basis.phi'], 1e-14) is None
    lam = O.eigenvalues(np.array([[2.0]]), 40)[0]
    assert mismatch(lam, k['basis.lam'], 1e-15) is None
    assert mismatch(O.matern_spectral(np.sqrt(lam), 1, 1, 1), k['basis.S'], 1e-14) is None
    assert mismatch(O.matern_spectral(np.sqrt(lam), 2.5, .7, 1.3), k['basis.S_nu2.5_l0.7_sf1.3'], 1e-14) is None
    assert abs(lam[2] - 5.551652475612764) < 1e-14


def test_kat_index_sets_bit_exact():
    k = load('kats')
    for (n, res, div) in k['index.cases']:
        offs = O.uniform_offsets(n, res, div)
        assert len(offs) == res + 1
        for j, o in enumerate(offs):
            ref = k['index.%d_%d_%d.L%d' % (n, res, div, j)]
            assert o.dtype == np.int64 and np.array_equal(o, ref)
    with pytest.raises(ValueError):
        O.uniform_offsets(10, 5, 2)


def test_kat_omega():
    k = load('kats')
    for lw, eta in zip(k['omega.log_omega_hat'], k['omega.ln_eta']):
        m = lw.shape[0]
        ref = np.exp(eta[:m, None] + eta[None, m:] + lw)
        assert mismatch(O.omega_fsolve(lw), ref, 1e-10) is None
        # the exact doubly-stochastic scaling differs from fsolve's answer by fsolve's own tolerance
        assert mismatch(O.omega_sinkhorn(lw), ref, 1e-6) is None


CASES = [
    ('c1_ci', 'ci', None), ('c1_fi', 'fi', None), ('n2000_ci', 'ci', None), ('n2000_fi', 'fi', None),
    ('c2_ci', 'ci', None), ('c2_fi', 'fi', None), ('n600_ci_snr', 'ci', dict(snr_ratio=10.)),
    # shared noise and / or bias (Posteriors.py:150-211, 414-475)
    ('shared_ci_nb', 'ci', dict(noise_region_specific=True, bias_region_specific=False)),
    ('shared_ci_sn', 'ci', dict(noise_region_specific=False, bias_region_specific=True)),
    ('shared_ci_ss', 'ci', dict(noise_region_specific=False, bias_region_specific=False)),
    ('shared_fi_nb', 'fi', dict(noise_region_specific=True, bias_region_specific=False)),
    ('shared_fi_sn', 'fi', dict(noise_region_specific=False, bias_region_specific=True)),
    ('shared_fi_ss', 'fi', dict(noise_region_specific=False, bias_region_specific=False)),
]


@pytest.mark.parametrize('name,mode,kw', CASES)
def test_sweeps_match_reference(name, mode, kw):
    g = load(name)
    x, y = g['x'], g['y']
    m = O.OracleMRGP(x, y, int(g['meta.M']), O.uniform_offsets(x.shape[0], int(g['meta.resolution']), 2),
                     mode=mode, **(kw or {}))
    assert_state_close(m.state(), expand_shared(split(g, 'k0.'), m.state()), 1e-13, skip=('kappa',))
    done = 0
    for k in g['meta.checkpoints']:
        for _ in range(int(k) - done):
            m.sweep()
        done = int(k)
        # kappa: only the unclamped eigenvalues enter the model; tiny trailing eigenvalues of rank-1 B
        # are roundoff (SURVEY.md §7 "roundoff-dependent branches"), compare them at absolute scale.
        st = m.state()
        ref = expand_shared(split(g, 'k%d.' % k), st)
        assert_state_close(st, ref, RTOL, skip=('kappa',), atol_abs=ATOL_ABS_SHARED if name.startswith('shared') else None)
        for key in ref:
            if key.endswith('kappa'):
                assert mismatch(st[key], ref[key], RTOL, atol_scale=1e-13) is None, key
    if 'pred.x' in g.files:
        xt = g['pred.x']
        res = int(g['meta.resolution'])
        assert mismatch(m.predict_mean(xt), g['pred.mean_global'], 1e-9) is None
        assert mismatch(m.predict_mean(xt, O.uniform_offsets(xt.shape[0], res, 2)), g['pred.mean_indexed'], 1e-9) is None
        assert mismatch(m.predict_var(xt), g['pred.var_global'], 1e-9) is None


def test_ci_upper_layers_are_inert():
    """Behavioural KAT: A and bias of every layer >= 1 stay exactly zero in ci mode."""
    g = load('c1_ci')
    for j in range(1, int(g['meta.resolution']) + 1):
        assert not np.any(g['k15.L%d.A' % j]) and not np.any(g['k15.L%d.bias_mean' % j])


def test_adaptive_intervals_match_reference():
    g = load('c1_ci_adaptive')
    x, y = g['x'], g['y']
    m = O.OracleMRGP(x, y, int(g['meta.M']), O.uniform_offsets(x.shape[0], int(g['meta.resolution']), 2),
                     mode='ci', adaptive=dict(use_prior=True, opt_interval_factor=(1, 1.2)))
    done = 0
    for k in g['meta.checkpoints']:
        for _ in range(int(k) - done):
            m.sweep()
        done = int(k)
        # fminbound stops at xtol=1e-5: identical calls, identical steps
        assert_state_close(m.state(), split(g, 'k%d.' % k), 1e-7, skip=('kappa',))


def test_elbo_terms_match_reference():
    g = load('c1_ci_elbo')
    x, y = g['x'], g['y']
    n_iter = g['lower_bound'].shape[0]
    m = O.OracleMRGP(x, y, int(g['meta.M']), O.uniform_offsets(x.shape[0], int(g['meta.resolution']), 2), mode='ci')
    m.fit(n_iter=n_iter, tol=1e-300, min_iter=n_iter)
    assert mismatch(np.array(m.lower_bound_terms), g['terms'], 1e-9) is None
    assert mismatch(np.array(m.lower_bound_layer), g['lower_bound_layer'], 1e-9) is None
    assert mismatch(np.array(m.lower_bound), g['lower_bound'], 1e-9) is None
