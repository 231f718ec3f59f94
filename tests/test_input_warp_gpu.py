"""(f).3 input-warp GP on the device against its NumPy twin (parity unpinned: GPy is absent, oracle/gp_rbf_oracle.py)."""
import random

import numpy as np
import pytest

from oracle import gp_rbf_oracle as G
import workloads

pytestmark = pytest.mark.gpu


def _warp(n, seed=0):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.standard_normal((n, 1)) ** 3, axis=0)
    z = np.linspace(x.min(), x.max(), n)[:, None]
    return x, z


def test_objective_and_gradient_match_the_twin():
    import torch
    from cimrgp_b200.RegressionInput import GP_RBF
    x, z = _warp(400)
    X, Y = (x - x.mean(0)) / x.std(0), (z - z.mean(0)) / z.std(0)
    D2 = G.sqdist(X, X)
    Xd, Yd = torch.as_tensor(X, device='cuda'), torch.as_tensor(Y, device='cuda')
    D2d = torch.cdist(Xd, Xd, compute_mode='donot_use_mm_for_euclid_dist') ** 2
    for theta in (np.log([1.3, 0.7, 0.02]), np.log([0.4, 0.05, 0.2])):
        f, g = GP_RBF._objective(torch, theta, D2d, Yd)
        fo, go = G.objective(theta, D2, Y)
        assert abs(f - fo) <= 1e-9 * max(1.0, abs(fo))
        assert np.allclose(g, go, rtol=1e-7, atol=1e-7 * np.abs(go).max())


def test_fit_and_predictions_match_the_twin():
    from cimrgp_b200.RegressionInput import GP_RBF
    x, z = _warp(500, 3)
    m, o = GP_RBF(), G.GPRBFOracle()
    m.fit([x, z])
    o.fit([x, z])
    assert abs(m.nll - o.nll) <= 1e-6 * max(1.0, abs(o.nll))
    xs = np.linspace(x.min(), x.max(), 1000)[:, None]
    assert np.abs(m.predict(xs) - o.predict(xs)).max() <= 1e-4 * (z.max() - z.min())


def test_adaptive_inputs_end_to_end():
    """MultiResolutionGaussianProcess(adaptive_inputs=True) fits its own warp model (Inputs.py:20-47) and predicts through it
    like a model that was handed the twin's (MRGP.py:731-741)."""
    from cimrgp_b200 import IndexSetUniform, LaplacianEigenpairs, MaternKernel, MultiResolutionGaussianProcess
    x, y = workloads.workload1(2000)
    args = (30, IndexSetUniform(2000, 4, 2), LaplacianEigenpairs(), MaternKernel(1, 1, 1))
    a = MultiResolutionGaussianProcess([x, y], *args, adaptive_inputs=True)
    xn = (x - x.mean()) / x.std() if a.standard_normalized_inputs else x
    o = G.GPRBFOracle()
    o.fit([xn, np.linspace(xn.min(), xn.max(), 2000)[:, None]])
    b = MultiResolutionGaussianProcess([x, y], *args, adaptive_inputs=True, input_model=o)
    a.fit(3, None)
    b.fit(3, None)
    xt = np.linspace(x.min(), x.max(), 500)[:, None]
    pa, pb = a.get_predicted_mean(xt), b.get_predicted_mean(xt)
    assert np.abs(pa - pb).max() <= 1e-3 * np.abs(pb).max()


def test_subsample_above_3000_points_draws_like_the_reference():
    """Inputs.py:24-47 on the global RNGs: random.uniform once, one numpy permutation per region, one for the rest."""
    from cimrgp_b200 import IndexSetUniform
    from cimrgp_b200.MRGP import MultiResolutionGaussianProcess as M
    n = 5000
    x, z = _warp(n, 4)
    random.seed(7)
    np.random.seed(7)
    models = M._learn_input_model(x, z, 0)
    after = (random.random(), np.random.rand())
    assert isinstance(models, list) and len(models) == 1 and models[0]._X.shape[0] == 3000
    # restatement of the draws
    random.seed(7)
    np.random.seed(7)
    rate = random.uniform(.1, .2)
    sets = IndexSetUniform(sample_length=n, resolution=1, divider=int(np.floor(rate * n))).index_set[-1]
    ids_l = [np.random.permutation(s)[0] for s in sets]
    np.random.permutation(np.delete(list(range(n)), ids_l))
    assert after == (random.random(), np.random.rand())
