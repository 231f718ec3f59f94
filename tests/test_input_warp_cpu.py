"""(f).3 input-warp GP: the NumPy twin (oracle/gp_rbf_oracle.py) and the host logic of cimrgp_b200/RegressionInput.py."""
import numpy as np
import pytest

from oracle import gp_rbf_oracle as G


def _warp(n, seed=0):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.standard_normal((n, 1)) ** 3, axis=0)
    z = np.linspace(x.min(), x.max(), n)[:, None]
    return x, z


def test_oracle_gradient_matches_finite_differences():
    x, z = _warp(150)
    X, Y = (x - x.mean(0)) / x.std(0), (z - z.mean(0)) / z.std(0)
    D2 = G.sqdist(X, X)
    for theta in (np.log([1.3, 0.7, 0.02]), np.log([0.4, 0.05, 0.2])):
        f, g = G.objective(theta, D2, Y)
        fd = [(G.objective(theta + 1e-6 * e, D2, Y)[0] - G.objective(theta - 1e-6 * e, D2, Y)[0]) / 2e-6 for e in np.eye(3)]
        assert np.allclose(g, fd, rtol=1e-5, atol=1e-6)


def test_oracle_fit_interpolates_a_monotone_warp():
    x, z = _warp(300, 1)
    o = G.GPRBFOracle()
    o.fit([x, z])
    assert np.all(np.isfinite(o.theta))
    assert np.abs(o.predict(x) - z).mean() < 0.02 * (z.max() - z.min())
    # maximum of the marginal likelihood: the gradient vanishes (or a bound is active)
    X, Y = (x - x.mean(0)) / x.std(0), (z - z.mean(0)) / z.std(0)
    f, g = G.objective(o.theta, G.sqdist(X, X), Y)
    at_bound = np.array([abs(t - lo) < 1e-9 or abs(t - hi) < 1e-9 for t, (lo, hi) in zip(o.theta, G.BOUNDS)])
    assert np.all((np.abs(g) < 1e-2 * max(1.0, abs(f))) | at_bound)


def test_regression_method_normalises_like_the_reference():
    from cimrgp_b200.RegressionInput import RegressionMethod

    class Identity(RegressionMethod):
        def _fit(self, train_data):
            self.seen = train_data
            return True

        def _predict(self, test_data):
            return test_data                     # labels == normalised inputs

    x, z = _warp(50)
    m = Identity()
    assert m.fit([x, z]) is True
    assert np.allclose(m.seen[0].mean(0), 0) and np.allclose(m.seen[0].std(0), 1)
    assert np.allclose(m.seen[1].mean(0), 0) and np.allclose(m.seen[1].std(0), 1)
    assert np.allclose(m.predict(x), (x - x.mean(0)) / x.std(0) * z.std(0) + z.mean(0))


def test_gp_rbf_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from cimrgp_b200.RegressionInput import GP_RBF
    x, z = _warp(20)
    with pytest.raises(RuntimeError):
        GP_RBF().fit([x, z])


def test_subsample_draws_follow_the_reference_order(monkeypatch):
    """Inputs.py:24-47 on the global RNGs (random.uniform once, one numpy permutation per region, one for the rest), with the
    GP fit itself stubbed out: 3000 distinct, sorted indices that hold one point of every region."""
    import random
    from cimrgp_b200 import IndexSetUniform
    from cimrgp_b200 import RegressionInput
    from cimrgp_b200.MRGP import MultiResolutionGaussianProcess as M
    seen = {}

    def fake_fit(self, train_data):
        seen['x'] = train_data[0]
        return True

    monkeypatch.setattr(RegressionInput.GP_RBF, 'fit', fake_fit)
    n = 5000
    x, z = _warp(n, 4)
    x = x + np.arange(n)[:, None] * 1e-9            # distinct values: the rows identify the indices
    random.seed(11)
    np.random.seed(11)
    models = M._learn_input_model(x, z, 0)
    after = (random.random(), np.random.rand())
    assert isinstance(models, list) and len(models) == 1 and seen['x'].shape == (3000, 1)
    random.seed(11)
    np.random.seed(11)
    rate = random.uniform(.1, .2)
    sets = IndexSetUniform(sample_length=n, resolution=1, divider=int(np.floor(rate * n))).index_set[-1]
    ids_l = [np.random.permutation(s)[0] for s in sets]
    rest = np.random.permutation(np.delete(list(range(n)), ids_l))[0:3000 - len(ids_l)]
    ids = np.sort(np.unique(list(rest) + ids_l))
    assert after == (random.random(), np.random.rand())
    assert np.array_equal(seen['x'], x[ids, :])
