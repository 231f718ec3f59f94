"""TEST INFRASTRUCTURE (only tests/ may import this): NumPy / SciPy twin of cimrgp_b200/RegressionInput.py::GP_RBF.

What the reference's input-warp model computes (RegressionInput.py:55-67: GPy.models.GPRegression with an RBF kernel, likelihood
variance 0.01 var(labels), model.optimize()): exact GP regression with type-II maximum likelihood.  GPy (third party, unpinned
upstream, absent from /root/reference and from this image) is restated from its published model, Rasmussen & Williams ch. 5.
PARITY UNPINNED: no GPy output exists to pin it to; the tests hold the CUDA-side class to this twin and this twin's gradient to
finite differences.
"""
import numpy as np
from scipy.linalg import cho_factor, cho_solve
from scipy.optimize import minimize


BOUNDS = [(-12.0, 12.0), (-12.0, 12.0), (float(np.log(1e-6)), 12.0)]


def sqdist(A, B):
    return ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)


def objective(theta, D2, Y):
    n, p = Y.shape
    v, ell, s2 = np.exp(theta)
    E = np.exp(-0.5 * D2 / ell ** 2)
    K = v * E + s2 * np.eye(n)
    c = cho_factor(K, lower=True)
    alpha = cho_solve(c, Y)
    nll = p * np.log(np.diag(c[0])).sum() + 0.5 * (Y * alpha).sum() + 0.5 * n * p * np.log(2 * np.pi)
    W = p * cho_solve(c, np.eye(n)) - alpha @ alpha.T
    return nll, np.array([0.5 * (W * E).sum() * v, 0.5 * (W * E * D2).sum() * v / ell ** 2, 0.5 * np.trace(W) * s2])


class GPRBFOracle(object):
    def fit(self, train_data):
        x, z = train_data
        self.data_mean, self.data_std = x.mean(0), x.std(0)
        self.labels_mean, self.labels_std = z.mean(0), z.std(0)
        X, Y = (x - self.data_mean) / self.data_std, (z - self.labels_mean) / self.labels_std
        D2 = sqdist(X, X)
        theta0 = np.log(np.array([1.0, 1.0, max(Y.var() * 0.01, 1e-12)]))
        res = minimize(objective, theta0, args=(D2, Y), jac=True, method='L-BFGS-B', bounds=BOUNDS, options={'maxiter': 1000})
        self.theta, self.nll = res.x, res.fun
        v, ell, s2 = np.exp(res.x)
        self.X, self.v, self.ell = X, v, ell
        self.alpha = cho_solve(cho_factor(v * np.exp(-0.5 * D2 / ell ** 2) + s2 * np.eye(len(X)), lower=True), Y)
        return True

    def predict(self, xs):
        Xs = (xs - self.data_mean) / self.data_std
        return (self.v * np.exp(-0.5 * sqdist(Xs, self.X) / self.ell ** 2)) @ self.alpha * self.labels_std + self.labels_mean
