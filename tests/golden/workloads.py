"""Synthetic workloads shared by the golden generator, the tests and bench.py.

The signal functions restate the reference example scripts' generators
(scripts/tests/ciMRGP_vs_fiMRGP.py:12-29, scripts/tests/GPRBF_vs_ciMRGP_vs_fiMRGP.py:16-35); script 1
sets no seed, so one is fixed here (10, as script 2).  NumPy's legacy global RNG is used so that the
reference run and every consumer see identical bytes.
"""
import numpy as np


def signal1(x):
    y1 = np.log(np.log(x) + abs(np.sin(x ** 2) * np.exp(np.sin(np.cos(2 * x)))))
    y2 = np.log(np.log(x) + abs(np.sin(-x ** 2 + 3 * x + 5) + np.log(1 + abs(np.cos(x ** 2)))))
    return np.array([y1, y2])


def signal2(x):
    y1 = np.log(np.log(x) + abs(np.sin(x ** 2) * np.exp(np.sin(np.cos(2 * x ** 2)))))
    y2 = np.log(np.log(x ** 2) + abs(np.tan(-x ** 2 + 3 * x + 5) + np.log(1 + abs(np.cos(x ** 2)))))
    return np.array([y1, y2])


def workload1(n, seed=10):
    """Script-1 data with 32 -> n samples on linspace(1, 3, n)."""
    np.random.seed(seed)
    x = np.atleast_2d(np.linspace(1, 3, n)).T
    y = signal1(x)[:, :, 0].T
    dy = 1 + 1 * np.random.random(y.shape)
    y = y + .1 * np.random.normal(0, dy)
    return x, y


def workload2(seed=10):
    """Script-2 data: 160 points = 4 x 40 of linspace(1, 4, 220), noise 0.5."""
    np.random.seed(seed)
    x_ = np.atleast_2d(np.linspace(1, 4, 220)).T
    x = np.stack([x_[0:40], x_[60:100], x_[120:160], x_[180:220]]).reshape(160, 1)
    y = signal2(x)[:, :, 0].T
    dy = 1 + 1 * np.random.random(y.shape)
    y = y + .5 * np.random.normal(0, dy)
    return x, y


class InterpInputModel(object):
    """Deterministic stand-in for the reference's GPy input-warp model (Inputs.py:22-49, RegressionInput.py:55-67;
    GPy is absent): maps a normalised input to the warped coordinate z by interpolating the training pairs
    (x_n, z_n), z = linspace(min x, max x, N) (Inputs.py:12).  SURVEY.md §8d, config 2 (ii)."""

    def __init__(self, x_train):
        xn = (x_train - np.mean(x_train, axis=0)) / np.std(x_train, axis=0)
        self.xn = xn[:, 0].copy()
        self.z = np.linspace(np.min(xn), np.max(xn), xn.shape[0])

    def predict(self, x):
        return np.atleast_2d(np.interp(np.asarray(x)[:, 0], self.xn, self.z)).T
