"""Basis functions and spectral densities: drop-in for the part of the reference's src/KernelClass.py that
the MRGP path uses (LaplacianEigenpairs.get_eigenpairs, MaternKernel.spectral).  These host versions exist
for API compatibility (collaborator objects, lazily materialised `phi_x`); the sweep evaluates the same
functions on the device (csrc/mrgp_math.cuh)."""
import numpy as np
from numpy import log, pi
from scipy.special import gammaln


class LaplacianEigenpairs(object):
    name = 'Laplacian'

    def get_eigenpairs(self, x, basis_id, basis_interval=None, per_dimension=False):
        # KernelClass.py:9-19
        x_dim = x.shape[1]
        if basis_interval is None:
            basis_interval = np.max(np.abs(x), axis=0)
        if len(basis_interval) != x_dim:
            raise ValueError('Basis interval should have the same dimensionality as the input.')
        eigen_function, eigen_value = self._learn(x, basis_interval, basis_id)
        if per_dimension is True:
            return eigen_function, eigen_value
        return np.prod(eigen_function, axis=1), np.sum(eigen_value)

    @staticmethod
    def _learn(x, basis_interval, basis_id):
        # KernelClass.py:21-37
        L = np.asarray(basis_interval, dtype=np.float64)[None, :]
        phi = (1. / np.sqrt(L)) * np.sin((np.pi * basis_id * (x + L)) / (2 * L))
        lam = np.power((np.pi * basis_id) / (2 * L[0]), 2)
        return phi, lam


class MaternKernel(object):
    name = 'Matern'

    def __init__(self, nu=1, l=1, sf=1):
        self.nu = nu
        self.l = l
        self.sf = sf

    def log_spectral(self, s):
        return self._matern_spectral(s)

    def spectral(self, s):
        return np.exp(self._matern_spectral(s))

    def _matern_spectral(self, s):
        # KernelClass.py:80-90
        nu, l, sf = self.nu, self.l, self.sf
        log_arg = log(2 * nu) - 2 * log(l)
        arg = np.exp(log_arg)
        return log(sf) + (0.5 * log(2 * pi)) + (nu * log_arg) + gammaln(nu + 0.5) - gammaln(nu) \
            - (nu + .5) * log(arg + s ** 2)
