"""Input-warp regressors of the reference (RegressionInput.py:1-67): `RegressionMethod` and `GP_RBF`.

The reference wraps GPy (`GPy.models.GPRegression` with an RBF kernel, `model.optimize()`); GPy is a third-party package that is
not part of the reference tree and not installed here, so this is a restatement of what that call computes - exact GP regression,
type-II maximum likelihood over (signal variance, length scale, noise variance) from GPy's starting point (1, 1, 0.01 var(labels)),
L-BFGS-B - not a binding.  PARITY UNPINNED: there is no GPy to compare with; the twin is `oracle/gp_rbf_oracle.py` (NumPy / SciPy).

Where it runs: on the GPU, in float64, through torch.linalg (cuSOLVER / cuBLAS: library calls, stated as such - this is the
one-off fit of at most 3000 points behind `adaptive_inputs=True` (Inputs.py:22-49), not the sweep).  No CPU fallback: without a CUDA
device `fit` raises.
"""
import numpy as np


class RegressionMethod(object):
    """RegressionInput.py:10-52: zero-mean, unit-variance normalisation of inputs and labels around `_fit` / `_predict`."""

    def __init__(self):
        self.preprocess = True

    def _preprocess(self, data, train):
        if train:
            inputs, labels = data
            self.data_mean = inputs.mean(axis=0)
            self.data_std = inputs.std(axis=0)
            self.labels_mean = labels.mean(axis=0)
            self.labels_std = labels.std(axis=0)
            return ((inputs - self.data_mean) / self.data_std, (labels - self.labels_mean) / self.labels_std)
        return (data - self.data_mean) / self.data_std

    def _reverse_trans_labels(self, labels):
        return labels * self.labels_std + self.labels_mean

    def fit(self, train_data):
        if self.preprocess:
            train_data = self._preprocess(train_data, True)
        return self._fit(train_data)

    def predict(self, test_data):
        if self.preprocess:
            test_data = self._preprocess(test_data, False)
        labels = self._predict(test_data)
        if self.preprocess:
            labels = self._reverse_trans_labels(labels)
        return labels

    def _fit(self, train_data):
        raise NotImplementedError

    def _predict(self, test_data):
        raise NotImplementedError


# log bounds of (signal variance, length scale, noise variance) for normalised data: the noise floor keeps the kernel matrix of
# clustered inputs factorisable in float64 (GPy has no bounds and adds jitter instead)
BOUNDS = [(-12.0, 12.0), (-12.0, 12.0), (float(np.log(1e-6)), 12.0)]


class GP_RBF(RegressionMethod):
    """RegressionInput.py:55-67.  theta = log(signal variance, length scale, noise variance)."""
    name = 'GP_RBF'
    max_iters = 1000          # GPy's default for optimize()

    def __init__(self, device=None):
        RegressionMethod.__init__(self)
        self.device = device

    def _device(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('GP_RBF runs on the GPU (torch.linalg in float64); no CUDA device is available and there is no CPU path')
        return torch.device('cuda', torch.cuda.current_device() if self.device is None else int(self.device))

    @staticmethod
    def _objective(torch, theta, D2, Y):
        """Negative log marginal likelihood and its gradient with respect to theta (Rasmussen & Williams eq. 5.8-5.9)."""
        n, p = Y.shape
        v, ell, s2 = (float(np.exp(t)) for t in theta)
        E = torch.exp(D2 * (-0.5 / (ell * ell)))                    # RBF correlations
        K = v * E
        K.diagonal().add_(s2)
        L, info = torch.linalg.cholesky_ex(K)
        for attempt in range(6):                                    # (GPy's jitchol: growing jitter on a failed factorisation)
            if int(info) == 0:
                break
            K.diagonal().add_(float(K.diagonal().mean()) * 1e-6 * 10.0 ** attempt)
            L, info = torch.linalg.cholesky_ex(K)
        if int(info) != 0:
            raise RuntimeError('GP_RBF: the kernel matrix is not positive definite')
        alpha = torch.cholesky_solve(Y, L)
        nll = 0.5 * p * 2.0 * torch.log(L.diagonal()).sum() + 0.5 * (Y * alpha).sum() + 0.5 * n * p * np.log(2.0 * np.pi)
        W = p * torch.cholesky_inverse(L) - alpha @ alpha.T         # dnll/dK = W / 2
        g_v = 0.5 * (W * E).sum() * v
        g_l = 0.5 * (W * E * D2).sum() * v / (ell * ell)
        g_s = 0.5 * W.diagonal().sum() * s2
        return float(nll), np.array([float(g_v), float(g_l), float(g_s)])

    def _fit(self, train_data):
        import torch
        from scipy.optimize import minimize
        inputs, labels = train_data
        dev = self._device()
        X = torch.as_tensor(np.ascontiguousarray(inputs, dtype=np.float64), device=dev)
        Y = torch.as_tensor(np.ascontiguousarray(labels, dtype=np.float64), device=dev)
        D2 = torch.cdist(X, X, p=2.0, compute_mode='donot_use_mm_for_euclid_dist') ** 2
        theta0 = np.log(np.array([1.0, 1.0, max(float(np.var(labels)) * 0.01, 1e-12)]))
        res = minimize(lambda t: self._objective(torch, t, D2, Y), theta0, jac=True, method='L-BFGS-B',
                       bounds=BOUNDS, options={'maxiter': self.max_iters})
        self.theta = np.asarray(res.x, dtype=np.float64)
        self.variance, self.lengthscale, self.noise_variance = (float(v) for v in np.exp(self.theta))
        self.nll = float(res.fun)
        K = self.variance * torch.exp(D2 * (-0.5 / self.lengthscale ** 2))
        K.diagonal().add_(self.noise_variance)
        self._X = X
        self._alpha = torch.cholesky_solve(Y, torch.linalg.cholesky(K))
        return True

    def _predict(self, test_data):
        import torch
        Xs = torch.as_tensor(np.ascontiguousarray(test_data, dtype=np.float64), device=self._X.device)
        Ks = self.variance * torch.exp(torch.cdist(Xs, self._X, p=2.0, compute_mode='donot_use_mm_for_euclid_dist') ** 2
                                       * (-0.5 / self.lengthscale ** 2))
        return (Ks @ self._alpha).cpu().numpy()
