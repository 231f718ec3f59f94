"""Device engine: one handle of the C ABI plus the torch tensors that own its device memory.

torch is used for plumbing only (device allocation, the CUDA stream, pinned host buffers); all arithmetic
of the hot path runs in the kernels of libcimrgp.so."""
import ctypes as C

import numpy as np

from . import _lib

_FIELD_SHAPES = None


def _dptr(arr):
    return arr.ctypes.data_as(C.POINTER(C.c_double))


class Engine(object):
    """State and phases of one ciMRGP / fiMRGP model on one GPU.

    x_norm (N, 1) normalised inputs and y (N, dy) observations as NumPy arrays (copied to the device from
    pinned host memory) or torch CUDA tensors (borrowed).  offsets: list of int64 arrays, one per layer.
    spectral: per layer None or (nu, l, sf).  interval_factor: per layer float.
    """

    def __init__(self, x_norm, y, offsets, n_basis, mode='ci', spectral=None, interval_factor=None,
                 noise_var0=1.0, ard_prior_influence=1.0, noise_region_specific=True, bias_region_specific=True,
                 device=0, n_ctas=0, intervals=None, chunk=None, defer_build=False, stream=None, workspace=None,
                 pinned=None):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.handle = None
        self.N = int(x_norm.shape[0])
        self.chunk = chunk            # (lo, hi) of the N_total samples owned by this handle, or None
        self.N_total = int(offsets[0][-1])
        self.dx = int(x_norm.shape[1]) if x_norm.ndim > 1 else 1
        self.dy = int(y.shape[1])
        self.M = int(n_basis)
        self.J = len(offsets)
        self.mode = mode
        self.offsets = [np.ascontiguousarray(o, dtype=np.int64) for o in offsets]
        self.R = [len(o) - 1 for o in self.offsets]
        lo, hi = chunk if chunk is not None else (0, 0)
        cfg = _lib.Config(_lib.ABI_VERSION, _lib.MODE_CI if mode == 'ci' else _lib.MODE_FI, self.N_total, self.dx, self.dy,
                          self.M, self.J, int(bool(noise_region_specific)), int(bool(bias_region_specific)),
                          int(device), int(n_ctas), int(lo), int(hi))
        ptrs, keep = _lib.offsets_arg(self.offsets)
        nreg = (C.c_int32 * self.J)(*self.R)
        out = C.c_void_p()
        rc = self.lib.mrgp_create(C.byref(cfg), ptrs, nreg, C.byref(out))
        if rc != 0:
            raise _lib.MrgpError(rc, self.lib.mrgp_last_error(None).decode())
        self.handle = out
        if not torch.cuda.is_available():
            raise _lib.MrgpError(_lib.ENODEVICE, 'no CUDA device: cimrgp_b200 has no CPU path')
        self.device = torch.device('cuda', int(device))
        # a batch of models (cimrgp_b200/batch.py) shares streams, one device allocation and one pinned staging area
        self.stream = stream if stream is not None else torch.cuda.Stream(device=self.device)
        nbytes = self.lib.mrgp_workspace_bytes(self.handle)
        if workspace is not None:
            if workspace.numel() < nbytes + 256:
                raise ValueError('workspace slice of %d bytes, %d needed' % (workspace.numel(), nbytes + 256))
            self.workspace = workspace
        else:
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        base = self.workspace.data_ptr()
        aligned = (base + 255) & ~255
        torch.cuda.synchronize(self.device)
        self._ck(self.lib.mrgp_set_stream(self.handle, C.c_void_p(self.stream.cuda_stream)))
        self._ck(self.lib.mrgp_bind_workspace(self.handle, C.c_void_p(aligned), nbytes))
        self._pinned = pinned
        self.set_data(x_norm, y)
        spectral = spectral if spectral is not None else [(1., 1., 1.)] * self.J
        interval_factor = interval_factor if interval_factor is not None else [1.0] * self.J
        self._interval_factor = interval_factor
        self._init_args = (float(noise_var0), float(ard_prior_influence))
        for j in range(self.J):
            sp = spectral[j]
            if sp is None:
                self._ck(self.lib.mrgp_set_spectral(self.handle, j, 0, 1., 1., 1.))
            else:
                self._ck(self.lib.mrgp_set_spectral(self.handle, j, 1, float(sp[0]), float(sp[1]), float(sp[2])))
            if not defer_build:
                self.build_basis(j, interval_factor[j], None if intervals is None else intervals[j])
        if not defer_build:
            self._ck(self.lib.mrgp_init_state(self.handle, float(noise_var0), float(ard_prior_influence)))
            self.synchronize()
        self._constructed = True

    @staticmethod
    def probe_workspace_bytes(offsets, n_basis, dy=2, mode='ci', n_ctas=0, device=0):
        """Bytes Engine will ask for (plus 256 for alignment) for a model of this shape: lets a caller carve many
        workspaces out of one allocation.  Host-only (plan construction), no device work."""
        lib = _lib.load()
        offs = [np.ascontiguousarray(o, dtype=np.int64) for o in offsets]
        cfg = _lib.Config(_lib.ABI_VERSION, _lib.MODE_CI if mode == 'ci' else _lib.MODE_FI, int(offs[0][-1]), 1, int(dy),
                          int(n_basis), len(offs), 1, 1, int(device), int(n_ctas), 0, 0)
        ptrs, keep = _lib.offsets_arg(offs)
        nreg = (C.c_int32 * len(offs))(*[len(o) - 1 for o in offs])
        out = C.c_void_p()
        rc = lib.mrgp_create(C.byref(cfg), ptrs, nreg, C.byref(out))
        if rc != 0:
            raise _lib.MrgpError(rc, lib.mrgp_last_error(None).decode())
        n = int(lib.mrgp_workspace_bytes(out))
        lib.mrgp_destroy(out)
        return n + 256

    # ------------------------------------------------------------------------------------------
    def _ck(self, rc):
        _lib.check(self.handle, rc)

    def close(self):
        if self.handle is not None:
            self.lib.mrgp_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_data(self, x_norm, y):
        """Inputs AND observations.  After construction this invalidates everything derived from x on the device
        (include/cimrgp.h, mrgp_set_data*): the basis of every layer is rebuilt here with the static interval rule
        (BasisInterval.py:15-16); the variational state is kept."""
        torch = self.torch
        rebuild = getattr(self, '_constructed', False)
        if isinstance(x_norm, np.ndarray):
            x_h = np.ascontiguousarray(x_norm, dtype=np.float64).reshape(self.N, self.dx)
            y_h = np.ascontiguousarray(y, dtype=np.float64).reshape(self.N, self.dy)
            if self._pinned is None:
                self._pinned = (torch.empty((self.N, self.dx), dtype=torch.float64).pin_memory(),
                                torch.empty((self.N, self.dy), dtype=torch.float64).pin_memory())
            self._pinned[0].numpy()[...] = x_h
            self._pinned[1].numpy()[...] = y_h
            self._ck(self.lib.mrgp_set_data_host(self.handle, C.c_void_p(self._pinned[0].data_ptr()),
                                                 C.c_void_p(self._pinned[1].data_ptr())))
            self.x_dev = self.y_dev = None
        else:
            self.x_dev = x_norm.to(device=self.device, dtype=torch.float64).contiguous()
            self.y_dev = y.to(device=self.device, dtype=torch.float64).contiguous()
            self._ck(self.lib.mrgp_set_data(self.handle, C.c_void_p(self.x_dev.data_ptr()),
                                            C.c_void_p(self.y_dev.data_ptr())))
        if rebuild:
            for j in range(self.J):
                self.build_basis(j, self._interval_factor[j])

    def set_observations(self, y):
        """New observations at the same inputs (NumPy array: staged in pinned memory; torch CUDA tensor: borrowed)."""
        torch = self.torch
        if isinstance(y, np.ndarray):
            if self._pinned is None:
                raise ValueError('this engine borrows device tensors: pass a CUDA tensor')
            self._pinned[1].numpy()[...] = np.ascontiguousarray(y, dtype=np.float64).reshape(self.N, self.dy)
            self.upload_observations()
        else:
            self.y_dev = y.to(device=self.device, dtype=torch.float64).contiguous()
            self._ck(self.lib.mrgp_set_observations(self.handle, C.c_void_p(self.y_dev.data_ptr())))

    def upload_observations(self):
        """Asynchronous H2D of the pinned y staging buffer on the engine's stream (16 B per sample)."""
        self._ck(self.lib.mrgp_set_observations_host(self.handle, C.c_void_p(self._pinned[1].data_ptr())))

    def prefetch_observations(self, y=None):
        """Double-buffered upload (mrgp_prefetch_observations_host): starts the H2D copy of the NEXT observations on the
        engine's copy stream and returns; they take effect at the next refresh_statistics().  y: NumPy array
        staged in the pinned buffer first, or None to send the pinned buffer as it is (the caller filled it - and must not
        touch it again before the data set has been taken over).  Loop of a pipelined consumer:
            eng.refresh_statistics(); eng.prefetch_observations(y_next); eng.sweep(1); ... read results ..."""
        if self._pinned is None:
            raise ValueError('this engine borrows device tensors: nothing to stage')
        if y is not None:
            self._ck(self.lib.mrgp_prefetch_sync(self.handle))     # the previous copy has left the staging buffer
            self._pinned[1].numpy()[...] = np.ascontiguousarray(y, dtype=np.float64).reshape(self.N, self.dy)
        self._ck(self.lib.mrgp_prefetch_observations_host(self.handle, C.c_void_p(self._pinned[1].data_ptr())))

    def build_basis(self, layer, interval_factor=1.0, intervals=None):
        if intervals is None:
            self._ck(self.lib.mrgp_build_basis(self.handle, layer, float(interval_factor), None))
        else:
            L = np.ascontiguousarray(intervals, dtype=np.float64).reshape(self.R[layer])
            self._ck(self.lib.mrgp_build_basis(self.handle, layer, float(interval_factor), _dptr(L)))

    # ------------------------------------------------------------------------------------------
    def set_adaptive_intervals(self, layer, use_prior=True, opt_interval_factor=(1., 1.2), enabled=True):
        """BasisInterval(use_prior, opt_interval_factor) of `layer` (-1: all layers): every sweep re-learns the
        layer's basis intervals after its bias / noise update (BasisInterval.py:18-134, MRGP.py:632-641)."""
        self._ck(self.lib.mrgp_set_adaptive_intervals(self.handle, int(layer), int(bool(enabled)), int(bool(use_prior)),
                                                      float(opt_interval_factor[0]), float(opt_interval_factor[1])))

    def learn_intervals(self, j):
        self._ck(self.lib.mrgp_learn_intervals(self.handle, int(j)))

    def interval_failures(self):
        out = C.c_uint64(0)
        self._ck(self.lib.mrgp_interval_failures(self.handle, C.byref(out)))
        return int(out.value)

    def refresh_statistics(self):
        """Layer-0 sufficient statistics of the fused ci sweep, now (mrgp_sweep does it on demand otherwise)."""
        self._ck(self.lib.mrgp_refresh_statistics(self.handle))

    def sweep(self, n_iter=1):
        self._ck(self.lib.mrgp_sweep(self.handle, int(n_iter)))

    def phase_a(self, j):
        self._ck(self.lib.mrgp_phase_a(self.handle, j))

    def axis_update(self, j):
        self._ck(self.lib.mrgp_axis_update(self.handle, j))

    def phase_b(self, j):
        self._ck(self.lib.mrgp_phase_b(self.handle, j))

    def bias_noise(self, j):
        self._ck(self.lib.mrgp_bias_noise(self.handle, j))

    def sweep_stepwise(self):
        """One sweep through the per-phase entry points (no graph); same arithmetic as sweep()."""
        for j in range(self.J):
            self.phase_a(j)
            self.axis_update(j)
            self.phase_b(j)
            self.bias_noise(j)

    def synchronize(self):
        self._ck(self.lib.mrgp_synchronize(self.handle))

    def elbo(self):
        out = np.zeros((self.J, 6))
        self._ck(self.lib.mrgp_elbo(self.handle, _dptr(out)))
        return out

    def set_fused(self, on):
        """False: the multi-kernel sweep (Sinkhorn / Newton omega solve) from now on; True: the fused sweep where it applies."""
        self._ck(self.lib.mrgp_set_fused(self.handle, 1 if on else 0))

    def elbo_async(self, slot):
        """Queue the lower bound (six terms per layer) and its copy into pinned host memory; elbo_result(slot) returns it.
        Two slots (0, 1): read the bound of step k after queueing step k + 1."""
        if getattr(self, '_elbo_pinned', None) is None:
            self._elbo_pinned = [self.torch.zeros((self.J, 6), dtype=self.torch.float64).pin_memory() for _ in range(2)]
        self._ck(self.lib.mrgp_elbo_async(self.handle, C.c_void_p(self._elbo_pinned[slot].data_ptr()), int(slot)))

    def elbo_result(self, slot):
        self._ck(self.lib.mrgp_elbo_wait(self.handle, int(slot)))
        return self._elbo_pinned[slot].numpy().copy()

    def launch_count(self):
        return int(self.lib.mrgp_launch_count(self.handle))

    def cholesky_count(self):
        return int(self.lib.mrgp_cholesky_count(self.handle))

    # ------------------------------------------------------------------------------------------
    def get(self, layer, field, shape):
        out = np.empty(shape, dtype=np.float64)
        self._ck(self.lib.mrgp_get_state(self.handle, layer, field, _dptr(out), out.size))
        return out

    def put(self, layer, field, value):
        v = np.ascontiguousarray(value, dtype=np.float64)
        self._ck(self.lib.mrgp_set_state(self.handle, layer, field, _dptr(v), v.size))

    def layer_state(self, j):
        """State of layer j with the reference's shapes (A, ytil as (R, dy, M))."""
        R, M, dy = self.R[j], self.M, self.dy
        F = _lib
        g = lambda f, shape: self.get(j, f, shape)
        st = {
            'L': g(F.F_L, (R, 1)), 'lam': g(F.F_LAMBDA, (R, M)), 'S': g(F.F_SPECTRAL, (R, M)), 'd': g(F.F_PHI2SUM, (R, M)),
            'scale_precision': g(F.F_SCALE_PRECISION, (R, M)), 'zeta': g(F.F_ZETA, (R, M)),
            'ytil': np.swapaxes(g(F.F_YTILDE, (R, M, dy)), 1, 2).copy(),
            'A': np.swapaxes(g(F.F_A, (R, M, dy)), 1, 2).copy(),
            'm2': g(F.F_M2, (R, M)), 'cm2': g(F.F_CM2, (R, M)),
            'noise_shape': g(F.F_NOISE_SHAPE, (R,)), 'noise_scale': g(F.F_NOISE_SCALE, (R,)),
            'noise_mean': g(F.F_NOISE_MEAN, (R,)), 'noise_log_mean': g(F.F_NOISE_LOG_MEAN, (R,)),
            'bias_prec': g(F.F_BIAS_PRECISION, (R,)), 'bias_mean': g(F.F_BIAS_MEAN, (R, dy)),
            'bias_var': g(F.F_BIAS_VAR, (R,)),
        }
        if self.mode == 'fi':
            st.update(self._axis_state(j, (R,)))
        return st

    def _axis_state(self, layer, lead):
        M, dy = self.M, self.dy
        F = _lib
        g = lambda f, shape: self.get(layer, f, lead + shape)
        return {
            'B': g(F.F_AXIS_B, (M, dy, dy)), 'kappa': g(F.F_AXIS_KAPPA, (M, dy)), 'rho': g(F.F_AXIS_RHO, (M, dy)),
            'logC': g(F.F_AXIS_LOGC, (M,)), 'axis_cov': g(F.F_AXIS_COV, (M, dy, dy)),
            'ard_shape': g(F.F_ARD_SHAPE, (M,)), 'ard_scale': g(F.F_ARD_SCALE, (M,)),
            'ard_mean': g(F.F_ARD_MEAN, (M,)), 'ard_log_mean': g(F.F_ARD_LOG_MEAN, (M,)),
        }

    def shared_state(self):
        st = self._axis_state(-1, ())
        st['omega'] = self.get(-1, _lib.F_OMEGA, (self.M, self.M))
        return st

    def latent(self, j):
        """(fbar (N, dy), fvar (N,)) of layer j (Stats.py:126-157)."""
        return self.get(j, _lib.F_FBAR, (self.N, self.dy)), self.get(j, _lib.F_FVAR, (self.N,))

    def state(self, latent=True):
        """Flat dict with the key names of oracle.mrgp_oracle.OracleMRGP.state()."""
        out = {}
        for j in range(self.J):
            for k, v in self.layer_state(j).items():
                out['L%d.%s' % (j, k)] = v
            if latent:
                fb, fv = self.latent(j)
                out['L%d.fbar' % j] = fb
                out['L%d.fvar' % j] = fv
        if self.mode == 'ci':
            for k, v in self.shared_state().items():
                out['S.' + k] = v
        return out

    # ------------------------------------------------------------------------------------------
    def predict_mean(self, x_test_norm, test_offsets=None):
        torch = self.torch
        xt = torch.as_tensor(np.ascontiguousarray(x_test_norm, dtype=np.float64).reshape(-1), device=self.device)
        out = torch.empty((xt.shape[0], self.dy), dtype=torch.float64, device=self.device)
        torch.cuda.synchronize(self.device)
        if test_offsets is None:
            self._ck(self.lib.mrgp_predict_mean(self.handle, C.c_void_p(xt.data_ptr()), xt.shape[0], None, 0,
                                                C.c_void_p(out.data_ptr())))
        else:
            ptrs, keep = _lib.offsets_arg(test_offsets)
            self._ck(self.lib.mrgp_predict_mean(self.handle, C.c_void_p(xt.data_ptr()), xt.shape[0], ptrs,
                                                len(test_offsets), C.c_void_p(out.data_ptr())))
        return out.cpu().numpy()

    def predict_var_indexed(self, x_test_norm, test_offsets):
        torch = self.torch
        xt = torch.as_tensor(np.ascontiguousarray(x_test_norm, dtype=np.float64).reshape(-1), device=self.device)
        out = torch.empty((xt.shape[0],), dtype=torch.float64, device=self.device)
        torch.cuda.synchronize(self.device)
        ptrs, keep = _lib.offsets_arg(test_offsets)
        self._ck(self.lib.mrgp_predict_var_indexed(self.handle, C.c_void_p(xt.data_ptr()), xt.shape[0], ptrs,
                                                   len(test_offsets), C.c_void_p(out.data_ptr())))
        return out.cpu().numpy()

    def predict_var(self, x_test_norm):
        torch = self.torch
        xt = torch.as_tensor(np.ascontiguousarray(x_test_norm, dtype=np.float64).reshape(-1), device=self.device)
        out = torch.empty((xt.shape[0],), dtype=torch.float64, device=self.device)
        torch.cuda.synchronize(self.device)
        self._ck(self.lib.mrgp_predict_var(self.handle, C.c_void_p(xt.data_ptr()), xt.shape[0], C.c_void_p(out.data_ptr())))
        return out.cpu().numpy()
