"""ctypes front end of oracle/mrgp_port.c, the multi-threaded C restatement of the ci sweep.

TEST INFRASTRUCTURE ONLY (CPU baseline of bench.py; held to oracle/mrgp_oracle.py by tests/test_port_c.py).
The product package never imports this module."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'mrgp_port.c')
LIB = os.path.join(HERE, 'libmrgp_port.so')
_lib = None


def build(force=False):
    """gcc -O3 -fopenmp (x86-64-v3: the library is built in one container and may run on another host)."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(['gcc', '-O3', '-march=x86-64-v3', '-fopenmp', '-shared', '-fPIC', '-o', LIB, SRC, '-lm'])
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.port_create.restype = C.c_void_p
        lib.port_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_int64)),
                                    C.POINTER(C.c_int32), C.c_double, C.c_double, C.c_double]
        lib.port_sweep.argtypes = [C.c_void_p]
        lib.port_get.restype = C.c_int64
        lib.port_get.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.port_omega_seconds.restype = C.c_double
        lib.port_omega_seconds.argtypes = [C.c_void_p]
        lib.port_threads.restype = C.c_int
        lib.port_destroy.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


class PortC(object):
    """ci model, static intervals, dy = 2, Matern(nu, l, sf) spectral density on every layer (MRGP.py:571-652)."""
    FIELDS = {'A': 0, 'noise_mean': 1, 'bias_mean': 2, 'cm2': 3, 'ytil': 4, 'noise_scale': 5}
    SHARED = {'B': 10, 'omega': 11, 'ard_mean': 12, 'logC': 13, 'ard_scale': 14}

    def __init__(self, x, y, n_basis, offsets, spectral=(1., 1., 1.)):
        self.lib = load()
        x = np.asarray(x, dtype=np.float64)
        # MRGP.py:278-295
        self.x = np.ascontiguousarray(((x - np.mean(x, 0)) / np.std(x, 0)).reshape(-1))
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        assert self.y.shape[1] == 2 and x.shape[1] == 1
        self.M, self.J = int(n_basis), len(offsets)
        self.offsets = [np.ascontiguousarray(o, dtype=np.int64) for o in offsets]
        self.R = [len(o) - 1 for o in self.offsets]
        ptrs = (C.POINTER(C.c_int64) * self.J)(*[o.ctypes.data_as(C.POINTER(C.c_int64)) for o in self.offsets])
        nreg = (C.c_int32 * self.J)(*self.R)
        self.h = self.lib.port_create(self.x.ctypes.data, self.y.ctypes.data, self.x.shape[0], self.M, self.J, ptrs, nreg,
                                      float(spectral[0]), float(spectral[1]), float(spectral[2]))
        if not self.h:
            raise ValueError('port_create failed')

    def sweep(self, n=1):
        for _ in range(n):
            self.lib.port_sweep(self.h)

    @property
    def threads(self):
        return int(self.lib.port_threads())

    @property
    def t_omega(self):
        return float(self.lib.port_omega_seconds(self.h))

    def get(self, layer, name):
        code = self.SHARED[name] if layer < 0 else self.FIELDS[name]
        n = self.lib.port_get(self.h, layer, code, None)
        out = np.empty(n)
        self.lib.port_get(self.h, layer, code, out.ctypes.data)
        return out

    def state(self):
        """Subset of OracleMRGP.state() with the same keys and shapes."""
        M, out = self.M, {}
        for j in range(self.J):
            R = self.R[j]
            out['L%d.A' % j] = np.swapaxes(self.get(j, 'A').reshape(R, M, 2), 1, 2)
            out['L%d.ytil' % j] = np.swapaxes(self.get(j, 'ytil').reshape(R, M, 2), 1, 2)
            out['L%d.cm2' % j] = self.get(j, 'cm2').reshape(R, M)
            out['L%d.noise_mean' % j] = self.get(j, 'noise_mean')
            out['L%d.noise_scale' % j] = self.get(j, 'noise_scale')
            out['L%d.bias_mean' % j] = self.get(j, 'bias_mean').reshape(R, 2)
        b = self.get(-1, 'B').reshape(M, 3)
        out['S.B'] = np.stack([b[:, 0], b[:, 1], b[:, 1], b[:, 2]], axis=1).reshape(M, 2, 2)
        out['S.omega'] = self.get(-1, 'omega').reshape(M, M)
        out['S.ard_mean'] = self.get(-1, 'ard_mean')
        out['S.ard_scale'] = self.get(-1, 'ard_scale')
        out['S.logC'] = self.get(-1, 'logC')
        return out

    def close(self):
        if self.h:
            self.lib.port_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
