"""Offline study of the omega solve: capture log omega_hat per (sweep, layer) from the oracle, then count Newton
iterations for different warm starts (numpy emulation of k_scale)."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import numpy as np
import workloads
from oracle import mrgp_oracle as O

N, RES, SWEEPS = 100000, 9, 8
x, y = workloads.workload1(N)
ora = O.OracleMRGP(x, y, 30, O.uniform_offsets(N, RES, 2), mode='ci', omega_solver='sinkhorn')
tables = []
orig = O.log_omega_hat
def spy(*a, **k):
    lw = orig(*a, **k); tables.append(lw.copy()); return lw
O.log_omega_hat = spy
import os
J = RES + 1
if os.path.exists('/tmp/omega_tables.npy'):
    tables = np.load('/tmp/omega_tables.npy')
else:
    for s in range(SWEEPS):
        ora.sweep()
    tables = np.array(tables).reshape(SWEEPS, J, 30, 30)
    np.save('/tmp/omega_tables.npy', tables)

def solve(lw, eta0, tol=1e-10, warm=True, max_it=46):
    M = lw.shape[0]
    K = lw - lw.max(1, keepdims=True)
    cs = K.max(0); K = np.exp(K - cs[None, :])
    v = np.exp(np.clip(eta0 + cs, -600, 600)) if warm else np.ones(M)
    n_warm = 0 if warm else 6
    err_prev = np.inf; errs = []
    for it in range(max_it):
        P = K * v[None, :]; P /= P.sum(1, keepdims=True); c = P.sum(0)
        err = np.max(np.abs(c - 1)); errs.append(err)
        if err < tol: break
        if not np.isfinite(err):
            v = np.ones(M); err_prev = np.inf; n_warm = it + 1 + 6; continue
        if it < n_warm or not (err < err_prev):
            v = v / c; err_prev = np.inf if it < n_warm else err; continue
        err_prev = err
        H = np.diag(c) - P.T @ P + 1.0 / M
        try:
            xs = np.linalg.solve(H, 1 - c)
        except np.linalg.LinAlgError:
            xs = np.full(M, np.nan)
        v = v * np.exp(np.clip(xs, -30, 30))
    return np.log(v) - cs, errs

for mode in ('prev_sweep', 'extrap', 'extrap_tol8'):
    eta = np.zeros((J, 30)); have = np.zeros(J, bool); last = None; eta2 = np.zeros((J, 30)); have2 = np.zeros(J, bool)
    print('==', mode)
    for s in range(SWEEPS):
        its = []; e0 = []
        for j in range(J):
            if mode == 'prev_sweep': e, w = eta[j], have[j]
            elif mode.startswith('extrap'): e, w = ((2 * eta[j] - eta2[j]) if have2[j] else eta[j]), have[j]
            elif mode == 'prev_layer': e, w = (last, True) if last is not None else (eta[j], False)
            else: e, w = eta[j], False
            new, errs = solve(tables[s, j], e, warm=w, tol=1e-8 if mode.endswith('tol8') else 1e-10)
            eta2[j] = eta[j]; have2[j] = have[j]
            eta[j] = new; have[j] = True; last = new
            its.append(len(errs)); e0.append(errs[0])
        print(s, its, ' '.join('%.0e' % v for v in e0))
