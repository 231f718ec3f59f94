import numpy as np
np.set_printoptions(linewidth=200, precision=3, suppress=False)
t = np.load('/root/repo/gpurun_out/tables.npz')
def solve(lw, eta0=None, tol=1e-10, max_it=46, verbose=False, sink_burst=1):
    M = lw.shape[0]
    K = lw - lw.max(1, keepdims=True)
    cs = K.max(0); K = np.exp(K - cs[None, :])
    warm = eta0 is not None
    v = np.exp(np.clip(eta0 + cs, -600, 600)) if warm else np.ones(M)
    n_warm = 0 if warm else 6
    err_prev = np.inf; errs = []; kinds = []
    it = 0; burst = 0
    while it < max_it:
        P = K * v[None, :]; P /= P.sum(1, keepdims=True); c = P.sum(0)
        err = np.max(np.abs(c - 1)); errs.append(err); it += 1
        if err < tol: break
        if not np.isfinite(err):
            v = np.ones(M); err_prev = np.inf; n_warm = it + 6; kinds.append('R'); continue
        if it - 1 < n_warm or not (err < err_prev) or burst > 0:
            v = np.clip(v / c, 1e-280, 1e280)
            if burst > 0: burst -= 1
            elif not (it - 1 < n_warm): burst = sink_burst - 1
            err_prev = np.inf if (it - 1 < n_warm or burst > 0) else err; kinds.append('S'); continue
        err_prev = err
        H = np.diag(c) - P.T @ P + 1.0 / M
        try:
            L = np.linalg.cholesky(H); xs = np.linalg.solve(H, 1 - c)
        except np.linalg.LinAlgError:
            xs = np.full(M, np.nan)
        v = np.clip(v * np.exp(np.clip(xs, -30, 30)), 1e-280, 1e280); kinds.append('N')
    return np.log(v) - cs, errs, ''.join(kinds)
eta = None
for s in range(1, 8):
    lw = t['lw_s%d_j1' % s]
    new, errs, kinds = solve(lw, eta)
    print('sweep', s, 'gpu iters', int(t['it_s%d_j1' % s]), 'emu evals', len(errs), kinds, ' '.join('%.1e' % e for e in errs[:14]))
    print('   lw range per row (max - 2nd max) min/median: %.2f %.2f; overall range %.1f' % (np.min(np.sort(lw,1)[:,-1]-np.sort(lw,1)[:,-2]), np.median(np.sort(lw,1)[:,-1]-np.sort(lw,1)[:,-2]), lw.max()-lw.min()))
    eta = new
