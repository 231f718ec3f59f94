"""TEST INFRASTRUCTURE (only tests/ may import this): NumPy twin of the omega solve of the fused sweep
(cimrgp_b200/csrc/chain.cu, omega_solve_warp): Anderson-accelerated Sinkhorn on x = log v.

The fixed point is the one of mrgp_oracle.omega_sinkhorn (the exact solution that the reference's fsolve call,
Stats.py:413-420, approximates): omega = diag(u) K diag(v) doubly stochastic, K = exp(shifted log omega_hat).  This file
restates HOW the device reaches it, step for step (shifts, residual, three differences, scaled 3 x 3 normal equations by
Cramer's rule, restart when the residual grows), so that the iteration counts and the result can be checked on the CPU.
"""
import numpy as np

TOL = 1e-10            # kOmegaTol: max |column sum - 1| with the rows normalised exactly
MAX_ANDERSON = 48      # kChainAndersonIters


def shifted_table(lw):
    """Row maxima, then column maxima: every row and column of K holds an entry close to 1 (shared_step)."""
    k = lw - lw.max(axis=1, keepdims=True)
    cshift = k.max(axis=0)
    return np.exp(k - cshift[None, :]), cshift


def evaluate(K, v):
    u = 1.0 / (K @ v)
    s = K.T @ u
    return u, s, v * s


def solve(lw, eta=None):
    """Returns omega, the evaluations used and log v - column shift (the next sweep's warm start)."""
    K, cshift = shifted_table(np.asarray(lw, dtype=np.float64))
    M = K.shape[0]
    x = np.zeros(M) if eta is None else np.clip(eta + cshift, -600.0, 600.0)
    xp = fp = np.zeros(M)
    dX, dF, G = np.zeros((3, M)), np.zeros((3, M)), np.zeros((3, 3))
    d = np.zeros(3)
    nh, have_prev, err_prev, evals = 0, False, np.inf, 0
    converged = False
    for _ in range(MAX_ANDERSON):
        v = np.exp(x)
        u, s, c = evaluate(K, v)
        evals += 1
        err = np.max(np.abs(c - 1.0))
        if err < TOL:
            converged = True
            break
        if not np.isfinite(err):
            x, nh, have_prev, err_prev = np.zeros(M), 0, False, np.inf
            continue
        fr = -np.log(s) - x
        if not err <= err_prev:
            nh, have_prev = 0, False
        B = np.zeros(3)
        if have_prev:
            dX[2], dX[1], dF[2], dF[1] = dX[1].copy(), dX[0].copy(), dF[1].copy(), dF[0].copy()
            G[2, 2], G[1, 2], G[1, 1] = G[1, 1], G[0, 1], G[0, 0]
            w = fr - fp
            dX[0] = x - xp
            mu = w.sum() / M
            G[0, 0] = (w * w).sum() - M * mu * mu
            G[0, 1], G[0, 2] = (w * dF[1]).sum(), (w * dF[2]).sum()
            B[:] = (fr * w).sum() - M * mu * mu, (fr * dF[1]).sum(), (fr * dF[2]).sum()
            dF[0] = w - mu
            fc = fr - mu
            nh = min(nh + 1, 3)
        else:
            fc = fr - fr.mean()
        xp, fp, have_prev = x.copy(), fc.copy(), True
        xn = x + fc
        if nh > 0 and G[0, 0] > 0.0:
            h1 = nh > 1 and G[1, 1] > 0.0
            h2 = nh > 2 and h1 and G[2, 2] > 0.0
            d0 = 1.0 / np.sqrt(G[0, 0])
            d1, d2 = (d[1] if h1 else 0.0), (d[2] if h2 else 0.0)
            a01, a02, a12 = G[0, 1] * d0 * d1, G[0, 2] * d0 * d2, G[1, 2] * d1 * d2
            r0, r1, r2 = B[0] * d0, B[1] * d1, B[2] * d2
            kD = 1.0 + 1e-7
            m00, m01, m02 = kD * kD - a12 * a12, a02 * a12 - kD * a01, a01 * a12 - kD * a02
            m11, m12, m22 = kD * kD - a02 * a02, a01 * a02 - kD * a12, kD * kD - a01 * a01
            det = kD * m00 + a01 * m01 + a02 * m02
            if det > 1e-12:
                g0 = (m00 * r0 + m01 * r1 + m02 * r2) / det * d0
                g1 = (m01 * r0 + m11 * r1 + m12 * r2) / det * d1
                g2 = (m02 * r0 + m12 * r1 + m22 * r2) / det * d2
                xn = xn - (g0 * (dX[0] + dF[0]) + g1 * (dX[1] + dF[1]) + g2 * (dX[2] + dF[2]))
            d[2], d[1] = d1, d0
        else:
            d[2], d[1] = d[1], 0.0
        x = np.clip(xn, -640.0, 640.0)
        err_prev = err
    v = np.exp(x)
    sweeps = 0
    while not converged and sweeps < 2000:      # last resort: plain Sinkhorn sweeps
        u, s, c = evaluate(K, v)
        if np.max(np.abs(c - 1.0)) < TOL:
            break
        v = np.clip(1.0 / s, 1e-280, 1e280)
        evals += 1
        sweeps += 1
    u, s, c = evaluate(K, v)
    return (K * v[None, :]) * u[:, None], evals, np.log(v) - cshift
