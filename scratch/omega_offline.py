"""Offline study of the omega solve on the tables of the headline run (gpurun_out/chain_tables.npz from scratch/dump_tables.py)."""
import numpy as np, sys
d = np.load('/root/repo/gpurun_out/chain_tables.npz')
T, IT = d['tables'], d['iters']          # (S, J, M, M), (S, J)
S, J, M, _ = T.shape
TOL = 1e-10
C_EVAL, C_NEWTON = 1.3, 21.0             # kcycles

def prep(lw):
    K = lw - lw.max(1, keepdims=True)
    cs = K.max(0)
    return np.exp(K - cs[None, :]), cs

def evalK(K, v):
    r = K @ v
    u = 1.0 / r
    s = K.T @ u
    c = v * s
    return u, s, c

def current(K, v, max_it=46, fast=0.33, below=0.2):
    ne = nn = 0; err_prev = np.inf; last = 0
    for it in range(max_it):
        u, s, c = evalK(K, v); ne += 1
        err = np.max(np.abs(c - 1))
        if err < TOL: break
        sink_fast = last == 1 and err < fast * err_prev
        newton_ok = last != 0 and err < below and not (last == 2 and not err < err_prev)
        if not (newton_ok and not sink_fast):
            v = 1.0 / s; err_prev = err; last = 1; continue
        err_prev = err; last = 2
        P = (K * v[None, :]) * u[:, None]
        H = np.diag(c) - P.T @ P + 1.0 / M
        x = np.linalg.solve(H, 1 - c); nn += 1
        v = v * np.exp(np.clip(x, -30, 30))
    return v, ne, nn

def sinkhorn(K, v, max_it=3000):
    ne = 0
    for it in range(max_it):
        u, s, c = evalK(K, v); ne += 1
        if np.max(np.abs(c - 1)) < TOL: break
        v = 1.0 / s
    return v, ne, 0

def lyusternik(K, v, max_it=400, every=3):
    """Sinkhorn on x = log v; after `every` plain steps extrapolate along the last step with the measured ratio."""
    ne = 0; x = np.log(v); hist = []
    errs = []
    for it in range(max_it):
        u, s, c = evalK(K, np.exp(x)); ne += 1
        err = np.max(np.abs(c - 1)); errs.append(err)
        if err < TOL: break
        xn = -np.log(s)
        dx = xn - x
        dx -= dx.mean()                      # scale-invariant direction
        hist.append(dx)
        x = xn
        if len(hist) >= 2 and (len(hist) % every) == 0:
            a, b = hist[-2], hist[-1]
            r = (a @ b) / (a @ a)
            if 0.05 < r < 0.98:
                x = x + (r / (1 - r)) * b
                hist = []
    return np.exp(x), ne, 0

def anderson(K, v, max_it=400, m=3):
    """Anderson acceleration (depth m) of the Sinkhorn map on x = log v."""
    ne = 0; x = np.log(v); X = []; G = []
    for it in range(max_it):
        u, s, c = evalK(K, np.exp(x)); ne += 1
        err = np.max(np.abs(c - 1))
        if err < TOL: break
        g = -np.log(s); f = g - x; f -= f.mean()
        X.append(x.copy()); G.append(f.copy())
        if len(G) > m + 1: X.pop(0); G.pop(0)
        if len(G) >= 2:
            dF = np.array([G[i + 1] - G[i] for i in range(len(G) - 1)]).T
            dX = np.array([X[i + 1] - X[i] for i in range(len(X) - 1)]).T
            gamma, *_ = np.linalg.lstsq(dF, f, rcond=None)
            xn = x + f - (dX + dF) @ gamma
        else:
            xn = x + f
        x = xn
    return np.exp(x), ne, 0

methods = {'current': current, 'sinkhorn': sinkhorn, 'lyusternik': lyusternik, 'anderson3': anderson}
sel = sys.argv[1:] or list(methods)
for name in sel:
    fn = methods[name]
    tot = np.zeros((S, J)); evs = np.zeros((S, J), int); nws = np.zeros((S, J), int)
    vprev = [None] * J
    for s in range(S):
        for j in range(J):
            K, cs = prep(T[s, j])
            if vprev[j] is None: v0 = np.ones(M)
            else:
                eta = vprev[j]                      # log v - cs of the previous sweep
                v0 = np.exp(np.clip(eta + cs, -600, 600))
            v, ne, nn = fn(K, v0)
            vprev[j] = np.log(v) - cs
            evs[s, j] = ne; nws[s, j] = nn; tot[s, j] = C_EVAL * ne + C_NEWTON * nn
    print('==', name, 'total kcycles per sweep (sweeps 4..23 mean): %.1f' % tot[4:24].sum(1).mean())
    for s in range(S):
        print('  s%2d evals %s | newton %s | kcyc %.0f%s' % (s, ' '.join('%3d' % v for v in evs[s]), ' '.join('%d' % v for v in nws[s]), tot[s].sum(),
              ('   gpu iters ' + ' '.join('%2d' % v for v in IT[s])) if name == 'current' else ''))

def anderson_ne(K, v, m=2, max_it=200, reg=1e-10, guard=True, mix_start=1):
    """Anderson acceleration as the kernel would run it: history of m differences, normal equations (m x m) with a
    relative ridge, restart (plain step, history dropped) when the residual grows."""
    ne = 0; x = np.clip(np.log(np.maximum(v, 1e-300)), -640, 640)
    xs = []; fs = []
    err_prev = np.inf
    nrestart = 0
    for it in range(max_it):
        u, s, c = evalK(K, np.exp(x)); ne += 1
        err = np.max(np.abs(c - 1))
        if err < TOL: break
        f = -np.log(s) - x; f -= f.mean()      # g(x) - x ; scale-free
        if guard and err > err_prev * 1.0 and len(fs) > 0:
            xs, fs = [], []; nrestart += 1      # restart from here with a plain step
        xs.append(x.copy()); fs.append(f.copy())
        if len(fs) > m + 1: xs.pop(0); fs.pop(0)
        k = len(fs) - 1
        if k >= 1 and it >= mix_start:
            dF = np.array([fs[i + 1] - fs[i] for i in range(k)]).T
            dX = np.array([xs[i + 1] - xs[i] for i in range(k)]).T
            A = dF.T @ dF; b = dF.T @ f
            A = A + reg * np.trace(A) / k * np.eye(k)
            try: gamma = np.linalg.solve(A, b)
            except np.linalg.LinAlgError: gamma = np.zeros(k)
            x = x + f - (dX + dF) @ gamma
        else:
            x = x + f
        err_prev = err
    return np.exp(x), ne, nrestart

for m_ in (1, 2, 3, 4):
    methods['and%d' % m_] = (lambda mm: (lambda K, v: anderson_ne(K, v, m=mm)))(m_)

if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'study':
    for name in ('and1', 'and2', 'and3', 'and4'):
        fn = methods[name]
        evs = np.zeros((S, J), int); rs = np.zeros((S, J), int)
        vprev = [None] * J
        for s in range(S):
            for j in range(J):
                K, cs = prep(T[s, j])
                v0 = np.ones(M) if vprev[j] is None else np.exp(np.clip(vprev[j] + cs, -600, 600))
                v, ne, nr = fn(K, v0)
                vprev[j] = np.log(v) - cs
                evs[s, j] = ne; rs[s, j] = nr
        cold = [fn(prep(T[s, j])[0], np.ones(M))[1] for s in (1, 5, 15, 25) for j in range(J)]
        print(name, 'evals/sweep mean (4..23): %.1f  max per solve %d  restarts %d | cold-start evals: max %d mean %.1f' % (
            evs[4:24].sum(1).mean(), evs.max(), rs.sum(), max(cold), np.mean(cold)))
        print('   s1 ', evs[1], ' s4 ', evs[4], ' s15 ', evs[15])
