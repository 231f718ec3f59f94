"""The algebra behind the sweep's closed-form layer statistics ("What the ci sweep no longer streams"),
checked on the CPU against the oracle's streamed sums: for a ci layer above the first the targets are inferred from
the layer's own posterior, so the P1 residual vanishes and the P4 / P5 sums are polynomials in the coefficients with
basis-only weights s = Phi^T 1, G = Phi^T Phi and D = sum phi_anc^2 over (region x coarser region) pieces."""
import numpy as np

from oracle import mrgp_oracle as O
import workloads


def _ragged_offsets(n, counts, seed):
    rng = np.random.RandomState(seed)
    offs = [np.array([0, n], dtype=np.int64)]
    for r in counts:
        cuts = np.sort(rng.choice(np.arange(1, n), size=r - 1, replace=False))
        offs.append(np.concatenate([[0], cuts, [n]]).astype(np.int64))
    return offs


def test_closed_form_sums_equal_the_streamed_ones():
    n, M = 3000, 12
    x, y = workloads.workload1(n)
    offsets = _ragged_offsets(n, (2, 5, 11), seed=3)          # regions straddle coarser regions
    m = O.OracleMRGP(x, y, M, offsets, mode='ci')
    m.sweep()
    m.sweep()
    rng = np.random.RandomState(7)
    for j in range(1, m.J):                                    # make the upper layers carry signal
        ly = m.layers[j]
        ly.A = 0.05 * rng.standard_normal(ly.A.shape)
        ly.bias_mean = 0.3 * rng.standard_normal(ly.bias_mean.shape)
    # one more sweep, recording what the streamed update saw for every layer
    seen = {}
    orig_bn, orig_sga = m._bias_noise, m._scale_given_axis

    def spy_sga(ly, yt, ard_mean):
        j = m.layers.index(ly)
        seen.setdefault(j, {})['A_old'] = ly.A.copy()
        seen[j]['b_old'] = ly.bias_mean.copy()
        orig_sga(ly, yt, ard_mean)
        seen[j]['ytil'] = ly.ytil.copy()

    def spy_bn(ly, yt, y_var):
        orig_bn(ly, yt, y_var)
        j = m.layers.index(ly)
        seen[j].update(sum_r=ly.sum_r.copy(), mean_term=ly.mean_term.copy(), var_f=ly.var_f.copy(),
                       var_au=ly.var_au.copy(), A_new=ly.A.copy())
    m._bias_noise, m._scale_given_axis = spy_bn, spy_sga
    m.sweep()
    for j in range(1, m.J):
        ly, s = m.layers[j], seen[j]
        off = ly.off
        # P1: Phi^T r == 0, hence y_tilde_i = d_i a_i (Posteriors.py:61-78 with inferred targets)
        want = ly.d[:, None, :] * s['A_old'] if s['A_old'].shape[1] == 2 else None
        assert want is not None
        assert np.allclose(s['ytil'], want, rtol=1e-9, atol=1e-12 * np.abs(want).max())
        dA = s['A_old'] - s['A_new']                           # (R, dy, M)
        for c in range(ly.R):
            rows = slice(off[c], off[c + 1])
            Phi = ly.Phi[rows]
            nn = Phi.shape[0]
            svec, G = Phi.sum(0), Phi.T @ Phi
            b = s['b_old'][c]
            sd = dA[c] @ svec                                  # (dy,)
            assert np.allclose(s['sum_r'][c], sd + nn * b, rtol=1e-9, atol=1e-10)
            quad = sum(dA[c, d] @ G @ dA[c, d] for d in range(2))
            assert np.isclose(s['mean_term'][c], quad + 2 * b @ sd + nn * b @ b, rtol=1e-9, atol=1e-10)
            assert np.isclose(s['var_au'][c], ly.d[c] @ ly.cm2[c], rtol=1e-12)
            fv = 0.0
            for jp in range(j):                                # pieces region x coarser region
                lp = m.layers[jp]
                a = np.searchsorted(lp.off, off[c], side='right') - 1
                while lp.off[a] < off[c + 1]:
                    lo, hi = max(off[c], lp.off[a]), min(off[c + 1], lp.off[a + 1])
                    D = np.sum(lp.Phi[lo:hi] ** 2, axis=0)
                    fv += (hi - lo) * lp.bias_var[a] + lp.cm2[a] @ D
                    a += 1
                    if a >= lp.R:
                        break
            assert np.isclose(s['var_f'][c], fv, rtol=1e-10)
