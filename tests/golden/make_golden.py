"""Generate the golden fixtures by running the UNMODIFIED reference (/root/reference/src).

Run in the build container only (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py [name ...]
Writes tests/golden/<name>.npz.  The flat key names are those of oracle.mrgp_oracle.OracleMRGP.state().
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
import workloads  # noqa: E402

R = ref_loader.load()


def ref_state(m):
    """Flatten the reference model's state (MRGP.py public attributes)."""
    out = {}
    fi = m.forced_independence
    for j in range(m.n_layers):
        p = 'L%d.' % j
        q, st = m.posterior_obj[j], m.stats_obj[j]
        out[p + 'L'] = np.array(m.train_basis_intervals[j])
        out[p + 'lam'] = np.array(m.lambda_[j])
        out[p + 'S'] = np.array(m.spectral_density_prior[j])
        out[p + 'd'] = np.array([np.sum(ph * ph, axis=0) for ph in m.phi_x[j]])
        out[p + 'scale_precision'] = np.array(q.scale_precision)
        out[p + 'zeta'] = np.array(q.scale_mean_zeta)
        out[p + 'ytil'] = np.array(q.scale_mean_y_tilde)
        out[p + 'A'] = np.array(st.scale_axis_mean)
        out[p + 'm2'] = np.array(st.scale_moment2)
        out[p + 'cm2'] = np.array(st.scale_axis_central_moment2)
        out[p + 'noise_shape'] = np.array(q.noise_gamma_shape, dtype=np.float64)
        out[p + 'noise_scale'] = np.array(q.noise_gamma_scale, dtype=np.float64)
        out[p + 'noise_mean'] = np.array(st.noise_mean, dtype=np.float64)
        out[p + 'noise_log_mean'] = np.array(st.noise_log_mean, dtype=np.float64)
        out[p + 'bias_prec'] = np.array(q.bias_normal_precision, dtype=np.float64)
        out[p + 'bias_mean'] = np.array(st.bias_mean, dtype=np.float64)
        out[p + 'bias_var'] = np.array(st.bias_var, dtype=np.float64)
        out[p + 'fbar'] = np.concatenate([np.asarray(a) for a in st.latent_f_mean])
        out[p + 'fvar'] = np.concatenate([np.asarray(a).ravel() for a in st.latent_f_var])
        if fi:
            out[p + 'B'] = np.array(q.axis_bingham_b)
            out[p + 'kappa'] = np.array(q.axis_bingham_kappa)
            out[p + 'rho'] = np.array(q.axis_bingham_rho)
            out[p + 'logC'] = np.array(q.axis_bingham_log_const)
            out[p + 'axis_cov'] = np.array(st.axis_cov)
            out[p + 'ard_shape'] = np.array(q.ard_gamma_shape)
            out[p + 'ard_scale'] = np.array(q.ard_gamma_scale)
            out[p + 'ard_mean'] = np.array(st.ard_mean)
            out[p + 'ard_log_mean'] = np.array(st.ard_log_mean)
    if not fi:
        q, st = m.shared_posterior, m.shared_stats
        out['S.B'] = np.array(q.axis_bingham_b)
        out['S.kappa'] = np.array(q.axis_bingham_kappa)
        out['S.rho'] = np.array(q.axis_bingham_rho)
        out['S.logC'] = np.array(q.axis_bingham_log_const)
        out['S.axis_cov'] = np.array(st.axis_cov)
        out['S.ard_shape'] = np.array(q.ard_gamma_shape)
        out['S.ard_scale'] = np.array(q.ard_gamma_scale)
        out['S.ard_mean'] = np.array(st.ard_mean)
        out['S.ard_log_mean'] = np.array(st.ard_log_mean)
        out['S.omega'] = np.array(st.omega)
    return out


def build(x, y, n_basis, resolution, fi, adaptive=False, divider=2, **kw):
    idx = R.IndexSetGenerator.IndexSetUniform(sample_length=x.shape[0], resolution=resolution, divider=divider)
    bi = R.BasisInterval.BasisInterval(opt_interval_factor=(1, 1.2)) if adaptive else None
    return R.MRGP.MultiResolutionGaussianProcess(
        train_xy=[x, y], n_basis=n_basis, index_set_obj=idx,
        basis_function_obj=R.KernelClass.LaplacianEigenpairs(),
        spectral_density_obj=R.KernelClass.MaternKernel(nu=1, l=1, sf=1),
        adaptive_inputs=kw.get('input_model') is not None, standard_normalized_inputs=True, basis_interval_obj=bi, interval_factor=1,
        forced_independence=fi, **kw)


def run_sweeps(name, x, y, n_basis, resolution, fi, checkpoints, adaptive=False, predict=None, **kw):
    t0 = time.time()
    m = build(x, y, n_basis, resolution, fi, adaptive, **kw)
    out = {'meta.N': x.shape[0], 'meta.M': n_basis, 'meta.resolution': resolution, 'meta.fi': int(fi),
           'meta.adaptive': int(adaptive), 'meta.checkpoints': np.array(checkpoints), 'x': x, 'y': y}
    for k, v in ref_state(m).items():
        out['k0.' + k] = v
    done = 0
    for k in checkpoints:
        for _ in range(k - done):
            if fi:
                m._independent_fit()
            else:
                m._fit()
        done = k
        for key, v in ref_state(m).items():
            out['k%d.%s' % (k, key)] = v
        print(name, 'sweep', k, '%.1fs' % (time.time() - t0), flush=True)
    if predict is not None:
        xt = predict
        out['pred.x'] = xt
        out['pred.mean_global'] = m.get_predicted_mean(xt)
        idx_t = R.IndexSetGenerator.IndexSetUniform(sample_length=xt.shape[0], resolution=resolution, divider=2)
        out['pred.mean_indexed'] = m.get_predicted_mean(xt, index_set_obj=idx_t)
        out['pred.var_global'] = m.get_central_moment2(xt)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)


def run_elbo(name, x, y, n_basis, resolution, n_iter, **kw):
    """fit(n_iter, tol) path: ELBO per layer and per term after every sweep (MRGP.py:373-399)."""
    m = build(x, y, n_basis, resolution, False, **kw)
    terms = []
    orig = m._compute_lower_bound

    def spy(prime_shared_posterior):
        t = np.zeros((m.n_layers, 6))
        for j in range(m.n_layers):
            t[j] = [m._data_likelihood(res=j), m._ll_scale_given_axis(res=j),
                    m._ll_axis(res=j, prime_shared_posterior=prime_shared_posterior),
                    m._ll_ard(res=j, prime_shared_posterior=prime_shared_posterior),
                    m._ll_bias(res=j), m._ll_noise(res=j)]
        terms.append(t)
        return orig(prime_shared_posterior=prime_shared_posterior)
    m._compute_lower_bound = spy
    m.fit(n_iter=n_iter, tol=1e-300, min_iter=n_iter)
    out = {'meta.N': x.shape[0], 'meta.M': n_basis, 'meta.resolution': resolution, 'x': x, 'y': y,
           'lower_bound': np.array(m.lower_bound, dtype=np.float64),
           'lower_bound_layer': np.array(m.lower_bound_layer, dtype=np.float64),
           'terms': np.array(terms)}
    for key, v in ref_state(m).items():
        out['final.' + key] = v
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'done', flush=True)


def run_kats(name):
    """Known-answer vectors for the primitives and index sets."""
    out = {}
    bc = R.computeRealBinghamConstant.logPartition_saddle
    rng = np.random.RandomState(7)
    kap2 = np.concatenate([np.zeros((1, 2)), np.array([[125.81942169, 17.94505309], [3., 1.], [1e-90, 0.],
                                                        [5e3, 5e3 - 1e-3], [2.5e6, 1.0]]),
                           np.sort(rng.gamma(1., 30., size=(40, 2)), axis=1)[:, ::-1]])
    lc, rho = bc(kap2)
    out['saddle2.kappa'], out['saddle2.logC'], out['saddle2.rho'] = kap2, lc, rho
    kap3 = np.concatenate([np.zeros((1, 3)), np.array([[3., 1., 0.]]),
                           np.sort(rng.gamma(1., 10., size=(20, 3)), axis=1)[:, ::-1]])
    lc, rho = bc(kap3)
    out['saddle3.kappa'], out['saddle3.logC'], out['saddle3.rho'] = kap3, lc, rho
    mats = [np.array([[2, .5], [.5, 1.]])]
    for _ in range(20):
        a = rng.randn(2, 3)
        mats.append(a @ a.T * rng.gamma(1., 20.))
    for _ in range(6):
        v = rng.randn(2, 1)
        mats.append(v @ v.T * rng.gamma(1., 20.))  # rank-1: exercises isPD / nearestPD
    mats = np.array(mats)
    kk, rr, ll, cc, bb = [], [], [], [], []
    for b in mats:
        sc = R.SanityCheck.SanityCheck()
        b2 = b if sc.isPD(b) else sc.nearestPD(b)
        bg = R.CommonDensities.Bingham(b2)
        kk.append(np.real(bg.kappa)); rr.append(np.real(bg.rho)); ll.append(np.real(bg.log_const))
        cc.append(np.real(np.dot(bg.rho * bg.axes, bg.axes.T))); bb.append(b2)
    out['bingham.B_in'], out['bingham.B'] = mats, np.array(bb)
    out['bingham.kappa'], out['bingham.rho'] = np.array(kk), np.array(rr)
    out['bingham.logC'], out['bingham.axis_cov'] = np.array(ll), np.array(cc)
    xk = np.array([[-1.5], [0.25], [1.0], [1.9999], [-2.0]])
    le = R.KernelClass.LaplacianEigenpairs()
    mk = R.KernelClass.MaternKernel(1, 1, 1)
    phi = np.array([le.get_eigenpairs(xk, basis_id=i, basis_interval=[2.0])[0] for i in range(1, 41)]).T
    lam = np.array([le.get_eigenpairs(xk, basis_id=i, basis_interval=[2.0])[1] for i in range(1, 41)])
    out['basis.x'], out['basis.L'], out['basis.phi'], out['basis.lam'] = xk, np.array([2.0]), phi, lam
    out['basis.S'] = np.array([mk.spectral(np.sqrt(l_)) for l_ in lam])
    mk2 = R.KernelClass.MaternKernel(2.5, 0.7, 1.3)
    out['basis.S_nu2.5_l0.7_sf1.3'] = np.array([mk2.spectral(np.sqrt(l_)) for l_ in lam])
    cases = [(32, 5, 2), (160, 7, 2), (100000, 7, 2), (1000000, 9, 2), (2048, 5, 2), (1000, 3, 3), (7, 0, 2),
             (100, 2, 5)]
    out['index.cases'] = np.array(cases)
    for (n, res, div) in cases:
        idx = R.IndexSetGenerator.IndexSetUniform(n, res, div)
        for j, regions in enumerate(idx.index_set):
            off = [regions[0][0]] + [r[-1] + 1 for r in regions]
            for r in regions:
                assert r == list(range(r[0], r[-1] + 1))
            out['index.%d_%d_%d.L%d' % (n, res, div, j)] = np.array(off, dtype=np.int64)
    # omega: captured (log_omega_hat, omega) pairs from a short ci run
    import Stats
    caps = []
    orig = Stats.fsolve

    def spy(func, x0, args):
        sol = orig(func, x0, args)
        caps.append((np.array(args), np.array(sol)))
        return sol
    Stats.fsolve = spy
    x, y = workloads.workload1(256)
    m = build(x, y, 30, 3, False)
    for _ in range(3):
        m._fit()
    Stats.fsolve = orig
    out['omega.log_omega_hat'] = np.array([c[0] for c in caps])
    out['omega.ln_eta'] = np.array([c[1] for c in caps])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'done', flush=True)


def run_predvar(name, x, y, n_basis, resolution, fi, n_sweeps, xt):
    """Index-set form of get_central_moment2 / get_test_likelihood (MRGP.py:863-932).  The call overwrites the
    latent functions of the model (MRGP.py:893-901), so it comes last."""
    m = build(x, y, n_basis, resolution, fi)
    for _ in range(n_sweeps):
        if fi:
            m._independent_fit()
        else:
            m._fit()
    out = {'meta.N': x.shape[0], 'meta.M': n_basis, 'meta.resolution': resolution, 'meta.fi': int(fi),
           'meta.sweeps': n_sweeps, 'x': x, 'y': y, 'pred.x': xt}
    idx_t = R.IndexSetGenerator.IndexSetUniform(sample_length=xt.shape[0], resolution=resolution, divider=2)
    yt = workloads.signal1(xt)[:, :, 0].T
    out['pred.y'] = yt
    out['pred.mean_indexed'] = m.get_predicted_mean(xt, index_set_obj=idx_t)
    out['pred.var_global'] = m.get_central_moment2(xt)
    out['pred.var_indexed'] = m.get_central_moment2(xt, index_set_obj=idx_t)
    out['pred.test_likelihood_indexed'] = np.array(m.get_test_likelihood([xt, yt], index_set_obj=idx_t))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'done', out['pred.var_indexed'][:3], float(out['pred.test_likelihood_indexed']), flush=True)


JOBS = {
    'kats': lambda: run_kats('kats'),
    'c1_ci': lambda: run_sweeps('c1_ci', *workloads.workload1(32), 30, 5, False, [1, 3, 15],
                                predict=np.atleast_2d(np.linspace(1, 3, 1000)).T),
    'c1_fi': lambda: run_sweeps('c1_fi', *workloads.workload1(32), 30, 5, True, [1, 3, 15],
                                predict=np.atleast_2d(np.linspace(1, 3, 1000)).T),
    'c1_ci_adaptive': lambda: run_sweeps('c1_ci_adaptive', *workloads.workload1(32), 30, 5, False, [1, 3],
                                         adaptive=True),
    'c1_ci_elbo': lambda: run_elbo('c1_ci_elbo', *workloads.workload1(32), 30, 5, 4),
    'c2_ci': lambda: run_sweeps('c2_ci', *workloads.workload2(), 40, 7, False, [1, 3]),
    'c2_fi': lambda: run_sweeps('c2_fi', *workloads.workload2(), 40, 7, True, [1, 3]),
    'n2000_ci': lambda: run_sweeps('n2000_ci', *workloads.workload1(2000), 30, 5, False, [1, 3]),
    'n2000_fi': lambda: run_sweeps('n2000_fi', *workloads.workload1(2000), 30, 5, True, [1, 3]),
    'predvar_ci': lambda: run_predvar('predvar_ci', *workloads.workload1(600), 20, 3, False, 3,
                                      np.atleast_2d(np.linspace(1, 3, 1000)).T),
    'predvar_fi': lambda: run_predvar('predvar_fi', *workloads.workload1(600), 20, 3, True, 3,
                                      np.atleast_2d(np.linspace(1, 3, 1000)).T),
    'shared_ci_nb': lambda: run_sweeps('shared_ci_nb', *workloads.workload1(600), 20, 3, False, [1, 3],
                                       noise_region_specific=True, bias_region_specific=False),
    'shared_ci_sn': lambda: run_sweeps('shared_ci_sn', *workloads.workload1(600), 20, 3, False, [1, 3],
                                       noise_region_specific=False, bias_region_specific=True),
    'shared_ci_ss': lambda: run_sweeps('shared_ci_ss', *workloads.workload1(600), 20, 3, False, [1, 3],
                                       noise_region_specific=False, bias_region_specific=False,
                                       predict=np.atleast_2d(np.linspace(1, 3, 500)).T),
    'shared_fi_nb': lambda: run_sweeps('shared_fi_nb', *workloads.workload1(600), 20, 3, True, [1, 3],
                                       noise_region_specific=True, bias_region_specific=False),
    'shared_fi_sn': lambda: run_sweeps('shared_fi_sn', *workloads.workload1(600), 20, 3, True, [1, 3],
                                       noise_region_specific=False, bias_region_specific=True),
    'shared_fi_ss': lambda: run_sweeps('shared_fi_ss', *workloads.workload1(600), 20, 3, True, [1, 3],
                                       noise_region_specific=False, bias_region_specific=False,
                                       predict=np.atleast_2d(np.linspace(1, 3, 500)).T),
    'shared_ci_ss_elbo': lambda: run_elbo('shared_ci_ss_elbo', *workloads.workload1(600), 20, 3, 3,
                                          noise_region_specific=False, bias_region_specific=False),
    'shared_ci_nb_elbo': lambda: run_elbo('shared_ci_nb_elbo', *workloads.workload1(600), 20, 3, 3,
                                          noise_region_specific=True, bias_region_specific=False),
    'c2_ci_warp': lambda: run_sweeps('c2_ci_warp', *workloads.workload2(), 40, 7, False, [1, 3],
                                     predict=np.atleast_2d(np.linspace(1, 4, 300)).T,
                                     input_model=workloads.InterpInputModel(workloads.workload2()[0])),
    'c1_ci_adaptive_elbo': lambda: run_elbo('c1_ci_adaptive_elbo', *workloads.workload1(32), 30, 5, 3, adaptive=True),
    'n600_ci_adaptive_elbo': lambda: run_elbo('n600_ci_adaptive_elbo', *workloads.workload1(600), 20, 3, 3, adaptive=True),
    # one series of BASELINE config 5 (N = 2048, 6 resolutions, seed 10 + s with s = 0 and s = 3)
    'c5_ci': lambda: run_sweeps('c5_ci', *workloads.workload1(2048, seed=10), 30, 5, False, [1, 3]),
    'c5_fi': lambda: run_sweeps('c5_fi', *workloads.workload1(2048, seed=13), 30, 5, True, [1, 3]),
    'n600_ci_snr_shared': lambda: run_sweeps('n600_ci_snr', *workloads.workload1(600), 20, 3, False, [1, 3],
                                             snr_ratio=10.),
}

if __name__ == '__main__':
    names = sys.argv[1:] or list(JOBS)
    for n in names:
        JOBS[n]()
