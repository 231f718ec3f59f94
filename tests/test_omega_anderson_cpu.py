"""The omega solve of the fused sweep, restated in NumPy (oracle/omega_anderson.py), against the exact scaling of the oracle
(mrgp_oracle.omega_sinkhorn, the fixed point the reference's fsolve call approximates, Stats.py:413-420)."""
import numpy as np

import workloads  # noqa: F401  (tests/golden on sys.path via conftest)
from oracle import mrgp_oracle as O
from oracle import omega_anderson as A


def _tables(seed, m, peak):
    rng = np.random.default_rng(seed)
    lw = rng.standard_normal((m, m)) * peak
    lw[np.arange(m), rng.permutation(m)] += 3.0 * peak      # a dominant permutation, as the model's tables have
    return lw


def test_same_fixed_point_as_the_exact_scaling():
    for seed, m, peak in ((0, 30, 1.0), (3, 8, 0.1), (5, 30, 1.5)):
        lw = _tables(seed, m, peak)
        om, evals, eta = A.solve(lw)
        ref = O.omega_sinkhorn(lw)
        assert np.abs(om.sum(0) - 1).max() < 1e-9 and np.abs(om.sum(1) - 1).max() < 1e-12
        assert np.abs(om - ref).max() < 1e-8, (seed, np.abs(om - ref).max())
        assert evals <= 48


def test_tables_of_a_model_converge_in_a_few_evaluations(monkeypatch):
    """The tables the solver meets in a model (captured from three oracle sweeps of the reference's workload): the twin
    reaches the oracle's exact scaling, from the previous sweep's solution, in at most 20 evaluations per layer."""
    import workloads
    seen = []
    exact = O.omega_sinkhorn
    monkeypatch.setattr(O, 'omega_sinkhorn', lambda lw, *a, **k: (seen.append(np.array(lw)), exact(lw, *a, **k))[1])
    x, y = workloads.workload1(4000)
    ora = O.OracleMRGP(x, y, 30, O.uniform_offsets(4000, 5, 2), mode='ci', omega_solver='sinkhorn')
    for _ in range(3):
        ora.sweep()
    J = 6
    assert len(seen) == 3 * J
    eta = [None] * J
    worst = 0
    for k, lw in enumerate(seen):
        j = k % J
        om, evals, eta[j] = A.solve(lw, eta[j])
        assert np.abs(om - exact(lw)).max() < 1e-8
        worst = max(worst, evals)
    assert worst <= 20, worst


def test_warm_start_and_acceleration_need_few_evaluations():
    lw = _tables(5, 30, 1.5)
    om, cold, eta = A.solve(lw)
    # a slowly drifting table (what consecutive sweeps see): a handful of evaluations from the previous solution
    drift = lw + 1e-3 * np.random.default_rng(6).standard_normal(lw.shape)
    om2, warm, _ = A.solve(drift, eta)
    assert warm <= cold
    # plain Sinkhorn on the same table needs many more sweeps than the accelerated iteration
    K, _ = A.shifted_table(lw)
    v, plain = np.ones(30), 0
    while plain < 5000:
        u, s, c = A.evaluate(K, v)
        if np.abs(c - 1).max() < A.TOL:
            break
        v = 1.0 / s
        plain += 1
    assert cold < plain


def test_budget_runs_out_on_unstructured_peaked_tables():
    """What the accelerated iteration does NOT do: Gaussian random tables with a wide spread (no model produces them) leave
    it - and the 2000 plain Sinkhorn sweeps behind it - short of the tolerance.  The count says so (>= 2040), and the host
    mirror then moves the model to the sweep with the Sinkhorn / Newton solver (MRGP.omega_solve_report)."""
    om, evals, _ = A.solve(_tables(6, 30, 3.0))
    assert evals >= 2040
    assert np.abs(om.sum(1) - 1).max() < 1e-12          # rows are exact in any case
