"""The C restatement of the ci sweep (oracle/mrgp_port.c, the CPU baseline of bench.py) against the NumPy oracle
(which is pinned to the reference's own outputs, tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from oracle import mrgp_oracle as O
from oracle import port_c
from parity import mismatch
import workloads


@pytest.mark.parametrize('n,res,m', [(3000, 4, 30), (2048, 5, 20), (777, 0, 8)])
def test_c_port_matches_numpy_oracle(n, res, m):
    x, y = workloads.workload1(n)
    offsets = O.uniform_offsets(n, res, 2)
    ora = O.OracleMRGP(x, y, m, offsets, mode='ci')
    port = port_c.PortC(x, y, m, offsets)
    for _ in range(3):
        ora.sweep()
    port.sweep(3)
    ref, got = ora.state(), port.state()
    for k, v in got.items():
        # omega: the reference's fsolve is converged to ~1e-8 (Stats.py:413), the port iterates to 1e-13
        assert mismatch(v, ref[k], 2e-6, atol_scale=1e-9 if k == 'S.omega' else 1e-12) is None, (k, mismatch(v, ref[k], 2e-6))
    assert port.threads >= 1
