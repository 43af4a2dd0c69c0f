"""CPU: the oracle's scalarAdvection restatement (oracle/advection.py; reference examples/scalarAdvection,
test/test_advection.cpp) -- properties the scheme guarantees, on the reference's 50x50x1-style mesh at small size."""
import numpy as np
import pytest

from foamadapter_b200.advection import advection_desc, init_fields
from oracle.advection import ScalarAdvectionOracle, init_fields as oracle_init_fields
from oracle.cpu import Mesh as OMesh


@pytest.fixture(scope="module")
def om():
    return OMesh.from_desc(advection_desc(24))


def test_initial_fields_match_between_product_host_code_and_oracle(om):
    C = om.C.reshape(-1, 3)
    U1, T1 = init_fields(C)
    U2, T2 = oracle_init_fields(C)
    assert np.array_equal(U1, U2) and np.array_equal(T1, T2)
    assert T1.max() <= 1.0 and T1.min() >= 0.0 and np.all(U1[:, 2] == 0.0)


@pytest.mark.parametrize("ddt", ["forwardEuler", "Runge-Kutta"])
def test_forward_euler_upwind_is_bounded_and_conservative(om, ddt):
    run = ScalarAdvectionOracle(om, dt=2e-3, endTime=0.2, scheme=1, ddt=ddt)
    m0 = float((run.T * om.V).sum())
    for _ in range(20):
        run.step()
    assert run.coNum[0] < 1.0
    assert run.T.min() >= -1e-15 and run.T.max() <= 1.0 + 1e-12          # upwind + CFL < 1: monotone
    # div form: what leaves through the (nearly closed) walls is the only mass change
    assert abs(float((run.T * om.V).sum()) - m0) <= 1e-6 * m0


def test_backward_euler_tracks_forward_euler_as_dt_shrinks(om):
    errs = []
    for dt in (2e-3, 5e-4):
        a = ScalarAdvectionOracle(om, dt=dt, endTime=0.2, scheme=1, ddt="forwardEuler")
        b = ScalarAdvectionOracle(om, dt=dt, endTime=0.2, scheme=1, ddt="backwardEuler", maxIter=50)
        for _ in range(int(round(0.02 / dt))):
            a.step(); b.step()
        assert all(s["finalResNorm"] <= 1e-12 * max(s["initResNorm"], 1e-300) or s["numIter"] == 50 for s in b.stats)
        errs.append(np.abs(a.T - b.T).max())
    assert errs[1] < 0.5 * errs[0] and errs[0] < 0.05   # first-order in dt
