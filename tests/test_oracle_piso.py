"""CPU tests of the oracle's neoIcoFoam restatement (no GPU): physical sanity of the PISO step the GPU path is
compared with, plus Serial vs OpenMP-atomics agreement of the glue functions."""
import numpy as np

from foamadapter_b200 import piso
from oracle.cpu import Mesh as OMesh
from oracle.piso import IcoFoamOracle


def test_cavity_step_is_divergence_free_and_bounded():
    d = piso.cavity_desc(12)
    om = OMesh.from_desc(d)
    o = IcoFoamOracle(om, nu=0.01, dt=5e-4)
    for _ in range(3):
        out = o.step()
    for st, hist in out:
        assert st["finalResNorm"] <= 1e-6 and st["numIter"] < 1000
        assert len(hist) == st["numIter"] + 1 and hist[0] >= hist[-1]
    cont = om.surface_integrate(o.phi) * om.V
    assert np.abs(cont).max() < 1e-5
    assert np.isfinite(o.U).all() and np.abs(o.U).max() <= 1.5
    # lid drives +x velocity in the top row, return flow below
    n = 12
    Ux = o.U[:, 0].reshape(n, n)
    assert Ux[-1].mean() > 0 and Ux[: n // 2].mean() < 0


def test_glue_serial_equals_parallel_within_rounding():
    d = piso.cavity_desc(6, True)
    om = OMesh.from_desc(d)
    rng = np.random.default_rng(0)
    vals = np.repeat(rng.uniform(0.5, 1.5, om.nnz)[:, None], 3, axis=1).copy()
    rhs, U = rng.uniform(-1, 1, (om.nC, 3)), rng.uniform(-1, 1, (om.nC, 3))
    rAU = om.rAU(vals)
    a, b = om.HbyA(vals, rhs, rAU, U, par=0), om.HbyA(vals, rhs, rAU, U, par=1)
    assert np.allclose(a, b, rtol=1e-12, atol=1e-14)
