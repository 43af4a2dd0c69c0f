"""GPU: BiCGStab (scalar and Vec3) and the scalarAdvection case (BASELINE.json configs[3]) through the C ABI against the
CPU oracle. Bars: the reference's own (test/test_advection.cpp:169,228): 1e-10 forward Euler, 1e-8 backward Euler; the
explicit path is in fact bit-identical (same kernels as test_explicit_gpu.py + an axpby)."""
import numpy as np
import pytest
import torch

from foamadapter_b200 import advection as adv, dsl, fvcc, la, mesh as M, ops
from oracle.advection import ScalarAdvectionOracle
from oracle.cpu import Mesh as OMesh, bicgstab as oracle_bicgstab

pytestmark = pytest.mark.gpu

dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
host = lambda t: t.detach().cpu().numpy()


def _advection_system(om, rng, dt=1e-3):
    """ddt + upwind div of a random flux: the nonsymmetric, diagonally dominant matrix backward Euler solves."""
    flux = rng.uniform(-1, 1, om.nF) * om.magSf
    bd = dict(value=np.zeros(om.nB), refValue=np.zeros(om.nB), valueFraction=np.zeros(om.nB), refGrad=np.zeros(om.nB))
    ls = om.empty_system(False)
    om.div_imp(ls, flux, bd, 1, 1.0, None)
    om.ddt_imp(ls, rng.uniform(0, 1, om.nC), dt, 1.0, None)
    return ls


@pytest.mark.parametrize("jacobi", [True, False])
@pytest.mark.parametrize("check_every", [1, 5])
def test_bicgstab_tracks_oracle_history(jacobi, check_every):
    d = M.MeshDesc.block(20, 14, 9)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    rng = np.random.default_rng(3)
    ls = _advection_system(om, rng, dt=2e-2)   # 9 / 14 iterations with / without Jacobi: well inside BiCGStab's stable range
    x0 = rng.uniform(-1, 1, om.nC)
    xo, so, ho = oracle_bicgstab(om.rowOffs, om.colIdxs, ls["values"], ls["rhs"], x0, jacobi=jacobi, max_iter=100, rel_tol=1e-12, max_hist=400)
    cfg = {"solver": "Ginkgo", "type": "solver::Bicgstab", "criteria": {"iteration": 100, "relative_residual_norm": 1e-12}}
    if jacobi:
        cfg["preconditioner"] = {"type": "preconditioner::Jacobi", "max_block_size": 1}
    gls = la.LinearSystem(gm)
    gls.values.copy_(dev(ls["values"])); gls.rhs.copy_(dev(ls["rhs"]))
    x = dev(x0.copy())
    st = la.Solver(cfg, check_every=check_every, history=True).solve(gls, x)
    assert so["numIter"] >= 3
    assert abs(st.initResNorm - so["initResNorm"]) <= 1e-13 * so["initResNorm"]
    assert abs(st.numIter - so["numIter"]) <= 1
    n = min(len(st.history), len(ho))
    sig = ho[:n] > 1e-8 * ho[0]
    assert np.allclose(st.history[:n][sig], ho[:n][sig], rtol=1e-6, atol=0)   # BiCGStab amplifies rounding more than CG
    assert np.allclose(host(x), xo, rtol=1e-8, atol=1e-10 * np.abs(xo).max())
    r = om.residual(ls["values"], ls["rhs"], host(x))
    assert np.linalg.norm(r) <= 2e-12 * np.linalg.norm(ls["rhs"])


def test_bicgstab_iteration_cap_and_converged_start():
    d = M.MeshDesc.block(12, 10, 8)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    rng = np.random.default_rng(4)
    ls = _advection_system(om, rng, dt=2e-2)
    gls = la.LinearSystem(gm); gls.values.copy_(dev(ls["values"])); gls.rhs.copy_(dev(ls["rhs"]))
    x = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
    cap = {"solver": "Ginkgo", "type": "solver::Bicgstab", "criteria": {"iteration": 2, "relative_residual_norm": 0.0}}
    st = la.Solver(cap).solve(gls, x)
    xo, so, _ = oracle_bicgstab(om.rowOffs, om.colIdxs, ls["values"], ls["rhs"], np.zeros(om.nC), jacobi=False, max_iter=2, rel_tol=0.0)
    assert st.numIter == so["numIter"] == 2 and np.allclose(host(x), xo, rtol=1e-10, atol=1e-14)
    full = {"solver": "Ginkgo", "type": "solver::Bicgstab", "criteria": {"iteration": 200, "relative_residual_norm": 1e-13}}
    la.Solver(full).solve(gls, x)
    before = host(x).copy()
    st = la.Solver({"solver": "Ginkgo", "type": "solver::Bicgstab", "criteria": {"iteration": 200, "relative_residual_norm": 1e-9}}).solve(gls, x)
    assert st.numIter == 0 and np.array_equal(host(x), before)


def test_vec3_solve_equals_three_scalar_solves():
    d = M.MeshDesc.block(10, 9, 8)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    rng = np.random.default_rng(8)
    ls = _advection_system(om, rng, dt=0.1)
    b3 = rng.uniform(-1, 1, (om.nC, 3))
    cfg = {"solver": "PBiCGStab", "preconditioner": "DIC", "tolerance": 0.0, "relTol": 1e-12, "maxIter": 100}
    v3 = la.LinearSystem(gm, 3)
    v3.values.copy_(dev(np.repeat(ls["values"][:, None], 3, axis=1))); v3.rhs.copy_(dev(b3))
    x3 = torch.zeros((om.nC, 3), dtype=torch.float64, device="cuda")
    stats = la.Solver(cfg).solve(v3, x3)
    assert len(stats) == 3
    for c in range(3):
        xo, so, _ = oracle_bicgstab(om.rowOffs, om.colIdxs, ls["values"], b3[:, c].copy(), np.zeros(om.nC), jacobi=True, max_iter=100, rel_tol=1e-12)
        assert abs(stats[c].numIter - so["numIter"]) <= 1
        assert np.allclose(host(x3)[:, c], xo, rtol=1e-8, atol=1e-11 * np.abs(xo).max())


# ---- scalarAdvection ------------------------------------------------------------------------------------------------
def _pair(n, three_d, ddt, dt, endTime, scheme="upwind", steps=10, fvSolution=None):
    d = adv.advection_desc(n, three_d)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    C = om.C.reshape(-1, 3)
    U, T = adv.init_fields_columns(C, n * n) if three_d else adv.init_fields(C)
    schemes = {"ddtSchemes": {"type": ddt, "Runge-Kutta-Method": "Forward-Euler"}, "divSchemes": {"div(phi,nfT)": f"Gauss {scheme}"}}
    app = adv.ScalarAdvection(gm, dt, endTime, fvSchemes=schemes, fvSolution=fvSolution, U=U, T=T)
    ref = ScalarAdvectionOracle(om, dt, endTime, scheme=ops.SCHEMES[scheme], ddt=ddt, U=U, T=T, maxIter=20, relTol=1e-14)
    for _ in range(steps):
        app.step(); ref.step()
    return app, ref, om


@pytest.mark.parametrize("scheme", ["upwind", "linear"])
@pytest.mark.parametrize("ddt", ["forwardEuler", "Runge-Kutta"])
def test_forward_euler_advection_50x50(scheme, ddt):
    # test/setup_advection: 50 x 50 x 1, the reference's bar is 1e-10 vs OpenFOAM (test_advection.cpp:169)
    app, ref, om = _pair(50, False, ddt, 1e-3, 0.1, scheme, steps=25)
    assert np.array_equal(host(app.phi0.internal), ref.phi0)
    assert np.array_equal(host(app.phi.internal), ref.phi)
    T = host(app.T.internal)
    assert np.abs(T - ref.T).max() <= 1e-10 * np.abs(ref.T).max()
    assert np.array_equal(T, ref.T)      # and in fact bit for bit
    co = host(app.coNum)
    assert co[0] == ref.coNum[0] and abs(co[1] - ref.coNum[1]) <= 1e-12 * ref.coNum[1]
    assert abs(app.t - ref.t) < 1e-15


def test_backward_euler_advection_50x50():
    # test_advection.cpp:176-228: Bicgstab + Jacobi, 20 iterations, 1e-14; bar 1e-8
    app, ref, om = _pair(50, False, "backwardEuler", 1e-3, 0.1, "upwind", steps=10)
    T = host(app.T.internal)
    assert np.abs(T - ref.T).max() <= 1e-8 * np.abs(ref.T).max()
    assert abs(app.stats.numIter - ref.stats[-1]["numIter"]) <= 1


def test_forward_euler_advection_64_cubed():
    app, ref, om = _pair(64, True, "forwardEuler", 5e-4, 0.1, "upwind", steps=4)
    assert np.array_equal(host(app.T.internal), ref.T)


def test_fused_forward_euler_step_equals_the_dsl_sequence():
    """fvk_div_forward_euler_s (T = old - dt * div written by the div kernel) vs dsl::solve's div -> source -> waxpby: same bits."""
    d = adv.advection_desc(24, True)
    C = OMesh.from_desc(d).C.reshape(-1, 3)
    U, T = adv.init_fields_columns(C, 24 * 24)
    for scheme in ("upwind", "linear"):
        schemes = {"ddtSchemes": {"type": "forwardEuler"}, "divSchemes": {"div(phi,nfT)": f"Gauss {scheme}"}}
        a = adv.ScalarAdvection(M.UnstructuredMesh(d), 5e-4, 0.1, fvSchemes=schemes, U=U, T=T)
        b = adv.ScalarAdvection(M.UnstructuredMesh(d), 5e-4, 0.1, fvSchemes=schemes, U=U, T=T, fuse_euler=False)
        assert a.fuse_euler and not b.fuse_euler
        for _ in range(6):
            a.step(); b.step()
            assert torch.equal(a.T.internal, b.T.internal) and torch.equal(a.T.boundary.value, b.T.boundary.value)
    with pytest.raises(Exception):   # in place is a race: rejected
        ops.div_forward_euler(a.mesh, a.phi.internal, a.T.internal, a.T.boundary.value, 1e-3, a.T.internal, ops.UPWIND)


def test_adjust_time_step_follows_setDeltaT():
    d = adv.advection_desc(32)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    app = adv.ScalarAdvection(gm, 1e-4, 1.0, adjustTimeStep=True, maxCo=0.1)
    dts = []
    for _ in range(5):
        app.step(); dts.append(app.dt)
    # auxiliary/setup.cpp:13-22: far below maxCo the step grows by the 1.2 cap
    assert np.allclose(np.array(dts) / np.array([1e-4] + dts[:-1]), 1.2)


def test_unknown_time_integrator_and_rk_tables():
    d = adv.advection_desc(8)
    gm = M.UnstructuredMesh(d)
    T = fvcc.VolumeField(gm, "nfT", 1, [("zeroGradient", 0.0)])
    phi = fvcc.SurfaceField(gm, "phi", 1)
    eqn = dsl.imp.ddt(T) + dsl.exp.div(phi, T)
    sch = {"divSchemes": {"div(phi,nfT)": "Gauss upwind"}}
    with pytest.raises(KeyError):
        dsl.solve(eqn, T, 0.0, 1e-3, dict(sch, ddtSchemes={"type": "crankNicolson"}), {})
    with pytest.raises(RuntimeError):   # sundials.hpp:59-78
        dsl.solve(eqn, T, 0.0, 1e-3, dict(sch, ddtSchemes={"type": "Runge-Kutta", "Runge-Kutta-Method": "Heun"}), {})
