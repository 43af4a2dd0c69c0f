"""GPU parity tests for the PISO glue kernels and the neoIcoFoam step (all through the C ABI) against the CPU oracle.

Tolerances: rAU/HbyA/flux/updateFaceVelocity/updateVelocity/setReference/diag are asserted BIT-EXACT (same arithmetic
order as the Serial oracle). A whole PISO step contains CG solves whose dot products are reduced in a different order
on the GPU; U and p are compared at 1e-8 / 1e-7 relative to their max norm, per-solve iteration counts within +-1 and
residual histories at 1e-6 relative while the residual is above 1e-8 of its start."""
import numpy as np
import pytest
import torch

from foamadapter_b200 import dsl, fvcc, la, mesh as M, ops, piso
from oracle.cpu import Mesh as OMesh
from oracle.piso import IcoFoamOracle
from tests.test_explicit_gpu import dev, host

pytestmark = pytest.mark.gpu

CASES = {"cavity2d_20": lambda: piso.cavity_desc(20), "cavity3d_8": lambda: piso.cavity_desc(8, True),
         "block_7x5x4": lambda: M.MeshDesc.block(7, 5, 4, 0.7, 0.3, 0.9)}


@pytest.fixture(scope="module", params=sorted(CASES))
def case(request):
    d = CASES[request.param]()
    return request.param, d, M.UnstructuredMesh(d), OMesh.from_desc(d)


def test_glue_kernels_bit_exact(case):
    name, d, gm, om = case
    rng = np.random.default_rng(21)
    vals1 = rng.uniform(0.5, 1.5, om.nnz)
    valsV = np.repeat(vals1[:, None], 3, axis=1).copy()
    rhsV, U, Ub = rng.uniform(-1, 1, (om.nC, 3)), rng.uniform(-1, 1, (om.nC, 3)), rng.uniform(-1, 1, (om.nB, 3))
    rAU_o = om.rAU(valsV)
    H_o = om.HbyA(valsV, rhsV, rAU_o, U)
    rAU = torch.empty(om.nC, dtype=torch.float64, device="cuda"); H = torch.empty((om.nC, 3), dtype=torch.float64, device="cuda")
    ops.rAU_HbyA(gm, dev(valsV), dev(rhsV), dev(U), rAU, H)
    assert np.array_equal(host(rAU), rAU_o) and np.array_equal(host(H), H_o)
    # flux
    ff_o, bv_o = om.flux(U, Ub)
    ff, bv = torch.empty(om.nF, dtype=torch.float64, device="cuda"), torch.empty(om.nB, dtype=torch.float64, device="cuda")
    ops.flux(gm, dev(U), dev(Ub), ff, bv)
    assert np.array_equal(host(ff), ff_o) and np.array_equal(host(bv), bv_o)
    # updateFaceVelocity
    ls = dict(values=vals1, bcMatrix=rng.uniform(-1, 1, om.nB), bcRhs=rng.uniform(-1, 1, om.nB))
    p, pred, predB = rng.uniform(-1, 1, om.nC), rng.uniform(-1, 1, om.nF), rng.uniform(-1, 1, om.nB)
    phi_o, phiB_o = om.update_face_velocity(ls, p, pred, predB)
    phi, phiB = torch.empty(om.nF, dtype=torch.float64, device="cuda"), torch.empty(om.nB, dtype=torch.float64, device="cuda")
    ops.update_face_velocity(gm, dev(vals1), dev(ls["bcMatrix"]), dev(ls["bcRhs"]), dev(p), dev(pred), dev(predB), phi, phiB)
    assert np.array_equal(host(phi), phi_o) and np.array_equal(host(phiB), phiB_o)
    # updateVelocity
    g = rng.uniform(-1, 1, (om.nC, 3))
    Uo = om.update_velocity(H_o, rAU_o, g)
    Un = torch.empty((om.nC, 3), dtype=torch.float64, device="cuda")
    ops.update_velocity(gm, dev(H_o), dev(rAU_o), dev(g), Un)
    assert np.array_equal(host(Un), Uo)
    # setReference + diag
    lso = dict(values=vals1.copy(), rhs=rng.uniform(-1, 1, om.nC))
    v, r = dev(lso["values"].copy()), dev(lso["rhs"].copy())
    om.set_reference(lso, om.nC // 2, 0.75)
    ops.set_reference(gm, om.nC // 2, 0.75, v, r)
    assert np.array_equal(host(v), lso["values"]) and np.array_equal(host(r), lso["rhs"])
    dg = torch.empty(om.nC, dtype=torch.float64, device="cuda")
    ops.diag(gm, v, dg)
    assert np.array_equal(host(dg), lso["values"][om.rowOffs[:-1] + om.diagOffset])
    # constrainHbyA
    if om.nB:
        mask = [i % 2 == 0 for i in range(gm.nPatches)]
        src, dst = rng.uniform(-1, 1, (om.nB, 3)), rng.uniform(-1, 1, (om.nB, 3))
        exp = dst.copy()
        for i, on in enumerate(mask):
            if on:
                exp[om.patchOffsets[i]:om.patchOffsets[i + 1]] = src[om.patchOffsets[i]:om.patchOffsets[i + 1]]
        t = dev(dst.copy())
        ops.copy_patches(gm, mask, dev(src), t)
        assert np.array_equal(host(t), exp)


@pytest.mark.parametrize("which", ["cavity2d_20", "cavity3d_8"])
def test_icofoam_steps_track_oracle(which):
    d = CASES[which]()
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    dt = 5e-4
    g = piso.IcoFoam(gm, nu=0.01, dt=dt, history=True, check_every=4)
    o = IcoFoamOracle(om, nu=0.01, dt=dt)
    assert np.array_equal(host(g.phi.internal), o.phi)
    for step in range(4):  # steps 0-1 eager, 2 captured + replayed, 3 replayed (CUDA-graph segments)
        gs = g.step()
        os_ = o.step()
        for (st, (so, ho)) in zip(gs, os_):
            assert abs(st.numIter - so["numIter"]) <= 1, (step, st.numIter, so["numIter"])
            assert abs(st.initResNorm - so["initResNorm"]) <= 1e-9 * max(so["initResNorm"], 1e-300)
            n = min(len(st.history), len(ho))
            sig = ho[:n] > 1e-8 * ho[0]
            assert np.allclose(st.history[:n][sig], ho[:n][sig], rtol=1e-6, atol=0)
            assert st.finalResNorm <= 1e-6
        U, p, phi = host(g.U.internal), host(g.p.internal), host(g.phi.internal)
        assert np.abs(U - o.U).max() <= 1e-8 * np.abs(o.U).max()
        assert np.abs(p - o.p).max() <= 1e-7 * max(np.abs(o.p).max(), 1e-30)
        assert np.abs(phi - o.phi).max() <= 1e-8 * np.abs(o.phi).max()
    # continuity: the corrected face flux is discretely divergence-free to solver tolerance
    div = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
    ops.surface_integrate(gm, g.phi.internal, div)
    assert float((div * dev(om.V)).abs().max()) < 1e-5
    co = host(g.coNum)
    assert np.isfinite(co).all() and co[0] > 0


def test_cuda_graph_segments_equal_eager_steps():
    """From the third step on IcoFoam replays the kernel-only segments between the linear solves as CUDA graphs: the
    fields must be bit-identical to the eager time loop, step after step."""
    d = piso.cavity_desc(12, True)
    a = piso.IcoFoam(M.UnstructuredMesh(d), nu=0.01, dt=5e-4, check_every=4, graphs=True, whole_step_graph=False)
    b = piso.IcoFoam(M.UnstructuredMesh(d), nu=0.01, dt=5e-4, check_every=4, graphs=False)
    for step in range(6):
        sa, sb = a.step(), b.step()
        assert [s.numIter for s in sa] == [s.numIter for s in sb]
        for x, y in ((a.U.internal, b.U.internal), (a.p.internal, b.p.internal), (a.phi.internal, b.phi.internal)):
            assert torch.equal(x, y), step
    assert a._captured is not None and len(a._captured) == 3 and b._captured is None


def test_whole_step_graph_with_captured_solves_equals_eager_steps():
    """Default mode: from the third step on a time step is ONE CUDA graph -- the pressure solves are conditional WHILE nodes
    whose stopping test runs on the device (no host round trip). Fields and iteration counts must equal the eager loop's, bit
    for bit, step after step (the graph runs the same kernels in the same order; iterations past the stop are no-ops)."""
    d = piso.cavity_desc(14, True)
    a = piso.IcoFoam(M.UnstructuredMesh(d), nu=0.01, dt=5e-4, check_every=4, graphs=True)
    b = piso.IcoFoam(M.UnstructuredMesh(d), nu=0.01, dt=5e-4, check_every=4, graphs=False)
    iters = []
    for step in range(8):
        sa, sb = a.step(), b.step()
        ia, ib = [s.numIter for s in sa], [s.numIter for s in sb]
        assert ia == ib, (step, ia, ib)
        iters.append(ia)
        for x, y in ((a.U.internal, b.U.internal), (a.p.internal, b.p.internal), (a.phi.internal, b.phi.internal)):
            assert torch.equal(x, y), step
        if step >= 2:
            assert abs(sa[0].finalResNorm - sb[0].finalResNorm) <= 1e-15 + 1e-12 * sb[0].finalResNorm
    assert a._whole is not None and a._captured is None
    # the solver's device-side log holds the iteration counts of the replayed solves in execution order (read without having
    # synchronised after every step); the first two steps run eagerly and are not in it
    log = a.solver.captured_log()
    flat = [n for i in iters[2:] for n in i]
    assert log == flat, (log, flat)
    assert a.solver.captured_log() == []   # reading resets it
    assert len({tuple(i) for i in iters[2:]}) > 1 and max(max(i) for i in iters[2:]) >= 3   # the loop really iterates, differently per step


def test_pdesolver_matches_dsl_solve_sequence(case):
    """PDESolver.solve == implicit assembly, rhs -= explicit*V, setReference, CG (dsl/solver.hpp:60-80)."""
    name, d, gm, om = case
    rng = np.random.default_rng(4)
    nP = gm.nPatches
    p = fvcc.VolumeField(gm, "p", 1, [("fixedValue", 1.0)] + [("zeroGradient", 0.0)] * (nP - 1))
    p.correctBoundaryConditions()
    gamma = fvcc.SurfaceField(gm, "rAUf", 1); gamma.internal.copy_(dev(rng.uniform(0.5, 1.5, om.nF)))
    fl = fvcc.SurfaceField(gm, "phiHbyA", 1); fl.internal.copy_(dev(rng.uniform(-1, 1, om.nF) * om.magSf))
    rt = dsl.RunTime(gm, 1.0, 0.0, piso.CAVITY_FVSCHEMES, {"solvers": {"p": {"solver": "PCG", "preconditioner": "DIC", "tolerance": 1e-10, "relTol": 0}}}, history=True)
    eq = dsl.PDESolver(dsl.imp.laplacian(gamma, p) - dsl.exp.div(fl), p, rt)
    st = eq.solve()
    # oracle
    pbd = om.correct_bcs([1] + [2] * (nP - 1), [1.0] + [0.0] * (nP - 1), np.zeros(om.nC))
    ls = om.empty_system(False)
    om.laplacian_imp(ls, host(gamma.internal), pbd, 1.0, None)
    om.rhs_sub_source(ls, om.surface_integrate(host(fl.internal), coeff=-1.0))
    assert np.array_equal(host(eq.ls.values), ls["values"]) and np.array_equal(host(eq.ls.rhs), ls["rhs"])
    xo, so, ho = om.cg(ls["values"], ls["rhs"], np.zeros(om.nC), jacobi=True, max_iter=1000, rel_tol=0.0, abs_tol=1e-10, max_hist=1002)
    assert abs(st.numIter - so["numIter"]) <= 1
    assert np.abs(host(p.internal) - xo).max() <= 1e-7 * max(np.abs(xo).max(), 1e-30)


# ---- compact Vec3 momentum system (values stored once) ------------------------------------------------------------------
@pytest.mark.parametrize("dims", [(7, 5, 4), (40, 11, 7), (33, 9, 3)])
def test_compact_momentum_system_equals_the_vec3_layout(dims):
    """fvk_assemble_vc / fvk_rAU_HbyA_c / valuesVec3 against the reference layout (fvk_assemble_v / fvk_rAU_HbyA) and the
    oracle, bit for bit; both kernel families (index-free and stencil-driven) and the old face-walk kernel (variant 5)."""
    import ctypes as C
    from foamadapter_b200._capi import lib
    d = M.MeshDesc.block(*dims, 0.7, 0.3, 0.9)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    rng = np.random.default_rng(5)
    U = fvcc.VolumeField(gm, "U", 3, [("fixedValue", (1.0, 0.0, 0.0)), ("noSlip", 0.0), ("zeroGradient", 0.0)])
    U_h = rng.uniform(-1, 1, (om.nC, 3))
    U.internal.copy_(dev(U_h)); U.correctBoundaryConditions()
    flux, nu, old = rng.uniform(-1, 1, om.nF), rng.uniform(0.5, 1.5, om.nF), rng.uniform(-1, 1, (om.nC, 3))
    terms = [dict(kind=ops.TERM_DIV, scheme=0, coeff=1.0, faceField=dev(flux)), dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=dev(nu)),
             dict(kind=ops.TERM_DDT, coeff=1.0, cellField=dev(old), dt=0.01)]
    ref = la.LinearSystem(gm, 3, zero=False)
    ops.assemble(gm, terms, U.boundary, ref.values, ref.rhs, ref.bcMatrix, ref.bcRhs)
    rAU0, H0 = torch.empty(om.nC, dtype=torch.float64, device="cuda"), torch.empty((om.nC, 3), dtype=torch.float64, device="cuda")
    lib().fvk_set_variant(C.c_int(5))          # the face-walk kernel (any mesh)
    ops.rAU_HbyA(gm, ref.values, ref.rhs, U.internal, rAU0, H0)
    lib().fvk_set_variant(C.c_int(0))
    Ho = om.HbyA(host(ref.values), host(ref.rhs), om.rAU(host(ref.values)), U_h)
    assert np.array_equal(host(H0), Ho)
    for affine in (1, 0):
        lib().fvk_set_affine(C.c_int(affine))
        cls = la.LinearSystem(gm, 3, zero=False, compact=True)
        assert cls.values.shape == (gm.nnz,)
        cls.values.fill_(float("nan"))
        ops.assemble(gm, terms, U.boundary, cls.values, cls.rhs, cls.bcMatrix, cls.bcRhs)
        assert torch.equal(cls.valuesVec3(), ref.values) and torch.equal(cls.rhs, ref.rhs)
        assert torch.equal(cls.bcMatrix, ref.bcMatrix) and torch.equal(cls.bcRhs, ref.bcRhs)
        for vals in (cls.values, ref.values):   # row-walk kernel on both layouts
            rAU, H = torch.empty_like(rAU0), torch.empty_like(H0)
            ops.rAU_HbyA(gm, vals, ref.rhs, U.internal, rAU, H)
            assert torch.equal(rAU, rAU0) and torch.equal(H, H0)
    lib().fvk_set_affine(C.c_int(1))


def test_piso_step_is_identical_with_and_without_the_compact_momentum_matrix():
    d = piso.cavity_desc(12, True)
    a = piso.IcoFoam(M.UnstructuredMesh(d), nu=0.01, dt=5e-4, compact_momentum=True, graphs=False)
    b = piso.IcoFoam(M.UnstructuredMesh(d), nu=0.01, dt=5e-4, compact_momentum=False, graphs=False)
    assert a.Uls.compact and not b.Uls.compact
    for _ in range(3):
        a.step(); b.step()
    assert torch.equal(a.U.internal, b.U.internal) and torch.equal(a.p.internal, b.p.internal) and torch.equal(a.phi.internal, b.phi.internal)


def test_momentum_predictor_step_matches_the_oracle():
    """momentumPredictor yes (neoIcoFoam.cpp:100-103): UEqn solved with BiCGStab + scalar Jacobi, component
    by component, before the PISO correctors."""
    import copy
    d = piso.cavity_desc(10, True)
    sol = copy.deepcopy(piso.CAVITY_FVSOLUTION)
    sol["PISO"]["momentumPredictor"] = True
    app = piso.IcoFoam(M.UnstructuredMesh(d), nu=0.01, dt=5e-3, fvSolution=sol, graphs=True)
    ref = IcoFoamOracle(OMesh.from_desc(d), nu=0.01, dt=5e-3, momentumPredictor=True)
    for step in range(4):   # steps 3 and 4 replay the CUDA-graph segments around the three linear solves
        app.step(); ref.step()
        assert [abs(a.numIter - b["numIter"]) <= 1 for a, b in zip(app.Ustats, ref.Ustats)] == [True] * 3
    assert any(s["numIter"] > 0 for s in ref.Ustats)
    assert np.abs(host(app.U.internal) - ref.U).max() <= 1e-7 * np.abs(ref.U).max()
    assert np.abs(host(app.p.internal) - ref.p).max() <= 1e-6 * np.abs(ref.p).max()


@pytest.mark.parametrize("dims", [(7, 5, 4), (40, 11, 7)])
def test_fused_epilogues_equal_the_two_kernel_forms(dims):
    """fvk_update_velocity_grad == fvk_grad_s + fvk_update_velocity; fvk_rhs_sub_surface_integrate_s == surfaceIntegrate into a
    zeroed source + rhs -= source * V; both on the index-free and the stencil-driven kernels."""
    import ctypes as C
    from foamadapter_b200._capi import lib
    d = M.MeshDesc.block(*dims, 0.7, 0.3, 0.9)
    gm = M.UnstructuredMesh(d)
    rng = np.random.default_rng(9)
    nC, nB, nF = gm.nCells, gm.nBoundaryFaces, gm.nFaces
    p, pB = dev(rng.uniform(-1, 1, nC)), dev(rng.uniform(-1, 1, nB))
    H, rAU = dev(rng.uniform(-1, 1, (nC, 3))), dev(rng.uniform(0.5, 2, nC))
    flux, rhs0, view = dev(rng.uniform(-1, 1, nF)), dev(rng.uniform(-1, 1, nC)), dev(rng.uniform(0.5, 2, nC))
    g = torch.empty((nC, 3), dtype=torch.float64, device="cuda")
    ops.grad(gm, p, pB, g)
    U_ref = torch.empty_like(g); ops.update_velocity(gm, H, rAU, g, U_ref)
    src = torch.zeros(nC, dtype=torch.float64, device="cuda")
    ops.surface_integrate(gm, flux, src, -1.0, view, ops.ADD)
    rhs_ref = rhs0.clone(); ops.rhs_sub_source(gm, src, rhs_ref)
    for affine in (1, 0):
        lib().fvk_set_affine(C.c_int(affine))
        U = torch.full_like(g, float("nan"))
        ops.update_velocity_grad(gm, H, rAU, p, pB, U)
        assert torch.equal(U, U_ref)
        rhs = rhs0.clone()
        ops.rhs_sub_surface_integrate(gm, flux, rhs, -1.0, view)
        assert torch.equal(rhs, rhs_ref)
    lib().fvk_set_affine(C.c_int(1))


def test_piso_step_is_identical_with_gamma_interpolated_inside_the_assembly():
    """laplacian(interpolate(rAU), p): the assembly kernels (index-free, stencil-driven, generic) evaluate the linear
    interpolate on the fly -- bit-identical to the materialised rAUf."""
    import ctypes as C
    from foamadapter_b200._capi import lib
    d = piso.cavity_desc(12, True)
    gm = M.UnstructuredMesh(d)
    rng = np.random.default_rng(2)
    rAU = fvcc.VolumeField(gm, "rAU", 1, [("extrapolated", 0.0)] * gm.nPatches)
    rAU.internal.copy_(dev(rng.uniform(0.5, 1.5, gm.nCells))); rAU.correctBoundaryConditions()
    p = fvcc.VolumeField(gm, "p", 1, [("fixedValue", 1.0)] + [("zeroGradient", 0.0)] * (gm.nPatches - 1)); p.correctBoundaryConditions()
    rAUf = fvcc.SurfaceInterpolation(gm, "linear").interpolate(rAU)
    ref = la.LinearSystem(gm, 1, zero=False)
    ops.assemble(gm, [dict(kind=ops.TERM_LAPLACIAN, coeff=1.0, faceField=rAUf.internal)], p.boundary, ref.values, ref.rhs, ref.bcMatrix, ref.bcRhs)
    lazy = [dict(kind=ops.TERM_LAPLACIAN, coeff=1.0, gammaCell=rAU.internal, gammaBoundary=rAU.boundary.value)]
    for affine in (1, 0):
        lib().fvk_set_affine(C.c_int(affine))
        ls = la.LinearSystem(gm, 1, zero=False); ls.values.fill_(float("nan"))
        ops.assemble(gm, lazy, p.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
        assert torch.equal(ls.values, ref.values) and torch.equal(ls.rhs, ref.rhs) and torch.equal(ls.bcMatrix, ref.bcMatrix)
    lib().fvk_set_affine(C.c_int(1))
    ls = la.LinearSystem(gm, 1, zero=True)     # accumulate = generic kernel
    ops.assemble(gm, lazy, p.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs, accumulate=True)
    assert torch.equal(ls.values, ref.values)
    a = piso.IcoFoam(M.UnstructuredMesh(d), nu=0.01, dt=5e-4, fuse_interpolate=True, graphs=False)
    b = piso.IcoFoam(M.UnstructuredMesh(d), nu=0.01, dt=5e-4, fuse_interpolate=False, graphs=False)
    for _ in range(3):
        a.step(); b.step()
    assert torch.equal(a.U.internal, b.U.internal) and torch.equal(a.p.internal, b.p.internal) and torch.equal(a.phi.internal, b.phi.internal)
