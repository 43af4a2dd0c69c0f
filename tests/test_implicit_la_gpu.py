"""GPU parity tests: fused implicit assembly (ddt/div/laplacian/source -> CSR), CSR SpMV / residual, BLAS-1 and the
Jacobi-CG solver, all through the C ABI, against the CPU oracle.

Tolerances: assembly, SpMV and residual are asserted BIT-EXACT against the Serial oracle (same per-entry and per-row
accumulation order, no FMA contraction). Dot products are reduced in a tree on the GPU, so reductions and everything
downstream of them (CG) are compared with rtol 1e-13 per dot and, for the CG residual history, 1e-9 relative on every
entry of the curve until the residual has dropped below 1e-10 of its start (rounding differences are amplified by CG)."""
import numpy as np
import pytest
import torch

from foamadapter_b200 import fvcc, la, mesh as M, ops
from oracle.cpu import Mesh as OMesh, cg as oracle_cg
from tests.test_explicit_gpu import MESHES, dev, host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=sorted(MESHES))
def case(request):
    d = MESHES[request.param]()
    return request.param, d, M.UnstructuredMesh(d), OMesh.from_desc(d)


def _bd(om, rng, vec):
    shp = lambda n: (n, 3) if vec else (n,)
    return dict(value=rng.uniform(1, 2, shp(om.nB)), refValue=rng.uniform(1, 2, shp(om.nB)),
                valueFraction=rng.integers(0, 2, om.nB).astype(np.float64), refGrad=rng.uniform(-1, 1, shp(om.nB)))


class _BD:
    def __init__(self, bd):
        self.value, self.refValue, self.valueFraction, self.refGrad = (dev(bd[k]) for k in ("value", "refValue", "valueFraction", "refGrad"))


@pytest.mark.parametrize("vec", [False, True])
@pytest.mark.parametrize("scheme", [0, 1])
@pytest.mark.parametrize("with_view", [False, True])
def test_fused_assembly_matches_sequential_operators(case, vec, scheme, with_view):
    name, d, gm, om = case
    rng = np.random.default_rng(7)
    shp = lambda n: (n, 3) if vec else (n,)
    flux = rng.uniform(-1, 1, om.nF)
    gamma = rng.uniform(0.5, 1.5, om.nF)
    old = rng.uniform(1, 2, shp(om.nC))
    k = rng.uniform(0.1, 1, om.nC)
    view = rng.uniform(0.5, 1.5, om.nC) if with_view else None
    bd = _bd(om, rng, vec)
    dt = 0.25
    # oracle: ddt + div - laplacian + source applied one after another on a zeroed system (dsl order:
    # spatial operators first, then temporal -> div, laplacian, source, ddt)
    ls = om.empty_system(vec)
    om.div_imp(ls, flux, bd, scheme, 1.0, view)
    om.laplacian_imp(ls, gamma, bd, -1.0, view)
    om.source_imp(ls, k, 2.0, view)
    om.ddt_imp(ls, old, dt, 1.0, view)
    gls = la.LinearSystem(gm, 3 if vec else 1, zero=False)
    gls.values.fill_(float("nan")); gls.rhs.fill_(float("nan"))  # fused mode must not read them
    dview = dev(view) if with_view else None
    terms = [dict(kind=ops.TERM_DIV, scheme=scheme, coeff=1.0, coeffView=dview, faceField=dev(flux)),
             dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, coeffView=dview, faceField=dev(gamma)),
             dict(kind=ops.TERM_SOURCE, coeff=2.0, coeffView=dview, cellField=dev(k)),
             dict(kind=ops.TERM_DDT, coeff=1.0, coeffView=dview, cellField=dev(old), dt=dt)]
    ops.assemble(gm, terms, _BD(bd), gls.values, gls.rhs, gls.bcMatrix, gls.bcRhs)
    assert np.array_equal(host(gls.values), ls["values"])
    assert np.array_equal(host(gls.rhs), ls["rhs"])
    if om.nB:
        # the last div/laplacian term (laplacian) owns the boundary coefficients
        assert np.array_equal(host(gls.bcMatrix), ls["bcMatrix"])
        assert np.array_equal(host(gls.bcRhs), ls["bcRhs"])
    # accumulate mode: apply a second laplacian on top of the assembled system
    om.laplacian_imp(ls, gamma, bd, 0.5, None)
    ops.assemble(gm, [dict(kind=ops.TERM_LAPLACIAN, coeff=0.5, faceField=dev(gamma))], _BD(bd), gls.values, gls.rhs,
                 gls.bcMatrix, gls.bcRhs, accumulate=True)
    assert np.array_equal(host(gls.values), ls["values"])
    assert np.array_equal(host(gls.rhs), ls["rhs"])


def test_cellwise_explicit_terms(case):
    name, d, gm, om = case
    rng = np.random.default_rng(3)
    f, o, k, src = (rng.uniform(1, 2, om.nC) for _ in range(4))
    s = dev(src.copy())
    ops.ddt_explicit(gm, dev(f), dev(o), 0.5, s)
    exp = src + (1.0 / 0.5) * (f - o) * om.V
    assert np.allclose(host(s), exp, rtol=1e-15, atol=0)
    rhs = dev(src.copy())
    ops.rhs_sub_source(gm, dev(f), rhs)
    assert np.array_equal(host(rhs), src - f * om.V)
    mi = torch.zeros(om.nB, dtype=torch.int32, device="cuda"); ri = torch.zeros_like(mi)
    if om.nB:
        ops.bc_coeff_indices(gm, mi, ri)
        assert np.array_equal(host(ri), om.faceCells)
        assert np.array_equal(host(mi), om.faceCells + om.diagOffset[om.faceCells])


def _poisson(om, rng):
    """SPD system: -laplacian + small ddt term (diagonally dominant), as the PISO pressure equation."""
    ls = om.empty_system(False)
    bd = dict(value=np.zeros(om.nB), refValue=np.zeros(om.nB), valueFraction=np.ones(om.nB), refGrad=np.zeros(om.nB))
    om.laplacian_imp(ls, np.ones(om.nF), bd, -1.0, None)  # Dirichlet walls -> non-singular
    om.ddt_imp(ls, rng.uniform(1, 2, om.nC), 1.0, 1e-3 * float(np.abs(ls["values"]).max()) / float(om.V.max()), None)
    return ls


def test_spmv_and_residual_bit_exact(case):
    name, d, gm, om = case
    rng = np.random.default_rng(11)
    vals = rng.uniform(-1, 1, om.nnz)
    x, b = rng.uniform(-1, 1, om.nC), rng.uniform(-1, 1, om.nC)
    sp = la.SparsityPattern.readOrCreate(gm)
    assert np.array_equal(host(la.spmv(sp, dev(vals), dev(x))), om.spmv(vals, x))
    assert np.array_equal(host(la.computeResidual(sp, dev(vals), dev(b), dev(x))), om.residual(vals, b, x))


def test_structured_spmv_bit_exact(case):
    """fvk_spmv_structured: columns of the regular rows from arithmetic when the plan proved the block topology, the
    generic rows / meshes otherwise -- always the reference's row sums, bit for bit."""
    name, d, gm, om = case
    rng = np.random.default_rng(12)
    vals, x = rng.uniform(-1, 1, om.nnz), rng.uniform(-1, 1, om.nC)
    assert np.array_equal(host(la.spmv_structured(gm, dev(vals), dev(x))), om.spmv(vals, x))


def test_structured_spmv_on_a_large_block_and_its_sub_domains():
    from foamadapter_b200.decomp import Decomposition
    g = M.MeshDesc.block(40, 24, 18, 1.0, 0.6, 0.45)
    gm, om = M.UnstructuredMesh(g), OMesh.from_desc(g)
    rng = np.random.default_rng(13)
    vals, x = rng.uniform(-1, 1, om.nnz), rng.uniform(-1, 1, om.nC)
    assert np.array_equal(host(la.spmv_structured(gm, dev(vals), dev(x))), om.spmv(vals, x))
    for r in range(2):
        dec = Decomposition(g, 2, r)
        lm = M.UnstructuredMesh(dec.desc)
        sp = la.SparsityPattern.readOrCreate(lm)
        lv, lx = rng.uniform(-1, 1, lm.nnz), rng.uniform(-1, 1, lm.nCells)
        assert torch.equal(la.spmv_structured(lm, dev(lv), dev(lx)), la.spmv(sp, dev(lv), dev(lx)))


def test_residual_known_answer():
    # src/NeoN/test/linearAlgebra/utilities.cpp:22-44
    ro = torch.tensor([0, 3, 6, 9], dtype=torch.int32, device="cuda")
    ci = torch.tensor([0, 1, 2] * 3, dtype=torch.int32, device="cuda")
    v = torch.arange(1, 10, dtype=torch.float64, device="cuda")
    x = torch.ones(3, dtype=torch.float64, device="cuda"); b = torch.full((3,), 2.0, dtype=torch.float64, device="cuda")
    res = torch.empty(3, dtype=torch.float64, device="cuda")
    import ctypes as C
    from foamadapter_b200._capi import check, lib, ptr
    check(lib().fvk_residual(C.c_int32(3), ptr(ro), ptr(ci), ptr(v), ptr(b), ptr(x), ptr(res), None))
    torch.cuda.synchronize()
    assert host(res).tolist() == [4.0, 13.0, 22.0]


def test_blas1():
    rng = np.random.default_rng(5)
    for n in (1, 31, 257, 100003):
        x, y = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        dx, dy = dev(x.copy()), dev(y)
        assert np.isclose(host(la.dot(dx, dy))[0], np.dot(x, y), rtol=1e-13, atol=1e-13)
        assert np.isclose(host(la.norm2(dx))[0], np.linalg.norm(x), rtol=1e-13)
        # determinism: bit-identical on repeat
        assert host(la.dot(dx, dy))[0] == host(la.dot(dx, dy))[0]
        assert np.array_equal(host(la.axpby(2.0, dx, -0.5, dev(y.copy()))), 2.0 * x + -0.5 * y)
        assert np.array_equal(host(la.add(dev(x.copy()), dy)), x + y)
        assert np.array_equal(host(la.sub(dev(x.copy()), dy)), x - y)
        assert np.array_equal(host(la.mul(dev(x.copy()), dy)), x * y)
        assert np.array_equal(host(la.scalarMul(dev(x.copy()), 3.0)), x * 3.0)
        assert np.array_equal(host(la.fill(dev(x.copy()), 1.5)), np.full(n, 1.5))


def test_cg_known_answer():
    # src/NeoN/test/linearAlgebra/ginkgo.cpp:95-124: tridiag [1, -0.1], b = (1,2,3), x0 = 0, 3 iterations
    ro = torch.tensor([0, 2, 5, 7], dtype=torch.int32, device="cuda")
    ci = torch.tensor([0, 1, 0, 1, 2, 1, 2], dtype=torch.int32, device="cuda")
    v = torch.tensor([1.0, -0.1, -0.1, 1.0, -0.1, -0.1, 1.0], dtype=torch.float64, device="cuda")
    b = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64, device="cuda")
    x = torch.zeros(3, dtype=torch.float64, device="cuda")
    s = la.Solver({"solver": "Ginkgo", "type": "solver::Cg", "criteria": {"iteration": 3, "relative_residual_norm": 1e-7}}, history=True)
    st = s.solve_csr(3, 3, ro.data_ptr(), ci.data_ptr(), v, b, x)
    assert st.numIter == 3
    assert abs(st.initResNorm - 3.741657386) < 1e-8
    assert st.finalResNorm < 1e-4
    assert np.allclose(host(x), [1.24489796, 2.44897959, 3.24489796], atol=1e-8)


@pytest.mark.parametrize("jacobi", [True, False])
@pytest.mark.parametrize("check_every", [1, 7])
def test_cg_tracks_oracle_history(case, jacobi, check_every):
    name, d, gm, om = case
    rng = np.random.default_rng(13)
    ls = _poisson(om, rng)
    b = rng.uniform(-1, 1, om.nC)
    x0 = rng.uniform(-1, 1, om.nC)
    abs_tol = 1e-9 * np.linalg.norm(b)
    xo, so, ho = oracle_cg(om.rowOffs, om.colIdxs, ls["values"], b, x0, jacobi=jacobi, max_iter=200, rel_tol=0.0, abs_tol=abs_tol, max_hist=300)
    cfg = {"solver": "Ginkgo", "type": "solver::Cg", "criteria": {"iteration": 200, "relative_residual_norm": 0.0, "absolute_residual_norm": abs_tol}}
    if jacobi:
        cfg["preconditioner"] = {"type": "preconditioner::Jacobi", "max_block_size": 1}
    gls = la.LinearSystem(gm)
    gls.values.copy_(dev(ls["values"])); gls.rhs.copy_(dev(b))
    x = dev(x0.copy())
    st = la.Solver(cfg, check_every=check_every, history=True).solve(gls, x)
    assert abs(st.initResNorm - so["initResNorm"]) <= 1e-13 * so["initResNorm"]
    assert abs(st.numIter - so["numIter"]) <= 1
    n = min(len(st.history), len(ho))
    sig = ho[:n] > 1e-10 * ho[0]
    assert np.allclose(st.history[:n][sig], ho[:n][sig], rtol=1e-9, atol=0)
    assert st.finalResNorm <= abs_tol or st.numIter == 200
    assert np.allclose(host(x), xo, rtol=1e-7, atol=1e-9 * np.abs(xo).max())
    assert len(st.history) == st.numIter + 1


def test_cg_is_run_to_run_deterministic():
    d = M.MeshDesc.block(24, 16, 12)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    rng = np.random.default_rng(1)
    ls = _poisson(om, rng)
    b = rng.uniform(-1, 1, om.nC)
    cfg = {"solver": "PCG", "preconditioner": "DIC", "tolerance": 1e-10, "relTol": 0.0}
    outs = []
    for _ in range(2):
        gls = la.LinearSystem(gm); gls.values.copy_(dev(ls["values"])); gls.rhs.copy_(dev(b))
        x = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
        st = la.Solver(cfg, history=True).solve(gls, x)
        outs.append((host(x).copy(), st.history.copy(), st.numIter))
    assert outs[0][2] == outs[1][2] and np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    r = om.residual(ls["values"], b, outs[0][0])
    assert np.linalg.norm(r) <= 1.01e-10


def test_solver_rejects_unknown_configuration():
    with pytest.raises(KeyError):
        la.Solver({"solver": "GAMG"})
    with pytest.raises(KeyError):
        la.Solver({"solver": "Ginkgo", "type": "solver::Bicg"})
    with pytest.raises(KeyError):  # DILU -> preconditioner::Ilu (fvSolution.cpp:56-62) is not on the hot path: no silent downgrade
        la.Solver({"solver": "PBiCGStab", "preconditioner": "DILU", "tolerance": 1e-6})


# ---- extension: multicolour DIC preconditioner -------------------------------------------------------
@pytest.mark.parametrize("dims", [(24, 16, 12), (17, 9, 5)])
def test_multicolour_dic_cg_tracks_the_oracle_and_beats_jacobi(dims):
    from oracle.cpu import cg_dic
    d = M.MeshDesc.block(*dims)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    rng = np.random.default_rng(2)
    ls = _poisson(om, rng)
    b = rng.uniform(-1, 1, om.nC)
    tol = 1e-9 * np.linalg.norm(b)
    xo, so, ho, colors = cg_dic(om.rowOffs, om.colIdxs, ls["values"], b, np.zeros(om.nC), max_iter=500, rel_tol=0.0, abs_tol=tol, max_hist=600)
    assert colors.max() == 1                                        # a hex block is two-colourable
    _, sj, _ = oracle_cg(om.rowOffs, om.colIdxs, ls["values"], b, np.zeros(om.nC), jacobi=True, max_iter=500, rel_tol=0.0, abs_tol=tol)
    assert so["numIter"] < 0.8 * sj["numIter"]                      # what the preconditioner is for
    gls = la.LinearSystem(gm); gls.values.copy_(dev(ls["values"])); gls.rhs.copy_(dev(b))
    x = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
    cfg = {"solver": "Ginkgo", "type": "solver::Cg", "preconditioner": {"type": "preconditioner::Ic"},
           "criteria": {"iteration": 500, "relative_residual_norm": 0.0, "absolute_residual_norm": tol}}
    st = la.Solver(cfg, check_every=3, history=True).solve(gls, x)
    assert abs(st.numIter - so["numIter"]) <= 1
    n = min(len(st.history), len(ho))
    sig = ho[:n] > 1e-8 * ho[0]
    assert np.allclose(st.history[:n][sig], ho[:n][sig], rtol=1e-7)
    assert np.allclose(host(x), xo, rtol=1e-7, atol=1e-9 * np.abs(xo).max())
    r = om.residual(ls["values"], b, host(x))
    assert np.linalg.norm(r) <= 1.01 * tol


def test_dic_needs_the_mesh_pattern_and_one_gpu():
    d = M.MeshDesc.block(6, 5, 4)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    cfg = {"solver": "Ginkgo", "type": "solver::Cg", "preconditioner": {"type": "preconditioner::Ic"}, "criteria": {"iteration": 10}}
    s = la.Solver(cfg)
    ls = _poisson(om, np.random.default_rng(0))
    ro, ci = dev(om.rowOffs.copy()), dev(om.colIdxs.copy())      # a foreign copy of the pattern: no colouring for it
    x = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
    from foamadapter_b200._capi import FvkError
    with pytest.raises(FvkError):
        s.solve_csr(om.nC, om.nC, ro.data_ptr(), ci.data_ptr(), dev(ls["values"]), dev(ls["rhs"] + 1.0), x)
