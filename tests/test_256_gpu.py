"""GPU vs oracle at 256^3 -- the size every north-star target is stated on (16.8 M cells, 50.1 M internal faces, 117 M
non-zeros): explicit operators, fused assembly, SpMV (generic and structured), a Jacobi-CG solve, and a 128^3 corner
sub-domain of the 2x2x2 decomposition (processor cuts on three sides). Bit-exact where the small-mesh tests are."""
import ctypes as C

import numpy as np
import pytest
import torch

from foamadapter_b200 import fvcc, la, mesh as M, ops
from foamadapter_b200._capi import lib
from foamadapter_b200.decomp import Decomposition
from oracle.cpu import Mesh as OMesh, cg as oracle_cg

pytestmark = pytest.mark.gpu
N = 256
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
host = lambda t: t.detach().cpu().numpy()
BCS = [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)]


@pytest.fixture(scope="module")
def big():
    d = M.MeshDesc.block(N, N, N, 0.1, 0.1, 0.01)
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    rng = np.random.Generator(np.random.MT19937(42))
    T_h = rng.uniform(1.0, 2.0, om.nC)
    flux_h = np.concatenate([np.arange(om.nI, dtype=np.float64), np.zeros(om.nB)])
    T = fvcc.VolumeField(gm, "T", 1, BCS)
    T.internal.copy_(dev(T_h)); T.correctBoundaryConditions()
    bd = om.correct_bcs([1, 1, 2], [10.5, 1.5, 0.0], T_h)
    yield dict(d=d, gm=gm, om=om, T=T, T_h=T_h, flux_h=flux_h, flux=dev(flux_h), bd=bd)


def test_sizes(big):
    om = big["om"]
    assert (om.nC, om.nI, om.nB, om.nnz) == (16777216, 50135040, 393216, 117047296)   # all < 2^31


def test_explicit_operators_bit_exact(big):
    gm, om, T, flux = big["gm"], big["om"], big["T"], big["flux"]
    phib = big["bd"]["value"]
    assert np.array_equal(host(T.boundary.value), phib)
    out = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
    ops.div(gm, flux, T.internal, T.boundary.value, out)
    assert np.array_equal(host(out), om.div(big["flux_h"], big["T_h"], phib, 0, par=0))
    ops.div(gm, flux, T.internal, T.boundary.value, out, scheme=ops.UPWIND)
    assert np.array_equal(host(out), om.div(big["flux_h"], big["T_h"], phib, 1, par=0))
    ops.laplacian(gm, T.internal, T.boundary.value, out)
    assert np.array_equal(host(out), om.laplacian(big["T_h"], phib))
    del out
    g = torch.empty((om.nC, 3), dtype=torch.float64, device="cuda")
    ops.grad(gm, T.internal, T.boundary.value, g)
    assert np.array_equal(host(g), om.grad(big["T_h"], phib))


def test_assembly_spmv_cg(big):
    gm, om, T, flux = big["gm"], big["om"], big["T"], big["flux"]
    gamma_h, old_h = np.ones(om.nF), big["T_h"] - 1.0
    ols = om.empty_system(False)
    om.div_imp(ols, big["flux_h"], big["bd"], 0, 1.0, None)
    om.laplacian_imp(ols, gamma_h, big["bd"], -1.0, None)
    om.ddt_imp(ols, old_h, 1.0, 1.0, None)
    ls = la.LinearSystem(gm, 1, zero=False)
    terms = [dict(kind=ops.TERM_DIV, scheme=0, coeff=1.0, faceField=flux), dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=dev(gamma_h)),
             dict(kind=ops.TERM_DDT, coeff=1.0, cellField=dev(old_h), dt=1.0)]
    for affine in (1, 0):   # the index-free kernel and the stencil-driven one
        lib().fvk_set_affine(C.c_int(affine))
        ls.values.fill_(float("nan")); ls.rhs.fill_(float("nan"))
        ops.assemble(gm, terms, T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
        assert np.array_equal(host(ls.values), ols["values"]) and np.array_equal(host(ls.rhs), ols["rhs"]), f"assembly (affine={affine})"
    lib().fvk_set_affine(C.c_int(1))
    assert np.array_equal(host(ls.bcMatrix), ols["bcMatrix"]) and np.array_equal(host(ls.bcRhs), ols["bcRhs"])
    # SpMV over the assembled matrix: generic tiled CSR and the structured variant (columns from arithmetic, int32 index math)
    rng = np.random.default_rng(1)
    x_h = rng.uniform(-1, 1, om.nC)
    y_ref = om.spmv(ols["values"], x_h, par=1)   # row sums are per-row: the parallel row loop is bit-identical to Serial
    x, y = dev(x_h), torch.empty(om.nC, dtype=torch.float64, device="cuda")
    la.spmv(la.SparsityPattern.readOrCreate(gm), ls.values, x, y)
    assert np.array_equal(host(y), y_ref)
    y.zero_()
    la.spmv_structured(gm, ls.values, x, y)
    assert np.array_equal(host(y), y_ref)
    # Jacobi-CG on the SPD part (-laplacian + ddt), 12 iterations, residual history against the oracle
    del ols
    pls = om.empty_system(False)
    om.laplacian_imp(pls, gamma_h, big["bd"], -1.0, None)
    om.ddt_imp(pls, old_h, 1.0, 1.0, None)
    ops.assemble(gm, terms[1:], T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
    assert np.array_equal(host(ls.values), pls["values"])
    xo, so, ho = oracle_cg(om.rowOffs, om.colIdxs, pls["values"], pls["rhs"], np.zeros(om.nC), jacobi=True, max_iter=12, rel_tol=0.0, abs_tol=0.0,
                           par=1, max_hist=20)
    xs = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
    cfg = {"solver": "Ginkgo", "type": "solver::Cg", "preconditioner": {"type": "preconditioner::Jacobi", "max_block_size": 1},
           "criteria": {"iteration": 12, "relative_residual_norm": 0.0, "absolute_residual_norm": 0.0}}
    st = la.Solver(cfg, history=True).solve(ls, xs)
    assert st.numIter == so["numIter"] == 12
    assert np.allclose(st.history, ho, rtol=1e-9, atol=0)
    assert np.allclose(host(xs), xo, rtol=1e-9, atol=1e-12 * np.abs(xo).max())


@pytest.mark.parametrize("rank", [0, 7])
def test_corner_subdomains_of_the_2x2x2_decomposition(big, rank):
    """128^3 owned cells + ghost layers: processor cuts on the upper (rank 0) / lower (rank 7) three sides."""
    om = big["om"]
    dec = Decomposition(big["d"], 8, rank, n=(2, 2, 2))
    lm = M.UnstructuredMesh(dec.desc)
    gid = dec.cellGlobal[: dec.nOwned]
    phib = big["bd"]["value"]
    f = dev(big["T_h"][dec.cellGlobal])
    lflux, lphib = dev(dec.scatter_faces(big["flux_h"])), dev(dec.scatter_boundary(phib, om.nI))
    o = torch.zeros(lm.nCells, dtype=torch.float64, device="cuda")
    ref = om.div(big["flux_h"], big["T_h"], phib, 0, par=0)
    assert np.array_equal(host(ops.div(lm, lflux, f, lphib, o))[: dec.nOwned], ref[gid])
    ref = om.laplacian(big["T_h"], phib)
    assert np.array_equal(host(ops.laplacian(lm, f, lphib, o))[: dec.nOwned], ref[gid])
    # assembly on the sub-domain: affine vs stencil-driven kernel, and the structured SpMV vs the generic one
    ls = la.LinearSystem(lm, 1, zero=False)
    lbd = {k: dev(dec.scatter_boundary(v, om.nI)) for k, v in big["bd"].items()}

    class BD:
        value, refValue, valueFraction, refGrad = lbd["value"], lbd["refValue"], lbd["valueFraction"], lbd["refGrad"]
    terms = [dict(kind=ops.TERM_DIV, scheme=0, coeff=1.0, faceField=lflux),
             dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=torch.ones(lm.nFaces, dtype=torch.float64, device="cuda")),
             dict(kind=ops.TERM_DDT, coeff=1.0, cellField=f - 1.0, dt=1.0)]
    res = []
    for affine in (1, 0):
        lib().fvk_set_affine(C.c_int(affine))
        ls.values.fill_(float("nan")); ls.rhs.fill_(float("nan"))
        ops.assemble(lm, terms, BD, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
        res.append((host(ls.values).copy(), host(ls.rhs).copy()))
    lib().fvk_set_affine(C.c_int(1))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert not np.isnan(res[0][0]).any()
    x = torch.rand(lm.nCells, dtype=torch.float64, device="cuda")
    y1 = la.spmv(la.SparsityPattern.readOrCreate(lm), ls.values, x)
    y2 = la.spmv_structured(lm, ls.values, x)
    assert torch.equal(y1, y2)
