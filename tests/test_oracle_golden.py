"""Pins the CPU oracle (oracle/fvo.cpp) against the reference's own golden vectors and known-answer
tests. CPU only."""
import numpy as np
import pytest

from foamadapter_b200.mesh import MeshDesc
from oracle.cpu import Mesh, cg
from tests.helpers import load_golden, neon_view, oracle_mesh_from_view


@pytest.fixture(scope="module")
def op_case():
    g = load_golden("setup_operator")
    v = neon_view(g)
    return g, v, oracle_mesh_from_view(v)


def _zero_gradient_bvalue(m, T):
    # zeroGradient -> fixedGradient 0: value = internal[faceCells] + 0 * (1/deltaCoeffs)
    return m.correct_bcs([2] * (len(m.patchOffsets) - 1), [0.0] * (len(m.patchOffsets) - 1), T)["value"]


@pytest.mark.parametrize("which,par", [("divT_Serial", 0), ("divT_OpenMP", 1), ("ofDivT", 0)])
def test_div_golden(op_case, which, par):
    # test/setup_operator/0/{T,phi,divT_*}: T = celli, phi_internal = facei, phi_boundary = 0
    g, v, m = op_case
    T = g["field_T"]
    assert np.array_equal(T, np.arange(25.0))
    phi = np.concatenate([g["field_phi"], np.zeros(m.nB)])
    assert np.array_equal(g["field_phi"], np.arange(40.0))
    div = m.div(phi, T, _zero_gradient_bvalue(m, T), scheme=0, par=par)
    np.testing.assert_allclose(div, g["field_" + which], rtol=5e-15, atol=0)


@pytest.mark.parametrize("which,par", [("gradT_Serial", 0), ("gradT_OpenMP", 1), ("ofGradT", 0)])
def test_grad_golden(op_case, which, par):
    g, v, m = op_case
    T = g["field_T"]
    grad = m.grad(T, _zero_gradient_bvalue(m, T), par=par)
    ref = g["field_" + which]
    # x,y to 1e-12 absolute on O(250) values; z is 1e-13 noise from the dropped empty faces and is
    # zeroed by OpenFOAM (the reference relaxes z to 1e-6 too, test/test_operators.cpp:79)
    np.testing.assert_allclose(grad[:, :2], ref[:, :2], rtol=0, atol=1e-12)
    np.testing.assert_allclose(grad[:, 2], ref[:, 2], rtol=0, atol=1e-6)


def test_sparsity_known_answer_1d():
    # src/NeoN/test/linearAlgebra/sparsityPattern.cpp:36-92
    m = Mesh.from_desc(MeshDesc.uniform_1d(10))
    assert list(m.diagOffset) == [0] + [1] * 9
    assert list(m.rowOffs) == [0, 2, 5, 8, 11, 14, 17, 20, 23, 26, 28]
    assert list(m.colIdxs[:11]) == [0, 1, 0, 1, 2, 1, 2, 3, 2, 3, 4]
    assert m.colIdxs.size == 10 + 2 * 9 and m.ownerOffset.size == 9 and m.neighbourOffset.size == 9


def test_sparsity_subset_of_cellcells_3d():
    # test/test_sparsityPattern.cpp:34-62 on the 3x3x3 fixture
    v = neon_view(load_golden("setup_stencil3D"))
    m = oracle_mesh_from_view(v)
    assert m.colIdxs.size == m.nC + 2 * m.nI and m.rowOffs.size == m.nC + 1
    own, nei = v["faceOwner"][: m.nI], v["faceNeighbour"]
    for c in range(m.nC):
        cols = set(m.colIdxs[m.rowOffs[c]: m.rowOffs[c + 1]])
        expect = {c} | set(nei[own == c]) | set(own[nei == c])
        assert cols == expect
    for f in range(m.nI):  # naming trap: neighbourOffset -> (row nei, col own)
        assert m.colIdxs[m.rowOffs[nei[f]] + m.neighbourOffset[f]] == own[f]
        assert m.colIdxs[m.rowOffs[own[f]] + m.ownerOffset[f]] == nei[f]


def test_cell_to_face_stencil_known_answer():
    # src/NeoN/test/finiteVolume/cellCentred/stencil/cellToFaceStencil.cpp:27-45
    m = Mesh.from_desc(MeshDesc.uniform_1d(5))
    seg, val = m.stencil()
    got = [list(val[seg[c]: seg[c + 1]]) for c in range(5)]
    assert got == [[0, 4], [0, 1], [1, 2], [2, 3], [3, 5]]


def test_residual_known_answer():
    # src/NeoN/test/linearAlgebra/utilities.cpp:22-44: [[1,2,3],[4,5,6],[7,8,9]].1 - 2 = (4,13,22)
    from oracle.cpu import call
    rowOffs = np.array([0, 3, 6, 9], np.int32)
    col = np.array([0, 1, 2] * 3, np.int32)
    vals = np.arange(1.0, 10.0)
    res = np.zeros(3)
    call("fvo_residual", 0, 3, rowOffs, col, vals, np.full(3, 2.0), np.ones(3), res)
    assert list(res) == [4.0, 13.0, 22.0]


def test_cg_known_answer():
    # src/NeoN/test/linearAlgebra/ginkgo.cpp:95-124
    rowOffs = np.array([0, 2, 5, 7], np.int32)
    col = np.array([0, 1, 0, 1, 2, 1, 2], np.int32)
    vals = np.array([1.0, -0.1, -0.1, 1.0, -0.1, -0.1, 1.0])
    x, st, hist = cg(rowOffs, col, vals, np.array([1.0, 2.0, 3.0]), np.zeros(3), jacobi=False, max_iter=100,
                     rel_tol=1e-7, abs_tol=0.0, max_hist=16)
    np.testing.assert_allclose(x, [1.24489796, 2.44897959, 3.24489796], atol=1e-8)
    assert st["numIter"] == 3
    assert abs(st["initResNorm"] - 3.741657386) < 1e-8
    assert st["finalResNorm"] < 1e-4
    assert len(hist) == 4 and hist[0] == st["initResNorm"]


def _bcs_1d(m, kind, first, last, vec):
    k = {"fixedValue": 1, "fixedGradient": 2}[kind]
    cst = [[first] * 3, [last] * 3] if vec else [first, last]
    return k, cst


@pytest.mark.parametrize("vec", [False, True])
@pytest.mark.parametrize("kind,first,last", [("fixedValue", 0.5, 10.5), ("fixedGradient", -10.0, 10.0)])
def test_laplacian_linear_field_is_zero(kind, first, last, vec):
    # src/NeoN/test/finiteVolume/cellCentred/operator/laplacianOperator.cpp:21-137
    m = Mesh.from_desc(MeshDesc.uniform_1d(10))
    phi = np.arange(1.0, 11.0)
    if vec:
        phi = np.repeat(phi[:, None], 3, 1).copy()
    k, cst = _bcs_1d(m, kind, first, last, vec)
    bd = m.correct_bcs([k, k], cst, phi)
    lap = m.laplacian(phi, bd["value"])
    assert np.abs(lap).max() < 1e-8
    if not vec:
        for coeff in (1.0, -0.5):
            ls = m.empty_system()
            m.laplacian_imp(ls, np.full(m.nF, 2.0), bd, coeff=coeff)
            res = m.residual(ls["values"], ls["rhs"], phi)
            assert np.abs(res).max() < 1e-8


@pytest.mark.parametrize("vec", [False, True])
def test_face_normal_grad_known_answer(vec):
    # src/NeoN/test/finiteVolume/cellCentred/faceNormalGradient/uncorrected.cpp:45-68
    m = Mesh.from_desc(MeshDesc.uniform_1d(10))
    phi, b = np.arange(1.0, 11.0), np.array([0.5, 10.5])
    if vec:
        phi, b = np.repeat(phi[:, None], 3, 1).copy(), np.repeat(b[:, None], 3, 1).copy()
    sn = m.face_normal_grad(phi, b)
    np.testing.assert_allclose(sn[: m.nI], 10.0, atol=1e-8)
    np.testing.assert_allclose(sn[m.nI], -10.0, atol=1e-8)   # left boundary
    np.testing.assert_allclose(sn[m.nI + 1], 10.0, atol=1e-8)  # right boundary


def test_interpolation_of_uniform_field():
    # src/NeoN/test/finiteVolume/cellCentred/interpolation/{linear,upwind}.cpp: uniform 1 -> 1 exactly
    m = Mesh.from_desc(MeshDesc.uniform_1d(10))
    one, b = np.ones(10), np.ones(2)
    assert np.array_equal(m.interpolate(one, b, scheme=0), np.ones(m.nF))
    assert np.array_equal(m.interpolate(one, b, scheme=1, faceFlux=np.ones(m.nF)), np.ones(m.nF))


def test_div_of_uniform_field_is_zero_1d():
    # src/NeoN/test/finiteVolume/cellCentred/operator/gaussGreenDiv.cpp:18-69
    m = Mesh.from_desc(MeshDesc.uniform_1d(10))
    flux = np.ones(m.nF)
    flux[m.nI] = -1.0  # left boundary face: outward normal -x
    div = m.div(flux, np.ones(10), np.ones(2), scheme=0)
    assert np.array_equal(div, np.zeros(10))


def test_ddt_and_source_implicit_identities():
    # test/test_implicitOperators.cpp:37-153: diag == coeff V/dt ; A.x - b identities
    v = neon_view(load_golden("setup_stencil3D"))
    m = oracle_mesh_from_view(v)
    rng = np.random.default_rng(1)
    old = rng.uniform(1, 2, m.nC)
    ls = m.empty_system()
    m.ddt_imp(ls, old, dt=0.5, coeff=2.0)
    diag = ls["values"][m.rowOffs[:-1] + m.diagOffset]
    np.testing.assert_allclose(diag, 2.0 * m.V / 0.5, rtol=1e-15)
    np.testing.assert_allclose(m.residual(ls["values"], ls["rhs"], old), 0.0, atol=1e-15)
    ls = m.empty_system()
    k = rng.uniform(1, 2, m.nC)
    m.source_imp(ls, k, coeff=1.0)
    np.testing.assert_allclose(ls["values"][m.rowOffs[:-1] + m.diagOffset], k * m.V, rtol=1e-15)
