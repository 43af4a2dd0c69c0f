"""GPU parity tests for the mesh handle and the explicit operators: CUDA path (through the C ABI)
vs the CPU oracle on the same seeded inputs, vs the reference's golden vectors, and size-independent
properties at the benchmark size.

Tolerance: north_star asks <= 1e-12 relative for fp64 explicit operators. The gather kernels sum in
the Serial executor's order without FMA contraction, so we assert BIT-EXACT equality with the
Serial oracle (rtol=0), which is stronger; TOL is kept for reductions whose order differs."""
import ctypes as C

import numpy as np
import pytest
import torch

from foamadapter_b200 import _capi, fvcc, mesh as M, ops
from oracle.cpu import Mesh as OMesh
from tests.helpers import FIXTURE_BLOCKS, load_golden, neon_view, oracle_mesh_from_view, renumbered_block

pytestmark = pytest.mark.gpu
TOL = 1e-12


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def host(t):
    return t.detach().cpu().numpy()


MESHES = {
    "1d10": lambda: M.MeshDesc.uniform_1d(10),
    "fix5x5x1": lambda: M.MeshDesc.block(*FIXTURE_BLOCKS["setup_operator"][0], *FIXTURE_BLOCKS["setup_operator"][1],
                                         patches=FIXTURE_BLOCKS["setup_operator"][2]),
    "fix3x3x3": lambda: M.MeshDesc.block(3, 3, 3, patches=FIXTURE_BLOCKS["setup_stencil3D"][2]),
    "7x5x4": lambda: M.MeshDesc.block(7, 5, 4, 0.7, 0.3, 0.9),
    "33x17x9": lambda: M.MeshDesc.block(33, 17, 9, 1.0, 0.5, 0.2),
    "1x1x1": lambda: M.MeshDesc.block(1, 1, 1),
    "cavity20": lambda: M.MeshDesc.block(20, 20, 1, 0.1, 0.1, 0.01, patches=M.PATCHES_CAVITY2D),
    # several 32x4x4 bricks per axis incl. ragged edge bricks (variant 0 = brick kernel)
    "70x11x9": lambda: M.MeshDesc.block(70, 11, 9, 1.0, 0.5, 0.2),
    # randomly renumbered cells: no block structure, the brick plan falls back to runs of consecutive cells
    "renum12x11x10": lambda: renumbered_block(12, 11, 10, 3),
    # non power-of-two brick (FVK_BRICK at mesh creation): the kernel's integer-division index path
    "20x10x11@brick8,3,5": lambda: M.MeshDesc.block(20, 10, 11, 1.0, 0.5, 0.55),
}


@pytest.fixture(scope="module", params=sorted(MESHES))
def case(request):
    import os
    d = MESHES[request.param]()
    brick = request.param.partition("@brick")[2]
    if brick:
        os.environ["FVK_BRICK"] = brick
    try:
        gm = M.UnstructuredMesh(d)
    finally:
        os.environ.pop("FVK_BRICK", None)
    return request.param, d, gm, OMesh.from_desc(d)


def _fields(om, seed, vec=False):
    rng = np.random.default_rng(seed)
    shp = lambda n: (n, 3) if vec else (n,)
    phi = rng.uniform(1, 2, shp(om.nC))
    phib = rng.uniform(1, 2, shp(om.nB))
    flux = rng.uniform(-1, 1, om.nF)
    view = rng.uniform(0.5, 1.5, om.nC)
    return phi, phib, flux, view


def test_mesh_handle_matches_oracle(case):
    name, d, gm, om = case
    assert (gm.nCells, gm.nInternalFaces, gm.nBoundaryFaces, gm.nnz) == (om.nC, om.nI, om.nB, om.nnz)
    # sparsity pattern: bit-exact
    assert np.array_equal(gm.to_host(M.ROW_OFFS), om.rowOffs)
    assert np.array_equal(gm.to_host(M.COL_IDXS), om.colIdxs)
    assert np.array_equal(gm.to_host(M.OWNER_OFFSET), om.ownerOffset)
    assert np.array_equal(gm.to_host(M.NEIGHBOUR_OFFSET), om.neighbourOffset)
    assert np.array_equal(gm.to_host(M.DIAG_OFFSET), om.diagOffset)
    seg, val = om.stencil()
    assert np.array_equal(gm.to_host(M.STENCIL_SEGMENTS), seg)
    assert np.array_equal(gm.to_host(M.STENCIL_VALUES), val)
    # geometry scheme computed on the device: bit-exact with the oracle
    assert np.array_equal(gm.to_host(M.WEIGHTS), om.w)
    assert np.array_equal(gm.to_host(M.DELTACOEFFS), om.dc)
    assert np.array_equal(gm.to_host(M.NONORTH_DELTACOEFFS), om.nodc)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("vec", [False, True])
@pytest.mark.parametrize("scheme", [0, 1])
def test_div(case, scheme, vec, variant):
    name, d, gm, om = case
    phi, phib, flux, view = _fields(om, 1, vec)
    _capi.lib().fvk_set_variant(variant)
    try:
        out = torch.zeros_like(dev(phi))
        ops.div(gm, dev(flux), dev(phi), dev(phib), out, scheme, 1.0, None, ops.SET)
        ref = om.div(flux, phi, phib, scheme)
        assert np.array_equal(host(out), ref)
        # Coeff with a view, accumulate-then-scale on a non-zero result (computeDiv semantics)
        pre = np.random.default_rng(5).uniform(-1, 1, phi.shape)
        out = dev(pre)
        ops.div(gm, dev(flux), dev(phi), dev(phib), out, scheme, -0.5, dev(view), ops.ACC_SCALE)
        ref = om.div(flux, phi, phib, scheme, coeff=-0.5, coeffView=view, res=pre.copy())
        assert np.array_equal(host(out), ref)
        # explicitOperation semantics: source += op(0)
        out = dev(pre)
        ops.div(gm, dev(flux), dev(phi), dev(phib), out, scheme, 2.0, None, ops.ADD)
        ref = pre + om.div(flux, phi, phib, scheme, coeff=2.0)
        assert np.array_equal(host(out), ref)
    finally:
        _capi.lib().fvk_set_variant(0)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6])
def test_grad(case, variant):
    name, d, gm, om = case
    phi, phib, _, _ = _fields(om, 2)
    _capi.lib().fvk_set_variant(variant)
    try:
        out = torch.empty((om.nC, 3), dtype=torch.float64, device="cuda")
        ops.grad(gm, dev(phi), dev(phib), out, ops.SET)
        assert np.array_equal(host(out), om.grad(phi, phib))
    finally:
        _capi.lib().fvk_set_variant(0)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("vec", [False, True])
def test_laplacian(case, vec, variant):
    name, d, gm, om = case
    phi, phib, _, view = _fields(om, 3, vec)
    _capi.lib().fvk_set_variant(variant)
    try:
        out = torch.zeros_like(dev(phi))
        ops.laplacian(gm, dev(phi), dev(phib), out, 1.0, None, ops.SET)
        assert np.array_equal(host(out), om.laplacian(phi, phib))
        out = torch.zeros_like(dev(phi))
        ops.laplacian(gm, dev(phi), dev(phib), out, 0.25, dev(view), ops.SET)
        assert np.array_equal(host(out), om.laplacian(phi, phib, coeff=0.25, coeffView=view))
    finally:
        _capi.lib().fvk_set_variant(0)


@pytest.mark.parametrize("vec", [False, True])
def test_affine_interior_kernel_equals_generic_kernel(vec):
    """A/B of the index-free kernel (regular cells of a block mesh) + list gather against the generic brick kernel: every operator and
    accumulation mode, bit for bit; and the plan really has an affine box for this mesh."""
    import ctypes as C
    d = M.MeshDesc.block(64, 24, 12, 1.0, 0.5, 0.2)
    info = (C.c_int32 * 5)()
    _capi.check(_capi.lib().fvk_brick_plan_affine_info(C.byref(d.c), info))
    assert info[0] == 1 and 0 < info[4] < d.nCells
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    phi, phib, flux, view = _fields(om, 11, vec)
    res = {}
    try:
        for affine in (1, 0):
            _capi.lib().fvk_set_affine(affine)
            r = []
            for scheme in (0, 1):
                for mode, coeff, vw in ((ops.SET, 1.0, None), (ops.ACC_SCALE, -0.5, dev(view)), (ops.ADD, 2.0, None)):
                    out = dev(np.random.default_rng(5).uniform(-1, 1, phi.shape))
                    ops.div(gm, dev(flux), dev(phi), dev(phib), out, scheme, coeff, vw, mode)
                    r.append(host(out).copy())
            out = dev(np.random.default_rng(6).uniform(-1, 1, phi.shape))
            ops.laplacian(gm, dev(phi), dev(phib), out, 0.7, dev(view), ops.ACC_SCALE)
            r.append(host(out).copy())
            if not vec:
                o3 = torch.empty((om.nC, 3), dtype=torch.float64, device="cuda")
                ops.grad(gm, dev(phi), dev(phib), o3, ops.SET)
                r.append(host(o3).copy())
            fl = np.random.default_rng(7).uniform(-1, 1, (om.nF, 3) if vec else om.nF)
            out = torch.zeros_like(dev(phi))
            ops.surface_integrate(gm, dev(fl), out)
            r.append(host(out).copy())
            res[affine] = r
    finally:
        _capi.lib().fvk_set_affine(1)
    assert len(res[0]) == len(res[1])
    for x, y in zip(res[0], res[1]):
        assert np.array_equal(x, y)
    assert np.array_equal(res[1][0], om.div(flux, phi, phib, 0))


@pytest.mark.parametrize("vec", [False, True])
def test_surface_integrate(case, vec):
    name, d, gm, om = case
    rng = np.random.default_rng(4)
    flux = rng.uniform(-1, 1, (om.nF, 3) if vec else om.nF)
    out = torch.zeros((om.nC, 3) if vec else om.nC, dtype=torch.float64, device="cuda")
    ops.surface_integrate(gm, dev(flux), out, -1.0, None, ops.SET)
    assert np.array_equal(host(out), om.surface_integrate(flux, coeff=-1.0))


@pytest.mark.parametrize("vec", [False, True])
def test_face_kernels(case, vec):
    name, d, gm, om = case
    phi, phib, flux, _ = _fields(om, 6, vec)
    shp = (om.nF, 3) if vec else (om.nF,)
    out = torch.empty(shp, dtype=torch.float64, device="cuda")
    ops.interpolate(gm, dev(phi), dev(phib), out, ops.LINEAR)
    assert np.array_equal(host(out), om.interpolate(phi, phib, 0))
    ops.interpolate(gm, dev(phi), dev(phib), out, ops.UPWIND, dev(flux))
    assert np.array_equal(host(out), om.interpolate(phi, phib, 1, flux))
    ops.face_normal_grad(gm, dev(phi), dev(phib), out)
    assert np.array_equal(host(out), om.face_normal_grad(phi, phib))
    if not vec:
        w = torch.empty(om.nF, dtype=torch.float64, device="cuda")
        wb = torch.empty(om.nB, dtype=torch.float64, device="cuda")
        ops.interpolation_weights(gm, w, wb, ops.UPWIND, dev(flux))
        rw, rwb = om.upwind_weights(flux)
        assert np.array_equal(host(w), rw) and np.array_equal(host(wb), rwb)
        ops.interpolation_weights(gm, w, wb, ops.LINEAR)
        assert np.array_equal(host(w), om.w) and np.array_equal(host(wb), om.bWeights)
        with pytest.raises(_capi.FvkError):  # upwind without a flux: "limited scheme require a faceFlux"
            ops.interpolate(gm, dev(phi), dev(phib), out, ops.UPWIND, None)


def test_conum(case):
    name, d, gm, om = case
    _, _, flux, _ = _fields(om, 7)
    res = host(ops.conum(gm, dev(flux), 1e-3))
    ref = om.conum(flux, 1e-3)
    assert res[0] == ref[0]                      # max: exact
    assert abs(res[1] - ref[1]) <= TOL * abs(ref[1])  # mean: reduction order differs


@pytest.mark.parametrize("vec", [False, True])
def test_boundary_conditions(case, vec):
    name, d, gm, om = case
    if om.nB == 0:
        pytest.skip("no boundary")
    rng = np.random.default_rng(8)
    nP = gm.nPatches
    kinds_all = ["fixedValue", "fixedGradient", "extrapolated", "calculated", "zeroGradient", "noSlip"]
    kinds = [kinds_all[(i + (1 if vec else 0)) % len(kinds_all)] for i in range(nP)]
    consts = [tuple(rng.uniform(-2, 2, 3)) if vec else float(rng.uniform(-2, 2)) for _ in range(nP)]
    f = fvcc.VolumeField(gm, "phi", 3 if vec else 1, list(zip(kinds, consts)))
    phi = rng.uniform(1, 2, (om.nC, 3) if vec else om.nC)
    f.internal.copy_(dev(phi))
    f.correctBoundaryConditions()
    okind = [fvcc._BC_KINDS[k] for k, _ in f.bcs]
    ref = om.correct_bcs(okind, [c for _, c in f.bcs], phi)
    for key in ("value", "refValue", "valueFraction", "refGrad"):
        assert np.array_equal(host(getattr(f.boundary, key)), ref[key]), key


def test_golden_div_grad_through_the_gpu():
    """test/setup_operator/0/{T,phi,divT_Serial,gradT_Serial}: the reference's own 16-digit vectors."""
    g = load_golden("setup_operator")
    v = neon_view(g)
    gm = M.UnstructuredMesh(M.MeshDesc.from_arrays(v, v["patch_names"]))
    T = fvcc.VolumeField(gm, "T", 1, [("zeroGradient", 0.0)])
    T.internal.copy_(dev(g["field_T"]))
    T.correctBoundaryConditions()
    phi = fvcc.SurfaceField(gm, "phi")
    phi.internal[: gm.nInternalFaces] = dev(g["field_phi"])
    div = torch.zeros(gm.nCells, dtype=torch.float64, device="cuda")
    fvcc.GaussGreenDiv(gm, "linear").div(div, phi, T, fvcc.Coeff(1.0))
    np.testing.assert_allclose(host(div), g["field_divT_Serial"], rtol=5e-15, atol=0)
    grad = fvcc.GaussGreenGrad(gm).grad(T)
    np.testing.assert_allclose(host(grad)[:, :2], g["field_gradT_Serial"][:, :2], rtol=0, atol=1e-12)
    np.testing.assert_allclose(host(grad)[:, 2], g["field_gradT_Serial"][:, 2], rtol=0, atol=1e-6)


def test_errors_are_loud(case):
    name, d, gm, om = case
    phi, phib, flux, _ = _fields(om, 9)
    with pytest.raises(_capi.FvkError):
        _capi.check(_capi.lib().fvk_div_s(gm.handle, 7, _capi.ptr(dev(flux)), _capi.ptr(dev(phi)), _capi.ptr(dev(phib)),
                                          C.c_double(1.0), None, _capi.ptr(dev(phi)), 0, None))
    with pytest.raises(ValueError):
        ops.div(gm, dev(flux)[:-1].contiguous(), dev(phi), dev(phib), dev(phi))


# ---- benchmark-size properties (128^3; the oracle takes seconds here, so it is used once) -----------
@pytest.fixture(scope="module")
def big():
    d = M.MeshDesc.block(128, 128, 128, 0.1, 0.1, 0.01)
    return d, M.UnstructuredMesh(d)


def test_128_parity_and_properties(big):
    d, gm = big
    om = OMesh.from_desc(d)
    rng = np.random.default_rng(42)
    phi = rng.uniform(1, 2, om.nC)
    flux = np.concatenate([np.arange(om.nI, dtype=np.float64), np.zeros(om.nB)])  # bench_explicitOperators.cpp:43-46
    T = fvcc.VolumeField(gm, "T", 1, [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)])
    T.internal.copy_(dev(phi))
    T.correctBoundaryConditions()
    phib = host(T.boundary.value)
    dphi, dflux, dphib = T.internal, dev(flux), T.boundary.value
    out = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
    for variant in (0, 1, 2, 3, 5):
        _capi.lib().fvk_set_variant(variant)
        ops.div(gm, dflux, dphi, dphib, out, ops.LINEAR)
        assert np.array_equal(host(out), om.div(flux, phi, phib, 0))
    _capi.lib().fvk_set_variant(0)
    ops.laplacian(gm, dphi, dphib, out)
    assert np.array_equal(host(out), om.laplacian(phi, phib))
    g = torch.empty((om.nC, 3), dtype=torch.float64, device="cuda")
    ops.grad(gm, dphi, dphib, g)
    assert np.array_equal(host(g), om.grad(phi, phib))
    # determinism: run-to-run bit-identical (no atomics)
    out2 = torch.zeros_like(out)
    ops.laplacian(gm, dphi, dphib, out2)
    assert torch.equal(out, out2)
    # conservation: sum_c V_c lap_c == sum over boundary faces of |Sf| snGrad_b (internal fluxes cancel)
    sn = torch.empty(om.nF, dtype=torch.float64, device="cuda")
    ops.face_normal_grad(gm, dphi, dphib, sn)
    lhs = float((out * dev(om.V)).sum())
    rhs = float((sn[om.nI:] * dev(om.magSf[om.nI:])).sum())
    assert abs(lhs - rhs) <= 1e-9 * abs(rhs)
    # linearity of div in phi (fixed flux): div(a*phi1 + phi2) == a*div(phi1) + div(phi2) to rounding
    phi2 = dev(rng.uniform(1, 2, om.nC))
    zb = torch.zeros_like(dphib)
    o1, o2, o3 = (torch.zeros_like(out) for _ in range(3))
    ops.div(gm, dflux, dphi, zb, o1)
    ops.div(gm, dflux, phi2, zb, o2)
    ops.div(gm, dflux, 3.0 * dphi + phi2, zb, o3)
    scale = float(o3.abs().max())
    assert float((o3 - (3.0 * o1 + o2)).abs().max()) <= 1e-12 * scale
    # gradient of a uniform field vanishes in the interior (closed cells)
    one = torch.ones(om.nC, dtype=torch.float64, device="cuda")
    ops.grad(gm, one, torch.ones(om.nB, dtype=torch.float64, device="cuda"), g)
    assert float(g.abs().max()) * float(dev(om.V).max()) <= 1e-12 * float(dev(om.magSf).max())
