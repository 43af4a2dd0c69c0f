"""GPU: the kernels on renumbered meshes (no block structure in the numbering: brick kernel over runs of consecutive cells,
stencil-driven assembly, generic SpMV) are bit-identical to the Serial oracle on the same mesh."""
import numpy as np
import pytest
import torch

from foamadapter_b200 import fvcc, la, mesh as M, ops
from oracle.cpu import Mesh as OMesh
from tests.helpers import renumbered_block

pytestmark = pytest.mark.gpu
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
host = lambda t: t.detach().cpu().numpy()


@pytest.mark.parametrize("method", ["random", "rcm", "morton"])
def test_operators_assembly_spmv_on_renumbered_meshes(method):
    g = renumbered_block(20, 13, 9, seed=4, box=(1.0, 0.7, 0.5))
    d = g if method == "random" else g.renumbered(method)[0]
    gm, om = M.UnstructuredMesh(d), OMesh.from_desc(d)
    assert not gm.size(M.AFFINE_TOPOLOGY)
    rng = np.random.default_rng(11)
    T = fvcc.VolumeField(gm, "T", 1, [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)])
    T_h = rng.uniform(1, 2, om.nC)
    T.internal.copy_(dev(T_h)); T.correctBoundaryConditions()
    bd = om.correct_bcs([1, 1, 2], [10.5, 1.5, 0.0], T_h)
    flux_h = rng.uniform(-1, 1, om.nF)
    flux = dev(flux_h)
    out = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
    for scheme in (0, 1):
        ops.div(gm, flux, T.internal, T.boundary.value, out, scheme=scheme)
        assert np.array_equal(host(out), om.div(flux_h, T_h, bd["value"], scheme))
    ops.laplacian(gm, T.internal, T.boundary.value, out)
    assert np.array_equal(host(out), om.laplacian(T_h, bd["value"]))
    g3 = torch.empty((om.nC, 3), dtype=torch.float64, device="cuda")
    ops.grad(gm, T.internal, T.boundary.value, g3)
    assert np.array_equal(host(g3), om.grad(T_h, bd["value"]))
    gamma_h, old_h = rng.uniform(0.5, 1.5, om.nF), rng.uniform(0, 1, om.nC)
    ols = om.empty_system(False)
    om.div_imp(ols, flux_h, bd, 1, 1.0, None); om.laplacian_imp(ols, gamma_h, bd, -1.0, None); om.ddt_imp(ols, old_h, 0.1, 1.0, None)
    ls = la.LinearSystem(gm, 1, zero=False)
    ops.assemble(gm, [dict(kind=ops.TERM_DIV, scheme=1, coeff=1.0, faceField=flux), dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=dev(gamma_h)),
                      dict(kind=ops.TERM_DDT, coeff=1.0, cellField=dev(old_h), dt=0.1)], T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
    assert np.array_equal(host(ls.values), ols["values"]) and np.array_equal(host(ls.rhs), ols["rhs"])
    x_h = rng.uniform(-1, 1, om.nC)
    y = la.spmv(la.SparsityPattern.readOrCreate(gm), ls.values, dev(x_h))
    assert np.array_equal(host(y), om.spmv(ols["values"], x_h))
    assert np.array_equal(host(la.spmv_structured(gm, ls.values, dev(x_h))), host(y))   # falls back to the generic kernel
