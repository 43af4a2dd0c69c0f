"""GPU tests of the decomposed path on ONE device: every rank's sub-mesh lives on the same GPU and the halo exchange is
emulated with device copies that follow the fvk_comm plan (the NCCL transport itself is exercised by tests/mgpu_check.py
under torchrun). Explicit operators on a sub-domain are asserted BIT-EXACT against the single-domain GPU result: ghost
cells + faceOrder keep the per-cell accumulation order of the undecomposed mesh."""
import numpy as np
import pytest
import torch

from foamadapter_b200 import la, mesh as M, ops
from foamadapter_b200.decomp import Decomposition
from oracle.cpu import Mesh as OMesh
from tests.test_explicit_gpu import dev, host

pytestmark = pytest.mark.gpu


def emulate_halo(decs, fields):
    """fields[r]: tensor of rank r (owned + ghost); copy owned values into the neighbours' ghost ranges."""
    for d, f in zip(decs, fields):
        for k, nb in enumerate(d.nbrRanks):
            o = decs[nb]
            kk = list(o.nbrRanks).index(d.rank)
            src = torch.from_numpy(o.sendCells[o.sendOff[kk]: o.sendOff[kk + 1]].astype(np.int64)).cuda()
            f[d.nOwned + int(d.recvOff[k]): d.nOwned + int(d.recvOff[k + 1])] = fields[nb][src]


@pytest.mark.parametrize("P", [2, 8])
@pytest.mark.parametrize("variant", [0, 5, 6])
def test_subdomain_operators_bit_exact(P, variant):
    from foamadapter_b200 import _capi
    g = M.MeshDesc.block(12, 10, 8, 1.2, 1.0, 0.8)
    gm, om = M.UnstructuredMesh(g), OMesh.from_desc(g)
    rng = np.random.default_rng(2)
    phi, phib, flux = rng.uniform(1, 2, om.nC), rng.uniform(1, 2, om.nB), rng.uniform(-1, 1, om.nF)
    U = rng.uniform(-1, 1, (om.nC, 3))
    _capi.lib().fvk_set_variant(variant)
    try:
        ref = {}
        out = torch.zeros(om.nC, dtype=torch.float64, device="cuda")
        ref["div"] = host(ops.div(gm, dev(flux), dev(phi), dev(phib), out)).copy()
        ref["lap"] = host(ops.laplacian(gm, dev(phi), dev(phib), out)).copy()
        o3 = torch.zeros((om.nC, 3), dtype=torch.float64, device="cuda")
        ref["grad"] = host(ops.grad(gm, dev(phi), dev(phib), o3)).copy()
        ff = torch.zeros(om.nF, dtype=torch.float64, device="cuda")
        ops.flux(gm, dev(U), dev(np.zeros((om.nB, 3))), ff)
        ref["flux"] = host(ff).copy()
        assert np.array_equal(ref["div"], om.div(flux, phi, phib, 0))
        decs = [Decomposition(g, P, r) for r in range(P)]
        meshes = [M.UnstructuredMesh(d.desc) for d in decs]
        # owned values only, ghosts by (emulated) halo exchange
        fields = []
        for d in decs:
            f = torch.zeros(d.nOwned + d.nGhost, dtype=torch.float64, device="cuda")
            f[: d.nOwned] = dev(phi[d.cellGlobal[: d.nOwned]])
            fields.append(f)
        emulate_halo(decs, fields)
        for d, lm, f in zip(decs, meshes, fields):
            assert lm.nOwned == d.nOwned and lm.nCells == d.nOwned + d.nGhost
            assert np.array_equal(host(f), phi[d.cellGlobal])
            gid = d.cellGlobal[: d.nOwned]
            lflux, lphib = dev(d.scatter_faces(flux)), dev(d.scatter_boundary(phib, om.nI))
            o = torch.full((lm.nCells,), float("nan"), dtype=torch.float64, device="cuda")
            assert np.array_equal(host(ops.div(lm, lflux, f, lphib, o))[: d.nOwned], ref["div"][gid])
            assert np.array_equal(host(ops.laplacian(lm, f, lphib, o))[: d.nOwned], ref["lap"][gid])
            o3 = torch.zeros((lm.nCells, 3), dtype=torch.float64, device="cuda")
            assert np.array_equal(host(ops.grad(lm, f, lphib, o3))[: d.nOwned], ref["grad"][gid])
            lf = torch.zeros(lm.nFaces, dtype=torch.float64, device="cuda")
            ops.flux(lm, dev(U[d.cellGlobal]), dev(np.zeros((lm.nBoundaryFaces, 3))), lf)
            assert np.array_equal(host(lf), ref["flux"][d.faceGlobal])
    finally:
        _capi.lib().fvk_set_variant(0)


@pytest.mark.parametrize("dims,P", [((160, 8, 6), 2), ((24, 20, 18), 8), ((12, 10, 8), 4)])
def test_tile_phases_split_the_operator(dims, P):
    """fvk_mesh_set_tile_phase: INTERIOR leaves the tiles that read ghost cells untouched (so it may run while the halo
    exchange is in flight), HALO computes exactly those; together they equal the one-pass result bit for bit."""
    from foamadapter_b200._capi import check, lib
    g = M.MeshDesc.block(*dims, 1.2, 1.0, 0.8)
    om = OMesh.from_desc(g)
    rng = np.random.default_rng(4)
    phi, phib, flux = rng.uniform(1, 2, om.nC), rng.uniform(1, 2, om.nB), rng.uniform(-1, 1, om.nF)
    for r in range(P):
        d = Decomposition(g, P, r)
        lm = M.UnstructuredMesh(d.desc)
        f = dev(phi[d.cellGlobal])
        lflux, lphib = dev(d.scatter_faces(flux)), dev(d.scatter_boundary(phib, om.nI))
        full = torch.full((lm.nCells,), float("nan"), dtype=torch.float64, device="cuda")
        ops.div(lm, lflux, f, lphib, full)
        full3 = torch.zeros((lm.nCells, 3), dtype=torch.float64, device="cuda")
        ops.grad(lm, f, lphib, full3)
        try:
            # interior phase with POISONED ghost values: nothing it computes may depend on them
            fp = f.clone()
            fp[d.nOwned:] = float("nan")
            split = torch.full((lm.nCells,), 7.0, dtype=torch.float64, device="cuda")
            split3 = torch.full((lm.nCells, 3), 7.0, dtype=torch.float64, device="cuda")
            check(lib().fvk_mesh_set_tile_phase(lm.handle, 1))
            ops.div(lm, lflux, fp, lphib, split)
            ops.grad(lm, fp, lphib, split3)
            inter = host(split).copy()
            assert not np.isnan(inter[: d.nOwned]).any()
            done = inter[: d.nOwned] != 7.0
            assert np.array_equal(inter[: d.nOwned][done], host(full)[: d.nOwned][done])
            check(lib().fvk_mesh_set_tile_phase(lm.handle, 2))
            ops.div(lm, lflux, f, lphib, split)
            ops.grad(lm, f, lphib, split3)
        finally:
            check(lib().fvk_mesh_set_tile_phase(lm.handle, 0))
        assert np.array_equal(host(split)[: d.nOwned], host(full)[: d.nOwned])
        assert np.array_equal(host(split3)[: d.nOwned], host(full3)[: d.nOwned])
        if d.nGhost and lm.nOwned >= 2048:
            assert done.any() and not done.all()  # a real split: some tiles on each side


def test_subdomain_assembly_and_spmv_rows():
    g = M.MeshDesc.block(9, 7, 5, 0.9, 0.7, 0.5)
    gm, om = M.UnstructuredMesh(g), OMesh.from_desc(g)
    rng = np.random.default_rng(5)
    gamma, x = rng.uniform(0.5, 1.5, om.nF), rng.uniform(-1, 1, om.nC)
    old = rng.uniform(1, 2, om.nC)
    bd = dict(value=np.zeros(om.nB), refValue=rng.uniform(1, 2, om.nB), valueFraction=np.ones(om.nB), refGrad=np.zeros(om.nB))

    class BD:
        def __init__(s, b): s.value, s.refValue, s.valueFraction, s.refGrad = (dev(b[k]) for k in ("value", "refValue", "valueFraction", "refGrad"))

    def assemble(mesh, gam, oldf, b):
        ls = la.LinearSystem(mesh, 1, zero=False)
        ops.assemble(mesh, [dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=dev(gam)), dict(kind=ops.TERM_DDT, coeff=1.0, cellField=dev(oldf), dt=0.5)],
                     BD(b), ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
        return ls
    gls = assemble(gm, gamma, old, bd)
    ref_Ax = host(la.spmv(la.SparsityPattern.readOrCreate(gm), gls.values, dev(x)))
    ref_rhs = host(gls.rhs)
    for r in range(4):
        d = Decomposition(g, 4, r)
        lm = M.UnstructuredMesh(d.desc)
        lbd = {k: d.scatter_boundary(v, om.nI) for k, v in bd.items()}
        ls = assemble(lm, d.scatter_faces(gamma), d.scatter_cells(old), lbd)
        gid = d.cellGlobal[: d.nOwned]
        assert np.array_equal(host(ls.rhs)[: d.nOwned], ref_rhs[gid])
        Ax = host(la.spmv(la.SparsityPattern.readOrCreate(lm), ls.values, dev(d.scatter_cells(x))))
        assert Ax.shape[0] == d.nOwned
        # sub-domain rows are [lower | diag | upper] with each half in GLOBAL face order: a ghost-owned face sits where the
        # undecomposed row has it, the products are summed in the same order -> the same bits
        assert np.array_equal(Ax, ref_Ax[gid])


@pytest.mark.parametrize("dims,P", [((24, 20, 18), 8), ((16, 12, 10), 4), ((40, 6, 6), 2)])
def test_every_rank_keeps_the_index_free_kernels(dims, P):
    """A sub-domain numbers its ghost-owned faces last; its CSR rows are nevertheless in stencil (= global face) order on EVERY
    rank -- also those with a processor patch on a lower side -- so the index-free assembly / rAU,HbyA / structured SpMV kernels
    serve all of them (round-2 finding: rows sorted by local face id left 6 of 8 ranks on the generic kernels)."""
    g = M.MeshDesc.block(*dims, 1.0, 1.0, 1.0)
    for r in range(P):
        lm = M.UnstructuredMesh(Decomposition(g, P, r).desc)
        assert lm.size(M.ROWS_IN_STENCIL_ORDER) == 1, (P, r)
        assert lm.size(M.AFFINE_TOPOLOGY) == 1, (P, r)

