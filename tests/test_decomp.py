"""CPU tests of the domain decomposition (host code of libfvk, no GPU): the decomposition map, the sub-mesh layout with
ghost cells, the halo plan, and -- with the CPU oracle running on every sub-mesh -- that a decomposed evaluation
reproduces the single-domain result. Includes a world_size-2 gloo run of the halo protocol."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from foamadapter_b200.decomp import Decomposition, default_split, simple_map
from foamadapter_b200.mesh import MeshDesc
from oracle.cpu import Mesh as OMesh

TOL = 1e-13  # sub-mesh oracle sums in local face order; the single-domain oracle in global order


def _fields(om, seed=0):
    rng = np.random.default_rng(seed)
    return rng.uniform(1, 2, om.nC), rng.uniform(1, 2, om.nB), rng.uniform(-1, 1, om.nF)


def test_simple_map_is_the_closed_form_rule():
    n = 12
    g = MeshDesc.block(n, n, n)
    for split in [(2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 2, 1)]:
        m = simple_map(g, split)
        i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        exp = np.zeros(n ** 3, dtype=np.int64)
        exp[(i + n * (j + n * k)).ravel()] = (i * split[0] // n + split[0] * (j * split[1] // n + split[1] * (k * split[2] // n))).ravel()
        assert np.array_equal(m, exp)


@pytest.mark.parametrize("P", [2, 4, 8])
def test_layout_and_halo_plan(P):
    g = MeshDesc.block(8, 6, 4, 0.8, 0.6, 0.4)
    decs = [Decomposition(g, P, r) for r in range(P)]
    owned = np.concatenate([d.cellGlobal[: d.nOwned] for d in decs])
    assert np.array_equal(np.sort(owned), np.arange(g.nCells))
    gown, gnei = g.array("faceOwner"), g.array("faceNeighbour")
    nIg = g.nInternalFaces
    seen_b = []
    for d in decs:
        L = d.desc
        o, n = L.array("faceOwner")[: L.nInternalFaces], L.array("faceNeighbour")
        fg = d.faceGlobal[: L.nInternalFaces]
        # local faces are the global faces, same orientation
        assert np.array_equal(d.cellGlobal[o], gown[fg]) and np.array_equal(d.cellGlobal[n], gnei[fg])
        # every face touches an owned cell; faces are sorted by local owner
        assert np.all((o < d.nOwned) | (n < d.nOwned)) and np.all(np.diff(o) >= 0)
        # owned cells ascending by global id; ghosts grouped by rank, ascending inside a group
        assert np.all(np.diff(d.cellGlobal[: d.nOwned]) > 0)
        for k in range(d.nNeighbours):
            gh = d.cellGlobal[d.nOwned + d.recvOff[k]: d.nOwned + d.recvOff[k + 1]]
            assert np.all(d.cellRank[gh] == d.nbrRanks[k]) and np.all(np.diff(gh) > 0)
            # the neighbour's send list for me, in global ids, is exactly my ghost list
            other = decs[d.nbrRanks[k]]
            kk = list(other.nbrRanks).index(d.rank)
            sent = other.cellGlobal[other.sendCells[other.sendOff[kk]: other.sendOff[kk + 1]]]
            assert np.array_equal(sent, gh)
        seen_b.append(d.faceGlobal[L.nInternalFaces:] - nIg)
        assert np.array_equal(L.array("faceCells"), o[:0].tolist() + list(np.searchsorted(d.cellGlobal[: d.nOwned], g.array("faceCells")[seen_b[-1]])))
    assert np.array_equal(np.sort(np.concatenate(seen_b)), np.arange(g.nBoundaryFaces))


@pytest.mark.parametrize("P", [2, 8])
def test_decomposed_oracle_matches_single_domain(P):
    g = MeshDesc.block(7, 6, 5, 0.7, 0.6, 0.5)
    om = OMesh.from_desc(g)
    phi, phib, flux = _fields(om)
    ref_div, ref_grad, ref_lap = om.div(flux, phi, phib, 0), om.grad(phi, phib), om.laplacian(phi, phib)
    x = np.random.default_rng(1).uniform(-1, 1, om.nC)
    ls = om.empty_system(False)
    bd = dict(value=phib, refValue=phib, valueFraction=np.ones(om.nB), refGrad=np.zeros(om.nB))
    om.laplacian_imp(ls, np.ones(om.nF), bd, -1.0, None)
    ref_Ax = om.spmv(ls["values"], x)
    for r in range(P):
        d = Decomposition(g, P, r)
        lm = OMesh.from_desc(d.desc)
        lphi, lflux = d.scatter_cells(phi), d.scatter_faces(flux)
        lphib = d.scatter_boundary(phib, om.nI)
        own = slice(0, d.nOwned)
        gid = d.cellGlobal[own]
        # geometry scheme recomputed on the sub-mesh equals the global one face by face
        assert np.array_equal(lm.w, om.w[d.faceGlobal]) and np.array_equal(lm.nodc, om.nodc[d.faceGlobal])
        for got, ref in ((lm.div(lflux, lphi, lphib, 0), ref_div), (lm.grad(lphi, lphib), ref_grad), (lm.laplacian(lphi, lphib), ref_lap)):
            assert np.allclose(got[own], ref[gid], rtol=TOL, atol=TOL * np.abs(ref).max())
        lls = lm.empty_system(False)
        lbd = {k: d.scatter_boundary(v, om.nI) for k, v in bd.items()}
        lm.laplacian_imp(lls, np.ones(lm.nF), lbd, -1.0, None)
        Ax = lm.spmv(lls["values"], d.scatter_cells(x))
        assert np.allclose(Ax[own], ref_Ax[gid], rtol=1e-12, atol=1e-12 * np.abs(ref_Ax).max())


def host_halo_exchange(dec: Decomposition, field: torch.Tensor):
    """The fvk_comm_halo_exchange protocol on CPU tensors over torch.distributed (gloo): pack the send cells per
    neighbour, post all receives straight into the ghost range, send, wait."""
    reqs = []
    for k, nb in enumerate(dec.nbrRanks):
        ghost = field[dec.nOwned + int(dec.recvOff[k]): dec.nOwned + int(dec.recvOff[k + 1])]
        reqs.append(dist.irecv(ghost, src=int(nb)))
    for k, nb in enumerate(dec.nbrRanks):
        cells = torch.from_numpy(dec.sendCells[dec.sendOff[k]: dec.sendOff[k + 1]].astype(np.int64))
        reqs.append(dist.isend(field[cells].contiguous(), dst=int(nb)))
    for r in reqs:
        r.wait()


def _gloo_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = MeshDesc.block(6, 5, 4, 0.6, 0.5, 0.4)
        om = OMesh.from_desc(g)
        phi, phib, flux = _fields(om, 3)
        d = Decomposition(g, world, rank)
        lm = OMesh.from_desc(d.desc)
        f = torch.zeros(d.nOwned + d.nGhost, dtype=torch.float64)
        f[: d.nOwned] = torch.from_numpy(phi[d.cellGlobal[: d.nOwned]])  # owned values only; ghosts arrive by halo exchange
        host_halo_exchange(d, f)
        assert np.array_equal(f.numpy(), phi[d.cellGlobal])
        got = lm.div(d.scatter_faces(flux), f.numpy(), d.scatter_boundary(phib, om.nI), 0)[: d.nOwned]
        # global reduction like fvk_comm_allreduce_sum
        s = torch.tensor([float(np.dot(got, got))], dtype=torch.float64)
        dist.all_reduce(s)
        ref = om.div(flux, phi, phib, 0)
        assert np.allclose(got, ref[d.cellGlobal[: d.nOwned]], rtol=TOL, atol=TOL * np.abs(ref).max())
        assert abs(s.item() - np.dot(ref, ref)) <= 1e-12 * np.dot(ref, ref)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2_halo_protocol():
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gloo_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret.get(0) and ret.get(1)


def test_external_cell_decomposition_file(tmp_path):
    """A cellDecomposition file as `decomposePar -cellDist` writes it (any method; here a deliberately irregular map)
    drives the same ghost-cell decomposition: every cell owned exactly once, halo plans consistent between ranks."""
    from foamadapter_b200.decomp import read_cell_decomposition
    g = MeshDesc.block(6, 5, 4)
    rng = np.random.default_rng(9)
    rank_of = (np.arange(g.nCells) * 3 // g.nCells).astype(np.int32)
    rank_of[rng.choice(g.nCells, 15, replace=False)] = rng.integers(0, 3, 15)      # ragged sub-domains
    f = tmp_path / "cellDecomposition"
    f.write_text("FoamFile\n{\n    version 2.0;\n    format ascii;\n    class labelList;\n    object cellDecomposition;\n}\n\n"
                 + f"{g.nCells}\n(\n" + "\n".join(map(str, rank_of)) + "\n)\n")
    m = read_cell_decomposition(f, g.nCells)
    assert np.array_equal(m, rank_of)
    decs = [Decomposition(g, 3, r, cellRank=m) for r in range(3)]
    owned = np.concatenate([d.cellGlobal[: d.nOwned] for d in decs])
    assert np.array_equal(np.sort(owned), np.arange(g.nCells))
    for d in decs:
        for k, nb in enumerate(d.nbrRanks):
            o = decs[nb]
            kk = list(o.nbrRanks).index(d.rank)
            sent = o.cellGlobal[o.sendCells[o.sendOff[kk]: o.sendOff[kk + 1]]]
            assert np.array_equal(sent, d.cellGlobal[d.nOwned + d.recvOff[k]: d.nOwned + d.recvOff[k + 1]])
    from foamadapter_b200._capi import FvkError
    with pytest.raises(FvkError):
        read_cell_decomposition(f, g.nCells + 1)
