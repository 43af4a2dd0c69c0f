"""CPU tests of the brick plan (host code of libfvk, no GPU): the plan of the default explicit-operator kernel is
replayed on the host exactly as k_gather_brick reads it and must reproduce, cell by cell, the (face, sign) sequence of
the reference's Serial accumulation order (gaussGreenDiv.cpp:46-67) -- on block meshes of awkward sizes, on every
sub-mesh of a decomposition (ghost cells, faceOrder key) and on a randomly renumbered mesh (no block structure)."""
import ctypes as C

import numpy as np
import pytest

from foamadapter_b200._capi import FvkError, check, lib
from foamadapter_b200.decomp import Decomposition
from foamadapter_b200.mesh import PATCHES_CAVITY2D, MeshDesc
from tests.helpers import renumbered_block


def selftest(desc):
    info = (C.c_int32 * 8)()
    bad = C.c_int64(-1)
    check(lib().fvk_brick_plan_selftest(C.byref(desc.c), info, C.byref(bad)))
    return list(info), bad.value


@pytest.mark.parametrize("dims", [(5, 5, 1), (3, 3, 3), (20, 20, 1), (33, 7, 5), (100, 3, 2), (1, 1, 1), (7, 1, 1),
                                  (1, 9, 4), (64, 48, 40), (40, 12, 9)])
def test_block_meshes(dims):
    d = MeshDesc.block(*dims)
    info, bad = selftest(d)
    assert bad == 0
    nT, nx, ny, nz, lx, by, bz, slots = info
    assert nx * ny * nz == d.nCells  # numbering detected (degenerate axes may be merged)
    assert lx * by * bz <= 128 and nT >= -(-d.nCells // 128)  # default: one cell per thread of a 128-thread block
    assert slots < 32768


def test_cavity_patches_with_empty_faces():
    info, bad = selftest(MeshDesc.block(20, 20, 1, patches=PATCHES_CAVITY2D))
    assert bad == 0


def test_brick_override(monkeypatch):
    monkeypatch.setenv("FVK_BRICK", "16,8,4")
    info, bad = selftest(MeshDesc.block(64, 32, 16))
    assert bad == 0 and info[4:7] == [16, 8, 4] and info[0] == 4 * 4 * 4
    monkeypatch.setenv("FVK_BRICK", "8,3,5")  # non power-of-two run/row counts: the kernel's division path
    info, bad = selftest(MeshDesc.block(20, 10, 11))
    assert bad == 0 and info[4:7] == [8, 3, 5]


@pytest.mark.parametrize("P", [2, 4, 8])
def test_decomposed_sub_meshes(P):
    g = MeshDesc.block(16, 12, 10)
    for r in range(P):
        dec = Decomposition(g, P, r)
        info, bad = selftest(dec.desc)
        assert bad == 0
        assert info[1] * info[2] * info[3] == dec.nOwned  # the owned block is still recognised


def test_unstructured_numbering_uses_runs_of_consecutive_cells():
    d = renumbered_block(12, 11, 10, 3)
    info, bad = selftest(d)
    assert bad == 0
    assert info[1:4] == [0, 0, 0] and info[0] == -(-d.nCells // 128)


def test_unsorted_faces_get_no_plan():
    g = MeshDesc.block(6, 5, 4)
    a = {k: g.array(k) for k in ("cellVolumes", "cellCentres", "faceAreas", "faceCentres", "magFaceAreas", "faceOwner",
                                 "faceNeighbour", "faceCells", "patchOffsets")}
    a.update(nCells=g.nCells, nInternalFaces=g.nInternalFaces, nBoundaryFaces=g.nBoundaryFaces, nPatches=g.nPatches)
    nI = g.nInternalFaces
    o, n = a["faceOwner"].copy(), a["faceNeighbour"].copy()
    o[:nI], n[:] = o[:nI][::-1].copy(), n[::-1].copy()  # faces in descending owner order
    a["faceOwner"], a["faceNeighbour"] = o, n
    with pytest.raises(FvkError) as e:
        selftest(MeshDesc.from_arrays(a))
    assert e.value.code == 5  # FVK_EUNSUPPORTED: operators keep the per-cell gather


def affine_info(desc):
    info = (C.c_int32 * 5)()
    check(lib().fvk_brick_plan_affine_info(C.byref(desc.c), info))
    return list(info)


def test_affine_topology_is_proven_only_where_it_holds():
    """k_gather_affine reads no index array: the plan must prove the closed-form topology cell by cell."""
    a = affine_info(MeshDesc.block(64, 48, 40))
    assert a[:4] == [1, 1, 1, 1] and a[4] == 64 * 48 * 40 - 62 * 46 * 38      # irregular = the outermost cell layer
    assert affine_info(MeshDesc.block(33, 7, 5))[0] == 1                      # ragged edge tiles are fine
    assert affine_info(MeshDesc.block(20, 20, 1, patches=PATCHES_CAVITY2D))[0] == 0   # 2-D: per-cell order differs
    assert affine_info(renumbered_block(12, 11, 10, 3))[0] == 0                # no block structure
    g = MeshDesc.block(64, 48, 40)
    lo, hi = affine_info(Decomposition(g, 2, 0).desc), affine_info(Decomposition(g, 2, 1).desc)
    assert lo[:4] == [1, 0, 1, 1]    # rank 0: its +x side is a processor cut (cells there own a face to a ghost)
    assert hi[:4] == [1, 1, 1, 1]    # rank 1: ghosts only on the lower side
    for P in (4, 8):
        for r in range(P):
            assert affine_info(Decomposition(MeshDesc.block(16, 12, 10), P, r).desc)[0] == 1


def test_random_blocks_and_their_sub_domains():
    """Randomised sweep (fixed seed): the plan replays the reference order on every block mesh and every sub-domain of a
    2/4/8-way decomposition; the affine topology is proven exactly when every extent is >= 3, and its irregular-cell list
    is the outermost layer."""
    rng = np.random.default_rng(0)
    for _ in range(12):
        dims = tuple(int(x) for x in rng.integers(1, 28, 3))
        g = MeshDesc.block(*dims)
        info, bad = selftest(g)
        assert bad == 0, dims
        a = affine_info(g)
        assert a[0] == (1 if min(dims) >= 3 else 0), (dims, a)
        if a[0]:
            assert a[4] == g.nCells - (dims[0] - 2) * (dims[1] - 2) * (dims[2] - 2)
        P = int(rng.choice([2, 4, 8]))
        if g.nCells >= 2 * P:
            for r in range(P):
                d = Decomposition(g, P, r)
                if d.nOwned:
                    assert selftest(d.desc)[1] == 0, (dims, P, r)


def test_structured_spmv_assumption_holds_row_by_row():
    """The structured SpMV computes the columns of a regular row instead of reading them; this replays the
    SparsityPattern on the host and checks every regular row of block meshes and of decomposition sub-domains."""
    def rows(desc):
        r = (C.c_int64 * 2)()
        check(lib().fvk_brick_plan_structured_rows(C.byref(desc.c), r))
        return list(r)
    g = MeshDesc.block(13, 9, 7)
    assert rows(g) == [1, 11 * 7 * 5]
    assert rows(MeshDesc.block(20, 20, 1, patches=PATCHES_CAVITY2D))[0] == 0      # not affine: generic SpMV
    assert rows(renumbered_block(12, 11, 10, 3))[0] == 0
    for P in (2, 4, 8):
        for r in range(P):
            ok, n = rows(Decomposition(MeshDesc.block(16, 12, 10), P, r).desc)
            assert ok == 1 and n > 0


def test_sparsity_pattern_host_builder_and_sub_domain_row_order():
    """fvk_sparsity_selftest runs the host function fvk_mesh_create uses for the SparsityPattern: on single-domain meshes its columns
    equal the reference's three serial face passes (sparsityPattern.cpp:21-143), offsets and diagonal positions are consistent, and
    -- the round-2 finding -- the rows of EVERY sub-domain of a decomposed block are in stencil order (each half of a row follows
    the global face order), which is what keeps the index-free assembly / rAU,HbyA kernels on all ranks."""
    def sp(desc):
        r = (C.c_int64 * 4)()
        check(lib().fvk_sparsity_selftest(C.byref(desc.c), r))
        return list(r)
    for dims in [(13, 9, 7), (20, 20, 1), (5, 1, 1), (3, 3, 3)]:
        g = MeshDesc.block(*dims, patches=PATCHES_CAVITY2D) if dims[2] == 1 and dims[1] > 1 else MeshDesc.block(*dims)
        rows, nnz, ordered, bad = sp(g)
        assert rows == g.nCells and nnz == g.nCells + 2 * g.nInternalFaces and ordered == 1 and bad == 0, dims
    rows, nnz, ordered, bad = sp(renumbered_block(12, 11, 10, 3))      # random cell order: still consistent, rows not in stencil order
    assert bad == 0
    g = MeshDesc.block(16, 12, 10)
    for P in (2, 4, 8):
        for r in range(P):
            d = Decomposition(g, P, r)
            rows, nnz, ordered, bad = sp(d.desc)
            assert ordered == 1 and bad == 0, (P, r)
            assert rows == d.desc.nCells and nnz == d.desc.nCells + 2 * d.desc.nInternalFaces

