"""CPU: fvk_renumber_* (reverse Cuthill-McKee / Morton cell order, faces re-sorted upper-triangular) -- the renumbered mesh is
the same mesh: operators evaluated by the oracle on it agree with the original mesh's through the maps (to summation order),
the orderings do what they are for (bandwidth / locality), and the brick plan tiles a Morton-ordered mesh compactly."""
import numpy as np
import pytest

from foamadapter_b200.mesh import MeshDesc
from oracle.cpu import Mesh as OMesh
from tests.helpers import renumbered_block


def bandwidth(d):
    nI = d.nInternalFaces
    return int(np.abs(d.array("faceNeighbour").astype(np.int64) - d.array("faceOwner")[:nI]).max())


@pytest.mark.parametrize("method", ["rcm", "morton"])
def test_renumbered_mesh_is_the_same_mesh(method):
    g = renumbered_block(12, 9, 7, seed=3, box=(1.2, 0.9, 0.7))       # a block mesh in a random cell order
    r, cmap, fmap, flip = g.renumbered(method)
    nC, nI, nB = g.nCells, g.nInternalFaces, g.nBoundaryFaces
    assert sorted(cmap) == list(range(nC)) and sorted(fmap) == list(range(nI + nB))
    own, nei = r.array("faceOwner"), r.array("faceNeighbour")
    assert np.all(own[:nI] < nei) and np.all(np.diff(own[:nI].astype(np.int64) * nC + nei) > 0)   # upper-triangular order
    assert np.array_equal(r.array("faceCells"), cmap[g.array("faceCells")]) and np.array_equal(fmap[nI:], np.arange(nI, nI + nB))
    assert np.array_equal(r.array("cellVolumes")[cmap], g.array("cellVolumes"))
    sgn = np.where(flip, -1.0, 1.0)
    assert np.array_equal(r.array("faceAreas")[fmap], g.array("faceAreas") * sgn[:, None])
    # the same operator on both meshes
    a, b = OMesh.from_desc(g), OMesh.from_desc(r)
    rng = np.random.default_rng(0)
    phi, phib, flux = rng.uniform(1, 2, nC), rng.uniform(1, 2, nB), rng.uniform(-1, 1, nI + nB)
    phi_r, flux_r = np.empty(nC), np.empty(nI + nB)
    phi_r[cmap] = phi
    flux_r[fmap] = flux * sgn
    for scheme in (0, 1):
        ref = a.div(flux, phi, phib, scheme)
        got = b.div(flux_r, phi_r, phib, scheme)
        assert np.allclose(got[cmap], ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    assert np.allclose(b.grad(phi_r, phib)[cmap], a.grad(phi, phib), rtol=1e-12, atol=1e-10)
    assert np.allclose(b.laplacian(phi_r, phib)[cmap], a.laplacian(phi, phib), rtol=1e-12, atol=1e-9)


def test_rcm_reduces_the_bandwidth_and_morton_restores_locality():
    g = renumbered_block(16, 16, 16, seed=1)
    assert bandwidth(g) > 3000                                  # random numbering: neighbours anywhere
    rcm, *_ = g.renumbered("rcm")
    assert bandwidth(rcm) <= 2 * 16 * 16 + 16                   # a few planes of the block
    mor, cmap, *_ = g.renumbered("morton")
    C = mor.array("cellCentres")
    # 64 consecutive cells of the Morton order form a 4x4x4 brick: their centres span 4 cells per axis
    for t in (0, 7, 40):
        blk = C[64 * t: 64 * (t + 1)]
        assert np.allclose(blk.max(0) - blk.min(0), 3.0 / 16.0)


def test_renumbering_rejects_bad_input():
    from foamadapter_b200._capi import FvkError
    g = MeshDesc.block(4, 3, 2)
    with pytest.raises(FvkError):
        g.renumbered(cellOldToNew=np.zeros(g.nCells, dtype=np.int32))   # not a permutation
    same, cmap, fmap, flip = g.renumbered(cellOldToNew=np.arange(g.nCells, dtype=np.int32))
    for k in ("faceOwner", "faceNeighbour", "faceAreas", "cellCentres"):
        assert np.array_equal(same.array(k), g.array(k))
    assert not flip.any() and np.array_equal(fmap, np.arange(g.nFaces))
