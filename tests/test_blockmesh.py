"""The product's block-hex mesh generator (stand-in for blockMesh + readOpenFOAMMesh) against the
polyMesh fixtures the reference commits: topology/ordering bit-exact, geometry to rounding. CPU only."""
import numpy as np
import pytest

from foamadapter_b200.mesh import MeshDesc
from tests.helpers import FIXTURE_BLOCKS, load_golden, neon_view


@pytest.mark.parametrize("case", sorted(FIXTURE_BLOCKS))
def test_topology_matches_fixture(case):
    dims, box, patches = FIXTURE_BLOCKS[case]
    g = load_golden(case)
    d = MeshDesc.block(*dims, *box, patches=patches, with_points=True)
    fp, po = d.poly()
    assert np.array_equal(fp, g["faces"])          # polyMesh/faces (point labels, order, orientation)
    assert np.array_equal(po, g["owner"])          # polyMesh/owner
    assert np.array_equal(d.array("faceNeighbour"), g["neighbour"])
    np.testing.assert_allclose(d.array("points"), g["points"], rtol=0, atol=4e-16)


@pytest.mark.parametrize("case", sorted(FIXTURE_BLOCKS))
def test_neon_view_matches_fixture_geometry(case):
    dims, box, patches = FIXTURE_BLOCKS[case]
    v = neon_view(load_golden(case))  # oracle numpy geometry on the fixture's own points/faces
    d = MeshDesc.block(*dims, *box, patches=patches)
    for k in ("nCells", "nInternalFaces", "nBoundaryFaces", "nPatches"):
        assert getattr(d, k) == v[k], k
    for k in ("faceOwner", "faceNeighbour", "faceCells", "patchOffsets"):
        assert np.array_equal(d.array(k), v[k]), k
    for k in ("cellVolumes", "cellCentres", "faceAreas", "faceCentres", "magFaceAreas", "bCf", "bCn", "bSf",
              "bMagSf", "bNf", "bDelta", "bWeights", "bDeltaCoeffs"):
        a, b = d.array(k), np.asarray(v[k])
        scale = np.abs(b).max()
        assert np.abs(a.reshape(b.shape) - b).max() <= 1e-12 * scale, k  # fixture points carry ~3e-16 absolute blockMesh noise


def test_empty_patch_faces_are_dropped():
    # the 5x5x1 fixture has 110 poly faces but NeoN sees nI=40, nB=20
    dims, box, patches = FIXTURE_BLOCKS["setup_operator"]
    d = MeshDesc.block(*dims, *box, patches=patches)
    assert (d.nCells, d.nInternalFaces, d.nBoundaryFaces, d.nPatches) == (25, 40, 20, 1)


def test_sizes_of_benchmark_meshes():
    # BASELINE.md §3 (closed forms; 128^3 generated for real)
    d = MeshDesc.block(128, 128, 128, 0.1, 0.1, 0.01)
    assert (d.nCells, d.nInternalFaces, d.nBoundaryFaces) == (2097152, 6242304, 98304)
    own, nei = d.array("faceOwner")[: d.nInternalFaces], d.array("faceNeighbour")
    assert np.all(own[1:] >= own[:-1]) and np.all(nei > own)  # upper-triangular order
    np.testing.assert_allclose(d.array("cellVolumes").sum(), 0.1 * 0.1 * 0.01, rtol=1e-12)


def test_rejects_bad_patch_spec():
    from foamadapter_b200._capi import FvkError
    with pytest.raises(FvkError):
        MeshDesc.block(2, 2, 2, patches=[("a", [0, 1, 2, 3, 4], False)])  # side 5 missing
    with pytest.raises(FvkError):
        MeshDesc.block(2, 2, 2, patches=[("a", [0, 0, 1, 2, 3, 4, 5], False)])  # duplicate


# ---- the oracle's own generator (oracle/blockmesh.cpp) vs the product's host generator: bit for bit ---------------------
import pytest  # noqa: E402

from foamadapter_b200.mesh import PATCHES_3DCUBE, PATCHES_CAVITY2D, PATCHES_CAVITY3D  # noqa: E402


@pytest.mark.parametrize("dims,box,patches", [((5, 5, 1), (0.1, 0.1, 0.01), PATCHES_CAVITY2D), ((7, 5, 3), (0.1, 0.1, 0.01), PATCHES_3DCUBE),
                                              ((4, 6, 5), (1.0, 2.0, 3.0), PATCHES_CAVITY3D), ((1, 1, 1), (1.0, 1.0, 1.0), PATCHES_3DCUBE),
                                              ((3, 1, 2), (1.0, 1.0, 1.0), PATCHES_3DCUBE), ((33, 17, 9), (0.3, 0.2, 0.1), PATCHES_3DCUBE)])
def test_oracle_generator_equals_product_generator(dims, box, patches):
    import numpy as np
    from foamadapter_b200.mesh import MeshDesc
    from oracle.cpu import Mesh
    a = Mesh.from_desc(MeshDesc.block(*dims, *box, patches=patches))
    b = Mesh.block(*dims, *box, patches=[(n, list(s), e) for n, s, e in patches])
    for k in ("owner", "neighbour", "faceCells", "V", "C", "Sf", "Cf", "magSf", "patchOffsets", "bSf", "bDeltaCoeffs", "bWeights", "w", "dc", "nodc",
              "rowOffs", "colIdxs", "ownerOffset", "neighbourOffset", "diagOffset"):
        x, y = np.ravel(getattr(a, k)), np.ravel(getattr(b, k))
        assert x.shape == y.shape and np.array_equal(x, y), k
