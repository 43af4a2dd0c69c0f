"""Shared test helpers: golden fixtures -> NeoN mesh view (via the oracle's numpy geometry)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from oracle import polymesh
from oracle.cpu import Mesh as OracleMesh

GOLDEN = Path(__file__).resolve().parent / "golden"

X0, X1, Y0, Y1, Z0, Z1 = range(6)
W2D = [("fixedWalls", [Y1, X0, X1, Y0], False), ("frontAndBack", [Z0, Z1], True)]
W3D = [("fixedWalls", [Y1, Y0, Z0, Z1], False), ("inlet", [X1], False), ("outlet", [X0], False)]
# (dims, box incl. scale, patches) of each reference fixture case (system/blockMeshDict)
FIXTURE_BLOCKS = {
    "setup_operator": ((5, 5, 1), (0.1, 0.1, 0.01), W2D),
    "setup_stencil3D": ((3, 3, 3), (1.0, 1.0, 1.0), W3D),
    "setup_pressureVelocityCoupling": ((3, 3, 3), (1.0, 1.0, 1.0), W3D),
    "setup_advection": ((50, 50, 1), (1.0, 1.0, 0.1), W2D),
    "setup_unstructuredMesh": ((3, 3, 1), (1.0, 1.0, 0.1), W2D),
    "setup_compatibility": ((3, 3, 1), (1.0, 1.0, 0.1), W2D),
}


def load_golden(name):
    return np.load(GOLDEN / f"{name}.npz")


def neon_view(g) -> dict:
    """What readOpenFOAMMesh (reference src/datastructures/meshAdapter.cpp:59-136) would produce
    from the polyMesh stored in golden file `g`; geometry by the oracle's numpy primitiveMesh."""
    points, faces, owner, nei = g["points"], g["faces"], g["owner"], g["neighbour"]
    nC, nI = int(owner.max()) + 1, len(nei)
    Cf, Sf = polymesh.face_geometry(points, list(faces))
    C, V = polymesh.cell_geometry(Cf, Sf, owner, nei, nC)
    keep = [i for i, t in enumerate(g["patch_types"]) if t != "empty"]
    bfaces = np.concatenate([np.arange(g["patch_start"][i], g["patch_start"][i] + g["patch_size"][i]) for i in keep])
    sel = np.concatenate([np.arange(nI), bfaces])
    nB = len(bfaces)
    fc = owner[bfaces].astype(np.int32)
    magSf = np.sqrt((Sf * Sf).sum(1))
    delta = Cf[bfaces] - C[fc]
    return dict(
        nCells=nC, nInternalFaces=nI, nBoundaryFaces=nB, nPatches=len(keep), nPoints=len(points), points=points,
        cellVolumes=V, cellCentres=C, faceAreas=Sf[sel], faceCentres=Cf[sel], magFaceAreas=magSf[sel],
        faceOwner=owner[sel].astype(np.int32), faceNeighbour=nei.astype(np.int32), faceCells=fc,
        bCf=Cf[bfaces], bCn=C[fc], bSf=Sf[bfaces], bMagSf=magSf[bfaces], bNf=Sf[bfaces] / magSf[bfaces, None],
        bDelta=delta, bWeights=np.ones(nB), bDeltaCoeffs=1.0 / np.sqrt((delta * delta).sum(1)),
        patchOffsets=np.concatenate([[0], np.cumsum(g["patch_size"][keep])]).astype(np.int32),
        patch_names=[str(g["patch_names"][i]) for i in keep],
    )


def oracle_mesh_from_view(v) -> OracleMesh:
    return OracleMesh(nCells=v["nCells"], owner=v["faceOwner"], neighbour=v["faceNeighbour"], faceCells=v["faceCells"],
                      V=v["cellVolumes"], C=v["cellCentres"], Sf=v["faceAreas"], Cf=v["faceCentres"],
                      magSf=v["magFaceAreas"], patchOffsets=v["patchOffsets"], bSf=v["bSf"],
                      bDeltaCoeffs=v["bDeltaCoeffs"], bWeights=v["bWeights"])


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / scale)


def renumbered_block(nx, ny, nz, seed, box=(1.0, 1.0, 1.0)):
    """A block mesh whose cells were renumbered at random, faces re-sorted into OpenFOAM's upper-triangular order
    (owner < neighbour, by owner then neighbour; a flipped face gets -Sf). Same geometry, no block structure in the
    numbering: exercises the unstructured path of every kernel."""
    from foamadapter_b200.mesh import MeshDesc
    g = MeshDesc.block(nx, ny, nz, *box)
    nC, nI, nB = g.nCells, g.nInternalFaces, g.nBoundaryFaces
    rng = np.random.default_rng(seed)
    perm = rng.permutation(nC).astype(np.int32)  # old cell -> new cell
    inv = np.argsort(perm)
    o, n = perm[g.array("faceOwner")[:nI]], perm[g.array("faceNeighbour")]
    flip = o > n
    lo, hi = np.minimum(o, n), np.maximum(o, n)
    order = np.lexsort((hi, lo))
    fsel = np.concatenate([order, nI + np.arange(nB)])
    Sf = g.array("faceAreas").reshape(-1, 3).copy()
    Sf[:nI][flip] *= -1.0
    fc = perm[g.array("faceCells")]
    a = dict(nCells=nC, nInternalFaces=nI, nBoundaryFaces=nB, nPatches=g.nPatches,
             cellVolumes=g.array("cellVolumes")[inv], cellCentres=g.array("cellCentres").reshape(-1, 3)[inv],
             faceAreas=Sf[fsel], faceCentres=g.array("faceCentres").reshape(-1, 3)[fsel],
             magFaceAreas=g.array("magFaceAreas")[fsel],
             faceOwner=np.concatenate([lo[order], fc]).astype(np.int32), faceNeighbour=hi[order].astype(np.int32),
             faceCells=fc.astype(np.int32), patchOffsets=g.array("patchOffsets").copy())
    for k in ("bCf", "bCn", "bSf", "bMagSf", "bNf", "bDelta", "bWeights", "bDeltaCoeffs"):
        a[k] = g.array(k).copy()  # g owns the storage behind array(); it dies with this frame
    return MeshDesc.from_arrays(a, g.patch_names)
