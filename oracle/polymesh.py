"""TEST INFRASTRUCTURE (oracle) -- not product code.

Minimal reader for OpenFOAM ASCII ``polyMesh`` directories and ASCII field files, plus a numpy
restatement of OpenFOAM's ``primitiveMesh`` geometry (face centres/areas by triangle fan about the
vertex average, cell centres/volumes by face pyramids about the face-centre average).

Used only by ``tests/`` and ``tests/golden/make_golden.py`` to turn the fixtures committed in the
reference (``test/setup_*/constant/polyMesh/*``, ``test/setup_operator/0/*``) into small ``.npz``
files, and to cross-check the product's C++ mesh generator.  The boundary flattening follows
``src/datastructures/meshAdapter.cpp:12-45,59-136`` of the reference (patches concatenated in patch
order, ``empty`` patches contribute zero faces because ``fvPatch::size()==0``).
"""
from __future__ import annotations

import re
from pathlib import Path

import numpy as np

_COMMENT_BLOCK = re.compile(r"/\*.*?\*/", re.S)
_COMMENT_LINE = re.compile(r"//.*?$", re.M)


def _strip(text: str) -> str:
    text = _COMMENT_BLOCK.sub("", text)
    text = _COMMENT_LINE.sub("", text)
    # drop the FoamFile header dictionary
    m = re.search(r"FoamFile\s*\{.*?\}", text, re.S)
    if m:
        text = text[: m.start()] + text[m.end():]
    return text


def _list_body(text: str) -> tuple[int, str]:
    """Return (n, body) of the first ``n ( ... )`` list in ``text``."""
    m = re.search(r"(\d+)\s*\(", text)
    if not m:
        raise ValueError("no list found")
    n = int(m.group(1))
    depth, i = 1, m.end()
    start = i
    while depth:
        ch = text[i]
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        i += 1
    return n, text[start: i - 1]


def read_label_list(path) -> np.ndarray:
    n, body = _list_body(_strip(Path(path).read_text()))
    a = np.array(body.split(), dtype=np.int32)
    assert a.size == n, (path, a.size, n)
    return a


def read_points(path) -> np.ndarray:
    n, body = _list_body(_strip(Path(path).read_text()))
    a = np.array(body.replace("(", " ").replace(")", " ").split(), dtype=np.float64)
    assert a.size == 3 * n
    return a.reshape(n, 3)


def read_faces(path) -> list[np.ndarray]:
    n, body = _list_body(_strip(Path(path).read_text()))
    faces = [np.array(m.group(2).split(), dtype=np.int32)
             for m in re.finditer(r"(\d+)\s*\(([^()]*)\)", body)]
    assert len(faces) == n
    return faces


def read_boundary(path) -> list[dict]:
    text = _strip(Path(path).read_text())
    _, body = _list_body(text)
    patches = []
    for m in re.finditer(r"(\w+)\s*\{(.*?)\}", body, re.S):
        d = {"name": m.group(1)}
        for key in ("type", "nFaces", "startFace"):
            mm = re.search(rf"\b{key}\s+([^;]+);", m.group(2))
            d[key] = mm.group(1).strip()
        d["nFaces"] = int(d["nFaces"])
        d["startFace"] = int(d["startFace"])
        patches.append(d)
    return patches


def read_internal_field(path) -> np.ndarray:
    """``internalField nonuniform List<scalar|vector> n ( ... )`` of an ASCII field file."""
    text = _strip(Path(path).read_text())
    m = re.search(r"internalField\s+nonuniform\s+List<(\w+)>", text)
    if not m:
        raise ValueError(f"{path}: no nonuniform internalField")
    kind = m.group(1)
    n, body = _list_body(text[m.end():])
    a = np.array(body.replace("(", " ").replace(")", " ").split(), dtype=np.float64)
    if kind == "vector":
        return a.reshape(n, 3)
    assert a.size == n
    return a


# --------------------------------------------------------------------------------------------
# geometry: OpenFOAM primitiveMeshFaceCentresAndAreas.C / primitiveMeshCellCentresAndVols.C
# --------------------------------------------------------------------------------------------

def face_geometry(points: np.ndarray, faces: list[np.ndarray]):
    nF = len(faces)
    Cf = np.zeros((nF, 3))
    Sf = np.zeros((nF, 3))
    for fi, f in enumerate(faces):
        p = points[f]
        n = len(f)
        if n == 3:
            Cf[fi] = (1.0 / 3.0) * (p[0] + p[1] + p[2])
            Sf[fi] = 0.5 * np.cross(p[1] - p[0], p[2] - p[0])
            continue
        fc = p[0].copy()
        for k in range(1, n):
            fc += p[k]
        fc /= n
        sumN = np.zeros(3)
        sumA = 0.0
        sumAc = np.zeros(3)
        for k in range(n):
            nxt = p[(k + 1) % n]
            cur = p[k]
            c = cur + nxt + fc
            nn = np.cross(nxt - cur, fc - cur)
            a = np.sqrt(nn @ nn)
            sumN += nn
            sumA += a
            sumAc += a * c
        if sumA < 1e-150:
            Cf[fi] = fc
        else:
            Cf[fi] = (1.0 / 3.0) * sumAc / sumA
            Sf[fi] = 0.5 * sumN
    return Cf, Sf


def cell_geometry(Cf, Sf, owner, neighbour, nCells):
    nI = len(neighbour)
    cEst = np.zeros((nCells, 3))
    cnt = np.zeros(nCells, dtype=np.int64)
    for f in range(len(owner)):
        cEst[owner[f]] += Cf[f]
        cnt[owner[f]] += 1
    for f in range(nI):
        cEst[neighbour[f]] += Cf[f]
        cnt[neighbour[f]] += 1
    cEst /= cnt[:, None]
    C = np.zeros((nCells, 3))
    V = np.zeros(nCells)
    for f in range(len(owner)):
        o = owner[f]
        pyr3 = Sf[f] @ (Cf[f] - cEst[o])
        pc = (3.0 / 4.0) * Cf[f] + (1.0 / 4.0) * cEst[o]
        C[o] += pyr3 * pc
        V[o] += pyr3
    for f in range(nI):
        n = neighbour[f]
        pyr3 = Sf[f] @ (cEst[n] - Cf[f])
        pc = (3.0 / 4.0) * Cf[f] + (1.0 / 4.0) * cEst[n]
        C[n] += pyr3 * pc
        V[n] += pyr3
    C /= V[:, None]
    V *= 1.0 / 3.0
    return C, V


def load_polymesh(case_dir) -> dict:
    """Read ``<case>/constant/polyMesh`` and build the NeoN view of it.

    Returns the arrays ``readOpenFOAMMesh`` would hand to ``NeoN::UnstructuredMesh``
    (meshAdapter.cpp:59-136): ``empty`` patches are dropped from the boundary arrays.
    """
    pm = Path(case_dir) / "constant" / "polyMesh"
    points = read_points(pm / "points")
    faces = read_faces(pm / "faces")
    owner = read_label_list(pm / "owner")
    neighbour = read_label_list(pm / "neighbour")
    patches = read_boundary(pm / "boundary")
    nCells = int(owner.max()) + 1
    nI = len(neighbour)
    Cf, Sf = face_geometry(points, faces)
    C, V = cell_geometry(Cf, Sf, owner, neighbour, nCells)
    keep = [p for p in patches if p["type"] != "empty"]
    bfaces = np.concatenate([np.arange(p["startFace"], p["startFace"] + p["nFaces"])
                             for p in keep]) if keep else np.zeros(0, dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum([p["nFaces"] for p in keep])]).astype(np.int32)
    nB = len(bfaces)
    faces_arr = np.full((len(faces), 4), -1, dtype=np.int32)
    for i, f in enumerate(faces):
        faces_arr[i, : len(f)] = f[:4]
    return dict(
        points=points, faces=faces_arr, owner_all=owner, neighbour=neighbour,
        nCells=nCells, nI=nI, nB=nB, nFacesPoly=len(faces),
        patch_names=[p["name"] for p in keep], patch_offsets=offsets,
        all_patch_names=[p["name"] for p in patches],
        all_patch_types=[p["type"] for p in patches],
        all_patch_start=np.array([p["startFace"] for p in patches], dtype=np.int32),
        all_patch_size=np.array([p["nFaces"] for p in patches], dtype=np.int32),
        bfaces=bfaces.astype(np.int32),
        C=C, V=V, Cf=Cf, Sf=Sf, magSf=np.sqrt((Sf * Sf).sum(1)),
        faceCells=owner[bfaces].astype(np.int32),
    )
