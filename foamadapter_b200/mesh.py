"""Host mesh descriptions and the device mesh handle (Python side of include/fvk.h).

`MeshDesc`  = the arrays FoamAdapter::readOpenFOAMMesh hands to NeoN::UnstructuredMesh
              (reference src/datastructures/meshAdapter.cpp:59-136), on the host.
`UnstructuredMesh` = the device-resident mesh + geometry scheme + stencil + sparsity pattern
              (NeoN::UnstructuredMesh, BasicGeometryScheme, CellToFaceStencil, la::SparsityPattern).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import MeshDesc as _CDesc, check, lib

# block sides
XMIN, XMAX, YMIN, YMAX, ZMIN, ZMAX = range(6)

# fvk_mesh_field
N_CELLS, N_INTERNAL_FACES, N_BOUNDARY_FACES, N_PATCHES, NNZ, N_OWNED_CELLS, ROWS_IN_STENCIL_ORDER, AFFINE_TOPOLOGY = range(8)
(CELL_VOLUMES, CELL_CENTRES, FACE_AREAS, FACE_CENTRES, MAG_FACE_AREAS, FACE_OWNER, FACE_NEIGHBOUR,
 FACE_CELLS, B_CF, B_CN, B_SF, B_MAGSF, B_NF, B_DELTA, B_WEIGHTS, B_DELTACOEFFS) = range(16, 32)
WEIGHTS, DELTACOEFFS, NONORTH_DELTACOEFFS = 48, 49, 50
STENCIL_SEGMENTS, STENCIL_VALUES = 64, 65
ROW_OFFS, COL_IDXS, OWNER_OFFSET, NEIGHBOUR_OFFSET, DIAG_OFFSET = range(80, 85)

_DTYPES = {FACE_OWNER: np.int32, FACE_NEIGHBOUR: np.int32, FACE_CELLS: np.int32,
           STENCIL_SEGMENTS: np.int32, STENCIL_VALUES: np.int32, ROW_OFFS: np.int32,
           COL_IDXS: np.int32, OWNER_OFFSET: np.uint8, NEIGHBOUR_OFFSET: np.uint8,
           DIAG_OFFSET: np.uint8}

_F64 = ["points", "cellVolumes", "cellCentres", "faceAreas", "faceCentres", "magFaceAreas", "bCf",
        "bCn", "bSf", "bMagSf", "bNf", "bDelta", "bWeights", "bDeltaCoeffs"]
_I32 = ["faceOwner", "faceNeighbour", "faceCells", "patchOffsets"]

# patch layouts of the reference's cases: (name, [sides], isEmpty)
PATCHES_3DCUBE = [("top", [YMAX], False), ("bottom", [YMIN], False),
                  ("sides", [XMIN, XMAX, ZMIN, ZMAX], False)]
# benchmarks/benchmarkSuite/templates/3DCube/system/blockMeshDict:42-70
PATCHES_CAVITY2D = [("movingWall", [YMAX], False), ("fixedWalls", [XMIN, XMAX, YMIN], False),
                    ("frontAndBack", [ZMIN, ZMAX], True)]
# tutorials/cavity/system/blockMeshDict
PATCHES_CAVITY3D = [("movingWall", [YMAX], False), ("fixedWalls", [XMIN, XMAX, YMIN, ZMIN, ZMAX], False)]


class MeshDesc:
    """Host-side mesh description; owns (or borrows) the arrays behind a C `fvk_mesh_desc`."""

    def __init__(self):
        self._c = None          # ctypes struct (by value) or pointer owned by the library
        self._owned_ptr = None  # fvk_mesh_desc* from fvk_blockmesh_create
        self._keep = {}
        self.patch_names: list[str] = []

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def block(cls, nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, patches=PATCHES_3DCUBE, with_points=False):
        self = cls()
        nsides = (C.c_int32 * len(patches))(*[len(p[1]) for p in patches])
        flat = [s for p in patches for s in p[1]]
        sides = (C.c_int32 * len(flat))(*flat)
        empty = (C.c_int32 * len(patches))(*[int(p[2]) for p in patches])
        out = C.POINTER(_CDesc)()
        check(lib().fvk_blockmesh_create(C.c_int32(nx), C.c_int32(ny), C.c_int32(nz), C.c_double(lx),
                                         C.c_double(ly), C.c_double(lz), C.c_int32(len(patches)),
                                         nsides, sides, empty, C.c_int32(int(with_points)),
                                         C.byref(out)))
        self._owned_ptr = out
        self._c = out.contents
        self.patch_names = [p[0] for p in patches if not p[2]]
        return self

    @classmethod
    def from_polymesh(cls, poly_mesh_dir):
        """Read an OpenFOAM ASCII `constant/polyMesh` directory (fvk_polymesh_read): the stand-in for
        FoamAdapter::readOpenFOAMMesh (src/datastructures/meshAdapter.cpp:59-136) without OpenFOAM."""
        self = cls()
        out = C.POINTER(_CDesc)()
        check(lib().fvk_polymesh_read(str(poly_mesh_dir).encode(), C.byref(out)))
        self._owned_ptr, self._owned_kind = out, "polymesh"
        self._c = out.contents
        name, typ = C.create_string_buffer(256), C.create_string_buffer(256)
        self.patch_names, self.patch_types = [], []
        for p in range(self._c.nPatches):
            check(lib().fvk_polymesh_patch(out, C.c_int32(p), name, C.c_int32(256), typ, C.c_int32(256)))
            self.patch_names.append(name.value.decode()); self.patch_types.append(typ.value.decode())
        return self

    def write_polymesh(self, poly_mesh_dir, patches=PATCHES_3DCUBE, types=None):
        """Write a block mesh created `with_points=True` as an OpenFOAM ASCII polyMesh directory (fvk_polymesh_write).
        `patches` is the layout the mesh was generated with; `types[name]` overrides the patch type
        (default: `empty` for empty patches, else `wall`)."""
        import os
        os.makedirs(poly_mesh_dir, exist_ok=True)
        fp, po = self.poly()
        nPoly, nI = fp.shape[0], self.nInternalFaces
        pts = np.ascontiguousarray(self.array("points"), dtype=np.float64)
        # side sizes follow from the faces' owner cells: recover them from the patch layout and the face count
        nei = np.ascontiguousarray(self.array("faceNeighbour"), dtype=np.int32)
        kept = self.array("patchOffsets")
        sizes, k = [], 0
        empty_total = (nPoly - nI) - int(kept[-1])
        n_empty = sum(1 for p in patches if p[2])
        for name_, sides, is_empty in patches:
            if is_empty:
                if n_empty != 1:
                    raise ValueError("write_polymesh: at most one empty patch")
                sizes.append(empty_total)
            else:
                sizes.append(int(kept[k + 1] - kept[k])); k += 1
        offs = np.arange(0, 4 * nPoly + 1, 4, dtype=np.int32)
        names = [p[0].encode() for p in patches]
        tys = [((types or {}).get(p[0]) or ("empty" if p[2] else "wall")).encode() for p in patches]
        arr = lambda lst: (C.c_char_p * len(lst))(*lst)
        fpc, poc = np.ascontiguousarray(fp, dtype=np.int32), np.ascontiguousarray(po, dtype=np.int32)
        check(lib().fvk_polymesh_write(str(poly_mesh_dir).encode(), C.c_int32(pts.shape[0]), pts.ctypes.data_as(C.c_void_p),
                                       C.c_int32(nPoly), offs.ctypes.data_as(C.c_void_p), fpc.ctypes.data_as(C.c_void_p),
                                       poc.ctypes.data_as(C.c_void_p), C.c_int32(nI), nei.ctypes.data_as(C.c_void_p),
                                       C.c_int32(len(patches)), arr(names), arr(tys), (C.c_int32 * len(sizes))(*sizes)))
        return poly_mesh_dir

    @classmethod
    def from_arrays(cls, a: dict, patch_names=None):
        """Build from numpy arrays keyed like the C struct (e.g. from a polyMesh fixture)."""
        self = cls()
        c = _CDesc()
        for k in ("nCells", "nInternalFaces", "nBoundaryFaces", "nPatches"):
            setattr(c, k, int(a[k]))
        c.nPoints = int(a.get("nPoints", 0))
        for k in _F64:
            if a.get(k) is not None:
                arr = np.array(a[k], dtype=np.float64, order="C", copy=True)  # own the storage: views of another desc dangle
                self._keep[k] = arr
                setattr(c, k, arr.ctypes.data_as(_capi.c_double_p))
        for k in _I32:
            if a.get(k) is not None:
                arr = np.array(a[k], dtype=np.int32, order="C", copy=True)
                self._keep[k] = arr
                setattr(c, k, arr.ctypes.data_as(_capi.c_int32_p))
        self._c = c
        self.patch_names = list(patch_names or [f"patch{i}" for i in range(c.nPatches)])
        return self

    def renumbered(self, method="morton", cellOldToNew=None):
        """(new MeshDesc, cellOldToNew, faceOldToNew, faceFlipped): the mesh with its cells in reverse Cuthill-McKee ("rcm") or
        Morton ("morton") order (or a given permutation), faces re-sorted into upper-triangular order (fvk_renumber_*) -- what
        OpenFOAM's `renumberMesh` does to a case before the solver runs. A cell field moves as new[cellOldToNew] = old; a face
        field as new[faceOldToNew] = old, negated where faceFlipped (face fluxes)."""
        if cellOldToNew is None:
            cellOldToNew = np.zeros(self.nCells, dtype=np.int32)
            check(lib().fvk_renumber_order(C.byref(self._c), C.c_int({"rcm": 0, "morton": 1}[method]), cellOldToNew.ctypes.data_as(C.c_void_p)))
        cellOldToNew = np.ascontiguousarray(cellOldToNew, dtype=np.int32)
        out = C.POINTER(_CDesc)()
        check(lib().fvk_renumber_apply(C.byref(self._c), cellOldToNew.ctypes.data_as(C.c_void_p), C.byref(out)))
        new = MeshDesc()
        new._c = out.contents
        new._owned_ptr, new._owned_kind = out, "renumbered"
        new.patch_names = list(self.patch_names)
        pf, pflip = C.POINTER(C.c_int32)(), C.POINTER(C.c_uint8)()
        check(lib().fvk_renumber_maps(out, C.byref(pf), C.byref(pflip)))
        nF = self.nFaces
        faceMap = np.ctypeslib.as_array(pf, shape=(nF,)).copy() if nF else np.zeros(0, np.int32)
        flipped = np.ctypeslib.as_array(pflip, shape=(nF,)).copy().astype(bool) if nF else np.zeros(0, bool)
        return new, cellOldToNew, faceMap, flipped

    @classmethod
    def uniform_1d(cls, nCells: int):
        """NeoN::create1DUniformMesh (src/NeoN/src/mesh/unstructured/unstructuredMesh.cpp:112-220):
        unit interval, nCells cells, internal faces i|i+1, then the left and the right boundary face
        as two one-face patches."""
        n = int(nCells)
        h = (1.0 - 0.0) / float(n)
        pts = np.zeros((n + 1, 3))
        pts[: n - 1, 0] = 0.0 + (np.arange(n - 1) + 1.0) * h
        pts[n - 1] = (0.0, 0.0, 0.0)
        pts[n] = (1.0, 0.0, 0.0)
        C_ = np.zeros((n, 3))
        C_[:, 0] = 0.5 * h + h * np.arange(n, dtype=np.float64)
        Sf = np.tile(np.array([1.0, 0.0, 0.0]), (n + 1, 1))
        Sf[n - 1] = (-1.0, 0.0, 0.0)
        owner = np.concatenate([np.arange(n - 1), [0, n - 1]]).astype(np.int32)
        delta = np.array([[0.0 - C_[0, 0], 0, 0], [1.0 - C_[n - 1, 0], 0, 0]])
        return cls.from_arrays(dict(
            nCells=n, nInternalFaces=n - 1, nBoundaryFaces=2, nPatches=2, nPoints=n + 1, points=pts,
            cellVolumes=np.full(n, h), cellCentres=C_, faceAreas=Sf, faceCentres=pts.copy(),
            magFaceAreas=np.ones(n + 1), faceOwner=owner, faceNeighbour=np.arange(1, n, dtype=np.int32),
            faceCells=np.array([0, n - 1], dtype=np.int32), bCf=pts[n - 1:].copy(), bCn=C_[[0, n - 1]].copy(),
            bSf=Sf[n - 1:].copy(), bMagSf=np.ones(2), bNf=Sf[n - 1:].copy(), bDelta=delta,
            bWeights=np.ones(2), bDeltaCoeffs=1.0 / np.abs(delta[:, 0]),
            patchOffsets=np.array([0, 1, 2], dtype=np.int32)), patch_names=["left", "right"])

    def __del__(self):
        if getattr(self, "_owned_ptr", None) is not None and _capi is not None and _capi._lib is not None:
            if getattr(self, "_owned_kind", "block") == "polymesh":
                _capi._lib.fvk_polymesh_destroy(self._owned_ptr)
            elif getattr(self, "_owned_kind", "block") == "renumbered":
                _capi._lib.fvk_renumber_destroy(self._owned_ptr)
            else:
                _capi._lib.fvk_blockmesh_destroy(self._owned_ptr)
            self._owned_ptr = None

    # -- access -----------------------------------------------------------------------------------
    @property
    def c(self):
        return self._c

    nCells = property(lambda s: s._c.nCells)
    nInternalFaces = property(lambda s: s._c.nInternalFaces)
    nBoundaryFaces = property(lambda s: s._c.nBoundaryFaces)
    nPatches = property(lambda s: s._c.nPatches)
    nFaces = property(lambda s: s._c.nInternalFaces + s._c.nBoundaryFaces)

    def _count(self, name):
        c = self._c
        nF, nB = c.nInternalFaces + c.nBoundaryFaces, c.nBoundaryFaces
        return {"points": 3 * c.nPoints, "cellVolumes": c.nCells, "cellCentres": 3 * c.nCells,
                "faceAreas": 3 * nF, "faceCentres": 3 * nF, "magFaceAreas": nF, "faceOwner": nF,
                "faceNeighbour": c.nInternalFaces, "faceCells": nB, "bCf": 3 * nB, "bCn": 3 * nB,
                "bSf": 3 * nB, "bMagSf": nB, "bNf": 3 * nB, "bDelta": 3 * nB, "bWeights": nB,
                "bDeltaCoeffs": nB, "patchOffsets": c.nPatches + 1}[name]

    def array(self, name) -> np.ndarray:
        """numpy view (no copy) of a host array of the description."""
        p = getattr(self._c, name)
        n = self._count(name)
        if not p or n == 0:
            return np.zeros(0, dtype=np.int32 if name in _I32 else np.float64)
        a = np.ctypeslib.as_array(p, shape=(n,))
        if name in ("points", "cellCentres", "faceAreas", "faceCentres", "bCf", "bCn", "bSf", "bNf", "bDelta"):
            a = a.reshape(-1, 3)
        return a

    def poly(self):
        """(facePoints [nPoly,4], polyOwner [nPoly]) of a block mesh created with points."""
        n = C.c_int32()
        fp = _capi.c_int32_p()
        po = _capi.c_int32_p()
        check(lib().fvk_blockmesh_poly(self._owned_ptr, C.byref(n), C.byref(fp), C.byref(po)))
        return (np.ctypeslib.as_array(fp, shape=(n.value, 4)), np.ctypeslib.as_array(po, shape=(n.value,)))


class UnstructuredMesh:
    """Device mesh handle (`fvk_mesh*`)."""

    def __init__(self, desc: MeshDesc):
        h = C.c_void_p()
        check(lib().fvk_mesh_create(C.byref(desc.c), C.byref(h)))
        self._h = h
        self.patch_names = list(desc.patch_names)
        self.nCells = self.size(N_CELLS)
        self.nInternalFaces = self.size(N_INTERNAL_FACES)
        self.nBoundaryFaces = self.size(N_BOUNDARY_FACES)
        self.nFaces = self.nInternalFaces + self.nBoundaryFaces
        self.nPatches = self.size(N_PATCHES)
        self.nnz = self.size(NNZ)
        self.nOwned = self.size(N_OWNED_CELLS)  # cells this rank owns (ghost cells of a decomposed mesh follow them)
        off = (C.c_int32 * (self.nPatches + 1))()
        check(lib().fvk_mesh_patch_offsets(self._h, off))
        self.patch_offsets = list(off)

    def __del__(self):
        if getattr(self, "_h", None) and _capi is not None and _capi._lib is not None:
            _capi._lib.fvk_mesh_destroy(self._h)
            self._h = None

    @property
    def handle(self):
        return self._h

    def set_tile_phase(self, phase: int):
        """fvk_mesh_set_tile_phase: 0 all cells, 1 only tiles that read no ghost cell, 2 only tiles that do."""
        check(lib().fvk_mesh_set_tile_phase(self._h, C.c_int(phase)))

    def size(self, field) -> int:
        v = C.c_int64()
        check(lib().fvk_mesh_size(self._h, C.c_int(field), C.byref(v)))
        return v.value

    def device_array(self, field) -> tuple[int, int]:
        """(device pointer, element count)."""
        p = C.c_void_p()
        n = C.c_int64()
        check(lib().fvk_mesh_array(self._h, C.c_int(field), C.byref(p), C.byref(n)))
        return (p.value or 0), n.value

    def to_host(self, field) -> np.ndarray:
        p, n = self.device_array(field)
        out = np.empty(n, dtype=_DTYPES.get(field, np.float64))
        if n:
            check(lib().fvk_memcpy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(p),
                                       C.c_size_t(out.nbytes), None))
            check(lib().fvk_stream_sync(None))
        return out
