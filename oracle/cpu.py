"""TEST INFRASTRUCTURE -- numpy front-end of the CPU oracle (oracle/fvo.cpp -> oracle/libfvo.so).

`Mesh` bundles the reference-order mesh arrays (the NeoN::UnstructuredMesh view) and derives
geometry-scheme / sparsity / stencil data with the oracle itself. Every method restates one
reference function; see fvo.cpp for file:line citations.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None


def build(force: bool = False) -> Path:
    so = HERE / "libfvo.so"
    if force or not so.exists() or so.stat().st_mtime < max((HERE / f).stat().st_mtime for f in ("fvo.cpp", "blockmesh.cpp", "Makefile")):
        subprocess.run(["make", "-C", str(HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(str(build()))
        _LIB.fvo_max_threads.restype = C.c_int
        _LIB.fvo_cg.restype = C.c_int
        _LIB.fvo_bicgstab.restype = C.c_int
        _LIB.fvo_cg_dic.restype = C.c_int
    return _LIB


def _arg(x):
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"], "oracle arrays must be contiguous"
        return C.c_void_p(x.ctypes.data)
    if isinstance(x, (bool, int, np.integer)):
        return C.c_int32(int(x))
    if isinstance(x, (float, np.floating)):
        return C.c_double(float(x))
    raise TypeError(type(x))


def call(name, *args):
    return getattr(lib(), name)(*[_arg(a) for a in args])


def max_threads() -> int:
    return lib().fvo_max_threads()


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Mesh:
    """Reference-order mesh arrays + oracle-derived geometry scheme, sparsity and stencil."""

    def __init__(self, *, nCells, owner, neighbour, faceCells, V, C, Sf, Cf, magSf, patchOffsets,
                 bSf=None, bDeltaCoeffs=None, bWeights=None):
        self.nC = int(nCells)
        self.nI = len(neighbour)
        self.nB = len(faceCells)
        self.nF = self.nI + self.nB
        cp_i = lambda a: np.array(a, dtype=np.int32, order="C", copy=True)
        cp_f = lambda a: np.array(a, dtype=np.float64, order="C", copy=True)
        self.owner = cp_i(owner[: self.nI])
        self.neighbour = cp_i(neighbour)
        self.faceCells = cp_i(faceCells)
        self.V, self.C, self.Sf, self.Cf, self.magSf = cp_f(V), cp_f(C), cp_f(Sf), cp_f(Cf), cp_f(magSf)
        self.patchOffsets = cp_i(patchOffsets)
        self.bSf = cp_f(bSf) if bSf is not None else cp_f(self.Sf.reshape(-1, 3)[self.nI:])
        if bDeltaCoeffs is None:
            d = self.Cf.reshape(-1, 3)[self.nI:] - self.C.reshape(-1, 3)[self.faceCells]
            bDeltaCoeffs = 1.0 / np.sqrt((d * d).sum(1))
        self.bDeltaCoeffs = cp_f(bDeltaCoeffs)
        self.bWeights = cp_f(bWeights) if bWeights is not None else np.ones(self.nB)
        self.nnz = self.nC + 2 * self.nI
        # geometry scheme
        self.w, self.dc, self.nodc = (np.zeros(self.nF) for _ in range(3))
        call("fvo_geometry_scheme", self.nI, self.nB, self.owner, self.neighbour, self.faceCells, self.C,
             self.Cf, self.Sf, self.magSf, self.w, self.dc, self.nodc)
        # sparsity
        self.rowOffs = np.zeros(self.nC + 1, np.int32)
        self.colIdxs = np.zeros(self.nnz, np.int32)
        self.ownerOffset = np.zeros(self.nI, np.uint8)
        self.neighbourOffset = np.zeros(self.nI, np.uint8)
        self.diagOffset = np.zeros(self.nC, np.uint8)
        call("fvo_sparsity", self.nC, self.nI, self.owner, self.neighbour, self.rowOffs, self.colIdxs,
             self.ownerOffset, self.neighbourOffset, self.diagOffset)

    # block sides: 0 x-min, 1 x-max, 2 y-min, 3 y-max, 4 z-min, 5 z-max; patches = [(name, [sides], isEmpty)]
    PATCHES_3DCUBE = [("top", [3], False), ("bottom", [2], False), ("sides", [0, 1, 4, 5], False)]

    @classmethod
    def block(cls, nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, patches=None):
        """Single hex block in blockMesh ordering with primitiveMesh geometry (oracle/blockmesh.cpp): the oracle's own
        mesh generator, so the CPU arms of bench.py never load the product library."""
        patches = patches or cls.PATCHES_3DCUBE
        nSides = i32([len(p[1]) for p in patches])
        sides = i32([s for p in patches for s in p[1]])
        empty = i32([1 if p[2] else 0 for p in patches])
        sz = np.zeros(4, np.int64)
        call("fvo_blockmesh_sizes", nx, ny, nz, len(patches), nSides, sides, empty, sz)
        nC, nI, nB, kept = (int(v) for v in sz)
        a = dict(owner=np.zeros(nI + nB, np.int32), neighbour=np.zeros(nI, np.int32), faceCells=np.zeros(nB, np.int32),
                 V=np.zeros(nC), C=np.zeros(nC * 3), Sf=np.zeros((nI + nB) * 3), Cf=np.zeros((nI + nB) * 3), magSf=np.zeros(nI + nB),
                 bSf=np.zeros(nB * 3), bDeltaCoeffs=np.zeros(nB), bWeights=np.zeros(nB), patchOffsets=np.zeros(kept + 1, np.int32))
        lib().fvo_blockmesh(C.c_int32(nx), C.c_int32(ny), C.c_int32(nz), C.c_double(lx), C.c_double(ly), C.c_double(lz),
                            C.c_int32(len(patches)), _arg(nSides), _arg(sides), _arg(empty), *[_arg(a[k]) for k in
                            ("owner", "neighbour", "faceCells", "V", "C", "Sf", "Cf", "magSf", "bSf", "bDeltaCoeffs", "bWeights", "patchOffsets")])
        return cls(nCells=nC, **a)

    @classmethod
    def from_desc(cls, d):
        """From a foamadapter_b200.mesh.MeshDesc (host arrays of the product's generator)."""
        a = d.array
        return cls(nCells=d.nCells, owner=a("faceOwner"), neighbour=a("faceNeighbour"), faceCells=a("faceCells"),
                   V=a("cellVolumes"), C=a("cellCentres"), Sf=a("faceAreas"), Cf=a("faceCentres"),
                   magSf=a("magFaceAreas"), patchOffsets=a("patchOffsets"), bSf=a("bSf"),
                   bDeltaCoeffs=a("bDeltaCoeffs"), bWeights=a("bWeights"))

    def stencil(self):
        seg = np.zeros(self.nC + 1, np.int32)
        val = np.zeros(2 * self.nI + self.nB, np.int32)
        call("fvo_cell_to_face_stencil", self.nC, self.nI, self.nB, self.owner, self.neighbour, self.faceCells, seg, val)
        return seg, val

    # ---- explicit ------------------------------------------------------------------------------
    def _topo(self):
        return (self.nC, self.nI, self.nB, self.owner, self.neighbour, self.faceCells, self.V)

    def div(self, faceFlux, phi, bvalue, scheme=0, coeff=1.0, coeffView=None, par=0, res=None):
        phi = f64(phi)
        vec = phi.ndim == 2
        res = np.zeros_like(phi) if res is None else res
        call("fvo_div_v" if vec else "fvo_div_s", par, scheme, *self._topo(), self.w, f64(faceFlux), phi, f64(bvalue),
             float(coeff), coeffView, res)
        return res

    def grad(self, phi, bvalue, par=0, res=None):
        res = np.zeros((self.nC, 3)) if res is None else res
        call("fvo_grad_s", par, *self._topo(), self.w, self.Sf, f64(phi), f64(bvalue), res)
        return res

    def laplacian(self, phi, bvalue, coeff=1.0, coeffView=None, par=0, res=None):
        phi = f64(phi)
        vec = phi.ndim == 2
        res = np.zeros_like(phi) if res is None else res
        call("fvo_laplacian_v" if vec else "fvo_laplacian_s", par, *self._topo(), self.magSf, self.nodc, phi,
             f64(bvalue), float(coeff), coeffView, res)
        return res

    def surface_integrate(self, flux, coeff=1.0, coeffView=None, par=0, res=None):
        flux = f64(flux)
        vec = flux.ndim == 2
        res = (np.zeros((self.nC, 3)) if vec else np.zeros(self.nC)) if res is None else res
        call("fvo_surface_integrate_v" if vec else "fvo_surface_integrate_s", par, *self._topo(), flux, float(coeff),
             coeffView, res)
        return res

    def interpolate(self, phi, bvalue, scheme=0, faceFlux=None, par=0):
        phi = f64(phi)
        vec = phi.ndim == 2
        out = np.zeros((self.nF, 3)) if vec else np.zeros(self.nF)
        call("fvo_interpolate_v" if vec else "fvo_interpolate_s", par, scheme, self.nI, self.nB, self.owner,
             self.neighbour, self.w, None if faceFlux is None else f64(faceFlux), phi, f64(bvalue), out)
        return out

    def upwind_weights(self, faceFlux):
        w, wb = np.zeros(self.nF), np.zeros(self.nB)
        call("fvo_upwind_weights", self.nI, self.nB, f64(faceFlux), w, wb)
        return w, wb

    def face_normal_grad(self, phi, bvalue, par=0):
        phi = f64(phi)
        vec = phi.ndim == 2
        out = np.zeros((self.nF, 3)) if vec else np.zeros(self.nF)
        call("fvo_face_normal_grad_v" if vec else "fvo_face_normal_grad_s", par, self.nI, self.nB, self.owner,
             self.neighbour, self.faceCells, self.nodc, phi, f64(bvalue), out)
        return out

    def conum(self, faceFlux, dt, par=0):
        out = np.zeros(2)
        call("fvo_conum", par, *self._topo(), f64(faceFlux), float(dt), out)
        return out

    def correct_bcs(self, kinds, consts, internal):
        """Returns dict(value, refValue, valueFraction, refGrad) after correctBoundaryConditions."""
        internal = f64(internal)
        vec = internal.ndim == 2
        shp = (self.nB, 3) if vec else (self.nB,)
        bd = dict(value=np.zeros(shp), refValue=np.zeros(shp), valueFraction=np.zeros(self.nB), refGrad=np.zeros(shp))
        cst = f64(np.asarray(consts, dtype=np.float64).reshape(len(kinds), 3 if vec else 1))
        call("fvo_correct_bcs_v" if vec else "fvo_correct_bcs_s", len(kinds), self.patchOffsets, i32(kinds), cst,
             self.faceCells, self.bDeltaCoeffs, internal, bd["value"], bd["refValue"], bd["valueFraction"], bd["refGrad"])
        return bd

    # ---- implicit ------------------------------------------------------------------------------
    def empty_system(self, vec=False):
        shp = (lambda n: (n, 3)) if vec else (lambda n: (n,))
        return dict(values=np.zeros(shp(self.nnz)), rhs=np.zeros(shp(self.nC)), bcMatrix=np.zeros(shp(self.nB)),
                    bcRhs=np.zeros(shp(self.nB)))

    def _imp(self):
        return (self.nI, self.nB, self.owner, self.neighbour, self.faceCells, self.rowOffs, self.diagOffset,
                self.ownerOffset, self.neighbourOffset)

    def div_imp(self, ls, faceFlux, bd, scheme=0, coeff=1.0, coeffView=None, par=0):
        vec = ls["values"].ndim == 2
        if scheme == 0:
            w, wb = self.w, self.bWeights
        else:
            w, wb = self.upwind_weights(faceFlux)
        call("fvo_div_imp_v" if vec else "fvo_div_imp_s", par, *self._imp(), f64(faceFlux), w, wb, self.bDeltaCoeffs,
             bd["valueFraction"], bd["refValue"], bd["refGrad"], float(coeff), coeffView, ls["values"], ls["rhs"],
             ls["bcMatrix"], ls["bcRhs"])

    def laplacian_imp(self, ls, gamma, bd, coeff=1.0, coeffView=None, par=0):
        vec = ls["values"].ndim == 2
        call("fvo_laplacian_imp_v" if vec else "fvo_laplacian_imp_s", par, *self._imp(), f64(gamma), self.nodc, self.magSf,
             bd["valueFraction"], bd["refValue"], bd["refGrad"], float(coeff), coeffView, ls["values"], ls["rhs"],
             ls["bcMatrix"], ls["bcRhs"])

    def ddt_imp(self, ls, oldField, dt, coeff=1.0, coeffView=None, par=0):
        vec = ls["values"].ndim == 2
        call("fvo_ddt_imp_v" if vec else "fvo_ddt_imp_s", par, self.nC, self.rowOffs, self.diagOffset, self.V,
             f64(oldField), float(dt), float(coeff), coeffView, ls["values"], ls["rhs"])

    def source_imp(self, ls, k, coeff=1.0, coeffView=None, par=0):
        vec = ls["values"].ndim == 2
        call("fvo_source_imp_v" if vec else "fvo_source_imp_s", par, self.nC, self.rowOffs, self.diagOffset, self.V,
             f64(k), float(coeff), coeffView, ls["values"])

    # ---- linear algebra --------------------------------------------------------------------------
    def spmv(self, values, x, par=0):
        y = np.zeros(self.nC)
        call("fvo_spmv", par, self.nC, self.rowOffs, self.colIdxs, f64(values), f64(x), y)
        return y

    def residual(self, values, b, x, par=0):
        r = np.zeros(self.nC)
        call("fvo_residual", par, self.nC, self.rowOffs, self.colIdxs, f64(values), f64(b), f64(x), r)
        return r

    # ---- PISO glue (FoamAdapter src/algorithms/pressureVelocityCoupling.cpp) -------------------------
    def rAU(self, valuesV):
        out = np.zeros(self.nC)
        call("fvo_rAU", self.nC, self.rowOffs, self.diagOffset, self.V, f64(valuesV), out)
        return out

    def HbyA(self, valuesV, rhsV, rAU, U, par=0):
        out = np.zeros((self.nC, 3))
        call("fvo_HbyA", par, self.nC, self.nI, self.owner, self.neighbour, self.rowOffs, self.ownerOffset,
             self.neighbourOffset, self.V, f64(valuesV), f64(rhsV), f64(rAU), f64(U), out)
        return out

    def flux(self, U, Ub, par=0):
        ff, bv = np.zeros(self.nF), np.zeros(self.nB)
        call("fvo_flux", par, self.nI, self.nB, self.owner, self.neighbour, self.w, self.Sf, self.bSf, f64(U), f64(Ub), ff, bv)
        return ff, bv

    def update_face_velocity(self, ls, p, predPhi, predPhiB, par=0):
        phi, phiB = np.zeros(self.nF), np.zeros(self.nB)
        call("fvo_update_face_velocity", par, self.nI, self.nB, self.owner, self.neighbour, self.faceCells, self.rowOffs,
             self.ownerOffset, self.neighbourOffset, ls["values"], ls["bcMatrix"], ls["bcRhs"], f64(p), f64(predPhi),
             f64(predPhiB), phi, phiB)
        return phi, phiB

    def update_velocity(self, HbyA, rAU, gradP, par=0):
        U = np.zeros((self.nC, 3))
        call("fvo_update_velocity", par, self.nC, f64(HbyA), f64(rAU), f64(gradP), U)
        return U

    def set_reference(self, ls, refCell, refValue):
        call("fvo_set_reference", int(refCell), float(refValue), self.rowOffs, self.diagOffset, ls["values"], ls["rhs"])

    def rhs_sub_source(self, ls, src, par=0):
        call("fvo_rhs_sub_source", par, self.nC, self.V, f64(src), ls["rhs"])

    def cg(self, values, b, x0, jacobi=True, max_iter=1000, rel_tol=0.0, abs_tol=1e-6, par=0, max_hist=0):
        return cg(self.rowOffs, self.colIdxs, values, b, x0, jacobi, max_iter, rel_tol, abs_tol, par, max_hist)

    def bicgstab(self, values, b, x0, jacobi=True, max_iter=1000, rel_tol=0.0, abs_tol=0.0, par=0, max_hist=0):
        return bicgstab(self.rowOffs, self.colIdxs, values, b, x0, jacobi, max_iter, rel_tol, abs_tol, par, max_hist)


def cg(rowOffs, colIdxs, values, b, x0, jacobi=True, max_iter=1000, rel_tol=0.0, abs_tol=1e-6, par=0, max_hist=0, fn="fvo_cg"):
    """Returns (x, dict(numIter, initResNorm, finalResNorm), history)."""
    x = f64(x0).copy()
    stats = np.zeros(3)
    hist = np.zeros(max(max_hist, 1))
    nh = getattr(lib(), fn)(C.c_int(par), C.c_int(int(jacobi)), C.c_int32(len(x)), _arg(i32(rowOffs)), _arg(i32(colIdxs)),
                      _arg(f64(values)), _arg(f64(b)), _arg(x), C.c_int(max_iter), C.c_double(rel_tol),
                      C.c_double(abs_tol), _arg(stats), _arg(hist) if max_hist else C.c_void_p(0), C.c_int(max_hist))
    return x, dict(numIter=int(stats[0]), initResNorm=stats[1], finalResNorm=stats[2]), hist[:nh]


def bicgstab(rowOffs, colIdxs, values, b, x0, jacobi=True, max_iter=1000, rel_tol=0.0, abs_tol=0.0, par=0, max_hist=0):
    """Ginkgo solver::Bicgstab restated (fvo_bicgstab); history holds every checked norm (||r|| and ||s|| alternate)."""
    return cg(rowOffs, colIdxs, values, b, x0, jacobi, max_iter, rel_tol, abs_tol, par, max_hist, fn="fvo_bicgstab")


def cg_dic(rowOffs, colIdxs, values, b, x0, max_iter=1000, rel_tol=0.0, abs_tol=1e-6, max_hist=0):
    """EXTENSION: CG with the multicolour DIC preconditioner (fvo_cg_dic). Returns (x, stats, history, colours)."""
    x = f64(x0).copy()
    stats, hist, colors = np.zeros(3), np.zeros(max(max_hist, 1)), np.zeros(len(x), np.int32)
    nh = lib().fvo_cg_dic(C.c_int32(len(x)), _arg(i32(rowOffs)), _arg(i32(colIdxs)), _arg(f64(values)), _arg(f64(b)), _arg(x), C.c_int(max_iter),
                          C.c_double(rel_tol), C.c_double(abs_tol), _arg(stats), _arg(hist) if max_hist else C.c_void_p(0), C.c_int(max_hist), _arg(colors))
    return x, dict(numIter=int(stats[0]), initResNorm=stats[1], finalResNorm=stats[2]), hist[:nh], colors
