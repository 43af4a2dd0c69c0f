"""Python mirror of NeoN::la (src/NeoN/include/NeoN/linearAlgebra/{sparsityPattern,CSRMatrix,linearSystem,
solver,utilities}.hpp): SparsityPattern, LinearSystem, Solver/SolverStats, computeResidual, spmv and the
Vector free functions. Bodies call the C ABI (include/fvk.h); tensors are device-memory handles only."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import mesh as _m
from . import ops
from ._capi import check, lib, ptr


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class SparsityPattern:
    """la::SparsityPattern (sparsityPattern.hpp:20-92); built once per mesh by fvk_mesh_create and cached there
    like SparsityPattern::readOrCreate (sparsityPattern.cpp:11-19). Arrays are views of the handle's device memory."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.rowOffs_ptr, _ = mesh.device_array(_m.ROW_OFFS)
        self.colIdxs_ptr, _ = mesh.device_array(_m.COL_IDXS)

    @staticmethod
    def readOrCreate(mesh):
        if getattr(mesh, "_sparsity", None) is None:
            mesh._sparsity = SparsityPattern(mesh)
        return mesh._sparsity

    def rowOffs(self): return self.mesh.to_host(_m.ROW_OFFS)
    def colIdxs(self): return self.mesh.to_host(_m.COL_IDXS)
    def ownerOffset(self): return self.mesh.to_host(_m.OWNER_OFFSET)
    def neighbourOffset(self): return self.mesh.to_host(_m.NEIGHBOUR_OFFSET)
    def diagOffset(self): return self.mesh.to_host(_m.DIAG_OFFSET)


class LinearSystem:
    """la::LinearSystem<scalar|Vec3, localIdx> + BoundaryCoefficients (linearSystem.hpp:36-186)."""

    def __init__(self, mesh, ncomp=1, device="cuda", zero=True, compact=False):
        """compact (Vec3 systems only; an HBM-layout choice of this implementation): the implicit operators give the three
        components of every matrix entry the same value, so `values` holds each entry once (double[nnz]);
        `valuesVec3()` materialises the reference's Vec3[nnz] layout bit for bit. rhs / bcMatrix / bcRhs stay Vec3."""
        self.mesh, self.ncomp = mesh, ncomp
        self.compact = bool(compact) and ncomp == 3
        self.sp = SparsityPattern.readOrCreate(mesh)
        mk = torch.zeros if zero else torch.empty
        shp = (lambda n: (n, 3)) if ncomp == 3 else (lambda n: (n,))
        self.values = mk((mesh.nnz,) if self.compact else shp(mesh.nnz), dtype=torch.float64, device=device)
        self.rhs = mk(shp(mesh.nCells), dtype=torch.float64, device=device)
        self.bcMatrix = torch.zeros(shp(mesh.nBoundaryFaces), dtype=torch.float64, device=device)
        self.bcRhs = torch.zeros(shp(mesh.nBoundaryFaces), dtype=torch.float64, device=device)

    def reset(self):
        self.values.zero_(); self.rhs.zero_()

    def valuesVec3(self):
        """LinearSystem<Vec3>::matrix().values() in the reference layout (a copy when the system is compact)."""
        if not self.compact:
            return self.values
        out = torch.empty((self.mesh.nnz, 3), dtype=torch.float64, device=self.values.device)
        check(lib().fvk_expand_vec3(C.c_int64(self.mesh.nnz), ptr(self.values), ptr(out), _stream()))
        ops._count()
        return out


def createEmptyLinearSystem(mesh, ncomp=1):
    """linearSystem.hpp:140-186."""
    return LinearSystem(mesh, ncomp, zero=True)


def spmv_structured(mesh, values, x, y=None):
    """y = A x over the mesh's own sparsity pattern with the structured fast path (fvk_spmv_structured)."""
    if y is None:
        y = torch.empty(mesh.nOwned, dtype=torch.float64, device=x.device)
    check(lib().fvk_spmv_structured(mesh.handle, ptr(values), ptr(x), ptr(y), _stream()))
    ops._count()
    return y


def spmv(sp: SparsityPattern, values, x, y=None, nRows=None):
    n = nRows if nRows is not None else sp.mesh.nOwned
    if y is None:
        y = torch.empty(n, dtype=torch.float64, device=x.device)
    check(lib().fvk_spmv(C.c_int32(n), C.c_void_p(sp.rowOffs_ptr), C.c_void_p(sp.colIdxs_ptr), ptr(values), ptr(x), ptr(y), _stream()))
    ops._count()
    return y


def computeResidual(sp: SparsityPattern, values, b, x, res=None):
    """la::computeResidual (utilities.cpp:11-35): res = A x - b."""
    n = sp.mesh.nOwned
    if res is None:
        res = torch.empty(n, dtype=torch.float64, device=x.device)
    check(lib().fvk_residual(C.c_int32(n), C.c_void_p(sp.rowOffs_ptr), C.c_void_p(sp.colIdxs_ptr), ptr(values), ptr(b), ptr(x), ptr(res), _stream()))
    ops._count()
    return res


@dataclass
class SolverStats:
    """la::SolverStats (solver.hpp:14-27)."""
    numIter: int
    initResNorm: float
    finalResNorm: float
    history: np.ndarray | None = None

    def print(self, name):
        print(f"Solver: {name} , Initial residual = {self.initResNorm} , Final residual = {self.finalResNorm} , No Iterations = {self.numIter}")


class CapturedSolverStats:
    """Statistics of a solve that was captured into a CUDA graph (fvk_solver_solve on a capturing stream): read from the
    solver's pinned slot after a replay; attribute access waits for the device (the replay must have been launched)."""

    def __init__(self, solver, slot):
        self._solver, self._slot = solver, slot
        self.history = None

    def _read(self):
        torch.cuda.current_stream().synchronize()
        st = _Stats()
        check(lib().fvk_solver_captured_stats(self._solver._h, C.c_int32(self._slot), C.byref(st)))
        return st

    @property
    def numIter(self): return self._read().numIter
    @property
    def initResNorm(self): return self._read().initResNorm
    @property
    def finalResNorm(self): return self._read().finalResNorm

    def snapshot(self):
        st = self._read()
        return SolverStats(st.numIter, st.initResNorm, st.finalResNorm)


class _Cfg(C.Structure):
    _fields_ = [("maxIter", C.c_int32), ("relTol", C.c_double), ("absTol", C.c_double), ("preconditioner", C.c_int32),
                ("checkEvery", C.c_int32), ("solverType", C.c_int32)]


class _Stats(C.Structure):
    _fields_ = [("numIter", C.c_int32), ("initResNorm", C.c_double), ("finalResNorm", C.c_double), ("nHistory", C.c_int32)]


# src/compatibility/fvSolution.cpp:22-28 (updateSolver) and :51-62 (updatePreconditioner)
SOLVER_MAP = {"PCG": "solver::Cg", "PBiCG": "solver::Bicg", "PBiCGStab": "solver::Bicgstab", "smoothSolver": "solver::Bicgstab",
              "GAMG": "solver::Multigrid"}
PRECONDITIONER_MAP = {
    "DIC": {"type": "preconditioner::Jacobi", "max_block_size": 1},
    "DILU": {"type": "preconditioner::Ilu", "reverse_apply": False, "factorization": {"type": "factorization::ParIlu"}},
}


def mapFvSolution(d: dict) -> dict:
    """FoamAdapter::mapFvSolution (src/compatibility/fvSolution.cpp:142-157): OpenFOAM solver entry -> Ginkgo-style
    dictionary, step for step: updateSolver (:19-46), updatePreconditioner (:48-105: a MISSING preconditioner becomes
    DIC -> Jacobi, `smoother` is dropped, DILU -> Ilu/ParIlu), updateCriteria (:107-139: iteration 1000 unless maxIter,
    relative/absolute norms only when relTol / tolerance are present). A dictionary with `configFile` is returned as is."""
    if "configFile" in d:
        return d
    out = {k: (dict(v) if isinstance(v, dict) else v) for k, v in d.items()}
    name = out.get("solver")
    if name in SOLVER_MAP:
        out["solver"], out["type"] = "Ginkgo", SOLVER_MAP[name]
    if "preconditioner" not in out:
        out["preconditioner"] = dict(PRECONDITIONER_MAP["DIC"])
    out.pop("smoother", None)
    pre = out["preconditioner"]
    if isinstance(pre, dict):
        if isinstance(pre.get("type"), dict):
            raise RuntimeError("GAMG is not supported in FoamAdapter, please use a different preconditioner.")
    elif pre in PRECONDITIONER_MAP:
        out["preconditioner"] = {k: (dict(v) if isinstance(v, dict) else v) for k, v in PRECONDITIONER_MAP[pre].items()}
    crit = out.setdefault("criteria", {})
    crit["iteration"] = 1000
    if "relTol" in out:
        crit["relative_residual_norm"] = float(out.pop("relTol"))
    if "maxIter" in out:
        crit["iteration"] = int(out.pop("maxIter"))
    if "tolerance" in out:
        crit["absolute_residual_norm"] = float(out.pop("tolerance"))
    return out


SOLVER_TYPES = {"solver::Cg": 0, "solver::Bicgstab": 1}


class Solver:
    """la::Solver(exec, dict) (solver.hpp:63-91) with a Ginkgo-style dictionary (ginkgo.hpp:95-108), i.e. what
    mapFvSolution emits or what the reference's tests pass directly (test/test_advection.cpp:125-131): solver::Cg or
    solver::Bicgstab, optional scalar Jacobi preconditioner, criteria iteration / relative_residual_norm /
    absolute_residual_norm. An OpenFOAM-style entry (no `type`) is mapped first. Everything else raises: there is no
    silent downgrade (preconditioner::Ilu, solver::Bicg, solver::Multigrid are not on the hot path).
    solve(ls, x) -> SolverStats like GinkgoSolver::solve (ginkgo.hpp:116-155)."""

    def __init__(self, config: dict, comm=None, check_every=8, history=False):
        cfg = config if "type" in config else mapFvSolution(config)
        if cfg.get("solver", "Ginkgo") != "Ginkgo":
            raise KeyError(f"la::SolverFactory has no solver '{cfg.get('solver')}'")  # RuntimeSelectionFactory::keyExistsOrError
        if cfg.get("type") not in SOLVER_TYPES:
            raise KeyError(f"solver type '{cfg.get('type')}' is not on the hot path (solver::Cg, solver::Bicgstab)")
        crit = cfg.get("criteria", {})
        pre = cfg.get("preconditioner")
        if pre in (None, "none"):
            precond = 0
        elif isinstance(pre, dict) and pre.get("type") == "preconditioner::Jacobi" and int(pre.get("max_block_size", 1)) == 1:
            precond = 1
        elif isinstance(pre, dict) and pre.get("type") == "preconditioner::Ic":
            # extension: Ginkgo's name for incomplete Cholesky selects the multicolour DIC of libfvk
            # (FVK_PRECOND_DIC; solver::Cg, one GPU). mapFvSolution never emits it: OpenFOAM's DIC maps to Jacobi like the reference.
            precond = 2
        else:
            raise KeyError(f"preconditioner {pre!r} is not supported (scalar preconditioner::Jacobi, or preconditioner::Ic = multicolour DIC)")
        self.cfg = _Cfg(int(crit.get("iteration", 1000)), float(crit.get("relative_residual_norm", 0.0)),
                        float(crit.get("absolute_residual_norm", 0.0)), precond, int(check_every), SOLVER_TYPES[cfg["type"]])
        self.type = cfg["type"]
        self.comm, self.history = comm, history
        self._h, self._shape, self._ghosts_current = None, None, False

    def _handle(self, nRows, nCols):
        if self._shape != (nRows, nCols):
            self.close()
            h = C.c_void_p()
            check(lib().fvk_solver_create(C.c_int32(nRows), C.c_int32(nCols), C.byref(self.cfg),
                                          self.comm.handle if self.comm is not None else None, C.byref(h)))
            self._h, self._shape, self._attached = h, (nRows, nCols), None
            if self._ghosts_current:
                check(lib().fvk_solver_set_ghosts_current(h, C.c_int32(1)))
        return self._h

    def close(self):
        if self._h is not None:
            lib().fvk_solver_destroy(self._h)
            self._h = None

    def captured_log(self):
        """iteration counts of the captured solves executed since the last call, in execution order (device-side log)"""
        out = (C.c_int32 * 8192)()
        n = C.c_int32(0)
        check(lib().fvk_solver_captured_log(self._h, out, C.c_int32(8192), C.byref(n)))
        return list(out[: n.value])

    def set_ghosts_current(self, on: bool = True):
        """fvk_solver_set_ghosts_current: every initial guess passed from now on has current ghost entries (skip that exchange)"""
        self._ghosts_current = bool(on)
        if self._h is not None:
            check(lib().fvk_solver_set_ghosts_current(self._h, C.c_int32(1 if on else 0)))

    @property
    def keeps_ghosts(self) -> bool:
        """fvk_solver_keeps_ghosts: the solution leaves the solver with current ghost entries (no exchange needed afterwards)"""
        if self._h is None:  # same rule as the library's, before the handle exists
            return self.comm is None or (self.type == "solver::Cg" and bool(getattr(self.comm, "p2p", False)))
        out = C.c_int32(0)
        check(lib().fvk_solver_keeps_ghosts(self._h, C.byref(out)))
        return bool(out.value)

    def reset_captures(self):
        """forget the slots of earlier captured solves (before capturing a new graph)"""
        if self._h is not None:
            check(lib().fvk_solver_reset_captures(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _launches(self, numIter):
        # CG: K1 + K2 per completed iteration, + the final K1 and the no-op K2 behind it, + the start-up SpMV (r0, ||b||^2 and, with
        # the mesh attached, 1 / diag; else one more kernel for it); BiCGStab: 5 per iteration
        if self.cfg.solverType == 0:
            return 2 * numIter + 3 + (0 if getattr(self, "_attached", None) is not None or self.cfg.preconditioner != 1 else 1)
        return 5 * numIter + 10

    def solve_csr(self, nRows, nCols, rowOffs_ptr, colIdxs_ptr, values, b, x, _mesh=None) -> SolverStats:
        h = self._handle(nRows, nCols)
        if _mesh is None and getattr(self, "_attached", None) is not None:
            check(lib().fvk_solver_attach_mesh(h, None))  # a foreign CSR: no structured shortcut
            self._attached = None
        st = _Stats()
        nh = (2 if self.cfg.solverType == 1 else 1) * self.cfg.maxIter + 2 if self.history else 0
        hist = np.zeros(max(nh, 1))
        check(lib().fvk_solver_solve(h, C.c_void_p(rowOffs_ptr), C.c_void_p(colIdxs_ptr), ptr(values), ptr(b), ptr(x),
                                     C.byref(st), hist.ctypes.data_as(C.c_void_p) if nh else None, C.c_int32(nh), _stream()))
        if st.numIter < 0:  # captured into a CUDA graph: the iteration is a conditional node, statistics come after a replay
            return CapturedSolverStats(self, -st.numIter - 1)
        ops._count(self._launches(st.numIter))
        return SolverStats(st.numIter, st.initResNorm, st.finalResNorm, hist[:st.nHistory] if nh else None)

    def _attach(self, m):
        h = self._handle(m.nOwned, m.nCells)
        if getattr(self, "_attached", None) is not m:  # structured SpMV fast path when the mesh plan allows it
            check(lib().fvk_solver_attach_mesh(h, m.handle))
            self._attached = m
        return h

    def solve(self, ls: LinearSystem, x):
        """Scalar system -> SolverStats; Vec3 system (identical components) -> one SolverStats per component."""
        m = ls.mesh
        h = self._attach(m)
        if ls.ncomp == 3:
            st3 = (_Stats * 3)()
            if ls.compact:
                check(lib().fvk_solver_solve_vec3c(h, C.c_void_p(ls.sp.rowOffs_ptr), C.c_void_p(ls.sp.colIdxs_ptr),
                                                   ptr(ls.values), ptr(ls.rhs), ptr(x), st3, _stream()))
            else:
                check(lib().fvk_solver_solve_vec3(h, C.c_int64(m.nnz), C.c_void_p(ls.sp.rowOffs_ptr), C.c_void_p(ls.sp.colIdxs_ptr),
                                                  ptr(ls.values), ptr(ls.rhs), ptr(x), st3, _stream()))
            ops._count(sum(self._launches(s.numIter) + 3 for s in st3) + 1)
            return [SolverStats(s.numIter, s.initResNorm, s.finalResNorm) for s in st3]
        return self.solve_csr(m.nOwned, m.nCells, ls.sp.rowOffs_ptr, ls.sp.colIdxs_ptr, ls.values, ls.rhs, x, _mesh=m)


# ---- Vector free functions (vectorFreeFunctions.cpp:18-106) ----------------------------------------
def _n(x): return C.c_int64(x.numel())
def fill(x, v): check(lib().fvk_vec_fill(_n(x), C.c_double(v), ptr(x), _stream())); ops._count(); return x
def scalarMul(x, a): check(lib().fvk_vec_scale(_n(x), C.c_double(a), ptr(x), _stream())); ops._count(); return x
def add(x, y): check(lib().fvk_vec_add(_n(x), ptr(x), ptr(y), _stream())); ops._count(); return x
def sub(x, y): check(lib().fvk_vec_sub(_n(x), ptr(x), ptr(y), _stream())); ops._count(); return x
def mul(x, y): check(lib().fvk_vec_mul(_n(x), ptr(x), ptr(y), _stream())); ops._count(); return x
def axpby(a, x, b, y): check(lib().fvk_vec_axpby(_n(y), C.c_double(a), ptr(x), C.c_double(b), ptr(y), _stream())); ops._count(); return y


def waxpby(a, x, b, y, w):
    """w = a*x + b*y"""
    check(lib().fvk_vec_waxpby(_n(w), C.c_double(a), ptr(x), C.c_double(b), ptr(y), ptr(w), _stream())); ops._count(); return w


def scaledCopy(a, x, out):
    """out = x * a"""
    check(lib().fvk_vec_scaled_copy(_n(x), C.c_double(a), ptr(x), ptr(out), _stream())); ops._count(); return out


def dot(x, y, out=None):
    out = out if out is not None else torch.empty(1, dtype=torch.float64, device=x.device)
    check(lib().fvk_dot(_n(x), ptr(x), ptr(y), ptr(out), _stream())); ops._count()
    return out


def norm2(x, out=None):
    out = out if out is not None else torch.empty(1, dtype=torch.float64, device=x.device)
    check(lib().fvk_norm2(_n(x), ptr(x), ptr(out), _stream())); ops._count()
    return out
