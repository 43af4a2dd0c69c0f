"""Python mirror of NeoN::dsl (src/NeoN/include/NeoN/dsl/{coeff,operator,spatialOperator,temporalOperator,
expression,implicit,explicit,solver}.hpp) and FoamAdapter::PDESolver (include/FoamAdapter/datastructures/
expression.hpp:23-179): same factory names (imp.ddt/div/laplacian/source, exp.div/laplacian/source), operator
arithmetic, assembly order and solve sequence; the bodies go to the fused CUDA kernels through the C ABI."""
from __future__ import annotations

from dataclasses import dataclass, field as _field

import torch

from . import fvcc, la, ops
from .fvcc import Coeff


class Operator:
    """dsl::SpatialOperator / TemporalOperator (type-erased in the reference). kind: ddt|div|laplacian|source|
    surfaceIntegrate; type: 'implicit' | 'explicit' (dsl/operator.hpp:29-64)."""

    def __init__(self, kind, type_, field=None, faceField=None, cellField=None, coeff=None):
        self.kind, self.type, self.field, self.faceField, self.cellField = kind, type_, field, faceField, cellField
        self.coeff = coeff or Coeff(1.0)
        self.scheme = None  # filled by read()

    def getType(self): return self.type
    def getName(self): return self.kind
    def getCoefficient(self): return self.coeff

    def _scaled(self, c):
        o = Operator(self.kind, self.type, self.field, self.faceField, self.cellField, self.coeff * c)
        o.scheme = self.scheme
        return o

    def __neg__(self): return self._scaled(-1.0)
    def __rmul__(self, c): return self._scaled(c)
    def __add__(self, rhs): return Expression([self]) + rhs
    def __sub__(self, rhs): return Expression([self]) - rhs

    def read(self, fvSchemes: dict):
        """DivOperator::read / LaplacianOperator::read (operators/divOperator.hpp:173-190, laplacianOperator.hpp:181-198):
        token lists like 'Gauss linear' / 'Gauss linear uncorrected' keyed by operator name."""
        if self.kind == "div" and self.field is not None:
            name = f"div({self.faceField.name},{self.field.name})"
            toks = fvSchemes.get("divSchemes", {}).get(name, "Gauss linear").split()
            if toks[0] != "Gauss":
                raise KeyError(f"unknown div scheme '{toks[0]}'")  # RuntimeSelectionFactory::keyExistsOrError
            if toks[1] not in ops.SCHEMES:
                raise KeyError(f"unknown interpolation scheme '{toks[1]}'")
            self.scheme = ops.SCHEMES[toks[1]]
        elif self.kind == "laplacian":
            name = f"laplacian({self.faceField.name},{self.field.name})"
            toks = fvSchemes.get("laplacianSchemes", {}).get(name, "Gauss linear uncorrected").split()
            if toks[0] != "Gauss" or (len(toks) > 2 and toks[2] != "uncorrected"):
                raise KeyError(f"unknown laplacian scheme '{' '.join(toks)}'")
            self.scheme = 0


class Expression:
    """dsl::Expression (dsl/expression.hpp:47-224): temporal + spatial operator lists; operator- multiplies the
    right-hand side's Coeff by -1 (:207-224)."""

    def __init__(self, operators=()):
        self.temporal = [o for o in operators if o.kind == "ddt"]
        self.spatial = [o for o in operators if o.kind != "ddt"]

    def _all(self): return self.temporal + self.spatial

    def __add__(self, rhs):
        r = rhs._all() if isinstance(rhs, Expression) else [rhs]
        return Expression(self._all() + r)

    def __sub__(self, rhs):
        r = rhs._all() if isinstance(rhs, Expression) else [rhs]
        return Expression(self._all() + [-o for o in r])

    def read(self, fvSchemes):
        for o in self._all():
            o.read(fvSchemes)

    # -- implicit: ONE fused kernel for all terms, in the reference's application order ---------------
    def implicit_terms(self, dt):
        terms = []
        for o in self.spatial + self.temporal:  # Expression::implicitOperation(ls) then (ls, t, dt)
            if o.type != "implicit":
                continue
            c = o.coeff
            if o.kind == "div":
                terms.append(dict(kind=ops.TERM_DIV, scheme=o.scheme or 0, coeff=c.value, coeffView=c.view, faceField=o.faceField.internal))
            elif o.kind == "laplacian":
                terms.append(dict(kind=ops.TERM_LAPLACIAN, coeff=c.value, coeffView=c.view, faceField=o.faceField.internal))
            elif o.kind == "source":
                terms.append(dict(kind=ops.TERM_SOURCE, coeff=c.value, coeffView=c.view, cellField=o.cellField))
            elif o.kind == "ddt":
                terms.append(dict(kind=ops.TERM_DDT, coeff=c.value, coeffView=c.view, cellField=o.field.oldTime().internal, dt=dt))
        return terms

    def assemble(self, t, dt, sp, ls, psi):
        terms = self.implicit_terms(dt)
        if not terms:
            ls.reset()
            return
        ops.assemble(psi.mesh, terms, psi.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs, accumulate=False)

    # -- explicit: source += op (Expression::explicitOperation, dsl/expression.hpp:69-78) --------------
    def has_explicit(self):
        return any(o.type == "explicit" for o in self._all())

    def explicitOperation(self, mesh, ncomp=1):
        shp = (mesh.nCells, 3) if ncomp == 3 else (mesh.nCells,)
        src = torch.zeros(shp, dtype=torch.float64, device="cuda")
        for o in self.spatial + self.temporal:
            if o.type != "explicit":
                continue
            c = o.coeff
            if o.kind == "surfaceIntegrate":
                ops.surface_integrate(mesh, o.faceField.internal, src, c.value, c.view, ops.ADD)
            elif o.kind == "div":
                ops.div(mesh, o.faceField.internal, o.field.internal, o.field.boundary.value, src, o.scheme or 0, c.value, c.view, ops.ADD)
            elif o.kind == "laplacian":
                ops.laplacian(mesh, o.field.internal, o.field.boundary.value, src, c.value, c.view, ops.ADD)
            elif o.kind == "source":
                ops.source_explicit(mesh, o.cellField, o.field.internal, src, c.value, c.view)
            elif o.kind == "ddt":
                raise NotImplementedError("explicit ddt needs dt; use ops.ddt_explicit")
        return src


class imp:
    """dsl::imp (dsl/implicit.hpp)."""
    @staticmethod
    def ddt(phi): return Operator("ddt", "implicit", field=phi)
    @staticmethod
    def div(faceFlux, phi): return Operator("div", "implicit", field=phi, faceField=faceFlux)
    @staticmethod
    def laplacian(gamma, phi): return Operator("laplacian", "implicit", field=phi, faceField=gamma)
    @staticmethod
    def source(coeff, phi): return Operator("source", "implicit", field=phi, cellField=coeff.internal if hasattr(coeff, "internal") else coeff)


class exp:
    """dsl::exp (dsl/explicit.hpp): div(flux) is SurfaceIntegrate."""
    @staticmethod
    def div(faceFlux, phi=None):
        if phi is None:
            return Operator("surfaceIntegrate", "explicit", faceField=faceFlux)
        return Operator("div", "explicit", field=phi, faceField=faceFlux)
    @staticmethod
    def laplacian(gamma, phi): return Operator("laplacian", "explicit", field=phi, faceField=gamma)
    @staticmethod
    def source(coeff, phi): return Operator("source", "explicit", field=phi, cellField=coeff.internal if hasattr(coeff, "internal") else coeff)


@dataclass
class RunTime:
    """FoamAdapter::RunTime (include/FoamAdapter/datastructures/runTime.hpp): t, dt and the dictionaries."""
    mesh: object
    dt: float
    t: float = 0.0
    fvSchemes: dict = _field(default_factory=dict)
    fvSolution: dict = _field(default_factory=dict)
    comm: object = None
    check_every: int = 8
    history: bool = False


class PDESolver:
    """FoamAdapter::PDESolver<T> (expression.hpp:23-179)."""

    def __init__(self, expr: Expression, psi, runTime: RunTime, ls: la.LinearSystem | None = None):
        self.psi, self.expr, self.rt = psi, expr, runTime
        self.sp = la.SparsityPattern.readOrCreate(psi.mesh)
        # the fused assembly writes every entry, so the system needs no zero-fill (createEmptyLinearSystem)
        self.ls = ls if ls is not None else la.LinearSystem(psi.mesh, psi.ncomp, zero=False)
        self.expr.read(runTime.fvSchemes)
        self.needReference, self.pRefCell, self.pRefValue = False, 0, 0.0
        self._solver = None

    def getField(self): return self.psi
    def sparsityPattern(self): return self.sp
    def linearSystem(self): return self.ls

    def assemble(self):
        self.expr.assemble(self.rt.t, self.rt.dt, self.sp, self.ls, self.psi)
        return self.ls

    def setReference(self, pRefCell, pRefValue):
        self.needReference, self.pRefCell, self.pRefValue = True, int(pRefCell), float(pRefValue)

    def prepare(self):
        """Everything of `solve` before the linear solver runs: implicit assembly, rhs -= explicit*V, post-assembly
        functors (SetReference), halo of the initial guess. Kernel launches only (CUDA-graph capturable)."""
        if self.psi.ncomp != 1:
            raise NotImplementedError("only the scalar (pressure) solve is on the hot path; momentumPredictor is 'no'")
        mesh = self.psi.mesh
        self.assemble()
        if self.expr.has_explicit():
            src = self.expr.explicitOperation(mesh, self.psi.ncomp)
            ops.rhs_sub_source(mesh, src, self.ls.rhs)
        if self.needReference:
            ops.set_reference(mesh, self.pRefCell, self.pRefValue, self.ls.values, self.ls.rhs)
        if self.rt.comm is not None:
            self.rt.comm.halo_exchange(self.psi.internal)

    def solve(self, solver: la.Solver | None = None) -> la.SolverStats:
        """iterativeSolveImpl as inferred from dsl::solve (dsl/solver.hpp:60-80): implicit assembly, rhs -= explicit*V,
        post-assembly functors (SetReference), la::Solver.solve."""
        self.prepare()
        if solver is None:
            cfg = self.rt.fvSolution.get("solvers", {}).get(self.psi.name)
            if cfg is None:
                raise KeyError(f"fvSolution.solvers has no entry for '{self.psi.name}'")
            solver = la.Solver(cfg, comm=self.rt.comm, check_every=self.rt.check_every, history=self.rt.history)
        return solver.solve(self.ls, self.psi.internal)
