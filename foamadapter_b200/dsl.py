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
                if isinstance(o.faceField, fvcc.InterpolatedSurfaceField):   # gamma interpolated on the fly
                    src = o.faceField.src
                    terms.append(dict(kind=ops.TERM_LAPLACIAN, coeff=c.value, coeffView=c.view, gammaCell=src.internal, gammaBoundary=src.boundary.value))
                else:
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

    def explicitOperationInto(self, mesh, src, fresh=False):
        """Expression::explicitOperation(source) (dsl/expression.hpp:55-65): the explicit SPATIAL operators added to an
        existing source vector (each operator: tmp = 0; op(tmp); source += tmp -- fused into one ADD-mode launch).
        fresh=True: `src` is the zero vector of explicitOperation(nCells) and has NOT been filled -- the first gather operator
        writes it (SET mode: 0 + x without the zero-fill and the read-back), the others add."""
        mode = ops.SET if fresh else ops.ADD
        for o in self.spatial:
            if o.type != "explicit":
                continue
            c = o.coeff
            if o.kind == "surfaceIntegrate":
                ops.surface_integrate(mesh, o.faceField.internal, src, c.value, c.view, mode)
            elif o.kind == "div":
                ops.div(mesh, o.faceField.internal, o.field.internal, o.field.boundary.value, src, o.scheme or 0, c.value, c.view, mode)
            elif o.kind == "laplacian":
                ops.laplacian(mesh, o.field.internal, o.field.boundary.value, src, c.value, c.view, mode)
            elif o.kind == "source":
                if mode == ops.SET:
                    src.zero_()
                ops.source_explicit(mesh, o.cellField, o.field.internal, src, c.value, c.view)
            mode = ops.ADD
        if mode == ops.SET:  # no explicit operator at all
            src.zero_()
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

    def prepare(self, exchange_guess=True):
        """Everything of `solve` before the linear solver runs: implicit assembly, rhs -= explicit*V, post-assembly
        functors (SetReference, scalar systems only: expression.hpp:120-132), halo of the initial guess. Kernel launches
        only (CUDA-graph capturable)."""
        mesh = self.psi.mesh
        self.assemble()
        if self.expr.has_explicit():
            ex = [o for o in self.expr.spatial + self.expr.temporal if o.type == "explicit"]
            if len(ex) == 1 and ex[0].kind == "surfaceIntegrate" and self.psi.ncomp == 1:
                # the pressure equation's `- exp::div(phiHbyA)`: source and rhs update in one pass, same bits
                ops.rhs_sub_surface_integrate(mesh, ex[0].faceField.internal, self.ls.rhs, ex[0].coeff.value, ex[0].coeff.view)
            else:
                src = self.expr.explicitOperation(mesh, self.psi.ncomp)
                ops.rhs_sub_source(mesh, src, self.ls.rhs)
        if self.needReference and self.psi.ncomp == 1:
            ops.set_reference(mesh, self.pRefCell, self.pRefValue, self.ls.values, self.ls.rhs)
        if exchange_guess and self.rt.comm is not None:
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


# ---- time integration (src/NeoN/include/NeoN/timeIntegration/*.hpp) and dsl::solve (dsl/solver.hpp:35-82) -------------
class TimeIntegratorBase:
    """timeIntegration::TimeIntegratorBase<SolutionVectorType> (timeIntegration.hpp:19-52): strategies registered by name
    (RuntimeSelectionFactory, core/runtimeSelectionFactory.hpp:193-413), selected by ddtSchemes.type. A subclass registers
    itself by defining `name` -- the Python spelling of `Register<Derived>`."""
    table: dict = {}
    name = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if cls.name:
            TimeIntegratorBase.table[cls.name] = cls

    def __init__(self, schemeDict: dict, solutionDict: dict, comm=None, check_every=8):
        self.schemeDict, self.solutionDict, self.comm, self.check_every = schemeDict, solutionDict, comm, check_every

    @classmethod
    def create(cls, key, schemeDict, solutionDict, **kw):
        if key not in cls.table:
            raise KeyError(f"unknown time integrator '{key}' (registered: {sorted(cls.table)})")  # keyExistsOrError
        return cls.table[key](schemeDict, solutionDict, **kw)

    def _source(self, eqn, sol):
        """zero source + the explicit spatial operators (Expression::explicitOperation(nCells), expression.hpp:48-65);
        the work vector is kept with the field instead of being allocated every step"""
        src = getattr(sol, "_tiSource", None)
        if src is None:
            src = sol._tiSource = torch.zeros_like(sol.internal)   # ghost slots stay zero: the operators write owned cells
        return eqn.explicitOperationInto(sol.mesh, src, fresh=True)


class ForwardEuler(TimeIntegratorBase):
    """forwardEuler.hpp:38-56: solution = old - source*dt; correctBoundaryConditions."""
    name = "forwardEuler"

    def solve(self, eqn, sol, t, dt):
        src = self._source(eqn, sol)
        old = sol.oldTime()
        la.waxpby(-dt, src, 1.0, old.internal, sol.internal)   # old - source*dt, bit for bit, in one pass
        sol.correctBoundaryConditions()


class RungeKutta(TimeIntegratorBase):
    """rungeKutta.cpp:34-57 with SUNDIALS ERKStep, fixed step, table ddtSchemes.`Runge-Kutta-Method`. Only the 1-stage
    Forward-Euler table is usable in the reference (sundials.hpp:59-78: Heun / Midpoint exit with "Currently unsupported");
    one ERK step with it is y + dt*f(t, y), f = -explicitOperation (sundials.hpp:196-215), i.e. forwardEuler without the
    boundary correction, followed by old = new (rungeKutta.cpp:55-56)."""
    name = "Runge-Kutta"

    def __init__(self, schemeDict, solutionDict, **kw):
        super().__init__(schemeDict, solutionDict, **kw)
        method = schemeDict.get("Runge-Kutta-Method")
        if method in ("Heun", "Midpoint"):
            raise RuntimeError("Currently unsupported until field time step-stage indexing resolved.")
        if method != "Forward-Euler":
            raise RuntimeError(f"Unsupported Runge-Kutta time integration method selectied: {method}.\n"
                               "Supported methods are: Forward-Euler, Heun, Midpoint.")

    def solve(self, eqn, sol, t, dt):
        old = sol.oldTime()
        src = self._source(eqn, sol)
        la.waxpby(-dt, src, 1.0, old.internal, sol.internal)
        old.internal.copy_(sol.internal)


class BackwardEuler(TimeIntegratorBase):
    """backwardEuler.hpp:41-60: the explicit source is evaluated and DROPPED (sic, :45), a fresh system takes the implicit
    spatial then temporal operators (:49-50; one fused assembly launch here), la::Solver(solutionDict).solve."""
    name = "backwardEuler"

    def solve(self, eqn, sol, t, dt):
        mesh = sol.mesh
        if eqn.has_explicit():
            self._source(eqn, sol)  # computed and dropped like the reference
        ls = getattr(sol, "_tiSystem", None)
        if ls is None:
            ls = sol._tiSystem = la.LinearSystem(mesh, sol.ncomp, zero=False)
        eqn.assemble(t, dt, ls.sp, ls, sol)
        solver = getattr(sol, "_tiSolver", None)
        if solver is None or solver._config is not self.solutionDict:
            solver = sol._tiSolver = la.Solver(self.solutionDict, comm=self.comm, check_every=self.check_every)
            solver._config = self.solutionDict
        self.stats = solver.solve(ls, sol.internal)  # (the distributed solver exchanges the ghosts of its initial guess)
        return self.stats


def solve(exp: Expression, solution, t, dt, fvSchemes: dict, fvSolution: dict, comm=None, check_every=8):
    """dsl::solve (dsl/solver.hpp:35-82). fvSolution is the solver dictionary of the field (Ginkgo-style, or OpenFOAM-style
    and then mapped like FoamAdapter::mapFvSolution). Multi-GPU (new here): `comm` exchanges the ghost values of the
    solution field before the operators read them and is handed to the linear solver."""
    if not exp.temporal and not exp.spatial:
        raise RuntimeError("No temporal or implicit terms to solve.")
    exp.read(fvSchemes)
    if comm is not None:
        comm.halo_exchange(solution.internal)
    if exp.temporal:
        ddt = fvSchemes.get("ddtSchemes", {})
        if "type" not in ddt:
            raise KeyError("Key type not found in dictionary")
        ti = TimeIntegratorBase.create(ddt["type"], ddt, fvSolution, comm=comm, check_every=check_every)
        return ti.solve(exp, solution, t, dt)
    mesh = solution.mesh
    ls = la.LinearSystem(mesh, solution.ncomp, zero=False)
    exp.assemble(t, dt, ls.sp, ls, solution)
    src = torch.zeros_like(solution.internal)
    exp.explicitOperationInto(mesh, src)
    ops.rhs_sub_source(mesh, src, ls.rhs)
    return la.Solver(fvSolution, comm=comm, check_every=check_every).solve(ls, solution.internal)
