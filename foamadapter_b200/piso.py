"""Python mirror of FoamAdapter's PISO helpers (src/algorithms/pressureVelocityCoupling.cpp) and of the
neoIcoFoam time loop (examples/neoIcoFoam/neoIcoFoam.cpp:80-180) on a synthetic lid-driven cavity. Function names
and argument meaning follow the reference; bodies call the CUDA kernels through the C ABI."""
from __future__ import annotations

import torch

from . import dsl, fvcc, la, ops
from . import mesh as _m
from .mesh import MeshDesc, PATCHES_CAVITY2D, PATCHES_CAVITY3D, UnstructuredMesh


def _extrapolated(mesh):
    return [("extrapolated", 0.0)] * mesh.nPatches


def computeRAU(expr: dsl.PDESolver, rAU: fvcc.VolumeField | None = None):
    """pressureVelocityCoupling.cpp:38-63"""
    mesh = expr.psi.mesh
    rAU = rAU or fvcc.VolumeField(mesh, "rAU", 1, _extrapolated(mesh))
    ops.rAU_HbyA(mesh, expr.ls.values, None, None, rAU.internal, None)
    return rAU


def computeRAUandHByA(expr: dsl.PDESolver, rAU=None, HbyA=None):
    """pressureVelocityCoupling.cpp:65-128 (one fused kernel + the two extrapolated BC launches)."""
    mesh = expr.psi.mesh
    rAU = rAU or fvcc.VolumeField(mesh, "rAU", 1, _extrapolated(mesh))
    HbyA = HbyA or fvcc.VolumeField(mesh, "HbyA", 3, _extrapolated(mesh))
    ops.rAU_HbyA(mesh, expr.ls.values, expr.ls.rhs, expr.psi.internal, rAU.internal, HbyA.internal)
    HbyA.correctBoundaryConditions()
    rAU.correctBoundaryConditions()
    return rAU, HbyA


def constrainHbyA(U: fvcc.VolumeField, p: fvcc.VolumeField, HbyA: fvcc.VolumeField):
    """pressureVelocityCoupling.cpp:14-36"""
    mask = [not U.assignable(i) for i in range(U.mesh.nPatches)]
    if any(mask):
        ops.copy_patches(U.mesh, mask, U.boundary.value, HbyA.boundary.value)


def flux(volField: fvcc.VolumeField, out: fvcc.SurfaceField | None = None):
    """pressureVelocityCoupling.cpp:215-267"""
    out = out or fvcc.SurfaceField(volField.mesh, "out", 1)
    ops.flux(volField.mesh, volField.internal, volField.boundary.value, out.internal, out.bvalue)
    return out


def updateFaceVelocity(predictedPhi: fvcc.SurfaceField, expr: dsl.PDESolver, phi: fvcc.SurfaceField):
    """pressureVelocityCoupling.cpp:131-197"""
    ls = expr.ls
    ops.update_face_velocity(phi.mesh, ls.values, ls.bcMatrix, ls.bcRhs, expr.psi.internal, predictedPhi.internal,
                             predictedPhi.bvalue, phi.internal, phi.bvalue)


def updateVelocity(HbyA, rAU, p, U, gradP=None):
    """pressureVelocityCoupling.cpp:199-213: U = HbyA - rAU * grad(p). One fused kernel (the gradient never goes to memory);
    passing a gradP buffer keeps the two-kernel form (GaussGreenGrad::grad, then the cell loop) -- same bits either way."""
    mesh = U.mesh
    if gradP is None:
        ops.update_velocity_grad(mesh, HbyA.internal, rAU.internal, p.internal, p.boundary.value, U.internal)
        return
    ops.grad(mesh, p.internal, p.boundary.value, gradP, ops.SET)
    ops.update_velocity(mesh, HbyA.internal, rAU.internal, gradP, U.internal)


# tutorials/cavity/system/{fvSchemes,fvSolution}
CAVITY_FVSCHEMES = {"ddtSchemes": {"type": "backwardEuler"}, "divSchemes": {"div(phi,U)": "Gauss linear"},
                    "laplacianSchemes": {"laplacian(nu,U)": "Gauss linear uncorrected", "laplacian(rAUf,p)": "Gauss linear uncorrected"}}
CAVITY_FVSOLUTION = {"solvers": {"p": {"solver": "PCG", "preconditioner": "DIC", "tolerance": 1e-6, "relTol": 0.0},
                                 "U": {"solver": "smoothSolver", "smoother": "symGaussSeidel", "tolerance": 1e-5, "relTol": 0.0}},
                     "PISO": {"nCorrectors": 2, "nNonOrthogonalCorrectors": 0, "momentumPredictor": False, "pRefCell": 0, "pRefValue": 0.0}}


def cavity_desc(n, three_d=False, L=0.1):
    """tutorials/cavity/system/blockMeshDict: N x N x 1 (scale 0.1, front/back empty) or the 3-D N^3 variant of
    BASELINE.json configs[4]."""
    if three_d:
        return MeshDesc.block(n, n, n, L, L, L, patches=PATCHES_CAVITY3D)
    return MeshDesc.block(n, n, 1, L, L, 0.1 * L, patches=PATCHES_CAVITY2D)


def mesh_rows_in_stencil_order(mesh) -> bool:
    """the compact momentum matrix needs fvk_rAU_HbyA_c, i.e. CSR rows laid out like the stencil (any mesh in OpenFOAM face order)"""
    return bool(mesh.size(_m.ROWS_IN_STENCIL_ORDER))


class IcoFoam:
    """neoIcoFoam (examples/neoIcoFoam/neoIcoFoam.cpp) on a lid-driven cavity: U = (1 0 0) on movingWall, noSlip on
    fixedWalls, p zeroGradient everywhere (tutorials/cavity/0.orig/{U,p}), nu uniform, momentumPredictor no."""

    def __init__(self, mesh: UnstructuredMesh, nu=0.01, dt=1e-4, fvSolution=None, fvSchemes=None, comm=None,
                 lid=(1.0, 0.0, 0.0), history=False, check_every=8, graphs=True, compact_momentum=True, whole_step_graph=True, fuse_interpolate=True):
        self.mesh = mesh
        fvSolution = fvSolution or CAVITY_FVSOLUTION
        self.rt = dsl.RunTime(mesh, dt, 0.0, fvSchemes or CAVITY_FVSCHEMES, fvSolution, comm, check_every, history)
        self.piso = fvSolution["PISO"]
        nP = mesh.nPatches
        self.U = fvcc.VolumeField(mesh, "U", 3, [("fixedValue", lid)] + [("noSlip", 0.0)] * (nP - 1))
        self.p = fvcc.VolumeField(mesh, "p", 1, [("zeroGradient", 0.0)] * nP)
        self.U.correctBoundaryConditions(); self.p.correctBoundaryConditions()
        self.nu = fvcc.SurfaceField(mesh, "nu", 1); self.nu.internal.fill_(nu); self.nu.bvalue.fill_(nu)
        self.phi = fvcc.SurfaceField(mesh, "phi", 1)
        flux(self.U, self.phi)  # createFields.H / createPhi.H: phi = linearInterpolate(U) & Sf
        # persistent work fields (the reference allocates these every corrector)
        self.rAU = fvcc.VolumeField(mesh, "rAU", 1, _extrapolated(mesh))
        self.HbyA = fvcc.VolumeField(mesh, "HbyA", 3, _extrapolated(mesh))
        # rAUf = linearInterpolate(rAU) (neoIcoFoam.cpp:117-124) is only ever the pressure laplacian's diffusivity: it is
        # interpolated inside the assembly kernel instead of being written and re-read (fuse_interpolate=False: materialised)
        self.rAUf = fvcc.InterpolatedSurfaceField(self.rAU, "rAUf") if fuse_interpolate else fvcc.SurfaceField(mesh, "rAUf", 1)
        self.phiHbyA = fvcc.SurfaceField(mesh, "phiHbyA", 1)
        self.Uls = la.LinearSystem(mesh, 3, zero=False, compact=compact_momentum and mesh_rows_in_stencil_order(mesh))
        self.pls = la.LinearSystem(mesh, 1, zero=False)
        self.linear = fvcc.SurfaceInterpolation(mesh, "linear")
        self.solver = la.Solver(fvSolution["solvers"]["p"], comm=comm, check_every=check_every, history=history)
        # momentumPredictor yes (neoIcoFoam.cpp:100-103): the Vec3 momentum system, component by component, with the solver
        # mapFvSolution gives for `U` (tutorials/cavity/system/fvSolution: smoothSolver -> solver::Bicgstab + scalar Jacobi)
        self.Usolver = (la.Solver(fvSolution["solvers"]["U"], comm=comm, check_every=check_every)
                        if self.piso.get("momentumPredictor", False) else None)
        self.stats, self.Ustats = [], None
        from ._capi import lib
        self._co = torch.empty(2, dtype=torch.float64, device="cuda")
        self._coScratch = torch.empty(lib().fvk_conum_scratch_bytes(mesh.handle) // 8, dtype=torch.float64, device="cuda")
        self.coNum = None
        self.graphs, self._captured, self._nsteps = bool(graphs), None, 0
        self.timing, self.segment_ms = False, None
        self.whole_step_graph, self._whole, self._whole_stats, self._whole_failed = bool(whole_step_graph), None, None, False

    def _halo(self, *ts):
        """processor-boundary exchange of one or several cell fields (one exchange for all of them)"""
        comm = self.rt.comm
        if comm is None:
            return
        if len(ts) == 1:
            comm.halo_exchange(ts[0])
        else:
            comm.halo_exchange_multi(list(ts))

    # ---- the step in segments; everything except the linear solvers is plain kernel launches --------------------------
    def _momentum(self):
        """neoIcoFoam.cpp:84-109 up to the momentum solve: old time level, CoNum, UEqn assembly."""
        rt, mesh, U, phi = self.rt, self.mesh, self.U, self.phi
        U.oldTime().internal.copy_(U.internal)                             # neoIcoFoam.cpp:84-85
        self.coNum = ops.conum(mesh, phi.internal, rt.dt, self._co, self._coScratch)  # :87 (device scalars; no host sync here)
        if self._nsteps == 0:
            # later steps: U's ghosts are current since the exchange that ended the previous step, p's since its last solve
            self._halo(U.internal, self.p.internal)
            if self.rt.comm is not None and self.solver.keeps_ghosts:
                self.solver.set_ghosts_current(True)  # p is only ever written by this solver: its guesses keep current ghosts
        self._UEqn = dsl.PDESolver(dsl.imp.ddt(U) + dsl.imp.div(phi, U) - dsl.imp.laplacian(self.nu, U), U, rt, ls=self.Uls)
        self._UEqn.assemble()                                              # :94-98 / :105-108

    def _pre(self):
        """neoIcoFoam.cpp:112-153: one PISO corrector up to (not including) pEqn.solve's linear solver."""
        U, p = self.U, self.p
        rAU, HbyA = computeRAUandHByA(self._UEqn, self.rAU, self.HbyA)     # :114
        constrainHbyA(U, p, HbyA)                                          # :115
        self._halo(rAU.internal, HbyA.internal)
        if not isinstance(self.rAUf, fvcc.InterpolatedSurfaceField):
            self.linear.interpolate(rAU, self.rAUf)                        # :117-124
        flux(HbyA, self.phiHbyA)                                           # :126
        self._p_prepare()

    def _p_prepare(self):
        """pEqn of one (non-orthogonal) corrector up to the linear solver (:140-153)."""
        rt, p = self.rt, self.p
        self._pEqn = dsl.PDESolver(dsl.imp.laplacian(self.rAUf, p) - dsl.exp.div(self.phiHbyA), p, rt, ls=self.pls)
        if self.piso.get("pRefCell", -1) >= 0 and (rt.comm is None or rt.comm.rank == self.piso.get("pRefRank", 0)):
            self._pEqn.setReference(self.piso["pRefCell"], self.piso["pRefValue"])  # :150-153
        self._pEqn.prepare(exchange_guess=False)  # the linear solver exchanges the ghosts of its initial guess itself

    def _post(self, last_nonorth: bool = True):
        """neoIcoFoam.cpp:156-167 after the linear solver."""
        U, p = self.U, self.p
        if not self.solver.keeps_ghosts:
            self._halo(p.internal)  # (solver::Cg over peer memory hands the solution back with current ghosts)
        p.correctBoundaryConditions()                                      # :156
        if last_nonorth:
            updateFaceVelocity(self.phiHbyA, self._pEqn, self.phi)         # :160
            updateVelocity(self.HbyA, self.rAU, p, U)                      # :166 (fused with the gradient)
            U.correctBoundaryConditions()                                  # :167
            self._halo(U.internal)

    def _plan(self):
        """One time step as a list of kernel-only segments (each becomes one CUDA graph) separated by the linear solves:
        [momentum] (solve U) [pre] solve p [post + pre] solve p ... [post]."""
        nC, nN = self.piso["nCorrectors"], self.piso["nNonOrthogonalCorrectors"]
        plan, cur = [], [self._momentum]
        if self.piso.get("momentumPredictor", False):
            plan += [("kernels", cur), ("solveU",)]                        # :100-103
            cur = [self._after_momentum_solve]
        for c in range(nC):
            for k in range(nN + 1):
                cur.append(self._pre if k == 0 else self._p_prepare)
                plan += [("kernels", cur), ("solveP",)]
                cur = [lambda k=k: self._post(k == nN)]
        plan.append(("kernels", cur))
        return plan

    def _after_momentum_solve(self):
        self.U.correctBoundaryConditions()
        self._halo(self.U.internal)

    def _graphs_allowed(self):
        comm = self.rt.comm
        return (self.graphs and self.piso["nCorrectors"] >= 1
                and (comm is None or comm.nRanks == 1 or comm.p2p))  # NCCL calls are kept out of captures

    def step(self):
        """One time step. The first two steps run eagerly (warm-up: lazy allocations, kernel attributes); from the third
        on the segments between the linear solves are replayed as CUDA graphs (all fields are persistent, dt is fixed,
        the peer-memory halo kernels keep their sequence numbers on the device), which removes ~60 host-driven launches
        per step. `graphs=False` or the NCCL transport keep the eager path."""
        rt = self.rt
        plan = self._plan()
        segs = [item[1] for item in plan if item[0] == "kernels"]
        use_graphs = self._graphs_allowed() and self._nsteps >= 2
        # whole-step graph: the pressure solves are captured too (conditional WHILE nodes, device-side stopping test), so a
        # time step is ONE graph launch with no host round trip. Needs solver::Cg, no residual history, no momentum solve.
        whole = (use_graphs and self.whole_step_graph and not self.timing and self.solver.type == "solver::Cg" and not self.solver.history
                 and not self.piso.get("momentumPredictor", False))
        if whole and self._whole is None and self._whole_failed is False:
            try:
                torch.cuda.synchronize()
                self.solver._attach(self.mesh)
                self.solver.reset_captures()
                g = torch.cuda.CUDAGraph()
                lazy = []
                with torch.cuda.graph(g):
                    for item in plan:
                        if item[0] == "kernels":
                            for f in item[1]:
                                f()
                        else:
                            lazy.append(self.solver.solve(self.pls, self.p.internal))
                self._whole, self._whole_stats = g, lazy
            except Exception as e:
                import sys
                print(f"[piso] whole-step CUDA-graph capture failed ({e!r}); using per-segment graphs", file=sys.stderr)
                self._whole, self._whole_failed = None, True
                torch.cuda.synchronize()
        if whole and self._whole is not None:
            self._whole.replay()
            self.stats.append(list(self._whole_stats))
            self.Ustats = None
            rt.t += rt.dt
            self._nsteps += 1
            return self.stats[-1]
        if use_graphs and self._captured is None:
            try:
                torch.cuda.synchronize()
                captured = []
                for seg in segs:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for f in seg:
                            f()
                    captured.append(g)
                self._captured = captured
            except Exception as e:  # not capturable on this setup: stay eager, loudly
                import sys
                print(f"[piso] CUDA-graph capture failed ({e!r}); continuing without graphs", file=sys.stderr)
                self.graphs, self._captured, use_graphs = False, None, False
                torch.cuda.synchronize()
        replay = use_graphs and self._captured
        self.stats.append([])
        self.Ustats = None
        k = 0
        evs = [] if self.timing else None
        for item in plan:
            if evs is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); evs.append((item[0], e))
            if item[0] == "kernels":
                if replay:
                    self._captured[k].replay()
                else:
                    for f in item[1]:
                        f()
                k += 1
            elif item[0] == "solveU":
                self.Ustats = self.Usolver.solve(self.Uls, self.U.internal)        # :102 UEqn.solve()
            else:
                self.stats[-1].append(self.solver.solve(self.pls, self.p.internal))   # :155
        if evs is not None:  # per-segment device times of this step (diagnostics: profiles/r2_piso_segments.jsonl)
            e = torch.cuda.Event(enable_timing=True); e.record(); evs.append(("end", e))
            torch.cuda.synchronize()
            self.segment_ms = [(evs[i][0], evs[i][1].elapsed_time(evs[i + 1][1])) for i in range(len(evs) - 1)]
        rt.t += rt.dt
        self._nsteps += 1
        return self.stats[-1]
