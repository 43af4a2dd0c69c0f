"""TEST INFRASTRUCTURE -- CPU restatement of one neoIcoFoam time step (reference examples/neoIcoFoam/neoIcoFoam.cpp:80-180
with FoamAdapter src/algorithms/pressureVelocityCoupling.cpp and PDESolver, include/FoamAdapter/datastructures/
expression.hpp:23-179), built from the oracle's per-function restatements in the reference's call order, Serial semantics.
Parity unpinned for whole steps: the reference has no committed PISO golden run; the per-function pins are in
tests/test_oracle_golden.py."""
from __future__ import annotations

import numpy as np

from .cpu import Mesh, bicgstab, cg

FIXED_VALUE, FIXED_GRADIENT, EXTRAPOLATED = 1, 2, 3


class IcoFoamOracle:
    def __init__(self, om: Mesh, nu=0.01, dt=1e-4, lid=(1.0, 0.0, 0.0), nCorrectors=2, tolerance=1e-6, relTol=0.0,
                 maxIter=1000, pRefCell=0, pRefValue=0.0, jacobi=True, momentumPredictor=False, Utolerance=1e-5):
        self.om, self.nu, self.dt = om, nu, dt
        self.nCorr, self.tol, self.relTol, self.maxIter, self.pRefCell, self.pRefValue, self.jacobi = (
            nCorrectors, tolerance, relTol, maxIter, pRefCell, pRefValue, jacobi)
        self.momentumPredictor, self.Utol, self.Ustats = momentumPredictor, Utolerance, []
        nP = len(om.patchOffsets) - 1
        self.Ukinds = [FIXED_VALUE] * nP
        self.Uconsts = [list(lid)] + [[0.0, 0.0, 0.0]] * (nP - 1)
        self.pkinds, self.pconsts = [FIXED_GRADIENT] * nP, [0.0] * nP
        self.ekinds = [EXTRAPOLATED] * nP
        self.U = np.zeros((om.nC, 3)); self.p = np.zeros(om.nC)
        self.Ubd = om.correct_bcs(self.Ukinds, self.Uconsts, self.U)
        self.pbd = om.correct_bcs(self.pkinds, self.pconsts, self.p)
        self.phi, self.phiB = om.flux(self.U, self.Ubd["value"])
        self.stats, self.hist = [], []

    def step(self):
        om, dt = self.om, self.dt
        oldU = self.U.copy()
        nuF = np.full(om.nF, self.nu)
        Uls = om.empty_system(True)
        om.div_imp(Uls, self.phi, self.Ubd, 0, 1.0, None)
        om.laplacian_imp(Uls, nuF, self.Ubd, -1.0, None)
        om.ddt_imp(Uls, oldU, dt, 1.0, None)
        if self.momentumPredictor:
            # neoIcoFoam.cpp:100-103 UEqn.solve(): the Vec3 system component by component (identical matrix components, A.3)
            # with what mapFvSolution makes of fvSolution.solvers.U: solver::Bicgstab + scalar Jacobi, absolute norm 1e-5
            vals0 = np.ascontiguousarray(Uls["values"][:, 0])
            self.Ustats = []
            for c in range(3):
                x, st, _ = bicgstab(om.rowOffs, om.colIdxs, vals0, np.ascontiguousarray(Uls["rhs"][:, c]), np.ascontiguousarray(self.U[:, c]),
                                    jacobi=True, max_iter=1000, rel_tol=0.0, abs_tol=self.Utol)
                self.U[:, c] = x
                self.Ustats.append(st)
            self.Ubd = om.correct_bcs(self.Ukinds, self.Uconsts, self.U)
        out = []
        for _ in range(self.nCorr):
            rAU = om.rAU(Uls["values"])
            HbyA = om.HbyA(Uls["values"], Uls["rhs"], rAU, self.U)
            Hbd = om.correct_bcs(self.ekinds, [[0.0] * 3] * len(self.ekinds), HbyA)
            rbd = om.correct_bcs(self.ekinds, [0.0] * len(self.ekinds), rAU)
            # constrainHbyA: all U patches are fixedValue (non-assignable)
            Hbd["value"][:] = self.Ubd["value"]
            rAUf = om.interpolate(rAU, rbd["value"], 0)
            phiH, phiHB = om.flux(HbyA, Hbd["value"])
            pls = om.empty_system(False)
            om.laplacian_imp(pls, rAUf, self.pbd, 1.0, None)
            src = np.zeros(om.nC)
            tmp = om.surface_integrate(phiH, coeff=-1.0)
            src += tmp
            om.rhs_sub_source(pls, src)
            if self.pRefCell >= 0:
                om.set_reference(pls, self.pRefCell, self.pRefValue)
            x, st, hist = cg(om.rowOffs, om.colIdxs, pls["values"], pls["rhs"], self.p, jacobi=self.jacobi, max_iter=self.maxIter,
                             rel_tol=self.relTol, abs_tol=self.tol, max_hist=self.maxIter + 2)
            self.p = x
            self.pbd = om.correct_bcs(self.pkinds, self.pconsts, self.p)
            self.phi, self.phiB = om.update_face_velocity(pls, self.p, phiH, phiHB)
            gradP = om.grad(self.p, self.pbd["value"])
            self.U = om.update_velocity(HbyA, rAU, gradP)
            self.Ubd = om.correct_bcs(self.Ukinds, self.Uconsts, self.U)
            out.append((st, hist))
        self.stats.append(out)
        return out
