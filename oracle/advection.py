"""TEST INFRASTRUCTURE -- CPU restatement of the reference's scalarAdvection case (BASELINE.json configs[3]):
examples/scalarAdvection/createFields.H:27-50 (fields), examples/scalarAdvection/scalarAdvection.cpp:57-83 and
test/test_advection.cpp:118-166,185-227 (time loop: old = T, phi = phi0 * cos(pi (t + dt/2) / endTime), computeCoNum,
dsl::solve with forwardEuler / backwardEuler / Runge-Kutta Forward-Euler), built from the oracle's per-function
restatements (oracle/fvo.cpp) with Serial semantics.

Parity status: the reference compares against OpenFOAM at 1e-10 (forward Euler) / 1e-8 (backward Euler)
(test_advection.cpp:169,228) but commits no golden field for this case, and the backward-Euler solve goes through Ginkgo's
BiCGStab (third party, not in the tree): whole-run parity is UNPINNED; what is pinned are the building blocks (div, ddt,
interpolation: tests/test_oracle_golden.py) and the properties tests/test_advection.py checks (conservation, bounds,
agreement of the two integrators as dt -> 0).
"""
from __future__ import annotations

import math

import numpy as np

from .cpu import Mesh, bicgstab

FIXED_GRADIENT = 2


def init_fields(C):
    """createFields.H:27-50 / test_advection.cpp:21-47 on cell centres C [n,3]: U = (-sin(2 pi y) sin^2(pi x),
    sin(2 pi x) sin^2(pi y), 0), T = exp(-0.5 (((x-0.5)/s)^2 + ((y-0.75)/s)^2)), s = 0.05. Plain libm calls, cell by cell
    (numpy's vectorised sin/exp may differ from libm in the last bit; the product computes these on the host the same way)."""
    n = len(C)
    U = np.zeros((n, 3))
    T = np.zeros(n)
    spread, pi = 0.05, math.pi
    for i in range(n):
        x, y = float(C[i, 0]), float(C[i, 1])
        U[i, 0] = -math.sin(2.0 * pi * y) * math.pow(math.sin(pi * x), 2.0)
        U[i, 1] = math.sin(2.0 * pi * x) * math.pow(math.sin(pi * y), 2.0)
        T[i] = math.exp(-0.5 * (math.pow((x - 0.5) / spread, 2.0) + math.pow((y - 0.75) / spread, 2.0)))
    return U, T


class ScalarAdvectionOracle:
    """One object = one run. scheme: 0 linear, 1 upwind (fvSchemes divSchemes `div(phi,nfT) Gauss upwind`);
    ddt: 'forwardEuler' | 'backwardEuler' | 'Runge-Kutta' (Forward-Euler table)."""

    def __init__(self, om: Mesh, dt, endTime, scheme=1, ddt="forwardEuler", jacobi=True, maxIter=20, relTol=1e-14, U=None, T=None):
        self.om, self.dt, self.endTime, self.scheme, self.ddt = om, float(dt), float(endTime), scheme, ddt
        self.jacobi, self.maxIter, self.relTol = jacobi, maxIter, relTol
        C = om.C.reshape(-1, 3)
        if U is None or T is None:
            U, T = init_fields(C)
        self.U, self.T = np.array(U, dtype=np.float64), np.array(T, dtype=np.float64)
        nP = len(om.patchOffsets) - 1
        self.kinds = [FIXED_GRADIENT] * nP           # zeroGradient on every patch (tutorials/scalarAdvection/0.orig/{T,U})
        Ubd = om.correct_bcs(self.kinds, [[0.0, 0.0, 0.0]] * nP, self.U)
        self.phi0, _ = om.flux(self.U, Ubd["value"])  # createFields.H:42-49: linearInterpolate(U) & Sf
        self.phi = self.phi0.copy()
        self.Tbd = om.correct_bcs(self.kinds, [0.0] * nP, self.T)
        self.t = 0.0
        self.coNum, self.stats = None, []

    def step(self):
        om, t, dt = self.om, self.t, self.dt
        old = self.T.copy()                                                   # scalarAdvection.cpp:57-58
        self.phi = self.phi0 * math.cos(math.pi * (t + 0.5 * dt) / self.endTime)  # :66-67
        self.coNum = om.conum(self.phi, dt)                                   # :70
        if self.ddt in ("forwardEuler", "Runge-Kutta"):
            src = np.zeros(om.nC)
            tmp = om.div(self.phi, self.T, self.Tbd["value"], self.scheme)    # exp::div: tmp = 0; div(tmp); source += tmp
            src += tmp
            self.T = old - src * dt                                           # forwardEuler.hpp:48
            if self.ddt == "forwardEuler":
                self.Tbd = om.correct_bcs(self.kinds, [0.0] * len(self.kinds), self.T)  # :49 (ERKStep does not correct BCs)
        elif self.ddt == "backwardEuler":
            ls = om.empty_system(False)                                        # backwardEuler.hpp:46-50
            om.div_imp(ls, self.phi, self.Tbd, self.scheme, 1.0, None)
            om.ddt_imp(ls, old, dt, 1.0, None)
            x, st, _ = bicgstab(om.rowOffs, om.colIdxs, ls["values"], ls["rhs"], self.T, jacobi=self.jacobi,
                                max_iter=self.maxIter, rel_tol=self.relTol, abs_tol=0.0)
            self.T = x
            self.stats.append(st)
        else:
            raise KeyError(self.ddt)
        self.t = t + dt
        return self.T
