"""scalarAdvection (reference examples/scalarAdvection/scalarAdvection.cpp, BASELINE.json configs[3]) on the B200 kernels:
fields of createFields.H:27-50, the time loop of scalarAdvection.cpp:52-95 / test/test_advection.cpp:118-166,185-227
(old = T; phi = phi0 * cos(pi (t + dt/2) / endTime); computeCoNum; optional setDeltaT; dsl::solve) through the Python
mirror of the DSL. Single GPU or one sub-domain per GPU (ghost cells + halo exchange of T inside dsl.solve)."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import dsl, fvcc, la, ops
from . import mesh as _m
from .mesh import MeshDesc, UnstructuredMesh

# tutorials/scalarAdvection/system/blockMeshDict: fixedWalls = y-max, x-min, x-max, y-min; frontAndBack empty
PATCHES_ADVECTION2D = [("fixedWalls", [3, 0, 1, 2], False), ("frontAndBack", [4, 5], True)]
# 3-D variant of BASELINE configs[3]: the z sides are walls too
PATCHES_ADVECTION3D = [("fixedWalls", [3, 0, 1, 2, 4, 5], False)]
# tutorials/scalarAdvection/system/fvSchemes (div(phi,nfT) Gauss upwind) + the integrators of test/test_advection.cpp
ADVECTION_FVSCHEMES = {"ddtSchemes": {"type": "forwardEuler"}, "divSchemes": {"div(phi,nfT)": "Gauss upwind"}}
# test/test_advection.cpp:176-183
ADVECTION_FVSOLUTION = {"solver": "Ginkgo", "type": "solver::Bicgstab",
                        "preconditioner": {"type": "preconditioner::Jacobi", "max_block_size": 1},
                        "criteria": {"iteration": 20, "relative_residual_norm": 1e-14}}


def advection_desc(n, three_d=False):
    """blockMeshDict of the tutorial: unit square, (NX NX 1) cells, 0.1 thick; or the n^3 unit cube."""
    if three_d:
        return MeshDesc.block(n, n, n, 1.0, 1.0, 1.0, patches=PATCHES_ADVECTION3D)
    return MeshDesc.block(n, n, 1, 1.0, 1.0, 0.1, patches=PATCHES_ADVECTION2D)


def init_fields(C: np.ndarray):
    """createFields.H:27-50, evaluated on the HOST like the reference does (OpenFOAM forAll loop over cell centres) with
    libm, cell by cell: device sin/exp differ from libm in the last bits and the fields are inputs, not part of the path."""
    n = len(C)
    U = np.zeros((n, 3))
    T = np.zeros(n)
    spread, pi = 0.05, math.pi
    sin, pw, ex = math.sin, math.pow, math.exp
    for i in range(n):
        x, y = float(C[i, 0]), float(C[i, 1])
        U[i, 0] = -sin(2.0 * pi * y) * pw(sin(pi * x), 2.0)
        U[i, 1] = sin(2.0 * pi * x) * pw(sin(pi * y), 2.0)
        T[i] = ex(-0.5 * (pw((x - 0.5) / spread, 2.0) + pw((y - 0.75) / spread, 2.0)))
    return U, T


def init_fields_columns(C: np.ndarray, nxy: int):
    """Same values for a block whose cells repeat in z (c = i + nx (j + ny k)): evaluate one x-y layer, tile it."""
    U0, T0 = init_fields(C[:nxy])
    reps = len(C) // nxy
    return np.tile(U0, (reps, 1)), np.tile(T0, reps)


class ScalarAdvection:
    def __init__(self, mesh: UnstructuredMesh, dt, endTime, fvSchemes=None, fvSolution=None, comm=None, U=None, T=None,
                 adjustTimeStep=False, maxCo=0.1, maxDeltaT=1.0, check_every=8, fuse_euler=True):
        self.mesh, self.dt, self.endTime, self.t = mesh, float(dt), float(endTime), 0.0
        self.fvSchemes = fvSchemes or ADVECTION_FVSCHEMES
        self.fvSolution = fvSolution or ADVECTION_FVSOLUTION
        self.comm, self.check_every = comm, check_every
        self.adjustTimeStep, self.maxCo, self.maxDeltaT = adjustTimeStep, maxCo, maxDeltaT
        nP = mesh.nPatches
        if U is None or T is None:
            U, T = init_fields(mesh.to_host(_m.CELL_CENTRES).reshape(-1, 3))
        zg = [("zeroGradient", 0.0)] * nP
        self.U = fvcc.VolumeField(mesh, "U", 3, zg)
        self.T = fvcc.VolumeField(mesh, "nfT", 1, zg)
        self.U.internal.copy_(torch.from_numpy(np.ascontiguousarray(U)))
        self.T.internal.copy_(torch.from_numpy(np.ascontiguousarray(T)))
        self.U.correctBoundaryConditions(); self.T.correctBoundaryConditions()
        self.phi0 = fvcc.SurfaceField(mesh, "phi", 1)
        self.phi = fvcc.SurfaceField(mesh, "phi", 1)
        ops.flux(mesh, self.U.internal, self.U.boundary.value, self.phi0.internal, self.phi0.bvalue)   # createFields.H:42-49
        self.phi.internal.copy_(self.phi0.internal)
        self._co = torch.empty(2, dtype=torch.float64, device="cuda")
        self._coScratch = None
        self.coNum = None
        implicit = self.fvSchemes["ddtSchemes"]["type"] == "backwardEuler"
        make = dsl.imp.div if implicit else dsl.exp.div
        # scalarAdvection.cpp:80 / test_advection.cpp:150-153,213-216
        self.eqn = dsl.imp.ddt(self.T) + make(self.phi, self.T)
        self.eqn.read(self.fvSchemes)
        self.stats = None
        # forwardEuler of `ddt(T) + div(phi, T)`: the old-time copy made at the top of the step IS the operand, so the div kernel can
        # write T = old - dt * div itself (fvk_div_forward_euler_s; no source vector, no separate update pass). fuse_euler=False keeps
        # the generic dsl::solve sequence -- same bits (tests/test_advection_gpu.py).
        self.fuse_euler = fuse_euler and self.fvSchemes["ddtSchemes"]["type"] == "forwardEuler"

    def step(self):
        t, dt = self.t, self.dt
        old = self.T.oldTime()
        old.internal.copy_(self.T.internal)                                          # scalarAdvection.cpp:57-58
        la.scaledCopy(math.cos(math.pi * (t + 0.5 * dt) / self.endTime), self.phi0.internal, self.phi.internal)  # :66-67
        if self._coScratch is None:
            from ._capi import lib
            self._coScratch = torch.empty(lib().fvk_conum_scratch_bytes(self.mesh.handle) // 8, dtype=torch.float64, device="cuda")
        self.coNum = ops.conum(self.mesh, self.phi.internal, dt, self._co, self._coScratch)  # :70 (device scalars)
        if self.adjustTimeStep:                                                     # :73-76, auxiliary/setup.cpp:13-22
            co = float(self._max_conum())
            fact = self.maxCo / (co + 1e-15)
            self.dt = min(min(min(fact, 1.0 + 0.1 * fact), 1.2) * dt, self.maxDeltaT)
        if self.fuse_euler:
            if self.comm is not None:
                self.comm.halo_exchange(old.internal)   # the operand's ghosts (dsl::solve would exchange T's)
            o = self.eqn.spatial[0]
            ops.div_forward_euler(self.mesh, self.phi.internal, old.internal, self.T.boundary.value, dt, self.T.internal, o.scheme or 0, o.coeff.value, o.coeff.view)
            self.T.correctBoundaryConditions()
            self.stats = None
        else:
            self.stats = dsl.solve(self.eqn, self.T, t, dt, self.fvSchemes, self.fvSolution, comm=self.comm, check_every=self.check_every)
        self.t = t + dt
        return self.stats

    def _max_conum(self):
        co = self.coNum.clone()
        if self.comm is not None and self.comm.nRanks > 1:
            self.comm.allreduce_max(co[:1])
        return co[0].item()
