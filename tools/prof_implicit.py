#!/usr/bin/env python
"""Tiny driver for ncu: the kernels of the implicit / solver / PISO-glue part of the path on an N^3 mesh.
   ncu --set full --clock-control none --import-source on -k regex:'k_assemble|k_spmv|k_cg_update|k_rAU' -o gpurun_out/x python tools/prof_implicit.py --mesh 256
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from foamadapter_b200 import fvcc, la, ops  # noqa: E402
from foamadapter_b200.mesh import MeshDesc, UnstructuredMesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", type=int, default=256)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--cg-iters", type=int, default=3)
args = ap.parse_args()
n = args.mesh
gm = UnstructuredMesh(MeshDesc.block(n, n, n, 0.1, 0.1, 0.01))
nC, nI, nB = gm.nCells, gm.nInternalFaces, gm.nBoundaryFaces
rng = np.random.default_rng(42)
bcs = [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)]
T = fvcc.VolumeField(gm, "T", 1, bcs)
T.internal.copy_(torch.from_numpy(rng.uniform(1, 2, nC)))
T.correctBoundaryConditions()
U = fvcc.VolumeField(gm, "U", 3, [("fixedValue", (1.0, 0.0, 0.0)), ("noSlip", 0.0), ("noSlip", 0.0)])
U.internal.copy_(torch.from_numpy(rng.uniform(-1, 1, (nC, 3))))
U.correctBoundaryConditions()
flux = torch.cat([torch.arange(nI, dtype=torch.float64), torch.zeros(nB, dtype=torch.float64)]).cuda()
gamma = torch.ones(nI + nB, dtype=torch.float64, device="cuda")
ls = la.LinearSystem(gm, 1, zero=False)
lsV = la.LinearSystem(gm, 3, zero=False)
terms = lambda f: [dict(kind=ops.TERM_DIV, scheme=0, coeff=1.0, faceField=flux), dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=gamma),
                   dict(kind=ops.TERM_DDT, coeff=1.0, cellField=f.internal - 1.0, dt=1.0)]
tS, tV = terms(T), terms(U)
sp = la.SparsityPattern.readOrCreate(gm)
x = torch.from_numpy(rng.uniform(-1, 1, nC)).cuda()
y = torch.empty_like(x)
rAU = torch.empty(nC, dtype=torch.float64, device="cuda")
HbyA = torch.empty((nC, 3), dtype=torch.float64, device="cuda")
for _ in range(args.reps):
    ops.assemble(gm, tS, T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
    ops.assemble(gm, tV, U.boundary, lsV.values, lsV.rhs, lsV.bcMatrix, lsV.bcRhs)
    la.spmv(sp, ls.values, x, y)
    la.spmv_structured(gm, ls.values, x, y)
    ops.rAU_HbyA(gm, lsV.values, lsV.rhs, U.internal, rAU, HbyA)
ops.assemble(gm, [tS[1], tS[2]], T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)
solver = la.Solver({"solver": "Ginkgo", "type": "solver::Cg", "preconditioner": {"type": "preconditioner::Jacobi", "max_block_size": 1},
                    "criteria": {"iteration": args.cg_iters, "relative_residual_norm": 0.0, "absolute_residual_norm": 0.0}}, check_every=args.cg_iters + 1)
xs = torch.zeros(nC, dtype=torch.float64, device="cuda")
solver.solve(ls, xs)
torch.cuda.synchronize()
print("done")
