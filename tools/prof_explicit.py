#!/usr/bin/env python
"""Tiny driver for ncu: runs the explicit gather kernels a few times on an N^3 mesh.
   ncu --set full --clock-control none --import-source on -k regex:k_gather -s <skip> -c <n> -o gpurun_out/prof python tools/prof_explicit.py --mesh 256 --variants 1 0
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from foamadapter_b200 import _capi, fvcc, ops  # noqa: E402
from foamadapter_b200.mesh import MeshDesc, UnstructuredMesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", type=int, default=256)
ap.add_argument("--variants", type=int, nargs="+", default=[0])
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
n = args.mesh
d = MeshDesc.block(n, n, n, 0.1, 0.1, 0.01)
gm = UnstructuredMesh(d)
nC, nI, nB = gm.nCells, gm.nInternalFaces, gm.nBoundaryFaces
T = fvcc.VolumeField(gm, "T", 1, [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)])
T.internal.copy_(torch.from_numpy(np.random.default_rng(42).uniform(1, 2, nC)))
T.correctBoundaryConditions()
flux = torch.cat([torch.arange(nI, dtype=torch.float64), torch.zeros(nB, dtype=torch.float64)]).cuda()
out = torch.zeros(nC, dtype=torch.float64, device="cuda")
out3 = torch.zeros((nC, 3), dtype=torch.float64, device="cuda")
for v in args.variants:
    _capi.lib().fvk_set_variant(v)
    for _ in range(args.reps):
        ops.div(gm, flux, T.internal, T.boundary.value, out)
        ops.grad(gm, T.internal, T.boundary.value, out3)
        ops.laplacian(gm, T.internal, T.boundary.value, out)
        ops.surface_integrate(gm, flux, out)
        ops.div(gm, flux, T.internal, T.boundary.value, out, scheme=ops.UPWIND)
        co = torch.empty(2, dtype=torch.float64, device="cuda")
        scratch = torch.empty(_capi.lib().fvk_conum_scratch_bytes(gm.handle) // 8, dtype=torch.float64, device="cuda")
        ops.conum(gm, flux, 1e-3, co, scratch)
torch.cuda.synchronize()
print("done")
