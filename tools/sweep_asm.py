#!/usr/bin/env python
"""Roofline sweep of the fused assembly kernels at N^3 (GPU box): tile shapes of k_assemble_affine (FVK_ASM_TILE, read per
call), the stencil-driven kernel for comparison, scalar and Vec3 systems, 1- and 2-face-term expressions. CUDA events, L2
flushed between launches. Usage: python tools/sweep_asm.py --mesh 256 --out gpurun_out/x.jsonl"""
import argparse
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import mesh_counts, peaks  # noqa: E402
from foamadapter_b200 import fvcc, la, ops  # noqa: E402
from foamadapter_b200._capi import lib  # noqa: E402
from foamadapter_b200.mesh import MeshDesc, UnstructuredMesh  # noqa: E402
from tools.roofline import timeit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "sweep_asm.jsonl"))
ap.add_argument("--tiles", nargs="*", default=["4,2,4", "4,2,3", "2,2,8", "2,2,6", "4,4,2", "8,2,2", "4,1,8"])
args = ap.parse_args()
n = args.mesh
peak, kind = peaks()
gm = UnstructuredMesh(MeshDesc.block(n, n, n, 0.1, 0.1, 0.01))
nC, nI, nB = mesh_counts(n)
rng = np.random.default_rng(42)
T = fvcc.VolumeField(gm, "T", 1, [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)])
T.internal.copy_(torch.from_numpy(rng.uniform(1, 2, nC))); T.correctBoundaryConditions()
U = fvcc.VolumeField(gm, "U", 3, [("fixedValue", (1.0, 0.0, 0.0)), ("noSlip", 0.0), ("noSlip", 0.0)])
U.internal.copy_(torch.from_numpy(rng.uniform(-1, 1, (nC, 3)))); U.correctBoundaryConditions()
flux = torch.cat([torch.arange(nI, dtype=torch.float64), torch.zeros(nB, dtype=torch.float64)]).cuda()
gamma = torch.ones(nI + nB, dtype=torch.float64, device="cuda")
ls, lsV = la.LinearSystem(gm, 1, zero=False), la.LinearSystem(gm, 3, zero=False)
oldS, oldV = T.internal - 1.0, U.internal - 1.0
terms = lambda old: [dict(kind=ops.TERM_DIV, scheme=0, coeff=1.0, faceField=flux), dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=gamma),
                     dict(kind=ops.TERM_DDT, coeff=1.0, cellField=old, dt=1.0)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
# algorithmic bytes: ddt+div+lap scalar 50/85/52; Vec3: values 7x24, rhs/old 24 each; laplacian only: gamma,dc,magSf 24 + 8 + 2
by = {"s_ddt_div_lap": 50 * nI + 85 * nC + 52 * nB, "v_ddt_div_lap": 50 * nI + (4 + 1 + 8 + 24 + 24 + 7 * 24) * nC + 52 * nB,
      "s_lap": 34 * nI + (4 + 1 + 8 + 7 * 8) * nC + 52 * nB}
cases = {"s_ddt_div_lap": lambda: ops.assemble(gm, terms(oldS), T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs),
         "v_ddt_div_lap": lambda: ops.assemble(gm, terms(oldV), U.boundary, lsV.values, lsV.rhs, lsV.bcMatrix, lsV.bcRhs),
         "s_lap": lambda: ops.assemble(gm, terms(oldS)[1:2], T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)}
rows = []
with open(args.out, "w") as fo:
    for tile in ["stencil"] + args.tiles:
        if tile == "stencil":
            lib().fvk_set_affine(C.c_int(0))
        else:
            lib().fvk_set_affine(C.c_int(1)); os.environ["FVK_ASM_TILE"] = tile
        for name, fn in cases.items():
            med, best = timeit(fn, args.reps, flush)
            r = {"mesh": n, "case": name, "tile": tile, "ms": med, "best_ms": best, "alg_bytes": by[name], "gbs": by[name] / med / 1e6,
                 "frac_of_" + kind: by[name] / med / 1e6 / peak}
            print(json.dumps(r), flush=True); fo.write(json.dumps(r) + "\n")
