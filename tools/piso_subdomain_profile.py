#!/usr/bin/env python
"""Kernel-level view of ONE rank's share of the decomposed PISO step, on one GPU (for ncu launch lists): the sub-domain of
rank R of the P-way decomposition of the n^3 cavity is stepped WITHOUT a communicator (ghost values stay stale, so the
numbers are timing-only; the pressure solve is capped at a few iterations).
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/piso_subdomain_profile.py --size 256 --ranks 8"""
import argparse
import copy
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from foamadapter_b200 import piso  # noqa: E402
from foamadapter_b200.decomp import Decomposition, default_split  # noqa: E402
from foamadapter_b200.mesh import UnstructuredMesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--ranks", type=int, default=8)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--iters", type=int, default=3)
args = ap.parse_args()
g = piso.cavity_desc(args.size, True)
dec = Decomposition(g, args.ranks, args.rank, n=default_split(args.ranks))
mesh = UnstructuredMesh(dec.desc)
sol = copy.deepcopy(piso.CAVITY_FVSOLUTION)
sol["solvers"]["p"] = {"solver": "PCG", "preconditioner": "DIC", "tolerance": 0.0, "relTol": 0.0, "maxIter": args.iters}
app = piso.IcoFoam(mesh, nu=0.01, dt=1e-4 * 20 / args.size, fvSolution=sol, check_every=args.iters + 1, graphs=False)
for _ in range(args.steps):
    app.step()
torch.cuda.synchronize()
print("cells", mesh.nOwned, "ghosts", mesh.nCells - mesh.nOwned)
