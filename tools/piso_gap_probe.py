#!/usr/bin/env python
"""Where does a small PISO step's time go? One GPU: (a) a standalone n^3 cavity, (b) rank 0's sub-domain of the 8-way decomposed
(2n)^3 cavity stepped without a communicator (timing only), both with the pressure solve capped at `--iters` iterations;
per-step time with a sync after every step vs back-to-back steps (launch latency hidden), whole-step graph vs eager."""
import argparse
import copy
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from foamadapter_b200 import piso  # noqa: E402
from foamadapter_b200.decomp import Decomposition  # noqa: E402
from foamadapter_b200.mesh import UnstructuredMesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--iters", type=int, default=0)
ap.add_argument("--all-ranks", action="store_true", help="probe every rank's sub-domain (whole-step graph only), not only rank 0's")
args = ap.parse_args()
n = args.size
sol = copy.deepcopy(piso.CAVITY_FVSOLUTION)
sol["solvers"]["p"] = {"solver": "PCG", "preconditioner": "DIC", "tolerance": 0.0 if args.iters else 1e30, "relTol": 0.0, "maxIter": max(args.iters, 1)}


def run(mesh, label, graphs):
    app = piso.IcoFoam(mesh, nu=0.01, dt=1e-4 * 20 / (2 * n), fvSolution=sol, check_every=4, graphs=graphs)
    for _ in range(4):
        app.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        e0.record(); app.step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    e0.record()
    for _ in range(40):
        app.step()
    e1.record(); torch.cuda.synchronize()
    its = [s.numIter for s in app.stats[-1]]
    print(json.dumps({"case": label, "graphs": "whole step" if app._whole is not None else ("segments" if app._captured else "eager"),
                      "cells": mesh.nOwned, "ghosts": mesh.nCells - mesh.nOwned, "ms_synced_per_step": round(float(np.median(ts)), 4),
                      "ms_back_to_back": round(e0.elapsed_time(e1) / 40, 4), "cg_iterations": its}), flush=True)


standalone = UnstructuredMesh(piso.cavity_desc(n, True))
for g in (True, False):
    run(standalone, f"standalone {n}^3", g)
del standalone
big = piso.cavity_desc(2 * n, True)
for r in (range(8) if args.all_ranks else (0,)):
    dec = Decomposition(big, 8, r, n=(2, 2, 2))
    sub = UnstructuredMesh(dec.desc)
    from foamadapter_b200 import mesh as _m
    flags = {"rowsInStencilOrder": sub.size(_m.ROWS_IN_STENCIL_ORDER), "affine": sub.size(_m.AFFINE_TOPOLOGY)}
    for g in ((True,) if args.all_ranks else (True, False)):
        run(sub, f"rank {r} of 8 of {2 * n}^3, no communicator, flags {flags}", g)
    del sub, dec
