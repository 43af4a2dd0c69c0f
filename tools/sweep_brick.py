#!/usr/bin/env python
"""Sweep of the brick kernel (GPU box): tile size / brick shape (plan, per mesh creation) x kernel configuration
(cells per thread, threads per block, resident blocks aimed at) on an N^3 mesh; CUDA events, L2 flushed between launches.
Usage: python tools/sweep_brick.py [--mesh 256] [--reps 10] [--out gpurun_out/sweep_brick.jsonl]"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import algorithmic_bytes, mesh_counts, peaks  # noqa: E402
from foamadapter_b200 import _capi, fvcc, ops  # noqa: E402
from foamadapter_b200.mesh import MeshDesc, UnstructuredMesh  # noqa: E402
from tools.roofline import timeit  # noqa: E402

PLANS = [  # (tile cells, brick shape, [(kernel: 1 = brick / 3 = affine, threads, resident blocks aimed at)]) -- instantiated combinations only
    (128, "16,4,2", [(1, 128, 8), (1, 128, 12), (3, 128, 6), (3, 128, 8), (3, 128, 12)]),
    (256, "16,4,4", [(1, 256, 4), (1, 256, 6), (3, 256, 3), (3, 256, 4), (3, 256, 6)]),
    (512, "16,8,4", [(1, 512, 2), (1, 512, 3), (3, 512, 2), (3, 512, 3)]),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "sweep_brick.jsonl"))
    ap.add_argument("--plans", type=int, nargs="*", default=None, help="indices into PLANS")
    args = ap.parse_args()
    n = args.mesh
    peak, kind = peaks()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    nC, nI, nB = mesh_counts(n)
    ab = algorithmic_bytes(n)
    rows = []
    L = _capi.lib()
    d = MeshDesc.block(n, n, n, 0.1, 0.1, 0.01)
    rng = np.random.default_rng(42)
    Th = rng.uniform(1, 2, nC)
    flux = torch.cat([torch.arange(nI, dtype=torch.float64), torch.zeros(nB, dtype=torch.float64)]).cuda()
    out = torch.zeros(nC, dtype=torch.float64, device="cuda")
    out3 = torch.zeros((nC, 3), dtype=torch.float64, device="cuda")
    plans = PLANS if args.plans is None else [PLANS[i] for i in args.plans]
    first = True
    for cells, shape, cfgs in plans:
        os.environ["FVK_BRICK_CELLS"] = str(cells)
        os.environ["FVK_BRICK"] = shape
        gm = UnstructuredMesh(d)
        T = fvcc.VolumeField(gm, "T", 1, [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)])
        T.internal.copy_(torch.from_numpy(Th))
        T.correctBoundaryConditions()
        phi, pb = T.internal, T.boundary.value
        kernels = {"div": lambda: ops.div(gm, flux, phi, pb, out, ops.LINEAR), "grad": lambda: ops.grad(gm, phi, pb, out3),
                   "laplacian": lambda: ops.laplacian(gm, phi, pb, out)}
        todo = [("brick", c) for c in cfgs]
        if first:
            todo.append(("gather", (0, 0, 0)))
            first = False
        for kind_, cfg in todo:
            L.fvk_set_variant(0 if kind_ == "brick" else 5)
            L.fvk_set_brick_config(*cfg)
            L.fvk_set_affine(1 if cfg[0] == 3 else 0)
            row = {"mesh": n, "kernel_kind": kind_, "cells": cells, "brick": shape, "cfg": list(cfg)}
            for name, fn in kernels.items():
                med, best = timeit(fn, args.reps, flush)
                row[name] = {"ms": round(med, 4), "frac": round(ab[name] / med / 1e6 / peak, 4)}
            rows.append(row)
            print(json.dumps(row), flush=True)
        L.fvk_set_variant(0)
        L.fvk_set_brick_config(0, 0, 0)
        L.fvk_set_affine(1)
        del gm, T, phi, pb
        torch.cuda.empty_cache()
    Path(args.out).parent.mkdir(exist_ok=True, parents=True)
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
