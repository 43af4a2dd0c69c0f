#!/usr/bin/env python
"""What the cell order does to the generic (any-mesh) kernels: an n^3 block mesh in (a) blockMesh order with the index-free
kernels, (b) blockMesh order with the generic kernels, (c) a RANDOM cell order, (d) that mesh after Morton renumbering, (e)
after reverse Cuthill-McKee. div / grad / laplacian (brick or per-cell gather), fused assembly, SpMV; L2 flushed, CUDA events.
    python tools/renumber_bench.py --mesh 128 --out gpurun_out/x.jsonl"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import alg_bytes_counts, peaks  # noqa: E402
from foamadapter_b200 import fvcc, la, ops  # noqa: E402
from foamadapter_b200._capi import lib  # noqa: E402
from foamadapter_b200.mesh import MeshDesc, UnstructuredMesh  # noqa: E402
from tests.helpers import renumbered_block  # noqa: E402
from tools.roofline import timeit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mesh", type=int, default=128)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "renumber_bench.jsonl"))
args = ap.parse_args()
n = args.mesh
peak, kind = peaks()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
block = MeshDesc.block(n, n, n, 0.1, 0.1, 0.01)
t0 = time.perf_counter(); rnd = renumbered_block(n, n, n, seed=1, box=(0.1, 0.1, 0.01)); t_rnd = time.perf_counter() - t0
t0 = time.perf_counter(); mor = rnd.renumbered("morton")[0]; t_mor = time.perf_counter() - t0
t0 = time.perf_counter(); rcm = rnd.renumbered("rcm")[0]; t_rcm = time.perf_counter() - t0
cases = [("blockMesh order, index-free kernels", block, 1), ("blockMesh order, generic kernels", block, 0), ("random cell order", rnd, 1),
         (f"random -> Morton ({t_mor:.1f} s host)", mor, 1), (f"random -> reverse Cuthill-McKee ({t_rcm:.1f} s host)", rcm, 1)]
with open(args.out, "w") as fo:
    for name, d, affine in cases:
        lib().fvk_set_affine(C.c_int(affine))
        gm = UnstructuredMesh(d)
        nC, nI, nB = gm.nCells, gm.nInternalFaces, gm.nBoundaryFaces
        ab = alg_bytes_counts(nC, nI, nB)
        rng = np.random.default_rng(42)
        T = fvcc.VolumeField(gm, "T", 1, [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)])
        T.internal.copy_(torch.from_numpy(rng.uniform(1, 2, nC))); T.correctBoundaryConditions()
        flux = torch.from_numpy(rng.uniform(-1, 1, nI + nB)).cuda()
        gamma = torch.ones(nI + nB, dtype=torch.float64, device="cuda")
        out, out3 = torch.zeros(nC, dtype=torch.float64, device="cuda"), torch.zeros((nC, 3), dtype=torch.float64, device="cuda")
        ls = la.LinearSystem(gm, 1, zero=False)
        terms = [dict(kind=ops.TERM_DIV, scheme=0, coeff=1.0, faceField=flux), dict(kind=ops.TERM_LAPLACIAN, coeff=-1.0, faceField=gamma),
                 dict(kind=ops.TERM_DDT, coeff=1.0, cellField=T.internal - 1.0, dt=1.0)]
        sp = la.SparsityPattern.readOrCreate(gm)
        x, y = torch.from_numpy(rng.uniform(-1, 1, nC)).cuda(), torch.empty(nC, dtype=torch.float64, device="cuda")
        row = {"mesh": n, "order": name, "affine_plan": bool(gm.size(7))}
        for key, fn in (("div", lambda: ops.div(gm, flux, T.internal, T.boundary.value, out)),
                        ("grad", lambda: ops.grad(gm, T.internal, T.boundary.value, out3)),
                        ("laplacian", lambda: ops.laplacian(gm, T.internal, T.boundary.value, out)),
                        ("assemble_ddt_div_lap", lambda: ops.assemble(gm, terms, T.boundary, ls.values, ls.rhs, ls.bcMatrix, ls.bcRhs)),
                        ("spmv", lambda: la.spmv(sp, ls.values, x, y))):
            med, _ = timeit(fn, args.reps, flush)
            row[key] = {"ms": round(med, 4), "frac_of_" + kind: round(ab[key] / med / 1e6 / peak, 3)}
        print(json.dumps(row), flush=True); fo.write(json.dumps(row) + "\n")
        del gm, T, flux, gamma, out, out3, ls, x, y
        torch.cuda.empty_cache()
lib().fvk_set_affine(C.c_int(1))
