#!/usr/bin/env python
"""Roofline harness (GPU box): times every explicit kernel variant on N^3 meshes with CUDA events,
L2 flushed between launches, and prints achieved algorithmic GB/s next to a device copy.
Usage: python tools/roofline.py [--mesh 128 256] [--variants 0 1 2 3] [--reps 20] [--out gpurun_out/roofline.jsonl]
"""
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


def timeit(fn, reps, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for a, b in ev:
        flush.zero_()  # 512 MB write > 126 MB L2
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return float(np.median(ts)), ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", type=int, nargs="+", default=[128, 256])
    ap.add_argument("--variants", type=int, nargs="+", default=[0, 5])
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "roofline.jsonl"))
    args = ap.parse_args()
    Path(args.out).parent.mkdir(exist_ok=True, parents=True)
    peak, kind = peaks()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    # device copy (same method as MEASURED_PEAKS): 1 GiB read + 1 GiB write
    a = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
    b = torch.empty_like(a)
    med, best = timeit(lambda: b.copy_(a), args.reps, flush)
    rows.append({"kernel": "torch_copy_1GiB", "ms": med, "gbs": 2 * a.numel() * 4 / med / 1e6, "best_gbs": 2 * a.numel() * 4 / best / 1e6})
    print(json.dumps(rows[-1]), flush=True)
    del a, b
    for n in args.mesh:
        d = MeshDesc.block(n, n, n, 0.1, 0.1, 0.01)
        gm = UnstructuredMesh(d)
        nC, nI, nB = mesh_counts(n)
        ab = algorithmic_bytes(n)
        ab["div_upwind"] = 16 * nI + 24 * nC + 28 * nB
        ab["surface_integrate"] = 16 * nI + 24 * nC + 12 * nB  # flux 8 + own/nei 8; V,coeff-less out
        rng = np.random.default_rng(42)
        T = fvcc.VolumeField(gm, "T", 1, [("fixedValue", 10.5), ("fixedValue", 1.5), ("zeroGradient", 0.0)])
        T.internal.copy_(torch.from_numpy(rng.uniform(1, 2, nC)))
        T.correctBoundaryConditions()
        flux = torch.cat([torch.arange(nI, dtype=torch.float64), torch.zeros(nB, dtype=torch.float64)]).cuda()
        out = torch.zeros(nC, dtype=torch.float64, device="cuda")
        out3 = torch.zeros((nC, 3), dtype=torch.float64, device="cuda")
        phi, pb = T.internal, T.boundary.value
        kernels = {
            "div": lambda: ops.div(gm, flux, phi, pb, out, ops.LINEAR),
            "div_upwind": lambda: ops.div(gm, flux, phi, pb, out, ops.UPWIND),
            "grad": lambda: ops.grad(gm, phi, pb, out3),
            "laplacian": lambda: ops.laplacian(gm, phi, pb, out),
            "surface_integrate": lambda: ops.surface_integrate(gm, flux, out),
        }
        for v in args.variants:
            _capi.lib().fvk_set_variant(v)
            for name, fn in kernels.items():
                med, best = timeit(fn, args.reps, flush)
                row = {"mesh": n, "variant": v, "kernel": name, "ms": med, "best_ms": best, "alg_bytes": ab[name],
                       "gbs": ab[name] / med / 1e6, "frac_of_" + kind: ab[name] / med / 1e6 / peak,
                       "face_ops_per_s": (nI + nB) / (med * 1e-3)}
                for k in ("FVK_BRICK", "FVK_BRICK_MINB"):
                    if os.environ.get(k):
                        row[k] = os.environ[k]
                rows.append(row)
                print(json.dumps(row), flush=True)
        _capi.lib().fvk_set_variant(0)
        del gm, d, T, flux, out, out3
        torch.cuda.empty_cache()
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
