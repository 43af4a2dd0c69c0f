#!/usr/bin/env python
"""HBM bandwidth vs number of concurrent read streams (GPU box): out = sum of NS arrays, one element per thread,
L2 flushed between launches. Usage: python tools/stream_probe.py [--mb 400] [--out gpurun_out/stream_probe.jsonl]"""
import argparse
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from foamadapter_b200._capi import check, lib  # noqa: E402
from tools.roofline import timeit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=400)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "stream_probe.jsonl"))
args = ap.parse_args()
n = args.mb * (1 << 20) // 8
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
arrs = [torch.rand(n, dtype=torch.float64, device="cuda") for _ in range(8)]
out = torch.empty(n, dtype=torch.float64, device="cuda")
L = lib()
L.fvk_probe_streams.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int64, C.c_void_p, C.c_void_p]
ptrs = (C.c_void_p * 8)(*[a.data_ptr() for a in arrs])
rows = []
med, best = timeit(lambda: out.copy_(arrs[0]), args.reps, flush)
rows.append({"kernel": "torch_copy", "ms": med, "gbs": 2 * n * 8 / med / 1e6})
print(json.dumps(rows[-1]), flush=True)
for ns in range(1, 9):
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    fn = lambda: check(L.fvk_probe_streams(ns, ptrs, n, out.data_ptr(), s))
    med, best = timeit(fn, args.reps, flush)
    rows.append({"kernel": "probe_streams", "read_streams": ns, "mb_per_stream": args.mb, "ms": med, "gbs": (ns + 1) * n * 8 / med / 1e6,
                 "best_gbs": (ns + 1) * n * 8 / best / 1e6})
    print(json.dumps(rows[-1]), flush=True)
Path(args.out).parent.mkdir(exist_ok=True, parents=True)
with open(args.out, "w") as f:
    for r in rows:
        f.write(json.dumps(r) + "\n")
