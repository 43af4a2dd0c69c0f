#!/usr/bin/env python
"""Top SASS lines of an .ncu-rep by stall samples / executed instructions (read here, no GPU).
Usage: python tools/ncu_hot.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
for b in blocks:
    h = {k: i for i, k in enumerate(b["hdr"])}
    S, E, SRC = h["# Samples"], h["Instructions Executed"], h["Source"]
    stalls = [k for k in b["hdr"] if k.startswith("stall_") and "Not Issued" not in k]
    tots = sum(float(r[S] or 0) for r in b["rows"]); tote = sum(float(r[E] or 0) for r in b["rows"])
    print("==", b["name"][:100], "samples", tots, "inst", tote)
    agg = {k: sum(float(r[h[k]] or 0) for r in b["rows"]) for k in stalls}
    print("  stall totals:", {k: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(b["rows"], key=lambda r: -float(r[S] or 0))[:N]:
        top = max(stalls, key=lambda k: float(r[h[k]] or 0))
        print(f"  {float(r[S] or 0)/max(tots,1)*100:5.1f}% smp  {float(r[E] or 0)/max(tote,1)*100:5.1f}% inst  {top:18s} {r[SRC][:100]}")
