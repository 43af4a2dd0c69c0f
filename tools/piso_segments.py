#!/usr/bin/env python
"""Per-kernel roofline table of ONE PISO step (profiles/r2_piso_segments.jsonl): the ncu launch list of the step (durations) joined with
each kernel's algorithmic bytes (the arrays the reference's algorithm for that stage reads / writes once, SURVEY.md 8d conventions:
owner / neighbour labels and matrix column indices count even where the index-free kernels do not read them), plus the per-segment
device times of the bench lines. Usage: piso_segments.py <launches.csv> <label> <nC> <nI> <nB> [bench.json ...]"""
import csv
import json
import re
import sys
from collections import OrderedDict

PEAK = 6541.8  # MEASURED_PEAKS.json hbm_gbs


def alg_bytes(nC, nI, nB):
    nF, nnz = nI + nB, nC + 2 * nI
    return OrderedDict([
        ("k_conum_regular|k_conum_list|k_conum_stage1", ("CoNum: |phi| over the faces of every cell, V", 8 * nF + 8 * nC)),
        ("k_assemble_affine<S3|k_assemble_fast<S3", ("UEqn = ddt + div(phi) - laplacian(nu), compact Vec3 matrix (values once, rhs Vec3)",
                                                      50 * nI + (4 + 1 + 8 + 24 + 24 + 7 * 8) * nC + 52 * nB)),
        ("k_rAU_HbyA_rows", ("rAU = 1/diag, HbyA = rAU (b - H(U)): matrix values + columns, rhs, U, V in; rAU, HbyA out", 12 * nnz + 88 * nC)),
        ("k_flux", ("phiHbyA = (HbyA_f . Sf): HbyA, Sf, weights, labels in; flux out", 24 * nC + 48 * nI + 56 * nB)),
        ("k_assemble_affine<S1|k_assemble_fast<S1", ("pEqn = laplacian(rAU_f, p), rAU interpolated inside: rAU, weights, deltaCoeffs, magSf, labels in; values, rhs out",
                                                      8 * nnz + 16 * nC + 32 * nI + 52 * nB)),
        ("SurfIntOp", ("pEqn rhs -= div(phiHbyA): flux, labels, V in; rhs in/out", 16 * nI + 24 * nC + 8 * nB)),
        ("k_spmv<2>", ("CG start-up r0 = b - A p, ||b||^2, 1/diag: values + columns, x, b in; r, dinv out", 12 * nnz + 36 * nC)),
        ("k_cg_update<1", ("CG first update: r, dinv in; z out; r.z, r.r", 24 * nC)),
        ("k_spmv<4>", ("CG iteration, K2: p = z + beta p, q = A p, p.q (this capture runs ONE iteration per solve; a second launch is the no-op behind the stop flag)", 12 * nnz + 28 * nC)),
        ("k_cg_update<0", ("CG iteration, K1: x += alpha p, r -= alpha q, z = r / diag, r.z, r.r", 64 * nC)),
        ("k_update_face_velocity", ("phi = phiHbyA - pEqn.flux(): flux, matrix coefficients, p, labels in; phi out", 8 * nC + 33 * nI + 40 * nB)),
        ("GradOp", ("U = HbyA - rAU grad(p) (gradient not stored): p, Sf, weights, labels, V, HbyA, rAU in; U out", 40 * nI + 72 * nC + 44 * nB)),
    ])


def main():
    path, label, nC, nI, nB = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    rows, hdr = [], None
    for r in csv.reader(open(path, errors="ignore")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            rows.append(dict(zip(hdr, r)))
    names = [re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"])).replace("<unnamed>::", "") for r in rows]
    dur = [float(r["Metric Value"].replace(",", "")) / 1e3 for r in rows]
    starts = [i for i, n in enumerate(names) if "k_conum_regular" in n] or [i for i, n in enumerate(names) if "k_conum_stage1" in n]
    a, b = starts[-2], starts[-1]
    step = list(zip(names[a:b], dur[a:b]))
    total = sum(d for _, d in step)
    table = alg_bytes(nC, nI, nB)
    seen = 0.0
    for key, (what, nbytes) in table.items():
        pats = key.split("|")
        sel = [(n, d) for n, d in step if any(p in n for p in pats)]
        if not sel:
            continue
        # CG kernels launched behind a raised stop flag return at once (a few us): not part of the per-launch figure
        noop = [x for x in sel if key.startswith(("k_spmv<4>", "k_cg_update<0")) and x[1] < 8.0]
        sel = [x for x in sel if x not in noop]
        seen += sum(d for _, d in noop)
        if not sel:
            print(json.dumps({"mesh": label, "stage": what, "note": f"{len(noop)} no-op launches behind the stop flag (0 iterations in this capture)",
                              "us_per_step": round(sum(d for _, d in noop), 1)}))
            continue
        us = sum(d for _, d in sel)
        # a stage that runs k times per step (two correctors): per-launch figures. The affine + irregular-cell kernels of one
        # assembly are one stage.
        groups = max(1, sum(1 for n, _ in sel if pats[0] in n))
        per = us / groups
        seen += us
        print(json.dumps({"mesh": label, "stage": what, "kernels": sorted({n[:60] for n, _ in sel}), "times_per_step": groups, "us_per_launch": round(per, 1),
                          "us_per_step": round(us, 1), "share_of_step": round(us / total, 3), "algorithmic_bytes": nbytes,
                          "GBs": round(nbytes / per / 1e3, 1), "frac_of_measured_peak": round(nbytes / per / 1e3 / PEAK, 3)}))
    print(json.dumps({"mesh": label, "stage": "everything else (boundary-condition kernels, reference cell, CoNum stage 2)",
                      "us_per_step": round(total - seen, 1), "share_of_step": round((total - seen) / total, 3), "launches": len(step) - 0}))
    print(json.dumps({"mesh": label, "stage": "TOTAL (ncu, one kernel at a time, caches flushed between launches)", "us_per_step": round(total, 1), "launches": len(step)}))
    for f in sys.argv[6:]:
        d = json.load(open(f))
        p = d.get("piso_256")
        if p:
            print(json.dumps({"bench_line": f.split("/")[-1], "n_gpus": d["n_gpus"], "ms_per_step_synced": p["ms_per_step"], "cg_iterations_per_solve": p["cg_iterations_per_solve"],
                              "back_to_back": p.get("back_to_back") and {k: p["back_to_back"][k] for k in ("steps", "ms_per_step", "cg_iterations_total")},
                              "segments_eager_ms": p.get("segments")}))


main()
