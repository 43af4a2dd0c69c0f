#!/bin/bash
OUT=gpurun_out/r2j
mkdir -p $OUT
FVK_CG_TIMING=1 timeout 300 python tools/roofline_la.py --mesh 128 --piso --reps 5 --out $OUT/la_plain.jsonl 2>&1 | grep -E "timing|pcg" | tail -4 | cut -c1-300
FVK_CG_FORCE_GW=1 FVK_CG_TIMING=1 timeout 300 python tools/roofline_la.py --mesh 128 --piso --reps 5 --out $OUT/la_gw.jsonl 2>&1 | grep -E "timing|pcg" | tail -4 | cut -c1-300
