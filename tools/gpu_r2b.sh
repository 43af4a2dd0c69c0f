#!/bin/bash
# r2b: ncu full capture of the brick kernel (div, grad) at 256^3
OUT=gpurun_out/r2b
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gather_brick -s 4 -c 2 -o $OUT/prof_brick python tools/prof_explicit.py --mesh 256 --variants 0 --reps 2 > $OUT/prof.log 2>&1
ls -la $OUT
