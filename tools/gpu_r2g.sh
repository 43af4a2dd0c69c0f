#!/bin/bash
OUT=gpurun_out/r2g
mkdir -p $OUT
timeout 900 python tools/sweep_brick.py --mesh 256 --reps 10 --pf 0 300 1200 2400 4800 9600 --out $OUT/sweep_pf256.jsonl 2> $OUT/sweep.err | cut -c1-300
timeout 900 python tools/sweep_brick.py --mesh 128 --reps 10 --plans 0 --pf 0 1200 2400 4800 --out $OUT/sweep_pf128.jsonl 2>> $OUT/sweep.err | cut -c1-300
tail -3 $OUT/sweep.err
