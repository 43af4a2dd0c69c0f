#!/bin/bash
OUT=gpurun_out/r2q
mkdir -p $OUT
timeout 300 python tools/piso_scaling.py --size 128 --steps 4 --warmup 2 > $OUT/piso_128.log 2>&1; grep PISO $OUT/piso_128.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_piso128.csv python tools/piso_scaling.py --size 128 --steps 1 --warmup 1 > $OUT/piso_under_ncu.log 2>&1
grep -c "gpu__time_duration" $OUT/launches_piso128.csv
