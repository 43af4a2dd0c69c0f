#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), ncu launch list of the bench command, ncu full capture of the explicit kernels.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 600 python bench.py --steps 500 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; tail -c 2500 $OUT/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gather -s 4 -c 4 -o $OUT/prof_explicit_256 python tools/prof_explicit.py --mesh 256 --variants 0 --reps 2 > $OUT/prof256.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gather -s 4 -c 4 -o $OUT/prof_explicit_128 python tools/prof_explicit.py --mesh 128 --variants 0 --reps 2 > $OUT/prof128.log 2>&1
timeout 300 python tools/roofline_la.py --mesh 128 256 --piso --reps 10 --out $OUT/roof_la.jsonl 2> $OUT/roof_la.err | cut -c1-260
ls -la $OUT
