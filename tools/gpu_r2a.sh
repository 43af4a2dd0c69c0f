#!/bin/bash
# round-2 first GPU pass: new parity tests, whole suite, smoke, ncu captures of the implicit-path kernels, roofline baseline
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_advection_gpu.py tests/test_implicit_la_gpu.py -m gpu -x -q > $OUT/pytest_new.log 2>&1; echo "pytest-new rc=$?" >> $OUT/pytest_new.log
tail -25 $OUT/pytest_new.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assemble|k_spmv|k_cg_update|k_rAU' -s 5 -c 9 -o $OUT/prof_implicit_256 python tools/prof_implicit.py --mesh 256 --reps 2 > $OUT/prof_implicit_256.log 2>&1; tail -2 $OUT/prof_implicit_256.log
timeout 400 python tools/roofline_la.py --mesh 128 256 --piso 128 --reps 10 --out $OUT/roof_la.jsonl 2> $OUT/roof_la.err | cut -c1-300
ls -la $OUT
