#!/bin/bash
# r2a: parity of the brick kernel + first roofline sweep (brick shape, resident blocks)
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_explicit_gpu.py tests/test_decomp_gpu.py -m gpu -x -q > $OUT/pytest_explicit.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_explicit.log
tail -5 $OUT/pytest_explicit.log
timeout 300 python tools/roofline.py --mesh 128 256 --variants 0 5 --out $OUT/roof_default.jsonl > $OUT/roof_default.log 2>&1
for mb in 2 4; do
  FVK_BRICK_MINB=$mb timeout 300 python tools/roofline.py --mesh 256 --variants 0 --reps 10 --out $OUT/roof_minb$mb.jsonl > $OUT/roof_minb$mb.log 2>&1
done
for br in 16,8,4 64,4,2 32,8,2 8,8,8 128,2,2; do
  FVK_BRICK=$br timeout 300 python tools/roofline.py --mesh 256 --variants 0 --reps 10 --out $OUT/roof_brick_$br.jsonl > $OUT/roof_brick_$br.log 2>&1
done
grep -h '"kernel": "div"\|"kernel": "grad"' $OUT/roof_*.jsonl | cut -c1-260
