#!/bin/bash
OUT=gpurun_out/r2m
mkdir -p $OUT
timeout 600 python -m pytest tests/test_explicit_gpu.py tests/test_decomp_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 300 python tools/roofline.py --mesh 128 256 --variants 0 --reps 10 --out $OUT/roof_affine.jsonl 2> $OUT/roof.err | cut -c1-200

