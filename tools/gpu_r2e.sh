#!/bin/bash
OUT=gpurun_out/r2e
mkdir -p $OUT
timeout 600 python -m pytest tests/test_explicit_gpu.py tests/test_decomp_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 900 python tools/sweep_brick.py --mesh 256 --reps 10 --out $OUT/sweep_brick.jsonl 2> $OUT/sweep.err | tee $OUT/sweep.log | cut -c1-300
tail -5 $OUT/sweep.err
