#!/bin/bash
OUT=gpurun_out/r2t
mkdir -p $OUT
timeout 200 python -m pytest tests/test_implicit_la_gpu.py tests/test_piso_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
timeout 120 python tools/roofline_la.py --mesh 128 256 --piso --reps 10 --out $OUT/roof_la.jsonl 2> $OUT/roof_la.err | cut -c1-200
