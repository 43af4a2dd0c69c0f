#!/bin/bash
OUT=gpurun_out/r2d
mkdir -p $OUT
timeout 600 python -m pytest tests/test_explicit_gpu.py tests/test_decomp_gpu.py tests/test_implicit_la_gpu.py tests/test_piso_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 900 python tools/sweep_brick.py --mesh 256 --reps 10 --out $OUT/sweep_brick.jsonl 2> $OUT/sweep.err | tee $OUT/sweep.log | cut -c1-300
tail -5 $OUT/sweep.err
timeout 300 python tools/roofline_la.py --mesh 256 --piso --reps 10 --out $OUT/roof_la.jsonl 2> $OUT/roof_la.err | cut -c1-300
FVK_ASM_GENERIC=1 timeout 300 python tools/roofline_la.py --mesh 256 --piso --reps 10 --out $OUT/roof_la_generic.jsonl 2> $OUT/roof_la_generic.err | cut -c1-300
