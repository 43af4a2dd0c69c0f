#!/bin/bash
OUT=gpurun_out/r2r
mkdir -p $OUT
timeout 600 python -m pytest tests/test_piso_gpu.py tests/test_implicit_la_gpu.py tests/test_cpp_host.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
for g in 1 0; do
  timeout 300 python tools/piso_scaling.py --size 128 --steps 6 --graphs $g > $OUT/piso_128_g$g.log 2>&1; grep -E "PISO|piso\]" $OUT/piso_128_g$g.log | cut -c1-500
done
timeout 300 python tools/piso_scaling.py --size 256 --steps 4 --graphs 1 > $OUT/piso_256_g1.log 2>&1; grep -E "PISO|piso\]" $OUT/piso_256_g1.log | cut -c1-500
timeout 300 python tools/roofline_la.py --mesh 128 256 --piso --reps 5 --out $OUT/roof_la.jsonl 2>&1 | grep pcg | cut -c1-200
