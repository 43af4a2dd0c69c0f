#!/bin/bash
OUT=gpurun_out/final
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? lines=$(wc -l < $OUT/bench.json)"; tail -c 1800 $OUT/bench.json
