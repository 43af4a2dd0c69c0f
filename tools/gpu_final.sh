#!/bin/bash
# last check of a round: the whole GPU suite, smoke(), the default bench line
OUT=gpurun_out/final
mkdir -p $OUT
timeout 200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
if [ "$1" != "--no-bench" ]; then
  timeout 200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? lines=$(wc -l < $OUT/bench.json)"; head -c 400 $OUT/bench.json
fi
