#!/bin/bash
OUT=gpurun_out/r2n
mkdir -p $OUT
timeout 600 python -m pytest tests/test_explicit_gpu.py tests/test_decomp_gpu.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 300 python tools/roofline.py --mesh 128 256 --variants 0 --reps 10 --out $OUT/roof_affine.jsonl 2> $OUT/roof.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?"
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu --no-extras > $OUT/bench_n1.json 2> $OUT/bench_n1.err
cut -c1-330 $OUT/bench_n2.json $OUT/bench_n1.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_check.py > $OUT/mgpu_check.log 2>&1; grep -E "MGPU|rc=|Error" $OUT/mgpu_check.log | cut -c1-300
