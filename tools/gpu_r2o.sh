#!/bin/bash
OUT=gpurun_out/r2o
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2000 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "rc=$?"

cut -c1-330 $OUT/bench_n2.json $OUT/bench_n2_nccl.json

