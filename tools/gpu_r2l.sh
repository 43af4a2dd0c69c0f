#!/bin/bash
# 8 GPUs: parity on both transports, PISO strong scaling 256^3 at N=8/4, bench weak scaling N=8/4
OUT=gpurun_out/r2l
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29531 tools/mgpu_check.py --perf > $OUT/mgpu_check8.log 2>&1; echo "rc=$?" >> $OUT/mgpu_check8.log
grep -E "MGPU|rc=" $OUT/mgpu_check8.log | cut -c1-1200
timeout 900 $TR --nproc-per-node 8 --master-port 29532 tools/piso_scaling.py --size 256 > $OUT/piso_n8_256.log 2>&1; grep PISO $OUT/piso_n8_256.log | cut -c1-600
timeout 900 $TR --nproc-per-node 4 --master-port 29533 tools/piso_scaling.py --size 256 > $OUT/piso_n4_256.log 2>&1; grep PISO $OUT/piso_n4_256.log | cut -c1-600
timeout 900 $TR --nproc-per-node 8 --master-port 29534 tools/piso_scaling.py --size 256 --transport nccl > $OUT/piso_n8_256_nccl.log 2>&1; grep PISO $OUT/piso_n8_256_nccl.log | cut -c1-600
timeout 600 $TR --nproc-per-node 8 --master-port 29535 bench.py --gpus 8 --steps 200 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench8 rc=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29536 bench.py --gpus 4 --steps 200 --warmup 5 > $OUT/bench_n4.json 2> $OUT/bench_n4.err; echo "bench4 rc=$?"
cut -c1-420 $OUT/bench_n8.json $OUT/bench_n4.json
