#!/bin/bash
# final multi-GPU numbers: PISO strong scaling 256^3 (CUDA graphs + peer-memory windows), bench weak scaling
OUT=gpurun_out/${1:-mgpu}
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29532 tools/piso_scaling.py --size 256 > $OUT/piso_n8_256.log 2>&1; grep -E "PISO|piso\]" $OUT/piso_n8_256.log | cut -c1-600
timeout 240 $TR --nproc-per-node 8 --master-port 29535 bench.py --gpus 8 --steps 2000 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench8 rc=$?"
timeout 240 $TR --nproc-per-node 4 --master-port 29533 tools/piso_scaling.py --size 256 > $OUT/piso_n4_256.log 2>&1; grep -E "PISO|piso\]" $OUT/piso_n4_256.log | cut -c1-600
timeout 240 $TR --nproc-per-node 4 --master-port 29536 bench.py --gpus 4 --steps 2000 --warmup 5 > $OUT/bench_n4.json 2> $OUT/bench_n4.err; echo "bench4 rc=$?"
timeout 240 $TR --nproc-per-node 2 --master-port 29537 tools/piso_scaling.py --size 256 > $OUT/piso_n2_256.log 2>&1; grep -E "PISO|piso\]" $OUT/piso_n2_256.log | cut -c1-600
cut -c1-330 $OUT/bench_n8.json $OUT/bench_n4.json
