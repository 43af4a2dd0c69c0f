#!/bin/bash
# 2 GPUs: transports parity + perf, bench N=2 with both transports
OUT=gpurun_out/r2h
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py --perf > $OUT/mgpu_check.log 2>&1; echo "rc=$?" >> $OUT/mgpu_check.log
grep -E "MGPU|Error|error|rc=" $OUT/mgpu_check.log | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > $OUT/bench_n2_p2p.json 2> $OUT/bench_n2_p2p.err; echo "rc=$?"
FVK_BENCH_TRANSPORT=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 5 > $OUT/bench_n2_nccl.json 2> $OUT/bench_n2_nccl.err; echo "rc=$?"
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu --no-extras > $OUT/bench_n1.json 2> $OUT/bench_n1.err
cut -c1-600 $OUT/bench_n2_p2p.json $OUT/bench_n2_nccl.json $OUT/bench_n1.json
tail -5 $OUT/bench_n2_p2p.err
