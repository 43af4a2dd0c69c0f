#!/bin/bash
OUT=gpurun_out/r2k
mkdir -p $OUT
for n in 128 256; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$((n/128)) tools/piso_scaling.py --size $n > $OUT/piso_n2_$n.log 2>&1; grep PISO $OUT/piso_n2_$n.log | cut -c1-600
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29529 tools/piso_scaling.py --size 256 --transport nccl > $OUT/piso_n2_256_nccl.log 2>&1; grep PISO $OUT/piso_n2_256_nccl.log | cut -c1-600
tail -3 $OUT/piso_n2_256.log
