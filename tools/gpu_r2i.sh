#!/bin/bash
OUT=gpurun_out/r2i
mkdir -p $OUT
FVK_CG_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py --perf > $OUT/mgpu_check.log 2>&1; echo "rc=$?" >> $OUT/mgpu_check.log
grep -E "MGPU|Error|error|rc=|timing" $OUT/mgpu_check.log | head -20
FVK_CG_TIMING=1 timeout 300 python tools/roofline_la.py --mesh 128 --piso --reps 5 --out $OUT/la_plain.jsonl 2>&1 | grep -E "timing|pcg" | tail -2 | cut -c1-300
