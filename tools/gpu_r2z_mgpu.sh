#!/bin/bash
# round-2 record pass on N GPUs (N = $1, default 8): multi-GPU parity check, then the bench line (weak-scaling headline + kernels_256 /
# piso_256 / advection_256 / parity on the decomposed 256^3 meshes). Usage: tools/gpu_r2z_mgpu.sh N [tag]
N=${1:-8}
TAG=${2:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo_n$N.txt 2>&1
timeout 400 $TR --nproc-per-node $N --master-port 29511 tests/mgpu_check.py > $OUT/mgpu_check_n$N.log 2>&1; grep -E "MGPU|Error|assert" $OUT/mgpu_check_n$N.log | cut -c1-300
timeout 600 $TR --nproc-per-node $N --master-port 29535 bench.py --gpus $N > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench$N rc=$?"
timeout 600 $TR --nproc-per-node $N --master-port 29536 bench.py --gpus $N --impl reference > $OUT/bench_n${N}_ref.json 2> $OUT/bench_n${N}_ref.err; echo "ref$N rc=$?"
cut -c1-400 $OUT/bench_n$N.json; tail -c 400 $OUT/bench_n$N.err
