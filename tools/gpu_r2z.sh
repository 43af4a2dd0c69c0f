#!/bin/bash
# round-2 record pass on ONE GPU: suite, smoke, both bench arms, ncu launch lists (bench, PISO 256^3 / one rank of 8, advection),
# ncu --set full of the implicit-path and explicit kernels at 256^3. Usage: tools/gpu_r2z.sh [tag]
TAG=${1:-r2z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 600 python bench.py --impl reference > $OUT/bench_n1_ref.json 2> $OUT/bench_n1_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench_n1.json
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 400 $NCU -c 600 --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/bench_under_ncu.log 2>&1
timeout 400 $NCU --log-file $OUT/launches_piso256.csv python tools/piso_subdomain_profile.py --size 256 --ranks 1 --rank 0 --steps 3 --iters 1 > /dev/null 2>&1
timeout 400 $NCU --log-file $OUT/launches_sub8_rank5.csv python tools/piso_subdomain_profile.py --size 256 --ranks 8 --rank 5 --steps 3 --iters 1 > /dev/null 2>&1
timeout 400 $NCU --log-file $OUT/launches_adv256.csv python tools/prof_advection.py 256 3 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_assemble|k_spmv|k_cg_update|k_rAU' -s 5 -c 10 -o $OUT/prof_implicit_256 python tools/prof_implicit.py --mesh 256 --reps 2 > $OUT/prof_implicit_256.log 2>&1; tail -1 $OUT/prof_implicit_256.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gather_affine|k_conum_stage1' -s 6 -c 12 -o $OUT/prof_explicit_256 python tools/prof_explicit.py --mesh 256 --reps 2 > $OUT/prof_explicit_256.log 2>&1; tail -1 $OUT/prof_explicit_256.log
timeout 400 python tools/roofline.py --mesh 128 256 --variants 0 --reps 20 --out $OUT/roofline_explicit.jsonl > /dev/null 2> $OUT/roof.err
timeout 400 python tools/roofline_la.py --mesh 128 256 --piso 128 --reps 10 --out $OUT/roofline_la.jsonl > /dev/null 2>> $OUT/roof.err
ls -la $OUT | head -40
