#!/bin/bash
# round 2, session 10: batch-32 reference golden, segmentation-branch kernels, stock-PyTorch-on-GPU extra arm; full suite + bench
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s10.log 2>&1; tail -6 $O/pytest_gpu_s10.log
timeout 700 python bench.py --steps 10 --warmup 3 > $O/bench_s10.json 2> $O/bench_s10.err; head -c 250 $O/bench_s10.json; tail -3 $O/bench_s10.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_s10.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_s10_reference_arm.json 2> $O/bench_s10_reference_arm.err; head -c 300 $O/bench_s10_reference_arm.json
timeout 200 python tools/bench_lsq.py > $O/lsq_stress_s10.jsonl 2> $O/lsq_stress_s10.err; tail -3 $O/lsq_stress_s10.err
