#!/bin/bash
# round 2, session 3: 3xTF32 weight gradient + full GPU suite + bench
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x -k "x3" -s > $O/pytest_x3_conv.log 2>&1; tail -25 $O/pytest_x3_conv.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s3.log 2>&1; tail -25 $O/pytest_gpu_s3.log
LANEFIT_FUSED_LOSS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-arm --conv-mode tf32x3 > $O/bench_x3_b.json 2> $O/bench_x3_b.err; cat $O/bench_x3_b.json; tail -3 $O/bench_x3_b.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_x3_b.json
