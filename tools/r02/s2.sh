#!/bin/bash
# round 2, session 2: first run of the 3xTF32 conv kernel
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_conv_tc_gpu.py -m gpu -q -x -k "x3" -s > $O/pytest_x3_conv.log 2>&1; tail -25 $O/pytest_x3_conv.log
timeout 600 python -m pytest tests/test_conv_tc_gpu.py tests/test_net_gpu.py -m gpu -q --maxfail=20 > $O/pytest_x3_net.log 2>&1; tail -25 $O/pytest_x3_net.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-arm --conv-mode tf32x3 > $O/bench_x3_a.json 2> $O/bench_x3_a.err; cat $O/bench_x3_a.json; tail -3 $O/bench_x3_a.err
