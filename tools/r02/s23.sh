#!/bin/bash
# round 2, session 23: ReLU byte masks (A/B), full suite
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s23.log 2>&1; tail -5 $O/pytest_gpu_s23.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s23_bits1.json 2> $O/bench_s23_bits1.err; head -c 230 $O/bench_s23_bits1.json; tail -3 $O/bench_s23_bits1.err
LANEFIT_RELU_BITS=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s23_bits0.json 2> $O/bench_s23_bits0.err; head -c 230 $O/bench_s23_bits0.json; tail -3 $O/bench_s23_bits0.err
timeout 300 python __graft_entry__.py --smoke > $O/smoke_s23.log 2>&1; tail -3 $O/smoke_s23.log | cut -c1-200
