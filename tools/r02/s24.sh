#!/bin/bash
# round 2, session 24: closing validation of HEAD -- full GPU suite, smoke, default bench line
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s24.log 2>&1; tail -4 $O/pytest_gpu_s24.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_s24.log 2>&1; tail -3 $O/smoke_s24.log | cut -c1-220
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s24.json 2> $O/bench_s24.err; head -c 230 $O/bench_s24.json; tail -3 $O/bench_s24.err
