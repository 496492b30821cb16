#!/bin/bash
# round 2, session 4: tcg x3 kernels, full suite, smoke, per-op timings, bench
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s4.log 2>&1; tail -12 $O/pytest_gpu_s4.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_s4.log 2>&1; tail -5 $O/smoke_s4.log
timeout 300 python tools/bench_ops.py > $O/ops_s4.jsonl 2> $O/ops_s4.err; tail -3 $O/ops_s4.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-arm > $O/bench_x3_c.json 2> $O/bench_x3_c.err; head -c 300 $O/bench_x3_c.json; tail -3 $O/bench_x3_c.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_x3_c.json
