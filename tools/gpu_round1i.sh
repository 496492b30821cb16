#!/bin/bash
# GPU session 13: fused BN1 backward reductions in the dgrad epilogue, AHEAD instantiations for residual-add launches, unrolled finalize/reduce loops.
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu --maxfail=60 -q > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tf32_graph_n1.json 2> gpurun_out/bench_tf32_graph_n1.err; echo "rc=$?" >> gpurun_out/bench_tf32_graph_n1.err
cat gpurun_out/bench_tf32_graph_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_tf32.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph --no-parity-arm > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out | tail -8
