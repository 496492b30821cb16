#!/bin/bash
# round 2, session 9: bias in the phase-2 epilogue registers, coalesced split reductions, robust flip gates, PDL default off,
# segmentation-branch kernels; full suite, smoke, bench (default arms incl. the real reference on the host cores), inference
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s9.log 2>&1; tail -6 $O/pytest_gpu_s9.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_s9.log 2>&1; tail -4 $O/smoke_s9.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_s9.json 2> $O/bench_s9.err; head -c 250 $O/bench_s9.json; tail -3 $O/bench_s9.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_s9.json
timeout 300 python tools/bench_infer.py > $O/infer_s9.jsonl 2> $O/infer_s9.err; cat $O/infer_s9.jsonl | cut -c1-200; tail -3 $O/infer_s9.err
timeout 300 python tools/bench_ops.py --modes tf32x3 > $O/ops_s9.jsonl 2> $O/ops_s9.err; tail -3 $O/ops_s9.err
