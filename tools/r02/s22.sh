#!/bin/bash
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 300 python -m pytest tests/test_net_gpu.py -m gpu -q -x -k "graphed_inference or eval" > $O/pytest_infer_s22.log 2>&1; tail -4 $O/pytest_infer_s22.log
timeout 300 python tools/bench_infer.py > $O/infer_s22.jsonl 2> $O/infer_s22.err; cut -c1-160 $O/infer_s22.jsonl; tail -3 $O/infer_s22.err
