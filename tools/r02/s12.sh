#!/bin/bash
# round 2, session 12: pre-masked residual gradient (lf_bn_bwd_apply_gated) A/B, LSQ forward back to scalar math; tests + bench
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s12.log 2>&1; tail -5 $O/pytest_gpu_s12.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s12_gm1.json 2> $O/bench_s12_gm1.err; head -c 250 $O/bench_s12_gm1.json; tail -3 $O/bench_s12_gm1.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_s12.json
LANEFIT_PREMASK_RES=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s12_gm0.json 2> $O/bench_s12_gm0.err; head -c 250 $O/bench_s12_gm0.json; tail -3 $O/bench_s12_gm0.err
timeout 200 python tools/bench_lsq.py > $O/lsq_stress_s12.jsonl 2> $O/lsq_stress_s12.err; tail -3 $O/lsq_stress_s12.err
timeout 300 python __graft_entry__.py --smoke > $O/smoke_s12.log 2>&1; tail -4 $O/smoke_s12.log
