#!/bin/bash
# round 2, session 14: one-operand AHEAD epilogue for the masked input-gradient launches (A/B), tests, racecheck subset
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s14.log 2>&1; tail -5 $O/pytest_gpu_s14.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s14_ahead1.json 2> $O/bench_s14_ahead1.err; head -c 250 $O/bench_s14_ahead1.json; tail -3 $O/bench_s14_ahead1.err
LANEFIT_X3_NOAHEAD1=1 timeout 400 python bench.py --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s14_ahead0.json 2> $O/bench_s14_ahead0.err; head -c 250 $O/bench_s14_ahead0.json; tail -3 $O/bench_s14_ahead0.err
timeout 300 python tools/bench_ops.py --modes tf32x3 > $O/ops_s14.jsonl 2> $O/ops_s14.err; tail -3 $O/ops_s14.err
timeout 500 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_conv_tc_gpu.py tests/test_lsq_gpu.py -m gpu -q -x \
    -k "x3_epilogues_and_dgrad and 64-16-128 or test_lsq_matches_reference_golden and bp_l2_d2_chol" > $O/sanitizer_racecheck_s14.log 2>&1; tail -6 $O/sanitizer_racecheck_s14.log
