#!/bin/bash
# round 2, session 6: lag-2 lo issue + 2 epilogue groups (conv x3), pipelined dY_lo (wgrad x3), LSQ batched loads,
# fused BN backward on the pre-activation; tests, ablation, per-op timings, bench, LSQ stress, sanitizer
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s6.log 2>&1; tail -8 $O/pytest_gpu_s6.log
timeout 300 python tools/x3_ablate.py > $O/x3_ablate_s6.jsonl 2> $O/x3_ablate_s6.err; cat $O/x3_ablate_s6.jsonl; tail -3 $O/x3_ablate_s6.err
timeout 300 python tools/bench_ops.py --modes tf32x3 > $O/ops_s6.jsonl 2> $O/ops_s6.err; tail -3 $O/ops_s6.err
timeout 200 python tools/bench_lsq.py > $O/lsq_stress_s6.jsonl 2> $O/lsq_stress_s6.err; tail -3 $O/lsq_stress_s6.err
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_x3_d.json 2> $O/bench_x3_d.err; head -c 400 $O/bench_x3_d.json; tail -3 $O/bench_x3_d.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_x3_d.json
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_conv_tc_gpu.py tests/test_lsq_gpu.py -m gpu -q -x \
    -k "x3_forward or x3_weight or x3_epilogues or golden" > $O/sanitizer_memcheck.log 2>&1; tail -5 $O/sanitizer_memcheck.log
