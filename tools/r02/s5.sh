#!/bin/bash
# round 2, session 5: where does the 3xTF32 conv kernel spend its time (ablation + ncu source view), LSQ stress after the
# balanced chunking, test re-run
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 300 python tools/x3_ablate.py > $O/x3_ablate_s5.jsonl 2> $O/x3_ablate_s5.err; cat $O/x3_ablate_s5.jsonl; tail -3 $O/x3_ablate_s5.err
timeout 300 python tools/bench_ops.py --modes tf32x3 tf32 > $O/ops_s5.jsonl 2> $O/ops_s5.err; tail -3 $O/ops_s5.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_tc_x3 -s 6 -c 2 -o $O/x3_conv_full \
    python tools/x3_ablate.py > $O/ncu_x3.log 2>&1; tail -3 $O/ncu_x3.log
timeout 200 python tools/bench_lsq.py > $O/lsq_stress_s5.jsonl 2> $O/lsq_stress_s5.err; tail -3 $O/lsq_stress_s5.err
timeout 600 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s5.log 2>&1; tail -6 $O/pytest_gpu_s5.log
