#!/bin/bash
# round 2, session 15: end-to-end training smoke (loss must fall), ncu of the final conv flavours, full suite, default bench
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
T=/tmp/ncu_r02; mkdir -p $T
timeout 300 python tools/train_synthetic.py > $O/train_synthetic_s15.json 2> $O/train_synthetic_s15.err; cat $O/train_synthetic_s15.json | cut -c1-600; tail -3 $O/train_synthetic_s15.err
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s15.log 2>&1; tail -5 $O/pytest_gpu_s15.log
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-parity-arm"
timeout 600 ncu --set full --clock-control none -k regex:"conv1d_tc_x3" --launch-skip 544 -c 136 -o $T/x3_conv_final $BENCH > $O/ncu_x3_final.log 2>&1; tail -2 $O/ncu_x3_final.log
python tools/ncu_summary.py $T/x3_conv_final.ncu-rep $O/ncu_x3_conv_final.json > /dev/null 2> $O/ncu_x3_conv_final.err
timeout 700 python bench.py --steps 10 --warmup 3 > $O/bench_s15.json 2> $O/bench_s15.err; head -c 250 $O/bench_s15.json; tail -3 $O/bench_s15.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_s15.json
