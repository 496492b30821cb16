#!/bin/bash
# round 2, session 7: after the x3 weight-slot deadlock fix (groups of fewer than LAG tiles): full suite, sanitizer, smoke,
# ncu --set full summaries of every kernel flavour of one bench step (processed on the box; only JSON comes back), bench
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
T=/tmp/ncu_r02; mkdir -p $T
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s7.log 2>&1; tail -8 $O/pytest_gpu_s7.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_s7.log 2>&1; tail -5 $O/smoke_s7.log
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_conv_tc_gpu.py tests/test_lsq_gpu.py -m gpu -q -x \
    -k "x3_forward or x3_weight or x3_epilogues or golden" > $O/sanitizer_memcheck.log 2>&1; tail -5 $O/sanitizer_memcheck.log
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-parity-arm"
# bench.py runs 3 warm-up steps + 1 counted step before the timed ones: skip them (204 x3 launches per step)
timeout 900 ncu --set full --clock-control none -k regex:"conv1d_tc_x3|wgrad_tc_x3" --launch-skip 816 -c 204 -o $T/x3_step $BENCH > $O/ncu_x3_step.log 2>&1
tail -2 $O/ncu_x3_step.log
python tools/ncu_summary.py $T/x3_step.ncu-rep $O/ncu_x3_step.json > /dev/null 2> $O/ncu_x3_step.err
timeout 900 ncu --set full --clock-control none -k regex:"bn_|lsq_|maxpool|outconv|conv_tcg|wgrad_tcg|reduce|pack_gather|backproj|conv_igemm|wgrad_small|nchw" --launch-skip 1200 -c 300 -o $T/misc_step $BENCH > $O/ncu_misc_step.log 2>&1
tail -2 $O/ncu_misc_step.log
python tools/ncu_summary.py $T/misc_step.ncu-rep $O/ncu_misc_step.json > /dev/null 2> $O/ncu_misc_step.err
# small reports WITH source for the dominant flavours (kept as .ncu-rep for the source page)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv1d_tc_x3" --launch-skip 826 -c 3 -o $O/x3_conv_src $BENCH > $O/ncu_x3_src.log 2>&1
for dt in fp32 bf16; do
timeout 300 ncu --set full --clock-control none -k regex:"lsq_" -c 6 -o $T/lsq_stress_$dt python tools/bench_lsq.py --lanes 6 --orders 3 --iters 1 --dtypes $dt > $O/ncu_lsq_stress_$dt.log 2>&1
python tools/ncu_summary.py $T/lsq_stress_$dt.ncu-rep $O/ncu_lsq_stress_L6_d3_$dt.json > /dev/null 2> $O/ncu_lsq_stress_$dt.err
done
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_x3_e.json 2> $O/bench_x3_e.err; head -c 300 $O/bench_x3_e.json; tail -3 $O/bench_x3_e.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_x3_e.json
ls -la $O $T
