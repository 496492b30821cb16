#!/bin/bash
# round 2, session 8: programmatic dependent launch (all kernels), Classification heads on the library kernels, packed fp32
# warp reduction in the LSQ forward; tests, smoke, PDL on/off A/B, LSQ stress, configs 3 / 3+clas / 4
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s8.log 2>&1; tail -8 $O/pytest_gpu_s8.log
timeout 300 python tools/r02/diag_block.py > $O/diag_block_s8.jsonl 2> $O/diag_block_s8.err; tail -3 $O/diag_block_s8.err
timeout 300 python __graft_entry__.py --smoke > $O/smoke_s8.log 2>&1; tail -4 $O/smoke_s8.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s8_pdl1.json 2> $O/bench_s8_pdl1.err; head -c 250 $O/bench_s8_pdl1.json; tail -3 $O/bench_s8_pdl1.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_s8_pdl1.json
LANEFIT_PDL=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s8_pdl0.json 2> $O/bench_s8_pdl0.err; head -c 250 $O/bench_s8_pdl0.json; tail -3 $O/bench_s8_pdl0.err
timeout 200 python tools/bench_lsq.py > $O/lsq_stress_s8.jsonl 2> $O/lsq_stress_s8.err; tail -3 $O/lsq_stress_s8.err
timeout 400 python bench.py --config 3 --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s8_c3.json 2> $O/bench_s8_c3.err; head -c 250 $O/bench_s8_c3.json; tail -3 $O/bench_s8_c3.err
timeout 400 python bench.py --config 3 --clas --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s8_c3clas.json 2> $O/bench_s8_c3clas.err; head -c 250 $O/bench_s8_c3clas.json; tail -3 $O/bench_s8_c3clas.err
cp gpurun_out/kernel_table_tf32x3_n1.json $O/kernel_table_s8_c3clas.json
timeout 400 python bench.py --config 4 --steps 10 --warmup 3 --no-parity-arm --no-cpu-baseline > $O/bench_s8_c4.json 2> $O/bench_s8_c4.err; head -c 250 $O/bench_s8_c4.json; tail -3 $O/bench_s8_c4.err
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_heads_gpu.py tests/test_conv_tc_gpu.py tests/test_lsq_gpu.py -m gpu -q -x \
    -k "head_matches or linear_kernels or x3_forward or x3_weight or x3_epilogues or golden" > $O/sanitizer_memcheck_s8.log 2>&1; tail -5 $O/sanitizer_memcheck_s8.log
