#!/bin/bash
# First GPU session: parity tests, bench, LSQ stress, ncu launch list + full capture of the LSQ kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu --maxfail=40 -q -x --deselect tests/test_net_gpu.py > gpurun_out/pytest_lsq.log 2>&1; echo "lsq tests rc=$?" >> gpurun_out/pytest_lsq.log
timeout 1200 python -m pytest tests/test_net_gpu.py -m gpu --maxfail=40 -q > gpurun_out/pytest_net.log 2>&1; echo "net tests rc=$?" >> gpurun_out/pytest_net.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python tools/bench_lsq.py > gpurun_out/lsq_stress.jsonl 2> gpurun_out/lsq_stress.err
timeout 600 python tools/bench_lsq.py --masked --dtypes fp32 --orders 2 > gpurun_out/lsq_stress_masked.jsonl 2>> gpurun_out/lsq_stress.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# ncu: full capture of the LSQ kernels at the stress size, then the launch list of one bench step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lsq_ -s 6 -c 4 -o gpurun_out/lsq_full \
    python tools/bench_lsq.py --batch 128 --iters 2 --orders 2 --lanes 6 --dtypes fp32 > gpurun_out/ncu_lsq.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_net_gpu.py -m gpu -q -x -k "block or output_conv" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
ls -la gpurun_out
tail -5 gpurun_out/pytest_lsq.log gpurun_out/pytest_net.log gpurun_out/smoke.log
cat gpurun_out/bench_n1.json
