#!/bin/bash
# GPU session 4: coalesced tc epilogue, LSQ NL kernels, small-tile wgrad, parallel reductions.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/tc_probe_*.npz
TC_VARIANT=2 timeout 900 python tools/tc_probe.py > gpurun_out/tc_probe.log 2>&1
cat gpurun_out/tc_probe.log
timeout 1800 python -m pytest tests -m gpu --maxfail=60 -q > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python tools/bench_block.py --modes tf32 --variant 2 --shapes 128,32,64,2 128,32,64,16 64,64,128,1 > gpurun_out/block_bench_v2.jsonl 2> gpurun_out/block_bench.err
timeout 600 python tools/bench_lsq.py > gpurun_out/lsq_stress.jsonl 2> gpurun_out/lsq_stress.err
timeout 300 python tools/bench_lsq.py --masked --dtypes fp32 --orders 2 > gpurun_out/lsq_stress_masked.jsonl 2>> gpurun_out/lsq_stress.err
timeout 900 python bench.py --steps 10 --warmup 3 --conv-mode tf32 > gpurun_out/bench_tf32_graph_n1.json 2> gpurun_out/bench_tf32_graph_n1.err; echo "rc=$?" >> gpurun_out/bench_tf32_graph_n1.err
timeout 900 python bench.py --steps 5 --warmup 3 --conv-mode fp32 --no-cpu-baseline > gpurun_out/bench_fp32_graph_n1.json 2> gpurun_out/bench_fp32_graph_n1.err; echo "rc=$?" >> gpurun_out/bench_fp32_graph_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:_tc_kernel -s 24 -c 6 -o gpurun_out/tc_full \
    python tools/bench_block.py --iters 1 --modes tf32 --shapes 128,32,64,2 > gpurun_out/ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lsq_fwd -s 3 -c 2 -o gpurun_out/lsq_fwd_full \
    python tools/bench_lsq.py --batch 128 --iters 2 --orders 2 --lanes 6 --dtypes fp32 > gpurun_out/ncu_lsq.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_tf32.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --conv-mode tf32 --no-graph > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out
cat gpurun_out/bench_tf32_graph_n1.json
