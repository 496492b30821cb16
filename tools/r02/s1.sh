#!/bin/bash
# round 2, session 1: tf32 conversion micro-experiment, fused loss validation, ncu evidence for the BN / LSQ kernels
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt
timeout 60 tools/micro/tf32_rounding > $O/tf32_rounding.jsonl 2>&1; cat $O/tf32_rounding.jsonl
LANEFIT_FUSED_LOSS=1 timeout 300 python -m pytest tests -m gpu -q -x -k "loss or full_path" > $O/pytest_fused_loss.log 2>&1; tail -5 $O/pytest_fused_loss.log
LANEFIT_FUSED_LOSS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-arm > $O/bench_tf32_fusedloss.json 2> $O/bench_tf32_fusedloss.err; cat $O/bench_tf32_fusedloss.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"bn_|lsq_|maxpool|outconv" -s 0 -c 60 -o $O/bn_lsq_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-parity-arm > $O/ncu_bn_lsq.log 2>&1
tail -3 $O/ncu_bn_lsq.log
ls -la $O
