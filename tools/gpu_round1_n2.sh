#!/bin/bash
# 2-GPU session: the driver's launch line for N=2 (torchrun, NCCL over NVLink), ours + reference arm.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_tf32_graph_n2.json 2> gpurun_out/bench_tf32_graph_n2.err; echo "rc=$?" >> gpurun_out/bench_tf32_graph_n2.err
cat gpurun_out/bench_tf32_graph_n2.json
tail -5 gpurun_out/bench_tf32_graph_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "rc=$?" >> gpurun_out/bench_ref_n2.err
cat gpurun_out/bench_ref_n2.json
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-parity-arm > gpurun_out/bench_tf32_graph_n1_b.json 2> gpurun_out/bench_tf32_graph_n1_b.err
cat gpurun_out/bench_tf32_graph_n1_b.json
