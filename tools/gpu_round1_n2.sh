#!/bin/bash
# 2-GPU session: the driver's launch line for N=2 (torchrun, NCCL over NVLink).  bench.py carries its own watchdog.
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-parity-arm --max-seconds 240 > gpurun_out/bench_tf32_graph_n2.json 2> gpurun_out/bench_tf32_graph_n2.err; echo "rc=$?" >> gpurun_out/bench_tf32_graph_n2.err
cat gpurun_out/bench_tf32_graph_n2.json
tail -5 gpurun_out/bench_tf32_graph_n2.err
