#!/bin/bash
# round 2, session 13 (8 GPUs): BASELINE config 4 (4 lanes, order 3, 320x640, global batch 256 = 32 per GPU, NCCL all-reduce in
# the graph) and config 2 at N = 8
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus_s13.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config 4 --steps 10 --warmup 3 --max-seconds 240 > $O/bench_s13_c4_n8.json 2> $O/bench_s13_c4_n8.err; head -c 300 $O/bench_s13_c4_n8.json; tail -3 $O/bench_s13_c4_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --max-seconds 240 > $O/bench_s13_c2_n8.json 2> $O/bench_s13_c2_n8.err; head -c 300 $O/bench_s13_c2_n8.json; tail -3 $O/bench_s13_c2_n8.err
