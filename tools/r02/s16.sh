#!/bin/bash
# round 2, session 16 (4 GPUs): the N=4 point of the weak-scaling table (config 2)
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 10 --warmup 3 --max-seconds 240 > $O/bench_s16_c2_n4.json 2> $O/bench_s16_c2_n4.err; grep -o '{"metric.*' $O/bench_s16_c2_n4.json | head -c 300; tail -2 $O/bench_s16_c2_n4.err
