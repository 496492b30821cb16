#!/bin/bash
# round 2, session 11 (2 GPUs): DDP correctness on hardware + the N=2 bench line (NCCL all-reduce inside the graph)
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus_s11.txt
timeout 400 python -m pytest tests/test_ddp_gpu.py -m gpu -q -x > $O/pytest_ddp_s11.log 2>&1; tail -4 $O/pytest_ddp_s11.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --max-seconds 240 > $O/bench_s11_n2.json 2> $O/bench_s11_n2.err; head -c 300 $O/bench_s11_n2.json; tail -3 $O/bench_s11_n2.err
