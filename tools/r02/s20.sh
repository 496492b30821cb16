#!/bin/bash
# round 2, session 20: full GPU suite + smoke on the then-final tree
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu_s20.log 2>&1; tail -4 $O/pytest_gpu_s20.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_s20.log 2>&1; tail -3 $O/smoke_s20.log
