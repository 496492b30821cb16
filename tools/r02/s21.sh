#!/bin/bash
set -x
mkdir -p gpurun_out/r02
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r02/smoke_s21.log 2>&1; tail -4 gpurun_out/r02/smoke_s21.log
