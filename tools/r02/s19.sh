#!/bin/bash
set -x
mkdir -p gpurun_out/r02
timeout 300 python tools/r02/mixed_wgrad_errors.py > gpurun_out/r02/mixed_wgrad_errors_s19.jsonl 2> gpurun_out/r02/mixed_wgrad_errors_s19.err; tail -2 gpurun_out/r02/mixed_wgrad_errors_s19.jsonl; tail -3 gpurun_out/r02/mixed_wgrad_errors_s19.err
