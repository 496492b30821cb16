#!/bin/bash
# round 2, session 18: experiment -- single-pass TF32 weight gradients inside the 3xTF32 mode: accuracy on the reference goldens
# and the bench extra
set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 300 python - > $O/mixed_wgrad_accuracy_s18.jsonl 2> $O/mixed_wgrad_accuracy_s18.err <<'PY'
import json, sys
sys.path.insert(0, ".")
from lanedetection_end2end_b200 import ops_net
from oracle import golden_check
for name in ("net_l2_d2", "net_l4_d3", "net_l2_d2_b32"):
    for single in (False, True):
        ops_net.WGRAD_SINGLE_PASS = single
        ok = True
        try:
            rep = golden_check.run_full_path(name, tol=1e-4, enforce=True)
        except AssertionError as e:
            ok = False
            rep = golden_check.run_full_path(name, tol=1e-4, enforce=False)
            rep["first_failed_gate"] = str(e)[:300]
        rep["wgrad_single_pass_tf32"] = single
        rep["all_reference_gates_pass"] = ok
        print(json.dumps(rep), flush=True)
ops_net.WGRAD_SINGLE_PASS = False
PY
cat $O/mixed_wgrad_accuracy_s18.jsonl | cut -c1-400; tail -3 $O/mixed_wgrad_accuracy_s18.err
timeout 700 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_s18.json 2> $O/bench_s18.err; head -c 200 $O/bench_s18.json; tail -3 $O/bench_s18.err
