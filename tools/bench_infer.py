#!/usr/bin/env python
"""Inference throughput: Net.forward in eval mode under torch.no_grad(), batch 32, 256x512, whole forward
replayed as one CUDA graph; BatchNorm-folded launches (ops_eval.py) against the unfused eval path.  One JSON line per
variant: images/s, ms per forward, launches per forward."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanedetection_end2end_b200 import _capi                     # noqa: E402
from lanedetection_end2end_b200.Networks import ERFNet            # noqa: E402
from lanedetection_end2end_b200.Networks.LSQ_layer import Net     # noqa: E402
from lanedetection_end2end_b200.Networks.utils import define_args  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--nclasses", type=int, default=2)
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    args = define_args().parse_args(["--image_dir", "x", "--gt_dir", "y", "--nclasses", str(a.nclasses), "--order", str(a.order),
                                     "--batch_size", str(a.batch)])
    torch.manual_seed(0)
    model = Net(args).cuda().eval()
    model.defer_status_check = True
    x = torch.rand(a.batch, 3, 256, 512, device="cuda")
    gt_line = torch.zeros(a.batch, 4)
    for fused in (True, False):
        ERFNet.EVAL_FUSED = fused
        with torch.no_grad():
            for _ in range(3):
                model(x, gt_line, True)
            torch.cuda.synchronize()
            l0 = _capi.LAUNCHES
            model(x, gt_line, True)
            launches = _capi.LAUNCHES - l0
            from lanedetection_end2end_b200.engine import GraphedInference
            gi = GraphedInference(model, x)
            g, out = gi.graph, gi.out
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
        print(json.dumps({"mode": "eval fused (BatchNorm folded)" if fused else "eval unfused", "batch": a.batch, "images_per_s": a.batch / ms * 1e3,
                          "ms_per_forward": ms, "launches_per_forward": launches, "beta0": out[0][0, :, 0].tolist()}), flush=True)
    ERFNet.EVAL_FUSED = True


if __name__ == "__main__":
    main()
