#!/usr/bin/env python
"""Where does the 3xTF32 conv kernel (csrc/conv_tc_x3.cu) spend its time?  Times lf_conv1d_tc_x3 (20 launches in a CUDA
graph, CUDA events) with parts switched off through lf_conv1d_tc_x3_set_debug:
bit0 = no epilogue body, bit2 = no in-place a_lo rewrite, bit3 = no lo MMAs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanedetection_end2end_b200 import _capi, ops_net as o  # noqa: E402

h = _capi.lib()
o.set_conv_mode("tf32x3")
for (N, C, H, W, vertical, dil) in [(32, 64, 64, 128, True, 1), (32, 128, 32, 64, False, 2), (32, 128, 32, 64, True, 8)]:
    x = torch.randn(N, H, W, C, device="cuda")
    kh, kw = (3, 1) if vertical else (1, 3)
    w = torch.randn(C, C, kh, kw, device="cuda") * 0.05
    mask = torch.randn(N, H, W, C, device="cuda")
    for name, kw_ in (("fwd", {}), ("dgrad+mask", {"mask_src": mask})):
        res = {}
        for dbg in (0, 1, 4, 8, 12, 13):
            h.lf_conv1d_tc_x3_set_debug(dbg)
            sgn = -1 if name != "fwd" else 1
            taps = [((sgn * (k - 1) * dil, 0) if vertical else (0, sgn * (k - 1) * dil)) for k in range(3)]
            wp = o.split_tf32(o.pack_tc_dgrad(w) if name != "fwd" else o.pack_tc_fwd(w))
            out = torch.empty_like(x)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                o.run_conv_tc(taps, x, wp, out, **kw_)
            torch.cuda.current_stream().wait_stream(side)
            with torch.cuda.graph(g):
                for _ in range(20):
                    o.run_conv_tc(taps, x, wp, out, **kw_)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            res["dbg%d" % dbg] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
        h.lf_conv1d_tc_x3_set_debug(0)
        print(json.dumps({"C": C, "HxW": [H, W], "vertical": vertical, "dil": dil, "op": name, "us_per_launch": res,
                          "GFLOP": 2 * N * H * W * 3 * C * C / 1e9}), flush=True)
