#!/usr/bin/env python
"""LSQ stress benchmark (BASELINE.json configs[4]): fused weight-map -> moments -> solve kernel and its
backward at B=128, 256x512, order in {2,3,4} x lanes in {2,4,6}.  CUDA events on the launching stream,
inputs (134-403 MB) larger than the 126 MB L2, so no flush is needed between iterations.
Prints one JSON line per (order, lanes, dtype) with achieved GB/s against the measured HBM peak.
ALGORITHMIC bytes: fwd = B*L*H*W*s, bwd = 2*B*L*H*W*s."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lanedetection_end2end_b200 import ops_lsq, _capi          # noqa: E402
from lanedetection_end2end_b200.Networks.LSQ_layer import ProjectiveGridGenerator  # noqa: E402
from lanedetection_end2end_b200.Networks.utils import get_homography  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--orders", type=int, nargs="+", default=[2, 3, 4])
    ap.add_argument("--lanes", type=int, nargs="+", default=[2, 4, 6])
    ap.add_argument("--dtypes", nargs="+", default=["fp32", "bf16"])
    ap.add_argument("--masked", action="store_true", help="also materialise the masked output")
    a = ap.parse_args()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peaks_path))["hbm_gbs"] if os.path.exists(peaks_path) else 6650.0
    H, W, B = 256, 512, a.batch
    M, _ = get_homography(256)
    grid = ProjectiveGridGenerator(torch.Size([1, 1, H, W]), torch.from_numpy(M).float().unsqueeze(0), False)
    tables = ops_lsq.grid_tables(grid, H, W, 255.0)
    dev = torch.device("cuda")
    h = _capi.lib()
    for dt in a.dtypes:
        for L in a.lanes:
            tdt = torch.float32 if dt == "fp32" else torch.bfloat16
            yy = torch.arange(H, device=dev).view(1, 1, H, 1).float()
            xx = torch.arange(W, device=dev).view(1, 1, 1, W).float()
            c = torch.rand(B, L, 1, 1, device=dev) * 300 + 100
            s = torch.rand(B, L, 1, 1, device=dev) * 1.5 - 0.75
            o = (torch.exp(-0.5 * ((xx - c - s * (yy - 128)) / 6.0) ** 2) + 0.02 * torch.rand(B, L, H, W, device=dev)).to(tdt)
            for order in a.orders:
                n = order + 1
                beta = torch.empty(B, L, n, dtype=torch.float64, device=dev)
                zinv = torch.empty(B, L, n, n, dtype=torch.float64, device=dev)
                status = torch.zeros(1, dtype=torch.int32, device=dev)
                masked = torch.empty(B, L, H, W, device=dev) if a.masked else None
                ws = torch.zeros(h.lf_lsq_workspace_bytes(B, L, H, W, order), dtype=torch.uint8, device=dev)
                gb = torch.randn(B, L, n, dtype=torch.float64, device=dev)
                d_o = torch.empty_like(o)
                p = _capi.ptr

                def fwd():
                    _capi.call("lf_lsq_fwd", p(o), _capi.dtype_id(o), p(tables.xtab), p(tables.ytab), p(tables.yrow), B, L, H, W,
                               order, 77, 1, 0.0, 0, p(beta), p(zinv), p(masked), p(status), p(ws), ws.numel(),
                               _capi.stream_ptr())

                def bwd():
                    _capi.call("lf_lsq_bwd", p(o), _capi.dtype_id(o), p(tables.xtab), p(tables.ytab), p(tables.yrow), B, L, H, W,
                               order, 77, 1, p(beta), p(zinv), p(gb), p(d_o), _capi.stream_ptr())

                res = {}
                for name, fn, nbytes in (("fwd", fwd, B * L * H * W * o.element_size() + (B * L * H * W * 4 if a.masked else 0)),
                                         ("bwd", bwd, 2 * B * L * H * W * o.element_size())):
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(a.iters):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / a.iters
                    res[name] = {"ms": ms, "GBps": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak, "bytes": nbytes}
                assert int(status.item()) == 0
                print(json.dumps({"bench": "lsq_stress", "dtype": dt, "B": B, "L": L, "order": order, "masked": a.masked,
                                  "peak_GBps": peak, **{k + "_" + kk: vv for k, v in res.items() for kk, vv in v.items()}}),
                      flush=True)


if __name__ == "__main__":
    main()
