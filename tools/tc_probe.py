#!/usr/bin/env python
"""Diagnostic probe for the tcgen05 kernels: each case runs in its own subprocess (a trap kills the
CUDA context), compares against the fp32 CUDA-core kernel on TF32-exact operands and dumps the first
tile of got/ref to gpurun_out/tc_probe_<case>.npz for offline analysis of swizzle / descriptor bugs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (kind, N, C, H, W, vertical, dil, weights)
    "fwd64_identity": ("fwd", 1, 64, 2, 128, False, 1, "identity"),
    "fwd64_center": ("fwd", 1, 64, 2, 128, False, 1, "center"),
    "fwd64_full": ("fwd", 2, 64, 8, 128, False, 1, "full"),
    "fwd64_vert": ("fwd", 2, 64, 8, 128, True, 2, "full"),
    "fwd128_identity": ("fwd", 1, 128, 4, 64, False, 1, "identity"),
    "fwd128_full": ("fwd", 2, 128, 32, 64, True, 4, "full"),
    "fwd128_many": ("fwd", 32, 128, 32, 64, False, 2, "full"),
    "fwd128_d16": ("fwd", 2, 128, 32, 64, False, 16, "full"),
    "fwd128_v8": ("fwd", 2, 128, 32, 64, True, 8, "full"),
    "fwd128_320": ("fwd", 1, 128, 40, 80, True, 4, "full"),
    "wgrad128": ("wgrad", 2, 128, 32, 64, False, 1, None),
    "wgrad128_vert": ("wgrad", 4, 128, 32, 64, True, 8, None),
    "wgrad64": ("wgrad", 2, 64, 16, 128, False, 1, None),
    "wgrad64_vert": ("wgrad", 2, 64, 16, 128, True, 1, None),
}


def run_case(name):
    import numpy as np
    import torch
    from lanedetection_end2end_b200 import ops_net as o
    kind, N, C, H, W, vertical, dil, wk = CASES[name]
    from lanedetection_end2end_b200 import _capi
    _capi.lib().lf_conv1d_tc_set_variant(int(os.environ.get("TC_VARIANT", "2")))
    g = torch.Generator().manual_seed(1)

    def tfx(t):
        return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)

    kh, kw = (3, 1) if vertical else (1, 3)
    out = {"name": name}
    if kind == "fwd":
        # structured input: value encodes (pixel, channel) exactly in TF32: small integers
        px = torch.arange(N * H * W).view(N, H, W, 1).float() % 512
        ch = torch.arange(C).view(1, 1, 1, C).float()
        x = (px + ch / 128.0).cuda() if wk == "identity" else tfx(torch.randn(N, H, W, C, generator=g)).cuda()
        w = torch.zeros(C, C, kh, kw)
        if wk == "identity":
            w.view(C, C, 3)[:, :, 1] = torch.eye(C)
        elif wk == "center":
            w.view(C, C, 3)[:, :, 1] = tfx(torch.randn(C, C, generator=g))
        else:
            w = tfx(torch.randn(C, C, kh, kw, generator=g) / (3 * C) ** 0.5)
        w = w.cuda()
        o.set_conv_mode("fp32")
        ref = o.conv3(x, w, vertical, dil, False)
        o.set_conv_mode("tf32")
        got = o.conv3(x, w, vertical, dil, False)
        torch.cuda.synchronize()
    else:
        x = tfx(torch.randn(N, H, W, C, generator=g)).cuda()
        dy = tfx(torch.randn(N, H, W, C, generator=g)).cuda()
        w = torch.zeros(C, C, kh, kw, device="cuda")
        o.set_conv_mode("fp32")
        ref, _ = o.wgrad3(x, dy, w, vertical, dil)
        o.set_conv_mode("tf32")
        got, _ = o.wgrad3(x, dy, w, vertical, dil)
        torch.cuda.synchronize()
    err = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
    out["rel_err"] = err
    out["nan"] = bool(torch.isnan(got).any())
    gf, rf = got.reshape(-1, got.shape[-1]) if kind == "fwd" else got.reshape(C, -1), None
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "tc_probe_%s.npz" % name),
                        got=got.detach().cpu().numpy().reshape(-1)[:128 * 128 * 4],
                        ref=ref.detach().cpu().numpy().reshape(-1)[:128 * 128 * 4], shape=np.array(got.shape))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in CASES:
        run_case(sys.argv[1])
    else:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        for name in CASES:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                                   timeout=180)
                tail = (r.stdout.strip().splitlines() or [""])[-1]
                print("%-18s rc=%d %s %s" % (name, r.returncode, tail, r.stderr.strip().splitlines()[-1:] if r.returncode else ""),
                      flush=True)
            except subprocess.TimeoutExpired:
                print("%-18s TIMEOUT" % name, flush=True)
