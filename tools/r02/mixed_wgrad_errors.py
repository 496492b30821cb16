#!/usr/bin/env python
"""Parameter-gradient errors (norm-wise, vs the reference's fp64 golden) of the default 3xTF32 mode, of the experiment with
single-pass TF32 weight gradients, and of the reference's own fp32 run -- per parameter, on the batch-32 golden."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lanedetection_end2end_b200 import ops_net                       # noqa: E402
from lanedetection_end2end_b200.Loss_crit import backprojection_loss  # noqa: E402
from oracle import golden_check, inputs                                # noqa: E402

name = "net_l2_d2_b32"
g = np.load(os.path.join(golden_check.GOLDEN, name + ".npz"))
meta = json.loads(str(g["meta"]))
L, order, B = meta["L"], meta["order"], meta["B"]
res = {}
for single in (False, True):
    ops_net.WGRAD_SINGLE_PASS = single
    model, args = golden_check.build_net(L, order, meta["mask_pct"], B)
    sd = model.state_dict()
    for k, v in inputs.make_erfnet_params(3, L, seed=meta["param_seed"]).items():
        sd[k] = torch.from_numpy(v)
    model.load_state_dict(sd)
    model = model.cuda().train()
    for m in model.modules():
        if hasattr(m, "dropout"):
            m.dropout.p = 0
    x = torch.from_numpy(inputs.make_images(B, 256, 512, seed=meta["image_seed"])).cuda()
    xgt_np, valid_np = inputs.make_loss_targets(B, 4, seed=meta["target_seed"])
    xgt, valid = torch.from_numpy(xgt_np).cuda(), torch.from_numpy(valid_np).cuda()
    out = model(x, torch.zeros(B, 4), True)
    crit = backprojection_loss(args)
    loss = sum(crit(out[l], xgt[:, l], valid[:, l])[0] for l in range(L)) / L
    loss.backward()
    torch.cuda.synchronize()
    res[single] = {n: p.grad.double().cpu().numpy().reshape(-1) for n, p in model.named_parameters() if p.grad is not None}
ops_net.WGRAD_SINGLE_PASS = False
gscale = max(g[k][2] for k in g.files if k.startswith("grad_f64/") and k.endswith("/stat"))
rows = []
for n in res[False]:
    k64, k32 = "grad_f64/" + n, "grad_f32/" + n
    if k64 + "/idx" not in g.files or not n.endswith("weight") or ".bn" in n:
        continue
    idx, ref = g[k64 + "/idx"], g[k64 + "/val"]
    scale = max(g[k64 + "/stat"][2], 1e-6 * gscale)
    rows.append({"param": n, "x3": float(np.abs(res[False][n][idx] - ref).max() / scale), "mixed": float(np.abs(res[True][n][idx] - ref).max() / scale),
                 "reference_fp32": float(np.abs(g[k32 + "/val"] - ref).max() / scale)})
for r in rows:
    print(json.dumps(r))
med = lambda key: float(np.median([r[key] for r in rows]))
print(json.dumps({"summary": "conv weight gradients, norm-wise error vs fp64 (sampled entries)", "n": len(rows),
                  "median": {k: med(k) for k in ("x3", "mixed", "reference_fp32")},
                  "max": {k: max(r[k] for r in rows) for k in ("x3", "mixed", "reference_fp32")},
                  "mixed_worse_than_reference_fp32_on": sum(r["mixed"] > r["reference_fp32"] for r in rows)}))
