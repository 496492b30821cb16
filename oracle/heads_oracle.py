"""CPU restatement of the Classification heads (TEST INFRASTRUCTURE ONLY -- never imported by the product package).

Follows BP/Networks/LSQ_layer.py:157-207 as plain functional torch: four conv -> BatchNorm2d (training statistics,
eps 1e-5 = nn.BatchNorm2d default, :164-180) -> ReLU stages (:194-197), MaxPool2d(2) (line, :199) or AvgPool2d((1, cols))
(horizon, :201), the NCHW flatten (:202) and the fully connected layers (:205-206 / :208).  Pinned against the
reference's own class by tests/test_oracle_vs_golden.py (tests/golden/clas_heads.npz)."""
import torch
import torch.nn.functional as F


def head_forward(P, x, kind):
    """P: dict name -> tensor (state_dict names of Classification), x: [B,128,32,64] NCHW; dtype follows x."""
    h = x
    for i in (1, 2, 3, 4):
        w, b = P["conv%d.weight" % i], P["conv%d.bias" % i]
        h = F.conv2d(h, w, b, padding=(w.shape[-1] - 1) // 2)
        h = F.relu(F.batch_norm(h, None, None, P["conv%d_bn.weight" % i], P["conv%d_bn.bias" % i], training=True, eps=1e-5))
    if kind == "line":
        f = F.max_pool2d(h, 2, 2).reshape(h.shape[0], -1)
        f = F.relu(F.linear(f, P["fully_connected1.weight"], P["fully_connected1.bias"]))
        return F.linear(f, P["fully_connected_line1.weight"], P["fully_connected_line1.bias"])
    f = h.mean(dim=3, keepdim=True).reshape(h.shape[0], -1)
    return F.linear(f, P["fully_connected_horizon.weight"], P["fully_connected_horizon.bias"])
