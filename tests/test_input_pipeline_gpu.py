"""lf_frame_preprocess (csrc/input_pipe.cu) against the Pillow golden bytes and the numpy restatement: bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import resize_oracle as ro
from oracle.make_resize_golden import CASES, make_frame
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_kernel_matches_pillow_golden_bit_exact(case):
    from lanedetection_end2end_b200.input_pipeline import FramePreprocessor
    name, h, w, R, seed = case
    g = np.load(os.path.join(GOLDEN, "resize_pil.npz"))
    a = make_frame(h, w, seed)
    pre = FramePreprocessor(R, frame_hw=(h, w))
    out = pre(torch.from_numpy(a).cuda().unsqueeze(0))
    assert out.shape == (1, 3, R, 2 * R) and out.dtype == torch.float32
    u8 = torch.round(out[0] * 255).to(torch.uint8).permute(1, 2, 0).contiguous().cpu().numpy()      # back to HWC bytes
    np.testing.assert_array_equal(u8.reshape(-1)[g[name + "/idx"]], g[name + "/val"])
    want = ro.preprocess(a, R)
    np.testing.assert_array_equal(out[0].cpu().numpy(), want)                                        # float bits included


def test_batch_flip_and_stem_layout():
    from lanedetection_end2end_b200.input_pipeline import FramePreprocessor
    frames = np.stack([make_frame(720, 1280, s) for s in (1, 2, 3)])
    pre = FramePreprocessor(256)
    fl = [False, True, False]
    x = pre(torch.from_numpy(frames).cuda(), flip=fl)
    y = pre(torch.from_numpy(frames).cuda(), flip=fl, layout="nhwc4")
    for n in range(3):
        np.testing.assert_array_equal(x[n].cpu().numpy(), ro.preprocess(frames[n], 256, flip=fl[n]))
    assert torch.equal(y[..., :3].permute(0, 3, 1, 2), x) and float(y[..., 3].abs().max()) == 0.0
    with pytest.raises(ValueError):
        pre(torch.zeros(1, 100, 100, 3, dtype=torch.uint8, device="cuda"))
