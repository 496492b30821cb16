"""Loader image path: the numpy restatement of Pillow's resample (oracle/resize_oracle.py) against the
golden bytes produced by Pillow itself (tests/golden/resize_pil.npz, oracle/make_resize_golden.py) and, where Pillow is
importable, against Pillow live; the product's host-side coefficient tables against the oracle's."""
import hashlib
import os

import numpy as np
import pytest

from oracle import resize_oracle as ro
from oracle.make_resize_golden import CASES, make_frame
from conftest import GOLDEN


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_pillow_golden(case):
    name, h, w, R, seed = case
    g = np.load(os.path.join(GOLDEN, "resize_pil.npz"))
    a = make_frame(h, w, seed)
    assert hashlib.sha256(a.tobytes()).hexdigest() == str(g[name + "/frame_sha256"])
    got = ro.resize_bilinear_u8(a[h - 640:], R, 2 * R)
    np.testing.assert_array_equal(got.reshape(-1)[g[name + "/idx"]], g[name + "/val"])
    assert hashlib.sha256(got.tobytes()).hexdigest() == str(g[name + "/sha256"])


def test_oracle_matches_pillow_live_with_flip_and_totensor():
    Image = pytest.importorskip("PIL.Image")
    a = make_frame(720, 1280, 5)
    ref = Image.fromarray(a).crop((0, 80, 1280, 720)).resize((512, 256), Image.BILINEAR).transpose(Image.FLIP_LEFT_RIGHT)
    want = np.transpose(np.asarray(ref).astype(np.float32) / np.float32(255), (2, 0, 1))
    np.testing.assert_array_equal(ro.preprocess(a, 256, flip=True), want)


@pytest.mark.parametrize("sizes", [(1280, 512), (640, 256), (640, 320), (1000, 512), (100, 250)])
def test_product_tables_equal_oracle_tables(sizes):
    from lanedetection_end2end_b200.input_pipeline import resample_tables
    b0, k0 = ro._coeffs(*sizes)
    b1, k1 = resample_tables(*sizes)
    np.testing.assert_array_equal(b0, b1)
    np.testing.assert_array_equal(k0, k1)
