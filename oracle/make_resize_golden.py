"""Golden fixture of the loader's image path, generated with Pillow itself (the reference's dependency) in the build
container: `python -m oracle.make_resize_golden` -> tests/golden/resize_pil.npz.  TEST INFRASTRUCTURE.
Stores, for seeded random frames, the sha256 and 8192 sampled bytes of PIL's crop + BILINEAR resize output."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "resize_pil.npz")
CASES = [("f720x1280_r256", 720, 1280, 256, 11), ("f720x1280_r320", 720, 1280, 320, 12), ("f700x1000_r256", 700, 1000, 256, 13)]


def make_frame(h, w, seed):
    """Smooth structure + noise so that both interpolation and rounding are exercised."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 127 + 100 * np.sin(xx / 37.0 + seed) * np.cos(yy / 23.0)
    img = base[..., None] + rng.integers(-40, 41, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    import PIL
    from PIL import Image
    out = {"pillow_version": PIL.__version__}
    for name, h, w, R, seed in CASES:
        a = make_frame(h, w, seed)
        ref = np.asarray(Image.fromarray(a).crop((0, h - 640, w, h)).resize((2 * R, R), Image.BILINEAR))
        idx = np.random.default_rng(seed + 100).integers(0, ref.size, 8192)
        out[name + "/sha256"] = hashlib.sha256(ref.tobytes()).hexdigest()
        out[name + "/idx"] = idx
        out[name + "/val"] = ref.reshape(-1)[idx]
        out[name + "/frame_sha256"] = hashlib.sha256(a.tobytes()).hexdigest()
    np.savez_compressed(GOLDEN, **out)
    print("wrote", GOLDEN, os.path.getsize(GOLDEN))


if __name__ == "__main__":
    main()
