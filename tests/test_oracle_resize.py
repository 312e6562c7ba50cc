"""oracle/resize_np.py == Pillow + torchvision through the reference's own transform factories
(D/infer/src/transform.py:20-43), live (container-only) and through tests/golden/resize_small.npz."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import refload, resize_np


def smooth_image(rng, h, w):
    """Seeded smooth-noise RGB frame (no ffmpeg / datasets here)."""
    low = rng.uniform(0, 255, (h // 8 + 2, w // 8 + 2, 3))
    ys, xs = np.linspace(0, low.shape[0] - 1.001, h), np.linspace(0, low.shape[1] - 1.001, w)
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    img = (low[y0][:, x0] * (1 - fy) * (1 - fx) + low[y0 + 1][:, x0] * fy * (1 - fx) +
           low[y0][:, x0 + 1] * (1 - fy) * fx + low[y0 + 1][:, x0 + 1] * fy * fx)
    return np.clip(img + rng.normal(0, 12, img.shape), 0, 255).astype(np.uint8)


CASES = [(90, 160, 56, 56), (36, 64, 96, 96), (75, 75, 75, 32), (50, 81, 50, 120), (360, 640, 224, 224)]


def test_golden_resize(golden_dir):
    g = np.load(os.path.join(golden_dir, "resize_small.npz"))
    for i in range(int(g["n"])):
        img, want = g[f"img{i}"], g[f"out{i}"]
        mean, std = g[f"mean{i}"], g[f"std{i}"]
        got = resize_np.preprocess(img, want.shape[1], want.shape[2], mean, std)
        np.testing.assert_array_equal(got, want)


def test_coefficients_sum_to_one():
    for a, b in ((640, 224), (360, 224), (64, 96), (100, 100)):
        bounds, kk, ksize = resize_np.coefficients(a, b)
        assert np.abs(kk.sum(1) - (1 << resize_np.PRECISION_BITS)).max() <= ksize
        assert (bounds[:, 0] >= 0).all() and (bounds.sum(1) <= a).all()


@pytest.mark.skipif(not refload.available(), reason="/root/reference not present")
def test_oracle_equals_reference_transforms():
    from PIL import Image
    spec = importlib.util.spec_from_file_location("_ref_transform", os.path.join(refload.D, "infer/src/transform.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    rng = np.random.default_rng(0)
    for h, w, oh, ow in CASES:
        img = smooth_image(rng, h, w)
        for factory, mean, std in ((tr.sscd_transform, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
                                   (tr.vit_transform, (0.5, 0.5, 0.5), (0.5, 0.5, 0.5))):
            want = factory(oh, ow)(Image.fromarray(img)).numpy()
            got = resize_np.preprocess(img, oh, ow, mean, std)
            np.testing.assert_array_equal(got, want)
