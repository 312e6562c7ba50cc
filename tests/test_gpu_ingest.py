"""Device frame preprocessing (ingest.FramePreprocessor -> csrc/resize.cu) is bit-identical to Pillow's bicubic resize +
torchvision ToTensor / Normalize: against the reference-written golden frames and the oracle restatement."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_golden_frames_bit_exact(golden_dir):
    from vsc22_submission_b200.ingest import FramePreprocessor
    g = np.load(os.path.join(golden_dir, "resize_small.npz"))
    for i in range(int(g["n"])):
        img, want = g[f"img{i}"], g[f"out{i}"]
        pre = FramePreprocessor(want.shape[1], want.shape[2], g[f"mean{i}"], g[f"std{i}"])
        got = pre(img[None])[0].cpu().numpy()
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("h,w,oh,ow", [(360, 640, 224, 224), (360, 640, 384, 384), (720, 1280, 256, 256), (100, 100, 100, 64),
                                       (64, 48, 64, 48), (33, 77, 130, 150)])
def test_batches_match_oracle(h, w, oh, ow):
    import torch
    from oracle import resize_np
    from test_oracle_resize import smooth_image
    from vsc22_submission_b200 import ingest
    rng = np.random.default_rng(h + w)
    frames = np.stack([smooth_image(rng, h, w) for _ in range(3)])
    frames[1] = rng.integers(0, 256, frames[1].shape, dtype=np.uint8)       # white noise: exercises the clipping
    frames[2, : h // 2] = 255
    frames[2, h // 2:] = 0                                                  # a hard edge: negative lobes over/undershoot
    for factory, mean, std in ((ingest.sscd_transform, ingest.IMAGENET_MEAN, ingest.IMAGENET_STD),
                               (ingest.vit_transform, (0.5,) * 3, (0.5,) * 3)):
        got = factory(oh, ow)(frames).cpu().numpy()
        for i in range(3):
            np.testing.assert_array_equal(got[i], resize_np.preprocess(frames[i], oh, ow, mean, std))
    # CUDA uint8 tensors and lists of per-frame arrays are accepted as well
    pre = ingest.sscd_transform(oh, ow)
    assert torch.equal(pre(torch.from_numpy(frames).cuda()), pre(list(frames)))
    with pytest.raises(AssertionError):
        pre(frames.astype(np.float32))
