"""The JPEG oracle (oracle/jpeg_np.py) pinned to what the reference's frame loop calls: Pillow's decoder (libjpeg-turbo),
``Image.open(io.BytesIO(b))`` at VSC22-Descriptor-Track-1st/infer/src/dataset.py:139-140.  Bit-identical on every case."""
import io
import os

import numpy as np
import pytest

from oracle import jpeg_np

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402


def synth(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 7.0 + yy / 13.0), 128 + 100 * np.cos(xx / 11.0 - yy / 5.0), (xx * 3 + yy * 5) % 256], axis=-1)
    return np.clip(img + rng.normal(0, 12, img.shape), 0, 255).astype(np.uint8)


def encode(img, **kw):
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", **kw)
    return buf.getvalue()


@pytest.mark.parametrize("size", [(16, 16), (45, 67), (33, 17), (8, 8), (1, 1), (17, 1), (1, 23), (9, 4), (40, 2)])
@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_oracle_equals_pillow(size, subsampling):
    rng = np.random.default_rng(size[0] * 100 + size[1])
    for q, opt in ((30, False), (75, True), (95, False)):
        data = encode(synth(rng, *size), quality=q, subsampling=subsampling, optimize=opt)
        ref = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
        np.testing.assert_array_equal(jpeg_np.decode(data), ref)


def test_oracle_restart_intervals_and_grey():
    rng = np.random.default_rng(5)
    for ss in (0, 2):
        data = encode(synth(rng, 45, 67), quality=80, subsampling=ss, restart_marker_blocks=3)
        assert b"\xff\xdd" in data
        np.testing.assert_array_equal(jpeg_np.decode(data), np.asarray(Image.open(io.BytesIO(data)).convert("RGB")))
    data = encode(synth(rng, 40, 56)[..., 0], quality=80)
    np.testing.assert_array_equal(jpeg_np.decode(data), np.asarray(Image.open(io.BytesIO(data)).convert("RGB")))


def test_oracle_rejects_what_is_out_of_scope():
    rng = np.random.default_rng(6)
    with pytest.raises(ValueError, match="progressive"):
        jpeg_np.decode(encode(synth(rng, 32, 32), quality=80, progressive=True))
    with pytest.raises(ValueError):
        jpeg_np.decode(b"not a jpeg")


def test_oracle_equals_committed_golden():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_small.npz"))
    for i in range(int(g["n"])):
        np.testing.assert_array_equal(jpeg_np.decode(g[f"jpg{i}"].tobytes()), g[f"rgb{i}"])
