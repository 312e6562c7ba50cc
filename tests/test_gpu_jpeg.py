"""JPEG frame decoding on the device (csrc/jpeg.cu) against Pillow itself, the committed goldens and the oracle: bit-identical
(the reference decodes every frame with ``Image.open(io.BytesIO(b))``, D/infer/src/dataset.py:137-141)."""
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available(), "GPU tests need a CUDA device"
    return t


def synth(rng, h, w):
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([128 + 100 * np.sin(xx / 7.0 + yy / 13.0), 128 + 100 * np.cos(xx / 11.0 - yy / 5.0), (xx * 3 + yy * 5) % 256], axis=-1)
    return np.clip(img + rng.normal(0, 12, img.shape), 0, 255).astype(np.uint8)


def encode(img, **kw):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", **kw)
    return buf.getvalue()


def pil_rgb(data):
    from PIL import Image
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def test_golden_files_decode_bit_exactly(torch, golden_dir):
    from vsc22_submission_b200 import ingest
    g = np.load(os.path.join(golden_dir, "jpeg_small.npz"))
    for i in range(int(g["n"])):
        got = ingest.decode_jpeg_frames([g[f"jpg{i}"].tobytes()]).cpu().numpy()[0]
        np.testing.assert_array_equal(got, g[f"rgb{i}"])
    # decode -> resize -> normalise = the reference's whole per-frame chain (vit_transform of transform.py:37-43)
    out = ingest.vit_transform(32, 32)([g["jpg0"].tobytes()]).cpu().numpy()[0]
    np.testing.assert_array_equal(out, g["chain_out"])


@pytest.mark.parametrize("subsampling", [0, 1, 2])
def test_batches_against_pillow(torch, subsampling):
    """A video's worth of equally sized frames in one call; sizes that are not multiples of the MCU; qualities; optimised
    Huffman tables (per-frame tables differ); restart intervals (one warp per interval)."""
    from vsc22_submission_b200 import ingest
    rng = np.random.default_rng(40 + subsampling)
    for (h, w), kw in [((90, 160), dict(quality=75)), ((45, 67), dict(quality=92, optimize=True)), ((17, 1), dict(quality=50)),
                       ((64, 48), dict(quality=80, restart_marker_blocks=2)), ((8, 8), dict(quality=30)), ((9, 4), dict(quality=85))]:
        files = [encode(synth(rng, h, w), subsampling=subsampling, **kw) for _ in range(5)]
        got = ingest.decode_jpeg_frames(files).cpu().numpy()
        for f, o in zip(files, got):
            np.testing.assert_array_equal(o, pil_rgb(f))


def test_grey_and_oracle_and_errors(torch):
    from oracle import jpeg_np
    from vsc22_submission_b200 import ingest
    rng = np.random.default_rng(44)
    grey = encode(synth(rng, 40, 56)[..., 0], quality=80)
    np.testing.assert_array_equal(ingest.decode_jpeg_frames([grey]).cpu().numpy()[0], pil_rgb(grey))
    col = encode(synth(rng, 40, 56), quality=80)
    np.testing.assert_array_equal(ingest.decode_jpeg_frames([col]).cpu().numpy()[0], jpeg_np.decode(col))
    with pytest.raises(RuntimeError, match="progressive"):
        ingest.decode_jpeg_frames([encode(synth(rng, 32, 32), quality=80, progressive=True)])
    with pytest.raises(RuntimeError, match="differs from frame 0"):
        ingest.decode_jpeg_frames([col, encode(synth(rng, 32, 32), quality=80)])
    with pytest.raises(RuntimeError, match="not a JPEG"):
        ingest.decode_jpeg_frames([b"garbage bytes"])
    assert ingest.decode_jpeg_frames([]).shape[0] == 0


def test_video_zip_mirror(torch, tmp_path):
    """``D_vsc.__getitem__`` (dataset.py:126-148): zip of JPEG frames -> [n, 3, h, w]; against Pillow + the oracle resize."""
    import zipfile
    from PIL import Image
    from vsc22_submission_b200 import ingest
    rng = np.random.default_rng(45)
    files = {f"{i:05d}.jpg": encode(synth(rng, 72, 128), quality=85, subsampling=2) for i in range(7)}
    zp = tmp_path / "Q100001.zip"
    with zipfile.ZipFile(zp, "w") as z:
        for name in reversed(sorted(files)):            # stored out of order: the loader sorts the names
            z.writestr(name, files[name])
    pre = ingest.sscd_transform(48, 48)
    got = ingest.video_zip_frames(str(zp), pre).cpu().numpy()
    want = pre(np.stack([pil_rgb(files[k]) for k in sorted(files)])).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    assert got.shape == (7, 3, 48, 48)
