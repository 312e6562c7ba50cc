"""Parity of the CUDA Swin-V2 encoder (through the C ABI) with the oracle and the reference-class golden."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available(), "GPU tests need a CUDA device"
    return t


def _rel_l2(a, b):
    return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-12)


def test_swin_matches_reference_class_golden(torch, golden_dir):
    """descriptors of the reference's own SwinTransformerV2 (swinv2.py:502) on seeded weights; bf16 operands:
    tolerance 2e-2 relative L2 per frame (measured value printed)."""
    from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SwinSpec, random_weights
    g = np.load(os.path.join(golden_dir, "swin_small.npz"))
    spec = SwinSpec(img=128, patch=4, embed=64, depths=(2, 2, 2, 2), heads=(2, 4, 8, 16), window=8,
                    pretrained_windows=(6, 6, 6, 3), out_dim=64)
    enc = B200SwinEncoder(spec, random_weights(spec, seed=0), max_frames=2).cuda().eval()
    out = enc(torch.from_numpy(g["frames"]).cuda()).cpu().numpy()          # 3 frames, max_frames 2: two chunks
    rel = _rel_l2(out, g["desc"])
    print("swin small rel L2 vs reference class:", rel)
    assert rel.max() < 8e-3          # bf16 mode: measured 4.6e-3 (the bf16-operand floor); fp32 mode: tests/test_gpu_exact.py


def test_swin_b_256_matches_oracle(torch):
    """The deployed configuration (config_v106.py: SwinV2-B 256, window 16) on 2 frames vs the fp32 oracle."""
    from oracle import swin_ref
    from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SWINV2_B_256, random_weights
    t = torch
    w = random_weights(SWINV2_B_256, seed=1)
    x = t.randn(2, 3, 256, 256, generator=t.Generator().manual_seed(3)).clamp(-1, 1)
    ref = swin_ref.forward(swin_ref.SWINV2_B_256, w, x).numpy()
    enc = B200SwinEncoder(SWINV2_B_256, w, max_frames=2).cuda().eval()
    out = enc(x.cuda()).cpu().numpy()
    rel = _rel_l2(out, ref)
    cos = (out * ref).sum(1) / (np.linalg.norm(out, axis=1) * np.linalg.norm(ref, axis=1))
    print("swinv2-b rel L2 vs fp32 oracle:", rel, "cos", cos)
    assert rel.max() < 8e-3 and cos.min() > 0.9999      # bf16 mode: measured 5.0e-3; the 1e-3 contract: test_gpu_parity_full.py
    # host API == device API, ragged batch
    out_h = enc.forward_host(x.numpy()[:1])
    assert np.abs(out_h - out[:1]).max() < 1e-5


def test_swin_rejects_cpu_and_bad_shapes(torch):
    from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SwinSpec, random_weights
    spec = SwinSpec(img=64, patch=4, embed=64, depths=(1, 1), heads=(2, 4), window=8, pretrained_windows=(0, 0), out_dim=32)
    enc = B200SwinEncoder(spec, random_weights(spec)).cuda()
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 3, 32, 32, device="cuda"))
    assert enc(torch.zeros(0, 3, 64, 64, device="cuda")).shape == (0, 32)
