"""Seam (A) with REAL artefacts: TorchScript files exported from the reference's own encoder classes exactly as the
reference exports its checkpoints (torch.jit.trace + torch.jit.save; tests/golden/make_golden.py `torchscript`).

CPU part (no GPU): the artefacts reproduce the stored outputs of the reference classes, every checkpoint is recognised
with the right architecture, and look-alike checkpoints are rejected.  GPU part: the unmodified call
``torch.jit.load(path)`` (D/infer/extract_ref_feats.py:24-27, extract_query_feats.py:77-92) returns a B200 encoder through
``install_jit_load_hook`` whose output equals what the TorchScript module itself returns."""
import os

import numpy as np
import pytest
import torch

ARTEFACTS = {   # tag -> (expected encoder class, spec fields to check)
    "clip_small": ("B200ViTEncoder", dict(img=64, patch=16, width=128, layers=2, heads=2, tail="tokens", pre_norm=True,
                                          act="quick_gelu", patch_bias=False)),
    "sscd_timm_small": ("B200ViTEncoder", dict(img=64, patch=32, width=128, layers=2, heads=2, tail="gem_conv_linear",
                                               pre_norm=False, act="gelu", patch_bias=True, out_dim=64, gem_hidden=2048)),
    "vit_hf_small": ("B200ViTEncoder", dict(img=64, patch=16, width=128, layers=2, heads=2, tail="gem_linear",
                                            pre_norm=False, act="gelu", patch_bias=True, out_dim=64)),
    "swin_small": ("B200SwinEncoder", dict(img=64, patch=4, embed=64, depths=(2, 1), heads=(2, 4), window=8,
                                           pretrained_windows=(6, 6), out_dim=32)),
}


def _rel(out, ref):
    out, ref = out.reshape(len(out), -1), ref.reshape(len(ref), -1)
    return np.linalg.norm(out - ref, axis=1) / np.linalg.norm(ref, axis=1)


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "jit_small.npz"))


@pytest.mark.parametrize("tag", sorted(ARTEFACTS))
def test_artefact_reproduces_the_reference_class_and_is_recognised(golden_dir, golden, tag):
    from vsc22_submission_b200 import encoder
    module = torch.jit.load(os.path.join(golden_dir, f"{tag}.torchscript.pt")).eval()
    with torch.no_grad():
        y = module(torch.from_numpy(golden[tag + "_frames"])).numpy()
    np.testing.assert_allclose(y, golden[tag + "_out"], rtol=1e-4, atol=1e-5)
    eps = encoder._jit_layer_norm_eps(module)
    enc = encoder.encoder_from_state_dict(dict(module.state_dict()), max_frames=4, ln_eps=eps)
    cls, fields = ARTEFACTS[tag]
    assert type(enc).__name__ == cls
    for k, v in fields.items():
        assert getattr(enc.spec, k) == v, (tag, k, getattr(enc.spec, k), v)
    assert eps == pytest.approx({"clip_small": 1e-5, "swin_small": 1e-5}.get(tag, 1e-6))
    assert enc.spec.ln_eps == pytest.approx(eps)


def test_lookalike_checkpoints_are_rejected(golden_dir):
    """A checkpoint with parameters the encoder would silently ignore (e.g. the `proj` of stock CLIP visual towers, which
    is None in the reference's CLIPModel, clip.py:126-134) must not be swapped in."""
    from vsc22_submission_b200 import encoder
    sd = dict(torch.jit.load(os.path.join(golden_dir, "clip_small.torchscript.pt")).state_dict())
    sd["proj"] = torch.zeros(128, 64)
    with pytest.raises(RuntimeError, match="does not use"):
        encoder.encoder_from_state_dict(sd)
    sd.pop("proj")
    sd.pop("ln_post.bias")
    with pytest.raises(KeyError):
        encoder.encoder_from_state_dict(sd)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("bf16", 6e-3), ("fp32", 1e-4)])
@pytest.mark.parametrize("tag", sorted(ARTEFACTS))
def test_jit_load_hook_swaps_in_the_b200_encoder(golden_dir, golden, tag, precision, tol):
    from vsc22_submission_b200 import encoder
    encoder.uninstall_jit_load_hook()
    encoder.install_jit_load_hook(max_frames=3, precision=precision)
    try:
        model = torch.jit.load(os.path.join(golden_dir, f"{tag}.torchscript.pt"))     # the reference's call, unmodified
        assert type(model).__name__ == ARTEFACTS[tag][0]
        model = model.cuda().eval()                                                   # extract_ref_feats.py:25 / extractor.py:25
        with torch.no_grad():
            out = model(torch.from_numpy(golden[tag + "_frames"]).cuda()).cpu().numpy()   # 4 frames, plan chunk 3
    finally:
        encoder.uninstall_jit_load_hook()
    rel = _rel(out, golden[tag + "_out"])
    print(tag, precision, "rel L2 vs the TorchScript module's own output:", rel.max())
    assert out.shape == golden[tag + "_out"].shape and rel.max() < tol


@pytest.mark.gpu
def test_hook_leaves_other_torchscript_modules_alone(tmp_path):
    from vsc22_submission_b200 import encoder

    class Other(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(4, 2)

        def forward(self, x):
            return self.fc(x)

    path = str(tmp_path / "other.pt")
    torch.jit.save(torch.jit.script(Other()), path)
    encoder.install_jit_load_hook()
    try:
        m = torch.jit.load(path)
    finally:
        encoder.uninstall_jit_load_hook()
    assert isinstance(m, torch.jit.ScriptModule)
