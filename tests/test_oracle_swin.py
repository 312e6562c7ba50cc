"""The Swin-V2 oracle (oracle/swin_ref.py) against the reference's own class: committed golden outputs
(tests/golden/swin_small.npz, made by tests/golden/make_golden.py from swinv2.py:502) and, when the reference
tree is present (authoring container), a live run of the reference class on a second configuration."""
import os

import numpy as np
import pytest
import torch

from oracle import refload, swin_ref

SMALL = swin_ref.SwinSpec(img=128, patch=4, embed=64, depths=(2, 2, 2, 2), heads=(2, 4, 8, 16), window=8,
                          pretrained_windows=(6, 6, 6, 3), out_dim=64)


def test_oracle_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "swin_small.npz"))
    w = swin_ref.init_weights(SMALL, seed=0)
    frames = torch.from_numpy(g["frames"])
    tokens = swin_ref.forward(SMALL, w, frames, return_tokens=True).numpy()
    desc = swin_ref.forward(SMALL, w, frames).numpy()
    np.testing.assert_allclose(tokens, g["tokens"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(desc, g["desc"], rtol=1e-5, atol=1e-5)


def test_bias_table_is_the_gathered_bias():
    w = swin_ref.init_weights(SMALL, seed=3)
    p = "layers.1.blocks.1.attn."
    ws, nH = 8, 4
    full = swin_ref.relative_position_bias(w, p, ws, 6, nH)
    tab = swin_ref.bias_table(w, p, ws, 6)
    ts = 2 * ws - 1
    for (i, j) in [(0, 0), (5, 60), (63, 0), (17, 42)]:
        yi, xi, yj, xj = i // ws, i % ws, j // ws, j % ws
        e = (yi - yj + ws - 1) * ts + (xi - xj + ws - 1)
        np.testing.assert_allclose(full[:, i, j].numpy(), tab[:, e].numpy(), rtol=1e-6)


def test_shift_rule_and_mask_shape():
    spec = swin_ref.SWINV2_B_256
    assert [spec.stage(i) for i in range(4)] == [(128, 64, 16), (256, 32, 16), (512, 16, 16), (1024, 8, 8)]
    assert [spec.shift(i, 1) for i in range(4)] == [8, 8, 0, 0]       # window == map: no shift (swinv2.py:223-226)
    m = swin_ref.shifted_window_mask(32, 16, 8)
    assert m.shape == (4, 256, 256) and float(m[0].abs().sum()) == 0.0 and float(m[3].min()) == -100.0


@pytest.mark.skipif(not refload.available(), reason="/root/reference is only present in the authoring container")
def test_oracle_matches_live_reference_class():
    mod = refload.swinv2_module()
    spec = swin_ref.SwinSpec(img=64, patch=4, embed=64, depths=(2, 1), heads=(2, 4), window=4, pretrained_windows=(0, 3),
                             out_dim=32)
    model = mod.SwinTransformerV2(img_size=spec.img, patch_size=spec.patch, embed_dim=spec.embed, depths=list(spec.depths),
                                  num_heads=list(spec.heads), window_size=spec.window,
                                  pretrained_window_sizes=list(spec.pretrained_windows), output_dim=spec.out_dim,
                                  p=spec.gem_p, pretrained=None)
    w = swin_ref.init_weights(spec, seed=5)
    model.load_state_dict(w, strict=False)
    model.eval()
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(2)).clamp(-1, 1)
    with torch.no_grad():
        ref = model(x)
    np.testing.assert_allclose(swin_ref.forward(spec, w, x).numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not refload.available(), reason="/root/reference is only present in the authoring container")
def test_spec_is_recovered_from_a_reference_state_dict():
    """install_jit_load_hook path: the architecture of a swinv2 checkpoint is read off its state dict
    (parameter shapes + the registered buffers relative_position_index / attn_mask / relative_coords_table)."""
    from vsc22_submission_b200.swin_encoder import param_names, spec_from_state_dict
    mod = refload.swinv2_module()
    for kw in (dict(img_size=128, patch_size=4, embed_dim=64, depths=[2, 2, 2, 2], num_heads=[2, 4, 8, 16], window_size=8,
                    pretrained_window_sizes=[6, 6, 6, 3], output_dim=64),
               dict(img_size=256, patch_size=4, embed_dim=128, depths=[2, 2, 2, 2], num_heads=[4, 8, 16, 32], window_size=16,
                    pretrained_window_sizes=[12, 12, 12, 6], output_dim=512)):
        model = mod.SwinTransformerV2(pretrained=None, **kw)
        sd = model.state_dict()
        spec = spec_from_state_dict(sd)
        assert (spec.img, spec.patch, spec.embed, list(spec.depths), list(spec.heads), spec.window,
                list(spec.pretrained_windows), spec.out_dim) == (
            kw["img_size"], kw["patch_size"], kw["embed_dim"], kw["depths"], kw["num_heads"], kw["window_size"],
            kw["pretrained_window_sizes"], kw["output_dim"])
        assert set(param_names(spec)) == {k for k, _ in model.named_parameters()}
