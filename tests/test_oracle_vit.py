"""Oracle (A): restated ViT forward against (i) the golden vectors produced by the reference's
own CLIPModel class and (ii) transformers.ViTModel for the timm flavour -- CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import vit_ref


def test_clip_flavour_matches_reference_class(golden_dir):
    g = np.load(os.path.join(golden_dir, "vit_clip_small.npz"))
    img, patch, width, layers, heads, out_dim = (int(v) for v in g["spec"])
    spec = vit_ref.VitSpec(img, patch, width, layers, heads, tail="gem_linear", out_dim=out_dim)
    w = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")}
    frames = torch.from_numpy(g["frames"])
    tokens = vit_ref.forward(spec, w, frames, return_tokens=True).numpy()
    desc = vit_ref.forward(spec, w, frames).numpy()
    np.testing.assert_allclose(tokens, g["tokens"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(desc, g["desc"], rtol=1e-4, atol=2e-5)


def test_timm_flavour_matches_hf_vit():
    """timm 0.6.12 VisionTransformer == HF ViTModel arithmetic (conv+bias, no ln_pre, eps 1e-6 as
    configured, erf-GELU, final LayerNorm, all tokens) -- the reference itself wraps HF ViTModel
    as backbone `VIT` (backbones/vit.py:10-54)."""
    transformers = pytest.importorskip("transformers")
    spec = vit_ref.VitSpec(img=64, patch=32, width=64, layers=2, heads=2, patch_bias=True, pre_norm=False,
                           act="gelu", ln_eps=1e-6, tail="tokens")
    cfg = transformers.ViTConfig(hidden_size=64, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                                 image_size=64, patch_size=32, hidden_act="gelu", layer_norm_eps=1e-6,
                                 hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, qkv_bias=True)
    torch.manual_seed(0)
    hf = transformers.ViTModel(cfg, add_pooling_layer=False).eval()
    sd = hf.state_dict()
    w = {"patch_w": sd["embeddings.patch_embeddings.projection.weight"],
         "patch_b": sd["embeddings.patch_embeddings.projection.bias"],
         "cls": sd["embeddings.cls_token"].reshape(-1), "pos": sd["embeddings.position_embeddings"][0],
         "ln_post_w": sd["layernorm.weight"], "ln_post_b": sd["layernorm.bias"]}
    for l in range(2):
        s, p = f"encoder.layer.{l}.", f"l{l}."
        w[p + "ln1_w"], w[p + "ln1_b"] = sd[s + "layernorm_before.weight"], sd[s + "layernorm_before.bias"]
        w[p + "qkv_w"] = torch.cat([sd[s + f"attention.attention.{n}.weight"] for n in ("query", "key", "value")])
        w[p + "qkv_b"] = torch.cat([sd[s + f"attention.attention.{n}.bias"] for n in ("query", "key", "value")])
        w[p + "proj_w"], w[p + "proj_b"] = sd[s + "attention.output.dense.weight"], sd[s + "attention.output.dense.bias"]
        w[p + "ln2_w"], w[p + "ln2_b"] = sd[s + "layernorm_after.weight"], sd[s + "layernorm_after.bias"]
        w[p + "fc1_w"], w[p + "fc1_b"] = sd[s + "intermediate.dense.weight"], sd[s + "intermediate.dense.bias"]
        w[p + "fc2_w"], w[p + "fc2_b"] = sd[s + "output.dense.weight"], sd[s + "output.dense.bias"]
    # HF adds pos to cls too and has no separate class/pos split: same as clip.py:146-151
    x = torch.randn(3, 3, 64, 64)
    with torch.no_grad():
        ref = hf(pixel_values=x).last_hidden_state
    out = vit_ref.forward(spec, w, x)
    np.testing.assert_allclose(out.numpy(), ref.numpy(), rtol=1e-4, atol=2e-5)


def test_flops_table():
    # 35.13 GFLOP/frame for ViT-B/16@224 (12 x 2.908 G + 0.231 G patch-embed)
    f = vit_ref.CLIP_B16_224.flops_per_frame()
    assert abs(f / 1e9 - 35.13) < 0.05
