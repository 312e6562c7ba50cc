"""torch-fp32 CPU restatement of the reference's ViT frame encoders (path A).

TEST INFRASTRUCTURE (see oracle/__init__.py).  One functional forward over a flat weight
dict covers the three ViT flavours on the reference's inference path:

* **CLIP flavour** -- ``D/train/train_vid_score/video/clip.py``:
    conv patch-embed, no bias (:105, :143) -> cat class token (:144-150) -> + positional
    (:151) -> ln_pre (:152) -> L x [x + MHA(ln_1 x); x + MLP(ln_2 x)] (:47-50) with
    nn.MultiheadAttention (:33, q scaled by d_h^-0.5, softmax over keys, out_proj) and
    QuickGELU ``x*sigmoid(1.702x)`` (:22-25) -> ln_post (:158); returns all tokens [N,T,W].
* **HF-ViT backbone tail** -- ``D/train/train_v106/vsc/baseline/model_factory/backbones/vit.py:42-54``:
    ``gem(tokens, p=3) = clamp(1e-6)^3 -> mean over tokens -> ^(1/3)`` then Linear(W -> out).
    (BASELINE config 2 = CLIP-flavour ViT-B/16@224 + this tail)
* **timm flavour (vit_v68)** -- timm==0.6.12 ``VisionTransformer`` as configured at
    ``D/train/train_v68/vsc/baseline/model_factory/backbones/sscd.py:78``
    (``timm.create_model('vit_base_patch32_384', global_pool='', num_classes=0)``): conv WITH
    bias, no ln_pre, LayerNorm eps 1e-6, exact-erf GELU, final norm, all 145 tokens incl. cls
    returned; head = GlobalGeMPool2d with Conv1d(768->2048,k=1) (sscd.py:30-40) +
    Linear(2048->512) (sscd.py:86).  timm itself is NOT in /root/reference (pinned
    ``timm==0.6.12``, dockerfile:33): its published forward is restated here and cross-checked
    against the installed ``transformers.ViTModel`` (tests/test_oracle_vit.py).

Pinning: CLIP flavour + gem tail are pinned against the reference's own ``CLIPModel`` class on
seeded weights (tests/golden/vit_clip_small.npz, made by tests/golden/make_golden.py which
imports the class from /root/reference).  The reference holds no encoder test and ships no
weights, so beyond that **parity is unpinned** for (A).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict

import torch
import torch.nn.functional as F


@dataclasses.dataclass
class VitSpec:
    img: int = 224
    patch: int = 16
    width: int = 768
    layers: int = 12
    heads: int = 12
    patch_bias: bool = False      # timm: True, CLIP: False
    pre_norm: bool = True         # CLIP ln_pre
    act: str = "quick_gelu"       # "quick_gelu" | "gelu"
    ln_eps: float = 1e-5          # timm ViT: 1e-6
    tail: str = "tokens"          # "tokens" | "gem_linear" | "gem_conv_linear"
    out_dim: int = 512
    gem_p: float = 3.0
    gem_hidden: int = 2048

    @property
    def tokens(self) -> int:
        return (self.img // self.patch) ** 2 + 1

    @property
    def head_dim(self) -> int:
        return self.width // self.heads

    def flops_per_frame(self) -> float:
        T, W, L = self.tokens, self.width, self.layers
        per_layer = 2 * T * W * (3 * W) + 2 * T * W * W + 4 * T * W * (4 * W) + 4 * T * T * W
        patch = 2 * (T - 1) * (3 * self.patch ** 2) * W
        tail = 0
        if self.tail == "gem_linear":
            tail = 2 * W * self.out_dim
        elif self.tail == "gem_conv_linear":
            tail = 2 * T * W * self.gem_hidden + 2 * self.gem_hidden * self.out_dim
        return float(L * per_layer + patch + tail)


CLIP_B16_224 = VitSpec(224, 16, 768, 12, 12, tail="gem_linear")          # BASELINE config 2
CLIP_L14_224 = VitSpec(224, 14, 1024, 24, 16, tail="tokens")             # shipped CLIP ViT-L/14
CLIP_L16_384 = VitSpec(384, 16, 1024, 24, 16, tail="tokens")             # BASELINE configs[3] (T = 577)
TIMM_B32_384 = VitSpec(384, 32, 768, 12, 12, patch_bias=True, pre_norm=False, act="gelu",
                       ln_eps=1e-6, tail="gem_conv_linear")              # vit_v68


def init_weights(spec: VitSpec, seed: int = 0, scale_std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the reference's init *distributions* (clip.py:107-124:
    class/pos ~ W^-0.5 * N(0,1); Linear ~ N(0, 0.02), bias 0; LayerNorm (1, 0); conv default
    kaiming-uniform).  LayerNorm affine and biases are additionally perturbed so that no
    parameter is a trivial 0/1 that would hide a wiring bug."""
    g = torch.Generator().manual_seed(seed)
    W, T, L = spec.width, spec.tokens, spec.layers

    def n(*shape, std=scale_std):
        return torch.randn(*shape, generator=g) * std

    w: Dict[str, torch.Tensor] = {}
    fan_in = 3 * spec.patch ** 2
    bound = 1.0 / math.sqrt(fan_in)
    w["patch_w"] = (torch.rand(W, 3, spec.patch, spec.patch, generator=g) * 2 - 1) * bound
    if spec.patch_bias:
        w["patch_b"] = n(W)
    w["cls"] = n(W, std=W ** -0.5)
    w["pos"] = n(T, W, std=W ** -0.5)
    if spec.pre_norm:
        w["ln_pre_w"] = 1.0 + n(W, std=0.1)
        w["ln_pre_b"] = n(W, std=0.1)
    for l in range(L):
        p = f"l{l}."
        w[p + "ln1_w"] = 1.0 + n(W, std=0.1)
        w[p + "ln1_b"] = n(W, std=0.1)
        w[p + "qkv_w"] = n(3 * W, W)
        w[p + "qkv_b"] = n(3 * W)
        w[p + "proj_w"] = n(W, W)
        w[p + "proj_b"] = n(W)
        w[p + "ln2_w"] = 1.0 + n(W, std=0.1)
        w[p + "ln2_b"] = n(W, std=0.1)
        w[p + "fc1_w"] = n(4 * W, W)
        w[p + "fc1_b"] = n(4 * W)
        w[p + "fc2_w"] = n(W, 4 * W)
        w[p + "fc2_b"] = n(W)
    w["ln_post_w"] = 1.0 + n(W, std=0.1)
    w["ln_post_b"] = n(W, std=0.1)
    if spec.tail == "gem_linear":
        w["head_w"] = n(spec.out_dim, W, std=W ** -0.5)
        w["head_b"] = n(spec.out_dim)
    elif spec.tail == "gem_conv_linear":
        w["gem_conv_w"] = n(spec.gem_hidden, W, std=W ** -0.5)
        w["gem_conv_b"] = n(spec.gem_hidden, std=0.2)
        w["head_w"] = n(spec.out_dim, spec.gem_hidden, std=spec.gem_hidden ** -0.5)
        w["head_b"] = n(spec.out_dim)
    return w


def from_clip_state_dict(sd: Dict[str, torch.Tensor], layers: int) -> Dict[str, torch.Tensor]:
    """Flat weight dict from a reference ``CLIPModel.state_dict()`` (clip.py:85-124 names)."""
    w = {"patch_w": sd["conv1.weight"], "cls": sd["class_embedding"], "pos": sd["positional_embedding"],
         "ln_pre_w": sd["ln_pre.weight"], "ln_pre_b": sd["ln_pre.bias"],
         "ln_post_w": sd["ln_post.weight"], "ln_post_b": sd["ln_post.bias"]}
    for l in range(layers):
        s, p = f"transformer.resblocks.{l}.", f"l{l}."
        w[p + "ln1_w"], w[p + "ln1_b"] = sd[s + "ln_1.weight"], sd[s + "ln_1.bias"]
        w[p + "qkv_w"], w[p + "qkv_b"] = sd[s + "attn.in_proj_weight"], sd[s + "attn.in_proj_bias"]
        w[p + "proj_w"], w[p + "proj_b"] = sd[s + "attn.out_proj.weight"], sd[s + "attn.out_proj.bias"]
        w[p + "ln2_w"], w[p + "ln2_b"] = sd[s + "ln_2.weight"], sd[s + "ln_2.bias"]
        w[p + "fc1_w"], w[p + "fc1_b"] = sd[s + "mlp.c_fc.weight"], sd[s + "mlp.c_fc.bias"]
        w[p + "fc2_w"], w[p + "fc2_b"] = sd[s + "mlp.c_proj.weight"], sd[s + "mlp.c_proj.bias"]
    return {k: v.detach().clone().float() for k, v in w.items()}


def _act(x, kind):
    if kind == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)          # clip.py:24
    return F.gelu(x)                                  # timm Mlp: nn.GELU (exact erf)


def gem_tokens(x, p=3.0, eps=1e-6):
    """backbones/vit.py:56-58."""
    return x.clamp(min=eps).pow(p).mean(dim=1).pow(1.0 / p)


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.bfloat16().float()


@torch.no_grad()
def forward(spec: VitSpec, w: Dict[str, torch.Tensor], frames: torch.Tensor,
            return_tokens: bool = False, precision: str = "fp32") -> torch.Tensor:
    """frames [N,3,H,W] float32 -> [N,T,W] (tail 'tokens') or [N,out_dim].

    ``precision="fp32"``: the reference's arithmetic (every op fp32).
    ``precision="bf16"``: the MATCHED-PRECISION oracle (SURVEY.md 8d "vs oracle in matching precision"): every
    matrix-product OPERAND is rounded to bf16 at the point where the CUDA encoder rounds it -- the patch pixels and every
    weight matrix, the LayerNorm outputs that feed a projection, q / k / v, the un-normalised softmax probabilities
    (the row sum stays fp32), the attention output, the activated MLP hidden -- while accumulation, biases, the residual
    stream, LayerNorm statistics, softmax and the GeM / Linear tail stay fp32, exactly like the reference under
    ``torch.autocast(bfloat16)`` would keep them."""
    assert precision in ("fp32", "bf16")
    r = _bf16 if precision == "bf16" else (lambda t: t)
    x = frames.float()
    N = x.shape[0]
    W, H, dh = spec.width, spec.heads, spec.head_dim
    x = F.conv2d(r(x), r(w["patch_w"]), w.get("patch_b"), stride=spec.patch)  # clip.py:143
    x = x.reshape(N, W, -1).permute(0, 2, 1)                                # :144-145
    x = torch.cat([w["cls"].expand(N, 1, W), x], dim=1)                     # :146-150
    x = x + w["pos"]                                                        # :151
    if spec.pre_norm:
        x = F.layer_norm(x, (W,), w["ln_pre_w"], w["ln_pre_b"], spec.ln_eps)  # :152
    T = x.shape[1]
    for l in range(spec.layers):
        p = f"l{l}."
        h = F.layer_norm(x, (W,), w[p + "ln1_w"], w[p + "ln1_b"], spec.ln_eps)
        qkv = r(F.linear(r(h), r(w[p + "qkv_w"]), w[p + "qkv_b"])).reshape(N, T, 3, H, dh)
        q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))      # [N,H,T,dh]
        if precision == "fp32":
            att = torch.softmax((q * dh ** -0.5) @ k.transpose(-1, -2), dim=-1)
            o = (att @ v).permute(0, 2, 1, 3).reshape(N, T, W)
        else:
            s_ = (q @ k.transpose(-1, -2)) * dh ** -0.5
            pr = torch.exp(s_ - s_.amax(dim=-1, keepdim=True))
            o = ((r(pr) @ v) / pr.sum(dim=-1, keepdim=True)).permute(0, 2, 1, 3).reshape(N, T, W)
        x = x + F.linear(r(o), r(w[p + "proj_w"]), w[p + "proj_b"])         # clip.py:48
        h = F.layer_norm(x, (W,), w[p + "ln2_w"], w[p + "ln2_b"], spec.ln_eps)
        h = _act(F.linear(r(h), r(w[p + "fc1_w"]), w[p + "fc1_b"]), spec.act)
        x = x + F.linear(r(h), r(w[p + "fc2_w"]), w[p + "fc2_b"])           # clip.py:49
    x = F.layer_norm(x, (W,), w["ln_post_w"], w["ln_post_b"], spec.ln_eps)  # :158
    if spec.tail == "tokens" or return_tokens:
        return x
    if spec.tail == "gem_linear":
        return F.linear(gem_tokens(x, spec.gem_p), w["head_w"], w["head_b"])
    if spec.tail == "gem_conv_linear":                                      # sscd.py:30-40, :86
        y = F.linear(r(x), r(w["gem_conv_w"]), w["gem_conv_b"])             # Conv1d k=1 over tokens
        g = y.clamp(min=1e-6).pow(spec.gem_p).mean(dim=1).pow(1.0 / spec.gem_p)
        return F.linear(g, w["head_w"], w["head_b"])
    raise ValueError(spec.tail)
