"""CPU restatement (torch fp32) of the reference's Swin-V2 frame encoder -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module (see
oracle/__init__.py); the product path never does.

Follows VSC22-Descriptor-Track-1st/train/train_v106/vsc/baseline/model_factory/backbones/swinv2.py
(identical copies under train_v107 / train_v115), configured by train_v106/config_v106.py:8-24:

  forward_features            :619-633   patch embed -> 4 stages -> norm -> gem(p=3) -> output_proj
  PatchEmbed                  :460-498   Conv2d(3, C, k=4, s=4) + LayerNorm
  BasicLayer                  :381-439   depth blocks (shift 0 / window//2 alternating) + PatchMerging
  SwinTransformerBlock        :204-311   x = x + norm1(W-MSA(x)); x = x + norm2(mlp(x))   (res-post-norm)
  WindowAttention             :72-185    cosine attention, clamped exp(logit_scale), 16*sigmoid(cpb_mlp) bias
  PatchMerging                :332-369   cat(x[0::2,0::2], x[1::2,0::2], x[0::2,1::2], x[1::2,1::2]) -> Linear -> LN
  gem                         :664-665   clamp(1e-6)^p -> mean over tokens -> ^(1/p)

Parity is pinned by tests/golden/swin_small.npz: outputs of the reference's own SwinTransformerV2
class on weights produced by ``init_weights`` below (tests/golden/make_golden.py).  Parameter names
are the reference's state-dict names.
"""
from __future__ import annotations

import dataclasses
import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclasses.dataclass
class SwinSpec:
    img: int = 256
    patch: int = 4
    embed: int = 128
    depths: Tuple[int, ...] = (2, 2, 18, 2)
    heads: Tuple[int, ...] = (4, 8, 16, 32)
    window: int = 16
    pretrained_windows: Tuple[int, ...] = (12, 12, 12, 6)
    out_dim: int = 512
    gem_p: float = 3.0
    ln_eps: float = 1e-5

    def stage(self, i: int):
        """(dim, resolution, effective window) of stage i (swinv2.py:223-226, 574-577)."""
        res = self.img // self.patch // (2 ** i)
        return self.embed * 2 ** i, res, min(self.window, res)

    def shift(self, i: int, j: int) -> int:
        _, res, _ = self.stage(i)
        if res <= self.window:           # window covers the map: no shift (swinv2.py:223-226)
            return 0
        return 0 if j % 2 == 0 else self.window // 2

    def flops_per_frame(self) -> float:
        fl = 2.0 * (self.img // self.patch) ** 2 * 3 * self.patch ** 2 * self.embed
        for i, depth in enumerate(self.depths):
            C, res, ws = self.stage(i)
            L, N = res * res, ws * ws
            fl += depth * (2.0 * L * C * 3 * C + 2.0 * L * C * C + 4.0 * L * N * C + 16.0 * L * C * C)
            if i + 1 < len(self.depths):
                fl += 2.0 * (L // 4) * 4 * C * 2 * C
        fl += 2.0 * self.embed * 2 ** (len(self.depths) - 1) * self.out_dim
        return fl


SWINV2_B_256 = SwinSpec()                                   # config_v106.py (swinv2_v106 / v107 / v115 checkpoints)
SWINV2_L_384 = SwinSpec(img=384, patch=4, embed=192, depths=(2, 2, 18, 2), heads=(6, 12, 24, 48), window=24,
                        pretrained_windows=(12, 12, 12, 6), out_dim=512)     # BASELINE configs[3]


def param_shapes(spec: SwinSpec) -> "OrderedDict[str, tuple]":
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["patch_embed.proj.weight"] = (spec.embed, 3, spec.patch, spec.patch)
    s["patch_embed.proj.bias"] = (spec.embed,)
    s["patch_embed.norm.weight"] = (spec.embed,)
    s["patch_embed.norm.bias"] = (spec.embed,)
    for i, depth in enumerate(spec.depths):
        C, _, _ = spec.stage(i)
        nH = spec.heads[i]
        for j in range(depth):
            p = f"layers.{i}.blocks.{j}."
            s[p + "norm1.weight"] = (C,); s[p + "norm1.bias"] = (C,)
            s[p + "attn.logit_scale"] = (nH, 1, 1)
            s[p + "attn.cpb_mlp.0.weight"] = (512, 2); s[p + "attn.cpb_mlp.0.bias"] = (512,)
            s[p + "attn.cpb_mlp.2.weight"] = (nH, 512)
            s[p + "attn.qkv.weight"] = (3 * C, C)
            s[p + "attn.q_bias"] = (C,); s[p + "attn.v_bias"] = (C,)
            s[p + "attn.proj.weight"] = (C, C); s[p + "attn.proj.bias"] = (C,)
            s[p + "norm2.weight"] = (C,); s[p + "norm2.bias"] = (C,)
            s[p + "mlp.fc1.weight"] = (4 * C, C); s[p + "mlp.fc1.bias"] = (4 * C,)
            s[p + "mlp.fc2.weight"] = (C, 4 * C); s[p + "mlp.fc2.bias"] = (C,)
        if i + 1 < len(spec.depths):
            s[f"layers.{i}.downsample.reduction.weight"] = (2 * C, 4 * C)
            s[f"layers.{i}.downsample.norm.weight"] = (2 * C,)
            s[f"layers.{i}.downsample.norm.bias"] = (2 * C,)
    Cf = spec.embed * 2 ** (len(spec.depths) - 1)
    s["norm.weight"] = (Cf,); s["norm.bias"] = (Cf,)
    s["output_proj.weight"] = (spec.out_dim, Cf); s["output_proj.bias"] = (spec.out_dim,)
    return s


def init_weights(spec: SwinSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random weights keyed by parameter NAME (any subset can be regenerated independently).
    The reference zero-initialises every res-post-norm (swinv2.py:452-457), which would turn each block into
    the identity; LayerNorm affines are therefore drawn around (1, 0) and the post-norms around 0.5."""
    w = {}
    for name, shape in param_shapes(spec).items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
        r = torch.randn(shape, generator=g)
        if name.endswith("logit_scale"):
            v = math.log(10.0) + 0.3 * r
        elif "norm" in name and name.endswith("weight"):
            v = (0.5 if ".blocks." in name else 1.0) + 0.1 * r
        elif "norm" in name and name.endswith("bias"):
            v = 0.05 * r
        elif name.endswith("bias"):
            v = 0.02 * r
        elif "cpb_mlp.0.weight" in name:
            v = 0.5 * r
        elif "cpb_mlp.2.weight" in name:
            v = 0.05 * r
        elif name == "patch_embed.proj.weight":
            v = r * (3 * spec.patch ** 2) ** -0.5
        else:
            v = r * shape[-1] ** -0.5
        w[name] = v.float()
    return w


def relative_position_bias(w, prefix: str, ws: int, pretrained_ws: int, nH: int) -> torch.Tensor:
    """16 * sigmoid(cpb_mlp(log-spaced relative coords))[index] -> [nH, N, N]  (swinv2.py:99-131, 165-170)."""
    rel = torch.arange(-(ws - 1), ws, dtype=torch.float32)
    table = torch.stack(torch.meshgrid([rel, rel], indexing="ij")).permute(1, 2, 0).contiguous()   # [2ws-1, 2ws-1, 2]
    table = table / ((pretrained_ws - 1) if pretrained_ws > 0 else (ws - 1))
    table = table * 8
    table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
    hidden = F.relu(F.linear(table.view(-1, 2), w[prefix + "cpb_mlp.0.weight"], w[prefix + "cpb_mlp.0.bias"]))
    tbl = F.linear(hidden, w[prefix + "cpb_mlp.2.weight"])                                         # [(2ws-1)^2, nH]
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij")).flatten(1)
    relc = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0) + (ws - 1)
    index = relc[:, :, 0] * (2 * ws - 1) + relc[:, :, 1]                                           # [N, N]
    bias = tbl[index.view(-1)].view(ws * ws, ws * ws, nH).permute(2, 0, 1)
    return 16 * torch.sigmoid(bias)


def bias_table(w, prefix: str, ws: int, pretrained_ws: int) -> torch.Tensor:
    """The same bias before the index gather: [nH, (2ws-1)^2] (what the CUDA plan uploads)."""
    rel = torch.arange(-(ws - 1), ws, dtype=torch.float32)
    table = torch.stack(torch.meshgrid([rel, rel], indexing="ij")).permute(1, 2, 0).contiguous()
    table = table / ((pretrained_ws - 1) if pretrained_ws > 0 else (ws - 1))
    table = table * 8
    table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / math.log2(8)
    hidden = F.relu(F.linear(table.view(-1, 2), w[prefix + "cpb_mlp.0.weight"], w[prefix + "cpb_mlp.0.bias"]))
    return (16 * torch.sigmoid(F.linear(hidden, w[prefix + "cpb_mlp.2.weight"]))).t().contiguous()


def shifted_window_mask(res: int, ws: int, shift: int):
    """[nW, N, N] of 0 / -100 (swinv2.py:232-255)."""
    if shift == 0:
        return None
    img = torch.zeros((res, res))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[hs, wsl] = cnt
            cnt += 1
    win = img.view(res // ws, ws, res // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = win[:, None, :] - win[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def _windows(x, ws):        # [B, H, W, C] -> [B*nW, ws*ws, C]   (window_partition :39-51)
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def _unwindows(xw, ws, H, W):   # inverse (window_reverse :54-69)
    B = xw.shape[0] // ((H // ws) * (W // ws))
    return xw.view(B, H // ws, W // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.bfloat16().float()


def window_attention(w, p: str, xw: torch.Tensor, nH: int, bias: torch.Tensor, mask, r=lambda t: t, matched=False):
    """swinv2.py:147-185.  ``matched``: operand rounding points of the CUDA encoder (see ``forward``)."""
    B_, N, C = xw.shape
    qkv_bias = torch.cat((w[p + "q_bias"], torch.zeros_like(w[p + "v_bias"]), w[p + "v_bias"]))
    qkv = F.linear(r(xw), r(w[p + "qkv.weight"]), qkv_bias).reshape(B_, N, 3, nH, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    scale = torch.clamp(w[p + "logit_scale"], max=math.log(1.0 / 0.01)).exp()
    if not matched:
        attn = F.normalize(q, dim=-1) @ F.normalize(k, dim=-1).transpose(-2, -1)
        attn = attn * scale
    else:      # the scale rides on q before the bf16 rounding (QKV-projection epilogue of the CUDA encoder)
        attn = r(F.normalize(q, dim=-1) * scale) @ r(F.normalize(k, dim=-1)).transpose(-2, -1)
        v = r(v)
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, nH, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, nH, N, N)
    if not matched:
        attn = attn.softmax(dim=-1)
        out = (attn @ v).transpose(1, 2).reshape(B_, N, C)
    else:      # un-normalised probabilities are the bf16 operand; the row sum stays fp32
        pr = torch.exp(attn - attn.amax(dim=-1, keepdim=True))
        out = ((r(pr) @ v) / pr.sum(dim=-1, keepdim=True)).transpose(1, 2).reshape(B_, N, C)
    return F.linear(r(out), r(w[p + "proj.weight"]), w[p + "proj.bias"])


@torch.no_grad()
def forward(spec: SwinSpec, w: Dict[str, torch.Tensor], frames: torch.Tensor, return_tokens: bool = False,
            precision: str = "fp32"):
    """``precision="fp32"``: the reference's arithmetic, bit for bit.  ``precision="bf16"``: the matched-precision oracle --
    every matrix-product operand (patch pixels, weight matrices, projection inputs, normalised q (with its logit scale) / k,
    v, un-normalised softmax probabilities, attention output, activated MLP hidden) rounded to bf16 where the CUDA encoder
    rounds it; accumulators, biases, CPB bias table, residual stream, LayerNorm, softmax statistics and the tail fp32."""
    assert precision in ("fp32", "bf16")
    matched = precision == "bf16"
    r = _bf16 if matched else (lambda t: t)
    eps = spec.ln_eps
    x = F.conv2d(r(frames.float()), r(w["patch_embed.proj.weight"]), w["patch_embed.proj.bias"], stride=spec.patch)
    x = x.flatten(2).transpose(1, 2)
    x = F.layer_norm(x, x.shape[-1:], w["patch_embed.norm.weight"], w["patch_embed.norm.bias"], eps)
    B = x.shape[0]
    for i, depth in enumerate(spec.depths):
        C, res, ws = spec.stage(i)
        nH = spec.heads[i]
        for j in range(depth):
            p = f"layers.{i}.blocks.{j}."
            shift = spec.shift(i, j)
            bias = relative_position_bias(w, p + "attn.", ws, spec.pretrained_windows[i], nH)
            mask = shifted_window_mask(res, ws, shift)
            h = x.view(B, res, res, C)
            if shift:
                h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
            a = window_attention(w, p + "attn.", _windows(h, ws), nH, bias, mask, r, matched)
            h = _unwindows(a, ws, res, res)
            if shift:
                h = torch.roll(h, shifts=(shift, shift), dims=(1, 2))
            x = x + F.layer_norm(h.reshape(B, res * res, C), (C,), w[p + "norm1.weight"], w[p + "norm1.bias"], eps)
            m = F.linear(r(F.gelu(F.linear(r(x), r(w[p + "mlp.fc1.weight"]), w[p + "mlp.fc1.bias"]))),
                         r(w[p + "mlp.fc2.weight"]), w[p + "mlp.fc2.bias"])
            x = x + F.layer_norm(m, (C,), w[p + "norm2.weight"], w[p + "norm2.bias"], eps)
        if i + 1 < len(spec.depths):
            h = x.view(B, res, res, C)
            h = torch.cat([h[:, 0::2, 0::2], h[:, 1::2, 0::2], h[:, 0::2, 1::2], h[:, 1::2, 1::2]], -1)
            h = F.linear(r(h.reshape(B, -1, 4 * C)), r(w[f"layers.{i}.downsample.reduction.weight"]))
            x = F.layer_norm(h, (2 * C,), w[f"layers.{i}.downsample.norm.weight"], w[f"layers.{i}.downsample.norm.bias"], eps)
    x = F.layer_norm(x, x.shape[-1:], w["norm.weight"], w["norm.bias"], eps)
    if return_tokens:
        return x
    g = x.clamp(min=1e-6).pow(spec.gem_p).mean(dim=1).pow(1.0 / spec.gem_p)
    return F.linear(g, w["output_proj.weight"], w["output_proj.bias"])
