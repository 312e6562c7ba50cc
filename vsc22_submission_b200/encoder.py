"""Seam (A): an ``nn.Module`` that quacks like the reference's ``torch.jit.load``-ed frame encoder.

Reference call sites: ``torch.jit.load(ckpt).to(dev)`` / ``.cuda()`` / ``.eval()`` /
``DDP(model, device_ids=[rank], find_unused_parameters=True)`` then ``model(frames)`` under
``torch.no_grad()`` -- D/infer/extract_ref_feats.py:24-27, D/infer/src/extractor.py:25,
D/infer/extract_query_feats.py:77-92,143-153, M/infer/infer_matching.py:84-133.

``forward(x: float32 CUDA [n,3,H,W]) -> float32 CUDA [n,out_dim]`` (SSCD flavour) or ``[n,T,W]`` (CLIP
flavour; the caller slices ``[:,0]``, extract_query_feats.py:149-150).  The arithmetic runs in
libvscb200.so (tcgen05 GEMMs, fused MHSA, fused LN / GeM tails) on PyTorch's *current stream*, so an
immediate ``.cpu()`` is ordered correctly.  No PyTorch ops are on the compute path and there is no
fallback: a CPU input raises.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import re
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib


@dataclasses.dataclass
class VitSpec:
    """Architecture of one ViT encoder (fields mirror vscb200_vit_spec in include/vscb200.h)."""
    img: int = 224
    patch: int = 16
    width: int = 768
    layers: int = 12
    heads: int = 12
    patch_bias: bool = False      # timm: True; CLIP conv1 has no bias (clip.py:105)
    pre_norm: bool = True         # CLIP ln_pre (clip.py:152)
    act: str = "quick_gelu"       # "quick_gelu" (clip.py:22-25) | "gelu" (timm)
    ln_eps: float = 1e-5          # timm ViT: 1e-6
    tail: str = "tokens"          # "tokens" | "gem_linear" (backbones/vit.py:42-58) | "gem_conv_linear" (sscd.py:30-40,86)
    out_dim: int = 512
    gem_p: float = 3.0
    gem_hidden: int = 2048
    # "bf16": tensor-core operands rounded to bf16 (throughput mode; == the matched-precision oracle within 1e-3);
    # "fp32": fp32-equivalent arithmetic (split-bf16 tcgen05 GEMMs, fp32 attention; == the reference's fp32 module
    #         within 1e-3 -- the reference calls its encoders without autocast, D/infer/src/extractor.py:25)
    precision: str = "bf16"

    @property
    def tokens(self) -> int:
        return (self.img // self.patch) ** 2 + 1

    def flops_per_frame(self) -> float:
        T, W, L = self.tokens, self.width, self.layers
        per_layer = 2 * T * W * (3 * W) + 2 * T * W * W + 4 * T * W * (4 * W) + 4 * T * T * W
        flops = L * per_layer + 2 * (T - 1) * (3 * self.patch ** 2) * W
        if self.tail == "gem_linear":
            flops += 2 * W * self.out_dim
        elif self.tail == "gem_conv_linear":
            flops += 2 * T * W * self.gem_hidden + 2 * self.gem_hidden * self.out_dim
        return float(flops)

    def to_c(self) -> "_lib.VitSpecC":
        return _lib.VitSpecC(
            img=self.img, patch=self.patch, width=self.width, layers=self.layers, heads=self.heads,
            patch_bias=int(self.patch_bias), pre_norm=int(self.pre_norm),
            act={"quick_gelu": _lib.ACT_QUICK_GELU, "gelu": _lib.ACT_GELU}[self.act],
            tail={"tokens": _lib.TAIL_TOKENS, "gem_linear": _lib.TAIL_GEM_LINEAR,
                  "gem_conv_linear": _lib.TAIL_GEM_CONV_LINEAR}[self.tail],
            out_dim=self.out_dim, gem_hidden=self.gem_hidden, ln_eps=self.ln_eps, gem_p=self.gem_p,
            precision=_lib.PRECISION[self.precision])


# Named configurations on the reference's inference path / in BASELINE.json
VIT_B16_224_GEM = VitSpec(224, 16, 768, 12, 12, tail="gem_linear")                    # BASELINE config 2
CLIP_L14_224 = VitSpec(224, 14, 1024, 24, 16, tail="tokens")                          # extract_query_feats.py:77
VIT_L16_384 = VitSpec(384, 16, 1024, 24, 16, tail="tokens")                           # BASELINE configs[3] "ViT-L ... 384^2" (T = 577)
VIT_V68 = VitSpec(384, 32, 768, 12, 12, patch_bias=True, pre_norm=False, act="gelu", ln_eps=1e-6,
                  tail="gem_conv_linear")                                              # infer_ref.sh vit_v68


def param_names(spec: VitSpec):
    names = ["patch_w", "cls", "pos", "ln_post_w", "ln_post_b"]
    if spec.patch_bias:
        names.append("patch_b")
    if spec.pre_norm:
        names += ["ln_pre_w", "ln_pre_b"]
    for l in range(spec.layers):
        names += [f"l{l}.{f}" for f in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b",
                                        "fc1_w", "fc1_b", "fc2_w", "fc2_b")]
    if spec.tail == "gem_linear":
        names += ["head_w", "head_b"]
    elif spec.tail == "gem_conv_linear":
        names += ["gem_conv_w", "gem_conv_b", "head_w", "head_b"]
    return names


def random_weights(spec: VitSpec, seed: int = 0, device="cpu") -> Dict[str, torch.Tensor]:
    """Seeded random-init weights of the given architecture (there are no shipped checkpoints:
    D/checkpoints/.gitkeep).  Distributions follow the reference's init (clip.py:107-124: class/pos
    ~ W^-0.5 N(0,1), Linear ~ N(0, 0.02)); LayerNorm affine is perturbed around (1, 0)."""
    g = torch.Generator().manual_seed(seed)
    W, T = spec.width, spec.tokens

    def n(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    shapes = {"patch_w": ((W, 3, spec.patch, spec.patch), (3 * spec.patch ** 2) ** -0.5), "patch_b": ((W,), 0.02),
              "cls": ((W,), W ** -0.5), "pos": ((T, W), W ** -0.5), "head_b": ((spec.out_dim,), 0.02),
              "gem_conv_w": ((spec.gem_hidden, W), W ** -0.5), "gem_conv_b": ((spec.gem_hidden,), 0.2)}
    w = {}
    for name in param_names(spec):
        f = name.split(".")[-1]
        if name in shapes:
            w[name] = n(*shapes[name][0], std=shapes[name][1])
        elif name == "head_w":
            cin = spec.gem_hidden if spec.tail == "gem_conv_linear" else W
            w[name] = n(spec.out_dim, cin, std=cin ** -0.5)
        elif f.startswith("ln") and f.endswith("_w"):
            w[name] = 1.0 + n(W, std=0.1)
        elif f.startswith("ln") and f.endswith("_b"):
            w[name] = n(W, std=0.1)
        else:
            rows = {"qkv": 3 * W, "proj": W, "fc1": 4 * W, "fc2": W}[f[:-2]]
            cols = 4 * W if f == "fc2_w" else W
            w[name] = n(rows, cols) if f.endswith("_w") else n(rows)
    return {k: v.to(device) for k, v in w.items()}


class B200ViTEncoder(nn.Module):
    """Frame encoder backed by a vscb200_vit plan.  Weights are held as fp32 buffers (so ``.to()`` /
    ``.cuda()`` / ``state_dict()`` behave) and packed to bf16 inside the plan on first use per device."""

    def __init__(self, spec: VitSpec, weights: Dict[str, torch.Tensor], max_frames: int = 256):
        super().__init__()
        self.spec = spec
        self.max_frames = int(max_frames)
        missing = [n for n in param_names(spec) if n not in weights]
        if missing:
            raise KeyError(f"B200ViTEncoder: missing weights {missing[:6]}{'...' if len(missing) > 6 else ''}")
        for n in param_names(spec):
            self.register_buffer("w_" + n.replace(".", "_"), weights[n].detach().float().contiguous().clone())
        # DDP refuses modules without a grad-requiring parameter (extract_ref_feats.py:26)
        self.ddp_anchor = nn.Parameter(torch.zeros(1))
        self._plan = None
        self._plan_device = None

    # ---- plan management -----------------------------------------------------------------------
    def _drop_plan(self):
        if self._plan is not None:
            _lib.lib().vscb200_vit_destroy(self._plan)
            self._plan, self._plan_device = None, None

    def __del__(self):
        try:
            self._drop_plan()
        except Exception:
            pass

    def _apply(self, fn, *a, **k):
        self._drop_plan()          # buffers may move; re-pack lazily
        return super()._apply(fn, *a, **k)

    def _ensure_plan(self, device: torch.device):
        if self._plan is not None and self._plan_device == device:
            return
        self._drop_plan()
        lib = _lib.lib()
        spec_c = self.spec.to_c()
        plan = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.vscb200_vit_create(C.byref(spec_c), self.max_frames, C.byref(plan)), "vit_create")
            stream = torch.cuda.current_stream(device).cuda_stream
            for n in param_names(self.spec):
                buf = getattr(self, "w_" + n.replace(".", "_"))
                if buf.device != device:
                    buf = buf.to(device)
                _lib.check(lib.vscb200_vit_set_param(plan, n.encode(), C.c_void_p(buf.data_ptr()), buf.numel(),
                                                     C.c_void_p(stream)), f"vit_set_param({n})")
            torch.cuda.current_stream(device).synchronize()
        self._plan, self._plan_device = plan, device

    # ---- the call the reference makes ----------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("B200ViTEncoder: input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.spec.img or x.shape[3] != self.spec.img:
            raise RuntimeError(f"B200ViTEncoder: expected [n,3,{self.spec.img},{self.spec.img}], got {tuple(x.shape)}")
        x = x.contiguous().float()
        self._ensure_plan(x.device)
        n = x.shape[0]
        if self.spec.tail == "tokens":
            out = torch.empty((n, self.spec.tokens, self.spec.width), dtype=torch.float32, device=x.device)
        else:
            out = torch.empty((n, self.spec.out_dim), dtype=torch.float32, device=x.device)
        if n:
            with torch.cuda.device(x.device):
                stream = torch.cuda.current_stream(x.device).cuda_stream
                _lib.check(_lib.lib().vscb200_vit_forward(self._plan, C.c_void_p(x.data_ptr()), n,
                                                          C.c_void_p(out.data_ptr()), C.c_void_p(stream)),
                           "vit_forward")
        return out

    def forward_host(self, frames, device: Optional[torch.device] = None):
        """numpy/CPU-tensor frames in, numpy descriptors out, copies inside the C ABI
        (the extractor's H2D + ``.cpu().numpy()`` of D/infer/src/extractor.py:16-30 in one call)."""
        import numpy as np
        device = torch.device(device or "cuda:0")
        self._ensure_plan(device)
        x = np.ascontiguousarray(frames.numpy() if isinstance(frames, torch.Tensor) else frames, dtype=np.float32)
        n = x.shape[0]
        per = int(_lib.lib().vscb200_vit_out_elems_per_frame(self._plan))
        out = np.empty((n, per), dtype=np.float32)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().vscb200_vit_forward_host(self._plan, x.ctypes.data_as(C.c_void_p), n,
                                                           out.ctypes.data_as(C.c_void_p)), "vit_forward_host")
        return out.reshape(n, self.spec.tokens, self.spec.width) if self.spec.tail == "tokens" else out


# ------------------------------------------------------------------------------------------------
# state-dict converters (reference parameter names -> flat names)
# ------------------------------------------------------------------------------------------------
def weights_from_clip_state_dict(sd: Dict[str, torch.Tensor], layers: int) -> Dict[str, torch.Tensor]:
    """``CLIPModel.state_dict()`` names (D/train/train_vid_score/video/clip.py:85-124)."""
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
    return w


def weights_from_sscd_timm_state_dict(sd: Dict[str, torch.Tensor], layers: int) -> Dict[str, torch.Tensor]:
    """``SSCDModel`` with a timm ViT backbone + GeM head, as exported by
    D/train/train_v68/torch2scripts.py:14-28 (names ``model.backbone.*`` / ``model.embeddings.*``)."""
    b = "model.backbone."
    w = {"patch_w": sd[b + "patch_embed.proj.weight"], "patch_b": sd[b + "patch_embed.proj.bias"],
         "cls": sd[b + "cls_token"].reshape(-1), "pos": sd[b + "pos_embed"][0],
         "ln_post_w": sd[b + "norm.weight"], "ln_post_b": sd[b + "norm.bias"],
         "gem_conv_w": sd["model.embeddings.0.conv.weight"].squeeze(-1), "gem_conv_b": sd["model.embeddings.0.conv.bias"],
         "head_w": sd["model.embeddings.1.weight"], "head_b": sd["model.embeddings.1.bias"]}
    for l in range(layers):
        s, p = f"{b}blocks.{l}.", f"l{l}."
        w[p + "ln1_w"], w[p + "ln1_b"] = sd[s + "norm1.weight"], sd[s + "norm1.bias"]
        w[p + "qkv_w"], w[p + "qkv_b"] = sd[s + "attn.qkv.weight"], sd[s + "attn.qkv.bias"]
        w[p + "proj_w"], w[p + "proj_b"] = sd[s + "attn.proj.weight"], sd[s + "attn.proj.bias"]
        w[p + "ln2_w"], w[p + "ln2_b"] = sd[s + "norm2.weight"], sd[s + "norm2.bias"]
        w[p + "fc1_w"], w[p + "fc1_b"] = sd[s + "mlp.fc1.weight"], sd[s + "mlp.fc1.bias"]
        w[p + "fc2_w"], w[p + "fc2_b"] = sd[s + "mlp.fc2.weight"], sd[s + "mlp.fc2.bias"]
    return w


def weights_from_hf_vit_state_dict(sd: Dict[str, torch.Tensor], layers: int) -> Dict[str, torch.Tensor]:
    """Backbone ``VIT`` = transformers.ViTModel + gem + Linear (D/train/train_v106/.../backbones/vit.py:10-58): names
    ``vit.embeddings.* / vit.encoder.layer.N.* / vit.layernorm.* / output_proj.*``; q, k, v are separate Linears."""
    e = "vit.embeddings."
    w = {"patch_w": sd[e + "patch_embeddings.projection.weight"], "patch_b": sd[e + "patch_embeddings.projection.bias"],
         "cls": sd[e + "cls_token"].reshape(-1), "pos": sd[e + "position_embeddings"][0],
         "ln_post_w": sd["vit.layernorm.weight"], "ln_post_b": sd["vit.layernorm.bias"],
         "head_w": sd["output_proj.weight"], "head_b": sd["output_proj.bias"]}
    for l in range(layers):
        s, p = f"vit.encoder.layer.{l}.", f"l{l}."
        w[p + "ln1_w"], w[p + "ln1_b"] = sd[s + "layernorm_before.weight"], sd[s + "layernorm_before.bias"]
        w[p + "qkv_w"] = torch.cat([sd[s + f"attention.attention.{n}.weight"] for n in ("query", "key", "value")])
        w[p + "qkv_b"] = torch.cat([sd[s + f"attention.attention.{n}.bias"] for n in ("query", "key", "value")])
        w[p + "proj_w"], w[p + "proj_b"] = sd[s + "attention.output.dense.weight"], sd[s + "attention.output.dense.bias"]
        w[p + "ln2_w"], w[p + "ln2_b"] = sd[s + "layernorm_after.weight"], sd[s + "layernorm_after.bias"]
        w[p + "fc1_w"], w[p + "fc1_b"] = sd[s + "intermediate.dense.weight"], sd[s + "intermediate.dense.bias"]
        w[p + "fc2_w"], w[p + "fc2_b"] = sd[s + "output.dense.weight"], sd[s + "output.dense.bias"]
    return w


class _Tracked(dict):
    """A state dict that remembers which entries a converter read."""

    def __init__(self, sd):
        super().__init__(sd)
        self.read = set()

    def __getitem__(self, k):
        self.read.add(k)
        return super().__getitem__(k)


# entries of a reference checkpoint that carry no arithmetic of the inference path: registered index / mask / coordinate
# buffers of Swin-V2 (recomputed from the spec), and the pooler of transformers.ViTModel (VIT.forward reads
# last_hidden_state only, backbones/vit.py:44)
_IGNORABLE = re.compile(r"(\.relative_position_index|\.relative_coords_table|\.attn_mask|^vit\.pooler\.)")


def _precision_default() -> str:
    import os
    p = os.environ.get("VSCB200_ENCODER_PRECISION", "bf16")
    if p not in ("bf16", "fp32"):
        raise RuntimeError(f"VSCB200_ENCODER_PRECISION must be 'bf16' or 'fp32', got {p!r}")
    return p


def encoder_from_state_dict(sd: Dict[str, torch.Tensor], max_frames: int = 256, ln_eps: Optional[float] = None,
                            precision: Optional[str] = None) -> nn.Module:
    """Recognise a reference checkpoint by its parameter names and build the matching encoder.  EVERY entry of the state
    dict must be accounted for (read by the converter, or a known arithmetic-free buffer): a checkpoint that is only
    similar to a supported architecture (an extra ``proj``, a different head) raises instead of producing plausible but
    wrong descriptors.  ``ln_eps``: LayerNorm epsilon when it is known from elsewhere (a TorchScript graph); ``precision``:
    "bf16" | "fp32" (default: $VSCB200_ENCODER_PRECISION, else "bf16")."""
    precision = precision or _precision_default()
    sd = _Tracked(sd)
    keys = set(sd.keys())

    def done(enc):
        extra = sorted(k for k in keys - sd.read if not _IGNORABLE.search(k))
        if extra:
            raise RuntimeError(f"checkpoint has parameters this encoder does not use: {extra[:6]}{'...' if len(extra) > 6 else ''}")
        return enc

    if "conv1.weight" in keys and "class_embedding" in keys:                    # CLIPModel
        width, _, patch, _ = sd["conv1.weight"].shape
        tokens = sd["positional_embedding"].shape[0]
        layers = 1 + max(int(m.group(1)) for k in keys if (m := re.match(r"transformer\.resblocks\.(\d+)\.", k)))
        img = int(round((tokens - 1) ** 0.5)) * patch
        spec = VitSpec(img, patch, width, layers, width // 64, tail="tokens", ln_eps=ln_eps or 1e-5, precision=precision)
        return done(B200ViTEncoder(spec, weights_from_clip_state_dict(sd, layers), max_frames))
    if "model.backbone.patch_embed.proj.weight" in keys and "model.embeddings.0.conv.weight" in keys:   # vit_v68
        width, _, patch, _ = sd["model.backbone.patch_embed.proj.weight"].shape
        tokens = sd["model.backbone.pos_embed"].shape[1]
        layers = 1 + max(int(m.group(1)) for k in keys if (m := re.match(r"model\.backbone\.blocks\.(\d+)\.", k)))
        img = int(round((tokens - 1) ** 0.5)) * patch
        spec = VitSpec(img, patch, width, layers, width // 64, patch_bias=True, pre_norm=False, act="gelu",
                       ln_eps=ln_eps or 1e-6, tail="gem_conv_linear", out_dim=sd["model.embeddings.1.weight"].shape[0],
                       gem_hidden=sd["model.embeddings.0.conv.weight"].shape[0], precision=precision)
        return done(B200ViTEncoder(spec, weights_from_sscd_timm_state_dict(sd, layers), max_frames))
    if "vit.embeddings.patch_embeddings.projection.weight" in keys and "output_proj.weight" in keys:   # backbone VIT (HF)
        width, _, patch, _ = sd["vit.embeddings.patch_embeddings.projection.weight"].shape
        tokens = sd["vit.embeddings.position_embeddings"].shape[1]
        layers = 1 + max(int(m.group(1)) for k in keys if (m := re.match(r"vit\.encoder\.layer\.(\d+)\.", k)))
        img = int(round((tokens - 1) ** 0.5)) * patch
        spec = VitSpec(img, patch, width, layers, width // 64, patch_bias=True, pre_norm=False, act="gelu",
                       ln_eps=ln_eps or 1e-12, tail="gem_linear", out_dim=sd["output_proj.weight"].shape[0], precision=precision)
        return done(B200ViTEncoder(spec, weights_from_hf_vit_state_dict(sd, layers), max_frames))
    if "patch_embed.proj.weight" in keys and "layers.0.blocks.0.attn.logit_scale" in keys:             # swinv2_v1xx
        import dataclasses as _dc

        from .swin_encoder import B200SwinEncoder, param_names, spec_from_state_dict
        spec = _dc.replace(spec_from_state_dict(sd), precision=precision, ln_eps=ln_eps or 1e-5)
        for n in param_names(spec):
            sd[n]
        return done(B200SwinEncoder(spec, sd, max_frames))
    raise RuntimeError("encoder_from_state_dict: unrecognised checkpoint (implemented: CLIP ViT, timm ViT + GeM head, "
                       "HF ViT + gem head, Swin-V2 -- the encoders on the reference's inference path)")


def _jit_layer_norm_eps(module) -> Optional[float]:
    """LayerNorm epsilon of a TorchScript module: a state dict does not carry it, the graph does (the constant fed to
    the first aten::layer_norm node).  None when the graph cannot be read."""
    try:
        for node in module.inlined_graph.nodes():
            if node.kind() == "aten::layer_norm":
                v = list(node.inputs())[4].toIValue()
                if v is not None:
                    return float(v)
    except Exception:
        pass
    return None


_orig_jit_load = None


def install_jit_load_hook(max_frames: int = 256, precision: Optional[str] = None):
    """Make the reference's unmodified ``torch.jit.load(ckpt)`` calls return a B200 encoder when the checkpoint is a
    recognised frame encoder (extract_ref_feats.py:24, extract_query_feats.py:77-92, infer_matching.py:84-117).
    Anything else -- unrecognised modules, checkpoints with parameters the encoder would ignore -- is returned
    untouched, with a log line either way."""
    import logging
    global _orig_jit_load
    if _orig_jit_load is not None:
        return
    _orig_jit_load = torch.jit.load
    log = logging.getLogger("vscb200")

    def _load(f, *a, **k):
        module = _orig_jit_load(f, *a, **k)
        try:
            enc = encoder_from_state_dict(dict(module.state_dict()), max_frames, ln_eps=_jit_layer_norm_eps(module),
                                          precision=precision)
        except (RuntimeError, KeyError) as e:
            log.warning("torch.jit.load(%s): kept the TorchScript module (%s)", f, e)
            return module
        log.info("torch.jit.load(%s): replaced by %s %s", f, type(enc).__name__, enc.spec)
        return enc

    torch.jit.load = _load


def uninstall_jit_load_hook():
    global _orig_jit_load
    if _orig_jit_load is not None:
        torch.jit.load = _orig_jit_load
        _orig_jit_load = None
