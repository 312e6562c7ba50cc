"""Seam (A), second encoder family: the reference's Swin-V2 frame encoder (``swinv2_v1xx`` TorchScript
checkpoints of D/infer/infer_ref.sh / extract_query_feats.py:81-92) behind the same ``nn.Module`` contract as
``encoder.B200ViTEncoder``: ``forward(x: float32 CUDA [n,3,H,W]) -> float32 CUDA [n,out_dim]`` on PyTorch's
current stream, ``.to()/.cuda()/.eval()/DDP`` friendly, no PyTorch op on the compute path, no CPU fallback.

Architecture: D/train/train_v106/vsc/baseline/model_factory/backbones/swinv2.py:502-633 configured by
config_v106.py:8-24.  Parameter names are the reference's state-dict names, so a reference checkpoint loads
without a converter; the arithmetic runs in libvscb200.so (csrc/swin.cu).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
import re
import zlib
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib


@dataclasses.dataclass
class SwinSpec:
    """Fields mirror vscb200_swin_spec (include/vscb200.h)."""
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
    precision: str = "bf16"        # "bf16" | "fp32" (fp32-equivalent arithmetic), see encoder.VitSpec.precision

    def stage(self, i: int):
        res = self.img // self.patch // (2 ** i)
        return self.embed * 2 ** i, res, min(self.window, res)

    def flops_per_frame(self) -> float:
        fl = 2.0 * (self.img // self.patch) ** 2 * 3 * self.patch ** 2 * self.embed
        for i, depth in enumerate(self.depths):
            Cw, res, ws = self.stage(i)
            L, N = res * res, ws * ws
            fl += depth * (2.0 * L * Cw * 3 * Cw + 2.0 * L * Cw * Cw + 4.0 * L * N * Cw + 16.0 * L * Cw * Cw)
            if i + 1 < len(self.depths):
                fl += 2.0 * (L // 4) * 4 * Cw * 2 * Cw
        return fl + 2.0 * self.embed * 2 ** (len(self.depths) - 1) * self.out_dim

    def to_c(self) -> "_lib.SwinSpecC":
        pad = lambda t: (C.c_int * 4)(*(list(t) + [0] * (4 - len(t))))
        return _lib.SwinSpecC(img=self.img, patch=self.patch, embed=self.embed, n_stages=len(self.depths),
                              depths=pad(self.depths), heads=pad(self.heads), window=self.window,
                              pretrained_windows=pad(self.pretrained_windows), out_dim=self.out_dim,
                              ln_eps=self.ln_eps, gem_p=self.gem_p, precision=_lib.PRECISION[self.precision])


SWINV2_B_256 = SwinSpec()          # config_v106.py: swinv2_v106 / v107 / v115
# BASELINE configs[3] "Swin-L ... 384^2": SwinV2-L, 24 x 24 windows (576 tokens), pre-trained at window 12
SWINV2_L_384 = SwinSpec(img=384, patch=4, embed=192, depths=(2, 2, 18, 2), heads=(6, 12, 24, 48), window=24,
                        pretrained_windows=(12, 12, 12, 6), out_dim=512)


def param_names(spec: SwinSpec) -> List[str]:
    names = ["patch_embed.proj.weight", "patch_embed.proj.bias", "patch_embed.norm.weight", "patch_embed.norm.bias"]
    per_block = ("norm1.weight", "norm1.bias", "attn.logit_scale", "attn.cpb_mlp.0.weight", "attn.cpb_mlp.0.bias",
                 "attn.cpb_mlp.2.weight", "attn.qkv.weight", "attn.q_bias", "attn.v_bias", "attn.proj.weight",
                 "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight",
                 "mlp.fc2.bias")
    for i, depth in enumerate(spec.depths):
        for j in range(depth):
            names += [f"layers.{i}.blocks.{j}.{f}" for f in per_block]
        if i + 1 < len(spec.depths):
            names += [f"layers.{i}.downsample.{f}" for f in ("reduction.weight", "norm.weight", "norm.bias")]
    return names + ["norm.weight", "norm.bias", "output_proj.weight", "output_proj.bias"]


def param_shape(spec: SwinSpec, name: str) -> tuple:
    m = re.match(r"layers\.(\d+)\.", name)
    i = int(m.group(1)) if m else 0
    Cw = spec.embed * 2 ** i
    nH = spec.heads[i] if m else 0
    Cf = spec.embed * 2 ** (len(spec.depths) - 1)
    tail = name.split(".blocks.")[-1].split(".", 1)[-1] if ".blocks." in name else name
    table = {"patch_embed.proj.weight": (spec.embed, 3, spec.patch, spec.patch), "patch_embed.proj.bias": (spec.embed,),
             "patch_embed.norm.weight": (spec.embed,), "patch_embed.norm.bias": (spec.embed,),
             "norm.weight": (Cf,), "norm.bias": (Cf,), "output_proj.weight": (spec.out_dim, Cf),
             "output_proj.bias": (spec.out_dim,)}
    if name in table:
        return table[name]
    if ".downsample." in name:
        return (2 * Cw, 4 * Cw) if name.endswith("reduction.weight") else (2 * Cw,)
    return {"norm1.weight": (Cw,), "norm1.bias": (Cw,), "norm2.weight": (Cw,), "norm2.bias": (Cw,),
            "attn.logit_scale": (nH, 1, 1), "attn.cpb_mlp.0.weight": (512, 2), "attn.cpb_mlp.0.bias": (512,),
            "attn.cpb_mlp.2.weight": (nH, 512), "attn.qkv.weight": (3 * Cw, Cw), "attn.q_bias": (Cw,),
            "attn.v_bias": (Cw,), "attn.proj.weight": (Cw, Cw), "attn.proj.bias": (Cw,),
            "mlp.fc1.weight": (4 * Cw, Cw), "mlp.fc1.bias": (4 * Cw,), "mlp.fc2.weight": (Cw, 4 * Cw),
            "mlp.fc2.bias": (Cw,)}[tail]


def random_weights(spec: SwinSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights keyed by parameter name (there are no shipped checkpoints).  The reference
    zero-initialises its res-post-norms (swinv2.py:452-457), which would make every block the identity, so the
    LayerNorm affines are drawn around (1, 0) and the post-norm gains around 0.5."""
    w = {}
    for name in param_names(spec):
        shape = param_shape(spec, name)
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


class B200SwinEncoder(nn.Module):
    """Frame encoder backed by a vscb200_swin plan (weights as fp32 buffers; packed to bf16 per device)."""

    def __init__(self, spec: SwinSpec, weights: Dict[str, torch.Tensor], max_frames: int = 64):
        super().__init__()
        self.spec = spec
        self.max_frames = int(max_frames)
        self._names = param_names(spec)
        missing = [n for n in self._names if n not in weights]
        if missing:
            raise KeyError(f"B200SwinEncoder: missing weights {missing[:6]}{'...' if len(missing) > 6 else ''}")
        for n in self._names:
            self.register_buffer("w_" + n.replace(".", "_"), weights[n].detach().float().contiguous().clone())
        self.ddp_anchor = nn.Parameter(torch.zeros(1))     # DDP needs a grad-requiring parameter (extract_ref_feats.py:26)
        self._plan = None
        self._plan_device = None

    def _drop_plan(self):
        if self._plan is not None:
            _lib.lib().vscb200_swin_destroy(self._plan)
            self._plan, self._plan_device = None, None

    def __del__(self):
        try:
            self._drop_plan()
        except Exception:
            pass

    def _apply(self, fn, *a, **k):
        self._drop_plan()
        return super()._apply(fn, *a, **k)

    def _ensure_plan(self, device: torch.device):
        if self._plan is not None and self._plan_device == device:
            return
        self._drop_plan()
        lib = _lib.lib()
        spec_c = self.spec.to_c()
        plan = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.vscb200_swin_create(C.byref(spec_c), self.max_frames, C.byref(plan)), "swin_create")
            stream = torch.cuda.current_stream(device).cuda_stream
            for n in self._names:
                buf = getattr(self, "w_" + n.replace(".", "_"))
                if buf.device != device:
                    buf = buf.to(device)
                _lib.check(lib.vscb200_swin_set_param(plan, n.encode(), C.c_void_p(buf.data_ptr()), buf.numel(),
                                                      C.c_void_p(stream)), f"swin_set_param({n})")
            torch.cuda.current_stream(device).synchronize()
        self._plan, self._plan_device = plan, device

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("B200SwinEncoder: input must be a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.spec.img or x.shape[3] != self.spec.img:
            raise RuntimeError(f"B200SwinEncoder: expected [n,3,{self.spec.img},{self.spec.img}], got {tuple(x.shape)}")
        x = x.contiguous().float()
        self._ensure_plan(x.device)
        n = x.shape[0]
        out = torch.empty((n, self.spec.out_dim), dtype=torch.float32, device=x.device)
        if n:
            with torch.cuda.device(x.device):
                stream = torch.cuda.current_stream(x.device).cuda_stream
                _lib.check(_lib.lib().vscb200_swin_forward(self._plan, C.c_void_p(x.data_ptr()), n,
                                                           C.c_void_p(out.data_ptr()), C.c_void_p(stream)), "swin_forward")
        return out

    def forward_host(self, frames, device: Optional[torch.device] = None):
        """numpy / CPU-tensor frames in, numpy descriptors out; copies inside the C ABI."""
        import numpy as np
        device = torch.device(device or "cuda:0")
        self._ensure_plan(device)
        x = np.ascontiguousarray(frames.numpy() if isinstance(frames, torch.Tensor) else frames, dtype=np.float32)
        out = np.empty((x.shape[0], self.spec.out_dim), dtype=np.float32)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().vscb200_swin_forward_host(self._plan, x.ctypes.data_as(C.c_void_p), x.shape[0],
                                                            out.ctypes.data_as(C.c_void_p)), "swin_forward_host")
        return out


def spec_from_state_dict(sd: Dict[str, torch.Tensor]) -> SwinSpec:
    """Recover the architecture from a reference checkpoint: parameter shapes plus the registered buffers
    ``relative_position_index`` (window), ``attn_mask`` (token-map side) and ``relative_coords_table``
    (pre-training window, swinv2.py:107-112)."""
    embed, _, patch, _ = sd["patch_embed.proj.weight"].shape
    n_stages = 1 + max(int(m.group(1)) for k in sd if (m := re.match(r"layers\.(\d+)\.", k)))
    depths, heads, pws, wss = [], [], [], []
    for i in range(n_stages):
        depths.append(1 + max(int(m.group(1)) for k in sd if (m := re.match(rf"layers\.{i}\.blocks\.(\d+)\.", k))))
        heads.append(int(sd[f"layers.{i}.blocks.0.attn.logit_scale"].shape[0]))
        ws = int(round(math.sqrt(sd[f"layers.{i}.blocks.0.attn.relative_position_index"].shape[0])))
        wss.append(ws)
        tmax = float(sd[f"layers.{i}.blocks.0.attn.relative_coords_table"].max())      # log2(8 (ws-1)/(pws-1) + 1) / 3
        pws.append(int(round(8.0 * (ws - 1) / (2.0 ** (3.0 * tmax) - 1.0))) + 1)
    window = wss[0]
    mask = sd.get("layers.0.blocks.1.attn_mask")
    res0 = window * int(round(math.sqrt(mask.shape[0]))) if mask is not None else window
    return SwinSpec(img=res0 * patch, patch=patch, embed=embed, depths=tuple(depths), heads=tuple(heads), window=window,
                    pretrained_windows=tuple(pws), out_dim=int(sd["output_proj.weight"].shape[0]))
