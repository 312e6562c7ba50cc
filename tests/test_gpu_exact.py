"""Building blocks of the encoders' fp32-equivalent mode, through the C ABI: the split-bf16 tcgen05 GEMM
(vscb200_gemm_split), the fp32 attention kernel (vscb200_attention_fp32), and the encoders in ``precision="fp32"`` on
small configurations against the fp32 oracle / the reference-class goldens at the north-star tolerance (1e-3)."""
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


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _split(torch, x):
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return hi.contiguous(), lo.contiguous()


def _rel(out, ref):
    out, ref = out.reshape(len(out), -1), ref.reshape(len(ref), -1)
    return np.linalg.norm(out - ref, axis=1) / np.linalg.norm(ref, axis=1)


def test_gemm_split_is_fp32_equivalent(torch):
    from vsc22_submission_b200 import _lib
    torch.manual_seed(0)
    for (M, N, K, epi, act) in [(128, 128, 64, 1, -1), (1000, 768, 768, 1, -1), (1576, 2304, 768, 0, -1), (333, 3072, 768, 0, 1),
                                (4096, 768, 3072, 2, -1), (85, 64, 128, 0, 0), (5000, 384, 96, 1, -1)]:
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda") * 0.05
        b = torch.randn(N, device="cuda")
        Ah, Al = _split(torch, A)
        Wh, Wl = _split(torch, W)
        res = torch.randn(M, N, device="cuda")
        if epi == 0:
            out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
            out_lo = torch.empty_like(out)
        else:
            out, out_lo = (res.clone() if epi == 2 else torch.empty((M, N), device="cuda")), None
        _lib.check(_lib.lib().vscb200_gemm_split(_p(Ah), _p(Al), _p(Wh), _p(Wl), _p(b), _p(out), _p(out_lo), M, N, K, K, K, N,
                                                 epi, act, None), "gemm_split")
        torch.cuda.synchronize()
        ref = A.double() @ W.double().T + b.double()
        if act == 0:
            ref = ref * torch.sigmoid(1.702 * ref)
        elif act == 1:
            ref = torch.nn.functional.gelu(ref)
        if epi == 2:
            ref = ref + res.double()
        got = out.float().double() + (out_lo.float().double() if out_lo is not None else 0.0)
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        assert err < 3e-5, (M, N, K, epi, act, err)


def test_attention_fp32_against_torch(torch):
    from vsc22_submission_b200 import _lib
    torch.manual_seed(3)
    for (n, T, H, hd, planes) in [(2, 17, 2, 64, 2), (3, 197, 12, 64, 2), (1, 577, 2, 64, 2), (2, 145, 3, 64, 1),
                                  (4, 64, 4, 32, 2), (2, 576, 2, 32, 1)]:
        W = H * hd
        qkv = torch.randn(n * T, 3 * W, device="cuda")
        qh, ql = _split(torch, qkv)
        out_h = torch.empty((n * T, W), dtype=torch.bfloat16, device="cuda")
        out_l = torch.empty_like(out_h) if planes == 2 else None
        _lib.check(_lib.lib().vscb200_attention_fp32(_p(qh), _p(ql) if planes == 2 else None, _p(out_h), _p(out_l), n, T, H, hd,
                                                     None), "attention_fp32")
        torch.cuda.synchronize()
        src = (qh.float() + ql.float()) if planes == 2 else qh.float()
        q, k, v = (src.double().reshape(n, T, 3, H, hd)[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        ref = (torch.softmax(q @ k.transpose(-1, -2) / hd ** 0.5, -1) @ v).permute(0, 2, 1, 3).reshape(n * T, W)
        got = out_h.float().double() + (out_l.float().double() if planes == 2 else 0.0)
        tol = 2e-5 if planes == 2 else 6e-3            # single plane: the output itself is rounded to bf16
        assert (got - ref).abs().max().item() < tol * ref.abs().max().item(), (n, T, H, hd, planes)


def test_vit_fp32_mode_matches_reference_class_golden(torch, golden_dir):
    """tokens / descriptors of the reference's own CLIPModel (+ gem/Linear tail) at the north-star tolerance."""
    from vsc22_submission_b200.encoder import B200ViTEncoder, VitSpec
    g = np.load(os.path.join(golden_dir, "vit_clip_small.npz"))
    img, patch, width, layers, heads, out_dim = (int(v) for v in g["spec"])
    w = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")}
    frames = torch.from_numpy(g["frames"]).cuda()
    enc = B200ViTEncoder(VitSpec(img, patch, width, layers, heads, tail="gem_linear", out_dim=out_dim, precision="fp32"), w,
                         max_frames=3).cuda().eval()
    rel = _rel(enc(frames).cpu().numpy(), g["desc"])
    enc_t = B200ViTEncoder(VitSpec(img, patch, width, layers, heads, tail="tokens", precision="fp32"),
                           {k: v for k, v in w.items() if not k.startswith("head")}, max_frames=8).cuda().eval()
    rel_t = _rel(enc_t(frames).cpu().numpy(), g["tokens"])
    print("ViT fp32-equivalent mode vs reference class: desc", rel.max(), "tokens", rel_t.max())
    assert rel.max() < 1e-4 and rel_t.max() < 1e-4              # measured 1.5e-6 / 3.6e-6


def test_swin_fp32_mode_matches_reference_class_golden(torch, golden_dir):
    from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SwinSpec, random_weights
    g = np.load(os.path.join(golden_dir, "swin_small.npz"))
    spec = SwinSpec(img=128, patch=4, embed=64, depths=(2, 2, 2, 2), heads=(2, 4, 8, 16), window=8,
                    pretrained_windows=(6, 6, 6, 3), out_dim=64, precision="fp32")
    enc = B200SwinEncoder(spec, random_weights(spec, seed=0), max_frames=2).cuda().eval()
    rel = _rel(enc(torch.from_numpy(g["frames"]).cuda()).cpu().numpy(), g["desc"])
    print("Swin-V2 fp32-equivalent mode vs reference class:", rel)
    assert rel.max() < 1e-4                                        # measured 8e-6


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_swin_windows_that_are_not_powers_of_two(torch, precision):
    """6 x 6 windows (the fp32 attention kernel in both modes), 12 x 12 and 24 x 24 windows (bf16 mode: the K-blocked tcgen05
    kernel, attention_kb.cu -- the windows of SwinV2-L@384, BASELINE configs[3]); shifted blocks included."""
    import dataclasses
    from oracle import swin_ref
    from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SwinSpec, random_weights
    for kw in (dict(img=96, patch=4, embed=64, depths=(2, 2), heads=(2, 4), window=6, pretrained_windows=(4, 4), out_dim=48),
               dict(img=96, patch=4, embed=64, depths=(2, 1), heads=(2, 4), window=12, pretrained_windows=(6, 6), out_dim=48),
               dict(img=192, patch=4, embed=64, depths=(2, 2), heads=(2, 4), window=24, pretrained_windows=(12, 12), out_dim=48)):
        spec = SwinSpec(**kw, precision=precision)
        w = random_weights(spec, seed=2)
        x = torch.randn(3, 3, kw["img"], kw["img"], generator=torch.Generator().manual_seed(5)).clamp(-1, 1)
        ref = swin_ref.forward(swin_ref.SwinSpec(**kw), w, x, precision=precision).numpy()
        out = B200SwinEncoder(spec, w, max_frames=2).cuda().eval()(x.cuda()).cpu().numpy()
        rel = _rel(out, ref)
        print("swin window", kw["window"], precision, "rel L2 vs the oracle of the same precision:", rel)
        assert rel.max() < (3e-3 if precision == "bf16" else 1e-4)     # measured 1.2e-3 / 3e-6
