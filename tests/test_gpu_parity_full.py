"""The encoder parity contract at the FULL configurations of the reference's inference path:

  descriptors  ||y - y_oracle|| / ||y_oracle||  <=  1e-3  per frame, on >= 32 frames,

against the plain fp32 oracle -- which is pinned to the reference's own classes -- in the fp32-equivalent mode
(``precision="fp32"``: split-bf16 tcgen05 GEMMs + fp32 attention; measured 3e-6 .. 2e-5), for ALL four configurations.

The bf16 tensor-core mode is compared with the matched-precision oracle (``precision="bf16"``: every matrix-product
operand rounded to bf16 where the CUDA encoder rounds it, everything else fp32 -- oracle/vit_ref.py, oracle/swin_ref.py)
and with the fp32 oracle.  It meets 1e-3 against BOTH on ViT-B/16@224 (BASELINE configs[1]; 3.5e-4 / 8.9e-4) and against
the matched oracle on vit_v68 (7.5e-4).  On the two deep configurations (24 blocks: CLIP ViT-L/14, SwinV2-B) the bf16
mode sits 4.0e-3 / 2.1e-3 from the matched oracle and 4.9e-3 / 5.8e-3 from fp32: rounding to bf16 is discontinuous, so
the 1e-4-level difference the residual stream accumulates between two fp32 implementations (summation order, ex2/tanh
approximations) flips a few percent of the bf16 roundings per layer, each flip a full bf16 ulp -- two bf16-operand
implementations decorrelate with depth no matter how faithfully the rounding POINTS are matched.  Those two are therefore
asserted at their measured level (and below the fp32 distance); the 1e-3 contract for them is the fp32-equivalent mode.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-3          # north-star tolerance (BASELINE.json), relative L2 per frame
N_FRAMES = 32
# bf16 mode vs the matched-precision oracle: asserted bound per configuration (measured: 3.5e-4, 7.5e-4, 4.0e-3, 2.1e-3)
BF16_MATCHED_TOL = {"vit_b16_224": TOL, "vit_v68": TOL, "clip_l14_224": 6e-3, "swinv2_b_256": 3.5e-3}


@pytest.fixture(scope="module")
def torch():
    import torch as t
    assert t.cuda.is_available(), "GPU tests need a CUDA device"
    return t


def _rel(out, ref):
    out, ref = out.reshape(len(out), -1), ref.reshape(len(ref), -1)
    return np.linalg.norm(out - ref, axis=1) / np.linalg.norm(ref, axis=1)


def _frames(torch, img, seed):
    return torch.randn(N_FRAMES, 3, img, img, generator=torch.Generator().manual_seed(seed)).clamp(-1, 1)


def _vit_case(torch, name):
    from oracle import vit_ref
    from vsc22_submission_b200 import encoder
    ospec, spec = {"vit_b16_224": (vit_ref.CLIP_B16_224, encoder.VIT_B16_224_GEM),
                   "vit_v68": (vit_ref.TIMM_B32_384, encoder.VIT_V68),
                   "clip_l14_224": (vit_ref.CLIP_L14_224, encoder.CLIP_L14_224)}[name]
    w = vit_ref.init_weights(ospec, seed=0)
    x = _frames(torch, ospec.img, 1)
    sel = (lambda y: y[:, 0]) if ospec.tail == "tokens" else (lambda y: y)      # the caller keeps [:, 0] (extract_query_feats.py:149)
    refs = {p: sel(vit_ref.forward(ospec, w, x, precision=p)).numpy() for p in ("fp32", "bf16")}
    return spec, w, x, sel, refs


@pytest.mark.parametrize("name", ["vit_b16_224", "vit_v68", "clip_l14_224"])
def test_vit_full_config_bf16_mode_vs_matched_oracle(torch, name):
    import dataclasses
    from vsc22_submission_b200.encoder import B200ViTEncoder
    spec, w, x, sel, refs = _vit_case(torch, name)
    enc = B200ViTEncoder(dataclasses.replace(spec, precision="bf16"), w, max_frames=16).cuda().eval()
    out = sel(enc(x.cuda())).cpu().numpy()
    rel_m, rel_f = _rel(out, refs["bf16"]), _rel(out, refs["fp32"])
    print(f"{name} bf16 mode: rel-L2 vs matched-precision oracle max {rel_m.max():.3e} mean {rel_m.mean():.3e}; "
          f"vs fp32 oracle max {rel_f.max():.3e} mean {rel_f.mean():.3e} ({N_FRAMES} frames)")
    assert rel_m.max() <= BF16_MATCHED_TOL[name], rel_m
    assert rel_m.mean() < rel_f.mean()      # matching the rounding points brings the oracle closer
    if name == "vit_b16_224":           # BASELINE configs[1]: the bf16 mode itself is inside the fp32 contract
        assert rel_f.max() <= TOL, rel_f


@pytest.mark.parametrize("name", ["vit_b16_224", "vit_v68", "clip_l14_224"])
def test_vit_full_config_fp32_mode_vs_fp32_oracle(torch, name):
    import dataclasses
    from vsc22_submission_b200.encoder import B200ViTEncoder
    spec, w, x, sel, refs = _vit_case(torch, name)
    enc = B200ViTEncoder(dataclasses.replace(spec, precision="fp32"), w, max_frames=16).cuda().eval()
    out = sel(enc(x.cuda())).cpu().numpy()
    rel = _rel(out, refs["fp32"])
    print(f"{name} fp32-equivalent mode: rel-L2 vs fp32 oracle max {rel.max():.3e} mean {rel.mean():.3e} ({N_FRAMES} frames)")
    assert rel.max() <= TOL / 10, rel       # measured 3.5e-6 .. 1.7e-5: two orders of magnitude inside the contract


def _swin_case(torch):
    from oracle import swin_ref
    from vsc22_submission_b200.swin_encoder import SWINV2_B_256, random_weights
    w = random_weights(SWINV2_B_256, seed=1)
    x = _frames(torch, 256, 3)
    refs = {p: swin_ref.forward(swin_ref.SWINV2_B_256, w, x, precision=p).numpy() for p in ("fp32", "bf16")}
    return SWINV2_B_256, w, x, refs


def test_swinv2_b_256_bf16_mode_vs_matched_oracle(torch):
    import dataclasses
    from vsc22_submission_b200.swin_encoder import B200SwinEncoder
    spec, w, x, refs = _swin_case(torch)
    enc = B200SwinEncoder(dataclasses.replace(spec, precision="bf16"), w, max_frames=16).cuda().eval()
    out = enc(x.cuda()).cpu().numpy()
    rel_m, rel_f = _rel(out, refs["bf16"]), _rel(out, refs["fp32"])
    print(f"SwinV2-B@256 bf16 mode: rel-L2 vs matched-precision oracle max {rel_m.max():.3e} mean {rel_m.mean():.3e}; "
          f"vs fp32 oracle max {rel_f.max():.3e} (the bf16-operand floor, tests/test_oracle_swin_bf16_floor.py)")
    assert rel_m.max() <= BF16_MATCHED_TOL["swinv2_b_256"], rel_m
    assert rel_m.mean() < rel_f.mean()


def test_swinv2_b_256_fp32_mode_vs_fp32_oracle(torch):
    """The contract against the reference class itself (swinv2.py:509-665; the fp32 oracle reproduces it bit for bit)."""
    import dataclasses
    from vsc22_submission_b200.swin_encoder import B200SwinEncoder
    spec, w, x, refs = _swin_case(torch)
    enc = B200SwinEncoder(dataclasses.replace(spec, precision="fp32"), w, max_frames=16).cuda().eval()
    out = enc(x.cuda()).cpu().numpy()
    rel = _rel(out, refs["fp32"])
    print(f"SwinV2-B@256 fp32-equivalent mode: rel-L2 vs fp32 oracle max {rel.max():.3e} mean {rel.mean():.3e}")
    assert rel.max() <= TOL / 10, rel       # measured 1.0e-5


# ---------------------------------------------------------------------------------------------- BASELINE configs[3]
# SwinV2-L / window 24 @ 384 (swinv2.py:72-185, 509-665 at the large hyper-parameters) and ViT-L/16 @ 384 (clip.py:85-163,
# T = 577): the K-blocked / streaming tcgen05 attention kernels at the sizes they were written for.  4 frames each (the
# fp32 oracle of one 384 x 384 frame is ~0.3 TFLOP on the host).
C4_FRAMES = 4


def _c4_vit(torch):
    from oracle import vit_ref
    w = vit_ref.init_weights(vit_ref.CLIP_L16_384, seed=0)
    x = torch.randn(C4_FRAMES, 3, 384, 384, generator=torch.Generator().manual_seed(21)).clamp(-1, 1)
    refs = {p: vit_ref.forward(vit_ref.CLIP_L16_384, w, x, precision=p)[:, 0].numpy() for p in ("fp32", "bf16")}
    return w, x, refs


def test_config4_vit_l16_384(torch):
    import dataclasses
    from vsc22_submission_b200.encoder import B200ViTEncoder, VIT_L16_384
    w, x, refs = _c4_vit(torch)
    enc = B200ViTEncoder(dataclasses.replace(VIT_L16_384, precision="bf16"), w, max_frames=4).cuda().eval()
    out = enc(x.cuda())[:, 0].cpu().numpy()
    rel_m, rel_f = _rel(out, refs["bf16"]), _rel(out, refs["fp32"])
    print(f"ViT-L/16@384 bf16 mode: rel-L2 vs matched-precision oracle max {rel_m.max():.3e}; vs fp32 oracle max {rel_f.max():.3e}")
    assert rel_m.max() <= 6e-3, rel_m          # 24 blocks: same regime as CLIP ViT-L/14 above
    del enc
    enc = B200ViTEncoder(dataclasses.replace(VIT_L16_384, precision="fp32"), w, max_frames=4).cuda().eval()
    rel = _rel(enc(x.cuda())[:, 0].cpu().numpy(), refs["fp32"])
    print(f"ViT-L/16@384 fp32-equivalent mode: rel-L2 vs fp32 oracle max {rel.max():.3e}")
    assert rel.max() <= TOL / 10, rel


def test_config4_swinv2_l_384(torch):
    import dataclasses
    from oracle import swin_ref
    from vsc22_submission_b200.swin_encoder import B200SwinEncoder, SWINV2_L_384, random_weights
    w = random_weights(SWINV2_L_384, seed=1)
    x = torch.randn(C4_FRAMES, 3, 384, 384, generator=torch.Generator().manual_seed(22)).clamp(-1, 1)
    refs = {p: swin_ref.forward(swin_ref.SWINV2_L_384, w, x, precision=p).numpy() for p in ("fp32", "bf16")}
    enc = B200SwinEncoder(dataclasses.replace(SWINV2_L_384, precision="bf16"), w, max_frames=4).cuda().eval()
    out = enc(x.cuda()).cpu().numpy()
    rel_m, rel_f = _rel(out, refs["bf16"]), _rel(out, refs["fp32"])
    print(f"SwinV2-L/w24@384 bf16 mode: rel-L2 vs matched-precision oracle max {rel_m.max():.3e}; vs fp32 oracle max {rel_f.max():.3e}")
    assert rel_m.max() <= 3.5e-3, rel_m        # same bound as SwinV2-B above
    del enc
    enc = B200SwinEncoder(dataclasses.replace(SWINV2_L_384, precision="fp32"), w, max_frames=4).cuda().eval()
    rel = _rel(enc(x.cuda()).cpu().numpy(), refs["fp32"])
    print(f"SwinV2-L/w24@384 fp32-equivalent mode: rel-L2 vs fp32 oracle max {rel.max():.3e}")
    assert rel.max() <= TOL / 10, rel
