#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE from /root/reference.

Run in the authoring container only (``python tests/golden/make_golden.py``); the fixtures are
committed because /root/reference does not travel to the GPU box.  Nothing here is copied from
the reference: its modules are imported in place through oracle/refload.py.

Fixtures
--------
vit_clip_small.npz   reference ``CLIPModel(64,16,128,2,2,.)`` (clip.py:85) + the VIT-backbone
                     gem/Linear tail (backbones/vit.py:42-58) on seeded weights: state dict (flat
                     names of oracle/vit_ref.py), input frames, output tokens, output descriptor.
search_small.npz     the reference's unmodified ``vsc`` package (CandidateGeneration / VideoIndex /
                     exhaustive_search / score_normalize) run over ``oracle.faiss_np`` on seeded
                     descriptors: kNN (D, I), range_search CSR, score-normalised descriptors,
                     and the final candidate list (query_id, ref_id, score).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import refload, vit_ref  # noqa: E402


def make_vit():
    clip = refload.clip_module()
    torch.manual_seed(0)
    spec = vit_ref.VitSpec(img=64, patch=16, width=128, layers=2, heads=2, tail="gem_linear", out_dim=64)
    model = clip.CLIPModel(spec.img, spec.patch, spec.width, spec.layers, spec.heads, 512)
    model.init_weights()
    # make LayerNorm affine / biases non-trivial (init_weights sets them to 1/0)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("bias") or "ln_" in name:
                p.add_(torch.randn(p.shape, generator=g) * 0.05)
    model.eval()
    head = torch.nn.Linear(spec.width, spec.out_dim)
    frames = torch.randn(5, 3, spec.img, spec.img, generator=g).clamp(-1, 1)
    with torch.no_grad():
        tokens = model(frames)
        desc = head(tokens.clamp(min=1e-6).pow(3.0).mean(dim=1).pow(1.0 / 3.0))   # vit.py:56-58,50
    w = vit_ref.from_clip_state_dict(model.state_dict(), spec.layers)
    w["head_w"], w["head_b"] = head.weight.detach().clone(), head.bias.detach().clone()
    out = {"w." + k: v.numpy() for k, v in w.items()}
    out.update(frames=frames.numpy(), tokens=tokens.numpy(), desc=desc.numpy(),
               spec=np.array([spec.img, spec.patch, spec.width, spec.layers, spec.heads, spec.out_dim]))
    np.savez_compressed(os.path.join(HERE, "vit_clip_small.npz"), **out)
    print("vit_clip_small.npz", tokens.shape, desc.shape)


def make_search():
    refload.vsc_package("D_infer")
    from vsc.baseline.score_normalization import score_normalize
    from vsc.candidates import CandidateGeneration, MaxScoreAggregation
    from vsc.index import VideoFeature
    import faiss  # -> oracle.faiss_np via the shim

    rng = np.random.default_rng(7)
    d = 64

    def videos(prefix, n, lo, hi):
        out = []
        for i in range(n):
            m = int(rng.integers(lo, hi))
            f = rng.standard_normal((m, d)).astype(np.float32)
            out.append(VideoFeature(video_id=f"{prefix}{i:06d}", timestamps=np.arange(m, dtype=np.float32),
                                    feature=f))
        return out

    queries, refs, noise = videos("Q", 6, 3, 9), videos("R", 12, 4, 12), videos("N", 10, 4, 12)
    # plant copies so that the candidate list has structure
    refs[3].feature[:3] = queries[1].feature[:3] + 0.05 * rng.standard_normal((3, d)).astype(np.float32)
    refs[7].feature[2:5] = queries[4].feature[1:4] + 0.05 * rng.standard_normal((3, d)).astype(np.float32)

    q_raw = np.concatenate([v.feature for v in queries])
    r_raw = np.concatenate([v.feature for v in refs])
    z_raw = np.concatenate([v.feature for v in noise])

    sn_q, sn_r = score_normalize(queries, refs, noise, beta=1.2)        # sscd_baseline.py:195-200
    sn_q_arr = np.concatenate([v.feature for v in sn_q])
    sn_r_arr = np.concatenate([v.feature for v in sn_r])

    index = faiss.index_factory(sn_r_arr.shape[1], "Flat", faiss.METRIC_INNER_PRODUCT)
    index.add(sn_r_arr)
    D10, I10 = index.search(sn_q_arr, 10)
    lims, Dr, Ir = index.range_search(sn_q_arr, 0.0)

    cg = CandidateGeneration(sn_r, MaxScoreAggregation())
    cands = cg.query(sn_q, global_k=40)                                  # sscd_baseline.py:98-100
    np.savez_compressed(
        os.path.join(HERE, "search_small.npz"),
        q_raw=q_raw, r_raw=r_raw, z_raw=z_raw,
        q_len=np.array([len(v) for v in queries]), r_len=np.array([len(v) for v in refs]),
        z_len=np.array([len(v) for v in noise]),
        sn_q=sn_q_arr, sn_r=sn_r_arr, D10=D10, I10=I10, lims=lims, Dr=Dr, Ir=Ir,
        cand_q=np.array([c.query_id for c in cands]), cand_r=np.array([c.ref_id for c in cands]),
        cand_s=np.array([c.score for c in cands], dtype=np.float32), global_k=np.array(40),
    )
    print("search_small.npz", sn_q_arr.shape, sn_r_arr.shape, len(cands), "candidates")
    refload.unload_vsc()


if __name__ == "__main__":
    assert refload.available(), "/root/reference is required to (re)generate the fixtures"
    make_vit()
    make_search()
