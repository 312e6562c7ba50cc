"""The encoder call patterns of the reference without their per-batch host round trips.

Mirrors, same names and argument meaning:

* ``extract_vsc_feat(model, dataloader, device)`` -- VSC22-Descriptor-Track-1st/infer/src/extractor.py:9-37: batches of
  ``(video_frames [B, S, 3, H, W], video_mask [B, S], video_id)`` from ``D_vsc.collate_fn`` (dataset.py:149-155);
  masked frames are gathered, encoded, and returned as ``(video ids per frame, features [N, D] numpy, timestamps)``.
  The reference synchronises twice per batch (``video_mask.sum`` on the device feeding Python ranges, and
  ``out_feature.cpu()``); here the mask arithmetic stays on the host (the mask arrives as a CPU tensor), uploads are
  asynchronous and the descriptors of ALL batches come back in one device-to-host copy at the end.
* ``single_infer(model, feature, len_threshold=48)`` -- extract_query_feats.py:143-153 / M/infer/infer_matching.py:122-133:
  the chunked call of one encoder over one video's frames (``[:, 0]`` when the module returns tokens).

Both work with any ``nn.Module``; with the B200 encoders (``encoder.B200ViTEncoder`` / ``swin_encoder.B200SwinEncoder``)
the chunking is done inside the plan, so ``single_infer`` passes the whole video in one call.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Tuple

import numpy as np
import torch


def extract_vsc_feat(model, dataloader: Iterable, device) -> Tuple[List[str], np.ndarray, np.ndarray]:
    feats, save_vids, save_timestamps = [], [], []
    on_gpu = torch.device(device).type == "cuda"
    for video_frames, video_mask, video_id in dataloader:
        mask_host = video_mask.bool() if video_mask.device.type == "cpu" else video_mask.bool().cpu()
        frame_num = mask_host.sum(dim=1).tolist()
        video_frames = video_frames.to(device, non_blocking=True)
        with torch.no_grad():
            flat_frames = video_frames[mask_host.to(device, non_blocking=True)]
            out_feature = model(flat_frames)
        assert out_feature.shape[0] == sum(frame_num)
        feats.append(out_feature.detach())
        for vid, n in zip(video_id, frame_num):
            save_vids.extend([vid] * n)
        save_timestamps.extend(range(n) for n in frame_num)
    if not feats:
        raise ValueError("need at least one array to concatenate")        # what np.concatenate([]) raises in the reference
    all_feats = torch.cat(feats)
    if on_gpu:
        host = torch.empty(all_feats.shape, dtype=all_feats.dtype, pin_memory=True)
        host.copy_(all_feats, non_blocking=True)
        torch.cuda.current_stream(device).synchronize()
        save_feat = host.numpy().copy()
    else:
        save_feat = all_feats.numpy()
    return save_vids, save_feat, np.concatenate([np.arange(len(r)) for r in save_timestamps])


def single_infer(model, feature: torch.Tensor, len_threshold: int = 48) -> np.ndarray:
    whole = getattr(model, "max_frames", None) is not None          # B200 encoders chunk inside the plan
    step = feature.shape[0] if whole and feature.shape[0] > 0 else len_threshold
    outs = []
    with torch.no_grad():
        for i in range(math.ceil(feature.shape[0] / step)):
            flat = model(feature[i * step:(i + 1) * step, ...])
            if flat.dim() == 3:
                flat = flat[:, 0]
            outs.append(flat)
    if not outs:
        raise ValueError("need at least one array to concatenate")
    return torch.cat(outs).cpu().numpy()
