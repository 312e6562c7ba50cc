"""Numpy restatement of the near-duplicate frame filter of the query extractor.

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for ensemble.near_dup_keep / csrc/ensemble.cu.

Restates VSC22-Descriptor-Track-1st/infer/extract_query_feats.py:190-199 (inside ``Main.process``; the same block runs
in the matching track's query path): rows normalised, ``sim_mat = feat @ feat.T - eye`` (float64 by promotion), frames
visited by descending column mean, a visited frame that has not been removed removes every frame whose similarity to it
exceeds ``FRAME_THRESHOLD``.  Pinned by tests/test_oracle_near_dup.py, which executes those very source lines read from
the reference file.
"""
import numpy as np


def keep_indices(features: np.ndarray, frame_threshold: float = 0.975) -> list:
    feat = features / np.linalg.norm(features, axis=1, keepdims=True)
    sim = np.matmul(feat, feat.T) - np.eye(len(feat))
    order = sim.mean(0).argsort()[::-1]
    removed = np.zeros(len(feat), bool)
    for i in order:
        if removed[i]:
            continue
        removed |= sim[i] > frame_threshold
    return [i for i in range(len(feat)) if not removed[i]]
