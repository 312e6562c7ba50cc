"""The evaluation entry point of the descriptor track with every stage on the device.

Mirrors ``search`` / ``localize_and_verify`` / ``match`` of
VSC22-Descriptor-Track-1st/infer/vsc/baseline/sscd_baseline.py:87-170 (what ``D/infer/eval.sh:12-16`` runs:
``python -m vsc.baseline.sscd_baseline --query_features ... --ref_features ... --output_path ... --overwrite``),
same constants and argument meaning:

* ``search``: ``CandidateGeneration(refs, MaxScoreAggregation()).query(queries, global_k=1200*|Q|)`` cut to
  ``25*|Q|`` candidates (sscd_baseline.py:87-101)  -> candidates.CandidateGeneration (csrc/global_topk.cu)
* ``localize_and_verify``: the first ``5*|Q|`` candidates through ``VCSLLocalizationMaxSim`` (score-normalised run,
  similarity_bias 0.5) or ``VCSLLocalizationCandidateScore`` on row-normalised features, ``model_type="TN"``,
  ``tn_max_step=5``, ``min_length=4``, batches of 512 (sscd_baseline.py:104-152)  -> localization.* (csrc/tn_align.cu)
* ``match``: both stages + ``candidates.csv`` / ``matches.csv`` with the reference's columns (metrics.py:65-70, 219-226)
* ``run``: ``main`` (sscd_baseline.py:178-211) on already loaded features, optional score normalisation with
  ``beta=1.2`` (search.score_normalize).
"""
from __future__ import annotations

import dataclasses
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import candidates as _cand
from . import localization as _loc
from . import search as _search


def search(queries: Sequence, refs: Sequence, retrieve_per_query: float = 1200.0, candidates_per_query: float = 25.0) -> list:
    cg = _cand.CandidateGeneration(refs, _cand.MaxScoreAggregation())
    num_to_retrieve = int(retrieve_per_query * len(queries))
    return cg.query(queries, global_k=num_to_retrieve, limit=int(candidates_per_query * len(queries)))


def _l2_normalize_rows(videos: Sequence) -> list:
    """``transform_features(videos, sklearn.preprocessing.normalize)`` (sscd_baseline.py:124-125) on the device."""
    x = _search._cat(videos, torch.device("cuda", torch.cuda.current_device()), "pipe_norm")
    n = x.norm(dim=1, keepdim=True)
    x = x / torch.where(n == 0, torch.ones_like(n), n)          # sklearn leaves all-zero rows untouched
    return _search._split(videos, _search._to_host(x, "pipe_norm_out"))


def localize_and_verify(queries: Sequence, refs: Sequence, candidates: Sequence, localize_per_query: float = 5.0,
                        score_normalization: bool = False) -> list:
    candidates = candidates[: int(len(queries) * localize_per_query)]
    if score_normalization:
        alignment = _loc.VCSLLocalizationMaxSim(queries, refs, model_type="TN", tn_max_step=5, min_length=4,
                                                concurrency=16, similarity_bias=0.5)
    else:
        alignment = _loc.VCSLLocalizationCandidateScore(_l2_normalize_rows(queries), _l2_normalize_rows(refs),
                                                        model_type="TN", tn_max_step=5, min_length=4, concurrency=16)
    matches, batch = [], 512
    for i in range(0, len(candidates), batch):
        matches.extend(alignment.localize_all(candidates[i:i + batch]))
    return matches


def write_candidates_csv(candidates: Sequence, path: str):
    import pandas as pd
    pd.DataFrame([{"query_id": c.query_id, "ref_id": c.ref_id, "score": c.score} for c in candidates],
                 columns=["query_id", "ref_id", "score"]).to_csv(path, index=False)


def write_matches_csv(matches: Sequence, path: str):
    import pandas as pd
    cols = ["query_id", "ref_id", "query_start", "query_end", "ref_start", "ref_end", "score"]
    df = pd.DataFrame([m._asdict() for m in matches], columns=cols)
    for c in cols[2:6]:
        df[c] = df[c].astype(np.float64)
    df.to_csv(path, index=False)


def match(queries: Sequence, refs: Sequence, output_path: str, score_normalization: bool = False) -> Tuple[str, str]:
    found = search(queries, refs)
    os.makedirs(output_path, exist_ok=True)
    candidate_file = os.path.join(output_path, "candidates.csv")
    write_candidates_csv(found, candidate_file)
    matches = localize_and_verify(queries, refs, found, score_normalization=score_normalization)
    matches_file = os.path.join(output_path, "matches.csv")
    write_matches_csv(matches, matches_file)
    return candidate_file, matches_file


def run(queries: Sequence, refs: Sequence, output_path: str, score_norm_refs: Optional[Sequence] = None,
        overwrite: bool = False) -> Tuple[str, str]:
    if os.path.exists(output_path) and not overwrite:
        raise Exception(f"Output path already exists: {output_path}. Do you want to --overwrite?")
    score_normalization = False
    if score_norm_refs:
        queries, refs = _search.score_normalize(queries, refs, score_norm_refs, beta=1.2)
        score_normalization = True
    return match(queries, refs, output_path, score_normalization=score_normalization)
