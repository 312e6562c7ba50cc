"""Environment-only activation of the B200 path under the UNMODIFIED reference (SURVEY.md 8b seam C).

``activate()`` -- called by ``shims/sitecustomize.py`` when ``VSCB200_ACTIVATE=1`` and ``<repo>/shims:<repo>`` lead
``PYTHONPATH``, or by a launcher -- arranges that, without touching a file under the reference tree:

* ``import faiss`` resolves to ``faiss_compat`` (the ``shims/faiss`` package; vsc/index.py:11, exhaustive_search.py:6, ...);
* ``torch.jit.load(ckpt)`` returns a B200 encoder for recognised checkpoints (extract_ref_feats.py:24, ...), installed
  lazily the first time ``torch.jit`` is used so that importing this module never initialises CUDA (the reference forks
  its DataLoader workers first, inference.py:1-17);
* right after the reference imports one of its own modules, the classes that have a device implementation are swapped
  in: ``vsc.index.VideoIndex``, ``vsc.candidates.{VideoIndex, CandidateGeneration, MaxScoreAggregation}``
  (candidates.py), ``vsc.baseline.localization.VCSLLocalization{,MaxSim,CandidateScore}`` (localization.py) and
  ``vsc.baseline.score_normalization.{score_normalize, query_score_normalize, ref_score_normalize}`` (search.py).

``D/infer/eval.sh`` -> ``python3 -m vsc.baseline.sscd_baseline ...`` then runs retrieval, score normalisation and
localisation on the GPU and writes the same ``candidates.csv`` / ``matches.csv``.
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import sys

_ACTIVE = False


def _patch_index(m):
    from . import candidates as b
    m.VideoIndex = b.VideoIndex


def _patch_candidates(m):
    from . import candidates as b
    m.VideoIndex, m.CandidateGeneration, m.MaxScoreAggregation = b.VideoIndex, b.CandidateGeneration, b.MaxScoreAggregation


def _patch_localization(m):
    from . import localization as b
    m.VCSLLocalization, m.VCSLLocalizationMaxSim = b.VCSLLocalization, b.VCSLLocalizationMaxSim
    m.VCSLLocalizationCandidateScore = b.VCSLLocalizationCandidateScore


def _patch_score_norm(m):
    from . import search as b
    m.score_normalize, m.query_score_normalize, m.ref_score_normalize = (b.score_normalize, b.query_score_normalize,
                                                                         b.ref_score_normalize)


PATCHES = {"vsc.index": _patch_index, "vsc.candidates": _patch_candidates,
           "vsc.baseline.localization": _patch_localization, "vsc.baseline.score_normalization": _patch_score_norm}


class _PatchingLoader(importlib.abc.Loader):
    def __init__(self, loader, patch):
        self._loader, self._patch = loader, patch

    def create_module(self, spec):
        return self._loader.create_module(spec)

    def exec_module(self, module):
        self._loader.exec_module(module)
        self._patch(module)


class _PostImportFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, name, path, target=None):
        patch = PATCHES.get(name)
        if patch is None:
            return None
        for finder in sys.meta_path:
            if finder is self or not hasattr(finder, "find_spec"):
                continue
            spec = finder.find_spec(name, path, target)
            if spec is not None and spec.loader is not None:
                spec.loader = _PatchingLoader(spec.loader, patch)
                return spec
        return None


def activate(max_frames: int = 256) -> None:
    global _ACTIVE
    if _ACTIVE:
        return
    _ACTIVE = True
    sys.meta_path.insert(0, _PostImportFinder())
    for name, patch in PATCHES.items():            # modules the process imported before activation
        if name in sys.modules:
            patch(sys.modules[name])
    from .encoder import install_jit_load_hook     # wraps torch.jit.load; no CUDA work until a checkpoint is loaded
    install_jit_load_hook(max_frames)
