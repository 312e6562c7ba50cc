"""Seam C: environment-only activation.  A child interpreter with <repo>/shims leading PYTHONPATH and
VSCB200_ACTIVATE=1 imports the UNMODIFIED reference modules and finds the device classes swapped in, `import faiss`
resolving to faiss_compat and torch.jit.load wrapped -- all without initialising CUDA (container-only)."""
import json
import os
import subprocess
import sys

import pytest

from oracle import refload

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys
import torch
import faiss
import vsc.index, vsc.candidates
import vsc.baseline.localization as L
import vsc.baseline.score_normalization as SN
from vsc.baseline import sscd_baseline
print(json.dumps({
    "faiss": faiss.IndexFlat.__module__,
    "VideoIndex": vsc.index.VideoIndex.__module__,
    "CandidateGeneration": sscd_baseline.CandidateGeneration.__module__,
    "MaxScoreAggregation": sscd_baseline.MaxScoreAggregation.__module__,
    "MaxSim": sscd_baseline.VCSLLocalizationMaxSim.__module__,
    "CandidateScore": sscd_baseline.VCSLLocalizationCandidateScore.__module__,
    "score_normalize": sscd_baseline.score_normalize.__module__,
    "query_score_normalize": SN.query_score_normalize.__module__,
    "jit_load": torch.jit.load.__module__,
    "cuda_initialized": torch.cuda.is_initialized(),
    "VideoFeature": vsc.index.VideoFeature.__module__,
}))
"""


def _run(activate: bool):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "shims"), REPO, os.path.join(REPO, "oracle", "refshim"),
                                         os.path.join(refload.D, "infer")])
    env.pop("VSCB200_ACTIVATE", None)
    if activate:
        env["VSCB200_ACTIVATE"] = "1"
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not refload.available(), reason="/root/reference not present")
def test_activation_swaps_the_device_classes_in():
    got = _run(True)
    pkg = "vsc22_submission_b200"
    assert got["faiss"] == f"{pkg}.faiss_compat"
    assert got["VideoIndex"] == got["CandidateGeneration"] == got["MaxScoreAggregation"] == f"{pkg}.candidates"
    assert got["MaxSim"] == got["CandidateScore"] == f"{pkg}.localization"
    assert got["score_normalize"] == got["query_score_normalize"] == f"{pkg}.search"
    assert got["jit_load"] == f"{pkg}.encoder"
    assert got["VideoFeature"] == "vsc.index"                  # boundary types stay the reference's
    assert got["cuda_initialized"] is False


@pytest.mark.skipif(not refload.available(), reason="/root/reference not present")
def test_without_the_variable_nothing_is_patched():
    got = _run(False)
    assert got["VideoIndex"] == "vsc.index" and got["CandidateGeneration"] == "vsc.candidates"
    assert got["MaxSim"] == "vsc.baseline.localization" and got["jit_load"].startswith("torch")
