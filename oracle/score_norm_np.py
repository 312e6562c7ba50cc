"""numpy restatement of the reference's score normalisation on plain arrays.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
``VSC22-Descriptor-Track-1st/infer/vsc/baseline/score_normalization.py``:

* ``low_var_dim``           :71-72   ``sn_features.var(axis=0).argmin()``
* ``drop_dim``              :73-78   ``np.delete(feature, low_var_dim, axis=1)``
* ``l2_normalize_rows``     :79-83   ``sklearn.preprocessing.normalize`` (rows, L2; zero rows stay 0)
* ``score_normalize``       :33-104  bias = -beta * mean(top-nk sims vs noise bank), appended as an
                                     extra query dim; refs get a constant 1 appended
* ``query_score_normalize`` :107-148 same, query side only, with the video-score gate
                                     (bias = -100 when video_score < threshold, :142-143)
* ``ref_score_normalize``   :150-192 reference side only
* ``score_normalize_v2``    VSC22-Matching-Track-1st/vsc/baseline/score_normalization.py:115-156

"Videos" are lists of ``[n_i, d]`` float32 arrays; the per-video loop of the reference
(:93-98) is a row-wise operation, so it is restated on the concatenation.  The reference has
no test for this function (parity unpinned there); it is pinned here by running
the reference's own function through ``oracle.faiss_np`` (tests/test_oracle_reference_pins.py,
container only) and by the committed fixture tests/golden/search_small.npz (arrays sn_q / sn_r: outputs of the
reference function written by tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

from . import faiss_np


def low_var_dim(noise: np.ndarray) -> int:
    return int(np.asarray(noise).var(axis=0).argmin())


def drop_dim(x: np.ndarray, dim: int) -> np.ndarray:
    return np.delete(x, dim, axis=1)


def l2_normalize_rows(x: np.ndarray) -> np.ndarray:
    """sklearn.preprocessing.normalize(x) for dense float input: x / max(||x||, tiny), with
    all-zero rows left untouched (sklearn's ``_handle_zeros_in_scale``)."""
    x = np.asarray(x)
    norms = np.sqrt(np.einsum("ij,ij->i", x, x))
    norms[norms == 0.0] = 1.0
    return x / norms[:, None]


def noise_bias(q: np.ndarray, noise: np.ndarray, beta: float, nk: int) -> np.ndarray:
    """[-beta * mean of the nk best inner products of each row of q against the noise bank]."""
    index = faiss_np.IndexFlat(noise.shape[1], faiss_np.METRIC_INNER_PRODUCT)
    index.add(noise)
    sim, _ = index.search(q, nk)
    return (-beta * sim[:, :nk].mean(axis=1, keepdims=True)).astype(np.float32)


def score_normalize(queries, refs, noise, l2_normalize=True, replace_dim=True, beta=1.0, nk=1):
    """-> (queries' [nq, d'], refs' [nr, d'], low_var_dim) with d' = d (replace_dim) or d+1."""
    q, r, z = (np.asarray(a, dtype=np.float32) for a in (queries, refs, noise))
    lvd = -1
    if replace_dim:
        lvd = low_var_dim(z)
        q, r, z = (drop_dim(a, lvd) for a in (q, r, z))
    if l2_normalize:
        q, r, z = (l2_normalize_rows(a) for a in (q, r, z))
    bias = noise_bias(q, z, beta, nk)
    q2 = np.concatenate([q, bias], axis=1)
    r2 = np.concatenate([r, np.ones_like(r[:, :1])], axis=1)
    return q2.astype(np.float32), r2.astype(np.float32), lvd


def query_score_normalize(queries, noise, gated_rows=None, low_var_dim_=0, l2_normalize=True,
                          replace_dim=True, beta=1.0, nk=1):
    """``gated_rows``: boolean [nq]; True where the owning video's score is below the
    threshold (score_normalization.py:142) => bias is the constant -100."""
    q, z = (np.asarray(a, dtype=np.float32) for a in (queries, noise))
    if replace_dim:
        q, z = (drop_dim(a, low_var_dim_) for a in (q, z))
    if l2_normalize:
        q, z = (l2_normalize_rows(a) for a in (q, z))
    bias = noise_bias(q, z, beta, nk)
    if gated_rows is not None:
        bias[np.asarray(gated_rows, bool)] = -100.0
    return np.concatenate([q, bias], axis=1).astype(np.float32)


def ref_score_normalize(refs, noise, l2_normalize=True, replace_dim=True):
    r, z = (np.asarray(a, dtype=np.float32) for a in (refs, noise))
    lvd = -1
    if replace_dim:
        lvd = low_var_dim(z)
        r = drop_dim(r, lvd)
    if l2_normalize:
        r = l2_normalize_rows(r)
    return np.concatenate([r, np.ones_like(r[:, :1])], axis=1).astype(np.float32), lvd


def score_normalize_v2(queries, refs, noise, beta=0.35, nk=10):
    """M/vsc/baseline/score_normalization.py:115-156: subtract beta * mean of the nk nearest
    (un-normalised) noise vectors from both sides, then L2-normalise."""
    q, r, z = (np.asarray(a, dtype=np.float32) for a in (queries, refs, noise))
    zn = l2_normalize_rows(z)
    index = faiss_np.IndexFlat(z.shape[1], faiss_np.METRIC_INNER_PRODUCT)
    index.add(zn)
    outs = []
    for x in (q, r):
        _, ids = index.search(l2_normalize_rows(x), nk)
        outs.append(l2_normalize_rows(x - z[ids].mean(1) * beta).astype(np.float32))
    return outs[0], outs[1]
