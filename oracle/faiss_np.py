"""Exact numpy restatement of the ``faiss`` subset the reference's hot path (B) calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The arithmetic of path (B) lives in the
third-party wheel **faiss-gpu (conda channel pytorch, version NOT pinned** --
``/root/reference/dockerfile:19``, ``setup.sh:18``), which is absent from /root/reference and
from this image.  IndexFlat is exact brute force (fp32 inner product / squared L2 followed by
a k-select), so its published semantics are restated here and parity is anchored on the
reference's own call sites and tests:

* ``index_factory(d, "Flat", metric)``      vsc/index.py:81
* ``index.add(x)``                          vsc/index.py:94
* ``index.search(x, k) -> (D, I)``          vsc/index.py:174, score_normalization.py:95,141,
                                            exhaustive_search.py:62, M/infer/infer_matching.py:232
* ``index.range_search(x, thr)``            exhaustive_search.py:74,126,246, infer_matching.py:235
* ``index.reset()`` / ``.ntotal`` / ``.metric_type``   exhaustive_search.py:39,60,63
* ``get_num_gpus`` / ``index_cpu_to_all_gpus`` / ``GpuMultipleClonerOptions``
                                            vsc/index.py:169-171, exhaustive_search.py:229-234
* ``ResultHeap`` / ``IndexFlat``            exhaustive_search.py:24-26 (knn_ground_truth)

Pinned by: the reference's unit tests run against this module
(tests/test_oracle_reference_pins.py: test_candidates.py:72-83 exact scores 2.0/1.0/0.25,
test_index.py:40-53 self-match under L2 in both search modes) and the committed fixtures in
tests/golden/ produced from it through the reference's unmodified ``vsc`` package.

Conventions fixed here (faiss leaves them open) and mirrored by the CUDA path:
  * scores are accumulated in float64 and rounded once to float32 (the CUDA kernel's split-bf16
    tensor-core path is fp32-equivalent; tests compare with a tie-aware checker);
  * ties are broken towards the lower database id;
  * ``search`` pads with id -1 and score -FLT_MAX (IP) / +FLT_MAX (L2) when k > ntotal
    (faiss HeapArray neutral element);
  * ``range_search`` uses strict ``>`` (IP) / ``<`` (L2) and lists each row's hits by
    ascending database id; ``lims`` is uint64.
"""
from __future__ import annotations

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1

_FLT_MAX = np.float32(3.4028234663852886e38)

# "f64": accumulate in float64, round once (the checker).  "f32": plain sgemm, which is what
# faiss-CPU IndexFlat executes -- used only when this module is TIMED as the CPU baseline.
_ACCUMULATE = "f64"


def set_accumulate(mode: str):
    global _ACCUMULATE
    assert mode in ("f64", "f32")
    _ACCUMULATE = mode


def _as_f32_2d(x, d):
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim != 2 or x.shape[1] != d:
        raise AssertionError(f"expected a [n, {d}] float32 array, got {x.shape}")
    return x


class IndexFlat:
    """Brute-force exact index (faiss ``IndexFlat`` / ``IndexFlatIP`` / ``IndexFlatL2``)."""

    def __init__(self, d: int, metric: int = METRIC_L2):
        self.d = int(d)
        self.metric_type = int(metric)
        self._xb = np.zeros((0, self.d), dtype=np.float32)
        self.is_trained = True

    @property
    def ntotal(self) -> int:
        return self._xb.shape[0]

    def add(self, x):
        x = _as_f32_2d(x, self.d)
        self._xb = np.concatenate([self._xb, x], axis=0)

    def reset(self):
        self._xb = np.zeros((0, self.d), dtype=np.float32)

    # -- scoring ---------------------------------------------------------------------------
    def _scores(self, xq: np.ndarray, lo: int = 0, hi: int | None = None) -> np.ndarray:
        """float32 score block [nq, hi-lo]: inner product, or squared L2 as |q|^2+|r|^2-2qr
        (faiss's BLAS formulation, clamped at 0)."""
        xb = self._xb[lo:hi]
        if _ACCUMULATE == "f32":
            ip = xq @ xb.T
        else:
            ip = (xq.astype(np.float64) @ xb.astype(np.float64).T)
        if self.metric_type == METRIC_INNER_PRODUCT:
            return ip.astype(np.float32, copy=False)
        qn = (xq.astype(np.float64) ** 2).sum(1)[:, None]
        rn = (xb.astype(np.float64) ** 2).sum(1)[None, :]
        return np.maximum(qn + rn - 2.0 * ip, 0.0).astype(np.float32)

    def search(self, x, k: int):
        xq = _as_f32_2d(x, self.d)
        nq, k = xq.shape[0], int(k)
        keep_max = self.metric_type == METRIC_INNER_PRODUCT
        D = np.full((nq, k), -_FLT_MAX if keep_max else _FLT_MAX, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        nb = self.ntotal
        if nb == 0 or nq == 0 or k == 0:
            return D, I
        kk = min(k, nb)
        bs = max(1, (1 << 26) // max(nb, 1))  # ~256 MB of float32 scores per block
        for q0 in range(0, nq, bs):
            S = self._scores(xq[q0:q0 + bs])
            key = -S if keep_max else S
            # stable argsort => ties resolved towards the lower database id
            order = np.argsort(key, axis=1, kind="stable")[:, :kk]
            D[q0:q0 + bs, :kk] = np.take_along_axis(S, order, axis=1)
            I[q0:q0 + bs, :kk] = order
        return D, I

    def range_search(self, x, thresh: float):
        xq = _as_f32_2d(x, self.d)
        nq = xq.shape[0]
        thresh = np.float32(thresh)
        keep_max = self.metric_type == METRIC_INNER_PRODUCT
        lims = np.zeros(nq + 1, dtype=np.uint64)
        Ds, Is = [], []
        nb = self.ntotal
        bs = max(1, (1 << 26) // max(nb, 1))
        for q0 in range(0, nq, bs):
            S = self._scores(xq[q0:q0 + bs]) if nb else np.zeros((len(xq[q0:q0 + bs]), 0), np.float32)
            mask = (S > thresh) if keep_max else (S < thresh)
            cnt = mask.sum(1)
            lims[q0 + 1:q0 + 1 + len(cnt)] = cnt
            rows, cols = np.nonzero(mask)  # row-major order => ascending id inside each row
            Ds.append(S[rows, cols])
            Is.append(cols.astype(np.int64))
        lims = np.cumsum(lims, dtype=np.uint64)
        D = np.concatenate(Ds) if Ds else np.zeros(0, np.float32)
        I = np.concatenate(Is) if Is else np.zeros(0, np.int64)
        return lims, D.astype(np.float32), I


class IndexFlatIP(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_INNER_PRODUCT)


class IndexFlatL2(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_L2)


def index_factory(d: int, description: str = "Flat", metric: int = METRIC_L2):
    if description != "Flat":
        raise RuntimeError(f"oracle faiss stand-in only implements 'Flat', got {description!r}")
    return IndexFlat(d, metric)


def get_num_gpus() -> int:
    """The oracle is the CPU path: the reference then takes its ``ngpu == 0`` branches
    (exhaustive_search.py:246, index.py:169)."""
    return 0


class GpuMultipleClonerOptions:
    def __init__(self):
        self.shard = False


def index_cpu_to_all_gpus(index, co=None, ngpu=-1):
    return index


class ResultHeap:
    """faiss.ResultHeap as used by exhaustive_search.knn_ground_truth (:24-44)."""

    def __init__(self, nq, k, keep_max=False):
        self.nq, self.k, self.keep_max = nq, k, keep_max
        self.D = np.full((nq, k), -_FLT_MAX if keep_max else _FLT_MAX, dtype=np.float32)
        self.I = np.full((nq, k), -1, dtype=np.int64)

    def add_result(self, D, I):
        D = np.concatenate([self.D, np.asarray(D, np.float32)], axis=1)
        I = np.concatenate([self.I, np.asarray(I, np.int64)], axis=1)
        key = -D if self.keep_max else D
        order = np.argsort(key, axis=1, kind="stable")[:, :self.k]
        self.D = np.take_along_axis(D, order, axis=1)
        self.I = np.take_along_axis(I, order, axis=1)

    def finalize(self):
        pass


def knn(xq, xb, k, metric=METRIC_L2):
    index = IndexFlat(xb.shape[1], metric)
    index.add(xb)
    return index.search(xq, k)
