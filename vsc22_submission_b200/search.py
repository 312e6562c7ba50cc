"""Device-resident similarity search and score normalisation (host side of seam B, one level above
``faiss_compat``): the reference's ``score_normalize`` family and kNN, with descriptors living in HBM.

Mirrors, with the same names / argument meaning / error behaviour:

* ``score_normalize`` / ``query_score_normalize`` / ``ref_score_normalize``
  -- VSC22-Descriptor-Track-1st/infer/vsc/baseline/score_normalization.py:33-104, 107-148, 150-192
  (operating on lists of ``VideoFeature``-like objects: ``.video_id``, ``.feature``; returned with
  ``dataclasses.replace`` exactly as the reference does)
* ``DeviceIndex`` -- faiss ``IndexFlat`` with torch CUDA tensors in and out (no host copies)

All arithmetic is in libvscb200.so; torch is used for device memory and streams only.  The reference
loops ``index.search`` once per query VIDEO (score_normalization.py:93-98 -- 8 295 bank scans at test
scale); here all query rows go through ONE batched search.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

METRIC_INNER_PRODUCT = _lib.METRIC_INNER_PRODUCT
METRIC_L2 = _lib.METRIC_L2


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32_cuda(x: torch.Tensor, what: str) -> torch.Tensor:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (no CPU fallback)")
    return x.contiguous().float()


class DeviceIndex:
    """Flat exact index over device-resident descriptors (faiss.IndexFlat semantics, tensors in/out)."""

    def __init__(self, d: int, metric: int = METRIC_INNER_PRODUCT, device: Optional[torch.device] = None):
        self.d, self.metric_type = int(d), int(metric)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().vscb200_index_create(self.d, self.metric_type, C.byref(self._ptr)), "index_create")

    def __del__(self):
        try:
            if self._ptr:
                _lib.lib().vscb200_index_destroy(self._ptr)
                self._ptr = None
        except Exception:
            pass

    @property
    def ntotal(self) -> int:
        return int(_lib.lib().vscb200_index_ntotal(self._ptr))

    def last_fallbacks(self) -> int:
        """Diagnostic: queries of the last small-k batch search that needed the exhaustive fp32 fallback (-1: none ran)."""
        return int(_lib.lib().vscb200_index_last_fallbacks(self._ptr))

    def set_id_offset(self, offset: int):
        """Global id of this shard's first row (bank sharding over ranks)."""
        _lib.check(_lib.lib().vscb200_index_set_id_offset(self._ptr, int(offset)), "index_set_id_offset")

    def add(self, x: torch.Tensor):
        x = _f32_cuda(x, "DeviceIndex.add")
        if x.dim() != 2 or x.shape[1] != self.d:
            raise AssertionError(f"add: expected [n, {self.d}], got {tuple(x.shape)}")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().vscb200_index_add(self._ptr, _p(x), x.shape[0], _stream(self.device)), "index.add")

    def add_sn(self, x: torch.Tensor, drop_dim, l2_normalize: bool, fill: float = 0.0, bias: Optional[torch.Tensor] = None):
        """``add(sn_transform(x, drop_dim, l2_normalize, fill, bias))`` in one pass over the raw rows: the transformed array
        is never materialised (score_normalization.py:73-83, 96-101 + vsc/index.py:94); the stored rows are bit-identical to
        the two-step form.  ``drop_dim``: int, or the device tensor of ``low_var_dim_device``."""
        x = _f32_cuda(x, "DeviceIndex.add_sn")
        dev_dim = isinstance(drop_dim, torch.Tensor)
        drops = dev_dim or 0 <= int(drop_dim) < x.shape[1]
        if x.dim() != 2 or x.shape[1] + (0 if drops else 1) != self.d:
            raise AssertionError(f"add_sn: rows of {tuple(x.shape)} do not transform to dimension {self.d}")
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().vscb200_index_add_sn(self._ptr, _p(x), x.shape[0], x.shape[1], -1 if dev_dim else int(drop_dim),
                                                       _p(drop_dim) if dev_dim else None, int(bool(l2_normalize)), float(fill),
                                                       _p(bias), _stream(self.device)), "index.add_sn")

    def reconstruct_n(self, i0: int, n: int) -> torch.Tensor:
        """Rows [i0, i0 + n) of the bank as stored (index.reconstruct_n)."""
        out = torch.empty((n, self.d), dtype=torch.float32, device=self.device)
        if n:
            _lib.check(_lib.lib().vscb200_index_reconstruct_n(self._ptr, int(i0), int(n), _p(out)), "index.reconstruct_n")
        return out

    def reset(self):
        _lib.check(_lib.lib().vscb200_index_reset(self._ptr), "index.reset")

    def search(self, q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        q = _f32_cuda(q, "DeviceIndex.search")
        if q.dim() != 2 or q.shape[1] != self.d:
            raise AssertionError(f"search: expected [n, {self.d}], got {tuple(q.shape)}")
        nq = q.shape[0]
        D = torch.empty((nq, k), dtype=torch.float32, device=self.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=self.device)
        if nq:
            if torch.cuda.current_device() == self.device.index:      # the per-video call pattern is host-latency sensitive
                _lib.check(_lib.lib().vscb200_index_search(self._ptr, _p(q), nq, int(k), _p(D), _p(I),
                                                           _stream(self.device)), "index.search")
            else:
                with torch.cuda.device(self.device):
                    _lib.check(_lib.lib().vscb200_index_search(self._ptr, _p(q), nq, int(k), _p(D), _p(I),
                                                               _stream(self.device)), "index.search")
        return D, I

    def global_search(self, q: torch.Tensor, global_k: int = 0, threshold: Optional[float] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """The ``global_k`` best (query row, bank row) pairs over ALL query rows -- the result of
        ``VideoIndex._global_threshold_knn_search`` (vsc/index.py:142-165) -- and/or every pair strictly better
        than ``threshold`` (infer_matching.py:229-247).  Returns (scores f32 [n], query rows i64 [n], bank rows
        i64 [n]) on the device, best first, ties by (query row, bank row); scores are exact fp32."""
        q = _f32_cuda(q, "DeviceIndex.global_search")
        if q.dim() != 2 or q.shape[1] != self.d:
            raise AssertionError(f"global_search: expected [n, {self.d}], got {tuple(q.shape)}")
        n = C.c_int64(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().vscb200_index_global_search(
                self._ptr, _p(q), q.shape[0], int(global_k), 0 if threshold is None else 1,
                float(0.0 if threshold is None else threshold), C.byref(n), _stream(self.device)), "index.global_search")
            sc = torch.empty(n.value, dtype=torch.float32, device=self.device)
            qi = torch.empty(n.value, dtype=torch.int64, device=self.device)
            ri = torch.empty(n.value, dtype=torch.int64, device=self.device)
            _lib.check(_lib.lib().vscb200_index_global_results(self._ptr, _p(sc), _p(qi), _p(ri), _stream(self.device)),
                       "index.global_results")
        return sc, qi, ri

    def global_video_pairs(self, q_offsets: torch.Tensor, r_offsets: torch.Tensor
                           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Reduce the pairs of the last ``global_search`` to (query video, ref video) candidates scored by their
        best frame pair, best first (vsc/candidates.py:24-40).  ``*_offsets``: int64 [n_videos + 1] first rows."""
        qo = q_offsets.to(self.device, torch.int64).contiguous()
        ro = r_offsets.to(self.device, torch.int64).contiguous()
        m = C.c_int64(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().vscb200_index_global_video_pairs(
                self._ptr, _p(qo), qo.numel() - 1, _p(ro), ro.numel() - 1, C.byref(m), _stream(self.device)),
                "index.global_video_pairs")
            sc = torch.empty(m.value, dtype=torch.float32, device=self.device)
            qv = torch.empty(m.value, dtype=torch.int64, device=self.device)
            rv = torch.empty(m.value, dtype=torch.int64, device=self.device)
            _lib.check(_lib.lib().vscb200_index_video_pair_results(self._ptr, _p(sc), _p(qv), _p(rv),
                                                                   _stream(self.device)), "index.video_pair_results")
        return sc, qv, rv

    def scores(self, q: torch.Tensor) -> torch.Tensor:
        """Dense [nq, ntotal] score matrix (localization.py:32-35 form)."""
        q = _f32_cuda(q, "DeviceIndex.scores")
        S = torch.empty((q.shape[0], self.ntotal), dtype=torch.float32, device=self.device)
        if S.numel():
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().vscb200_index_scores(self._ptr, _p(q), q.shape[0], _p(S), S.stride(0),
                                                           _stream(self.device)), "index.scores")
        return S


# ------------------------------------------------------------------------------------------------
# exchange format of partial top-k results between bank shards (csrc/merge.cu)
# ------------------------------------------------------------------------------------------------
def pack_topk(D: torch.Tensor, I: torch.Tensor, keep_max: bool = True) -> torch.Tensor:
    """(D f32 [nq,k], I i64 [nq,k]) -> int64 [nq,k] keys ``(order-preserving score bits << 32) | ~id`` (0 = padding):
    scores and ids travel in ONE all-gather and a larger key (as uint64) is a better result, ties to the lower id."""
    D, I = _f32_cuda(D, "pack_topk"), I.contiguous()
    keys = torch.empty(D.shape, dtype=torch.int64, device=D.device)
    if D.numel():
        with torch.cuda.device(D.device):
            _lib.check(_lib.lib().vscb200_topk_pack(_p(D), _p(I), D.shape[0], D.shape[1], int(bool(keep_max)), _p(keys),
                                                    _stream(D.device)), "topk_pack")
    return keys


def pack_topk_into(keys: torch.Tensor, col0: int, D: torch.Tensor, I: torch.Tensor, keep_max: bool = True) -> None:
    """``keys[:, col0:col0 + k] = pack_topk(D, I)`` in one kernel (keys int64 [nq, ld], contiguous)."""
    D, I = _f32_cuda(D, "pack_topk_into"), I.contiguous()
    assert keys.dim() == 2 and keys.dtype == torch.int64 and keys.is_contiguous() and keys.shape[0] == D.shape[0]
    if D.numel():
        with torch.cuda.device(D.device):
            _lib.check(_lib.lib().vscb200_topk_pack_cols(_p(D), _p(I), D.shape[0], D.shape[1], int(bool(keep_max)), _p(keys),
                                                         keys.shape[1], int(col0), _stream(D.device)), "topk_pack_cols")


def merge_packed_topk(keys: torch.Tensor, k: int, keep_max: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """keys int64 [parts, nq, kin] (the all-gather layout) -> the k best per query (D f32 [nq,k], I i64 [nq,k])."""
    assert keys.dim() == 3 and keys.dtype == torch.int64 and keys.is_cuda
    keys = keys.contiguous()
    parts, nq, kin = keys.shape
    D = torch.empty((nq, k), dtype=torch.float32, device=keys.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=keys.device)
    if nq:
        with torch.cuda.device(keys.device):
            _lib.check(_lib.lib().vscb200_topk_merge(_p(keys), parts, nq, kin, int(k), int(bool(keep_max)), _p(D), _p(I),
                                                     _stream(keys.device)), "topk_merge")
    return D, I


def merge_packed_topk_cols(keys: torch.Tensor, col0: int, kin: int, k: int, keep_max: bool = True,
                           bias: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``merge_packed_topk`` over columns [col0, col0 + kin) of keys [parts, nq, ld]; ``bias`` f32 [nq] (or [nq, 1]) is
    added to every merged score of its row."""
    assert keys.dim() == 3 and keys.dtype == torch.int64 and keys.is_cuda and keys.is_contiguous()
    parts, nq, ld = keys.shape
    D = torch.empty((nq, k), dtype=torch.float32, device=keys.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=keys.device)
    if nq:
        with torch.cuda.device(keys.device):
            _lib.check(_lib.lib().vscb200_topk_merge_cols(_p(keys), parts, nq, ld, int(col0), int(kin), int(k), int(bool(keep_max)),
                                                          _p(bias), _p(D), _p(I), _stream(keys.device)), "topk_merge_cols")
    return D, I


def col_sums(x: torch.Tensor, sum_in: Optional[torch.Tensor] = None, inv_n: float = 0.0) -> torch.Tensor:
    """float64 [d]: column sums of x, or (``sum_in`` given) sums of squared deviations from ``sum_in * inv_n``."""
    x = _f32_cuda(x, "col_sums")
    out = torch.empty((x.shape[1],), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vscb200_col_sums(_p(x), x.shape[0], x.shape[1], _p(sum_in), float(inv_n), _p(out),
                                               _stream(x.device)), "col_sums")
    return out


def var_argmin_device(ss: torch.Tensor) -> torch.Tensor:
    """First minimum of a float64 [d] vector, left on the device (int32 [1])."""
    out = torch.empty((1,), dtype=torch.int32, device=ss.device)
    with torch.cuda.device(ss.device):
        _lib.check(_lib.lib().vscb200_var_argmin_dev(_p(ss), ss.numel(), _p(out), _stream(ss.device)), "var_argmin_dev")
    return out


def col_moments_local(x: torch.Tensor) -> torch.Tensor:
    """float64 [3d]: (sum x | sum (x - local mean)^2 | (sum x)^2 / n) per column of a bank shard -- all-reduce it over the
    shards, then ``var_argmin_moments`` (one collective instead of two)."""
    x = _f32_cuda(x, "col_moments_local")
    out = torch.empty((3 * x.shape[1],), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().vscb200_col_moments_local(_p(x), x.shape[0], x.shape[1], _p(out), _stream(x.device)), "col_moments_local")
    return out


def var_argmin_moments(m3: torch.Tensor, n_total: int) -> torch.Tensor:
    """First column of minimum variance from the all-reduced moments, left on the device (int32 [1])."""
    out = torch.empty((1,), dtype=torch.int32, device=m3.device)
    with torch.cuda.device(m3.device):
        _lib.check(_lib.lib().vscb200_var_argmin_moments(_p(m3), float(n_total), m3.numel() // 3, _p(out), _stream(m3.device)),
                   "var_argmin_moments")
    return out


# ------------------------------------------------------------------------------------------------
# score normalisation on device tensors
# ------------------------------------------------------------------------------------------------
def low_var_dim(noise: torch.Tensor) -> int:
    """``sn_features.var(axis=0).argmin()`` (score_normalization.py:72)."""
    noise = _f32_cuda(noise, "low_var_dim")
    out = C.c_int(0)
    with torch.cuda.device(noise.device):
        _lib.check(_lib.lib().vscb200_low_var_dim(_p(noise), noise.shape[0], noise.shape[1], C.byref(out),
                                                  _stream(noise.device)), "low_var_dim")
    return int(out.value)


def low_var_dim_device(noise: torch.Tensor) -> torch.Tensor:
    """``low_var_dim`` left on the device (int32 tensor [1]) -- no host synchronisation; feed it to ``sn_transform``."""
    noise = _f32_cuda(noise, "low_var_dim_device")
    out = torch.empty((1,), dtype=torch.int32, device=noise.device)
    with torch.cuda.device(noise.device):
        _lib.check(_lib.lib().vscb200_low_var_dim_dev(_p(noise), noise.shape[0], noise.shape[1], _p(out),
                                                      _stream(noise.device)), "low_var_dim_dev")
    return out


def sn_transform(x: torch.Tensor, drop_dim, l2_normalize: bool, fill: float = 0.0,
                 bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[l2norm(delete(x, drop_dim)), last] in one pass (score_normalization.py:73-83, 96-101).
    drop_dim < 0: nothing dropped, output gains a column.  ``drop_dim`` may be the device tensor of
    ``low_var_dim_device`` (a column is always dropped then)."""
    x = _f32_cuda(x, "sn_transform")
    n, d = x.shape
    if isinstance(drop_dim, torch.Tensor):
        out = torch.empty((n, d), dtype=torch.float32, device=x.device)
        if n:
            with torch.cuda.device(x.device):
                _lib.check(_lib.lib().vscb200_sn_transform_dev(_p(x), n, d, _p(drop_dim), int(bool(l2_normalize)), float(fill),
                                                               _p(bias), _p(out), _stream(x.device)), "sn_transform_dev")
        return out
    dout = d if 0 <= drop_dim < d else d + 1
    out = torch.empty((n, dout), dtype=torch.float32, device=x.device)
    if n:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().vscb200_sn_transform(_p(x), n, d, int(drop_dim), int(bool(l2_normalize)), float(fill),
                                                       _p(bias), _p(out), _stream(x.device)), "sn_transform")
    return out


def bias_from_topk(D: torch.Tensor, beta: float, nk: int) -> torch.Tensor:
    """bias[row] = -beta * mean(D[row, :nk])  (score_normalization.py:96)."""
    D = _f32_cuda(D, "bias_from_topk")
    bias = torch.empty((D.shape[0],), dtype=torch.float32, device=D.device)
    if D.shape[0]:
        with torch.cuda.device(D.device):
            _lib.check(_lib.lib().vscb200_sn_bias(_p(D), D.shape[0], D.shape[1], int(nk), float(beta), _p(bias),
                                                  _stream(D.device)), "sn_bias")
    return bias


def noise_bias(q_t: torch.Tensor, z_t: torch.Tensor, beta: float, nk: int) -> torch.Tensor:
    """bias[row] = -beta * mean(top-nk inner products of q_t[row] against the noise bank z_t)."""
    ix = DeviceIndex(z_t.shape[1], METRIC_INNER_PRODUCT, z_t.device)
    ix.add(z_t)
    D, _ = ix.search(q_t, nk)
    return bias_from_topk(D, beta, nk)


def sn_index(x: torch.Tensor, drop_dim, l2_normalize: bool, fill: float) -> "DeviceIndex":
    """Inner-product index over ``sn_transform(x, drop_dim, l2_normalize, fill)`` built in one pass over ``x``."""
    x = _f32_cuda(x, "sn_index")
    drops = isinstance(drop_dim, torch.Tensor) or 0 <= int(drop_dim) < x.shape[1]
    ix = DeviceIndex(x.shape[1] + (0 if drops else 1), METRIC_INNER_PRODUCT, x.device)
    ix.add_sn(x, drop_dim, l2_normalize, fill=fill)
    return ix


def score_normalized_search(q: torch.Tensor, r: torch.Tensor, z: torch.Tensor, k: int, l2_normalize=True, replace_dim=True,
                            beta=1.0, nk=1) -> Tuple[torch.Tensor, torch.Tensor]:
    """``score_normalize`` followed by the top-k search of the normalised queries against the normalised references
    (score_normalization.py:33-104 + vsc/index.py:174), everything resident on the device: the two banks go through
    ``DeviceIndex.add_sn`` (no transformed copy), the queries through ``sn_transform``.  Same D, I as
    ``score_normalize_tensors`` + ``DeviceIndex.add`` + ``search``."""
    lvd = low_var_dim_device(z) if replace_dim else -1
    q_0 = sn_transform(q, lvd, l2_normalize, fill=0.0)
    Dz, _ = sn_index(z, lvd, l2_normalize, 0.0).search(q_0, nk)
    q_0[:, -1] = bias_from_topk(Dz, beta, nk)      # = sn_transform(q, lvd, l2_normalize, bias=...): only the last column differs
    return sn_index(r, lvd, l2_normalize, 1.0).search(q_0, k)


def score_normalize_tensors(q: torch.Tensor, r: Optional[torch.Tensor], z: torch.Tensor, l2_normalize=True,
                            replace_dim=True, beta=1.0, nk=1, low_var_dim_: Optional[int] = None,
                            gated_rows: Optional[torch.Tensor] = None):
    """Array form of score_normalize: -> (q' , r' or None, low_var_dim)."""
    lvd = -1
    if replace_dim:
        lvd = low_var_dim_device(z) if low_var_dim_ is None else int(low_var_dim_)    # stays on the device: no sync
    # the noise search ignores the appended column (0 on both sides); the noise bank goes straight into its index
    q_0 = sn_transform(q, lvd, l2_normalize, fill=0.0)
    Dz, _ = sn_index(z, lvd, l2_normalize, 0.0).search(q_0, nk)
    bias = bias_from_topk(Dz, beta, nk)
    if gated_rows is not None:
        bias = torch.where(gated_rows.to(bias.device), torch.full_like(bias, -100.0), bias)
    q_t = sn_transform(q, lvd, l2_normalize, bias=bias)
    r_t = sn_transform(r, lvd, l2_normalize, fill=1.0) if r is not None else None
    return q_t, r_t, lvd


def score_normalize_v2_tensors(q: torch.Tensor, r: torch.Tensor, z: torch.Tensor, l2_normalize: bool = True,
                               beta: float = 0.35, nk: int = 10) -> Tuple[torch.Tensor, torch.Tensor]:
    """Array form of the matching track's ``score_normalizev2`` (M/vsc/baseline/score_normalization.py:115-156): every
    query / reference row loses beta x the mean of its nk nearest (cosine) noise rows, then is L2-normalised."""
    q, r, z = (_f32_cuda(t, n) for t, n in ((q, "queries"), (r, "refs"), (z, "noise")))
    d = z.shape[1]
    unit = (lambda t: sn_transform(t, -1, True)[:, :d].contiguous()) if l2_normalize else (lambda t: t)
    ix = DeviceIndex(d, METRIC_INNER_PRODUCT, z.device)
    ix.add(unit(z))
    outs = []
    for x in (q, r):
        _, ids = ix.search(unit(x), nk)
        out = torch.empty_like(x)
        _lib.check(_lib.lib().vscb200_sn2_adapt(_p(x), _p(z), _p(ids), x.shape[0], d, nk, float(beta), 1 if l2_normalize else 0,
                                                _p(out), _stream(x.device)), "sn2_adapt")
        outs.append(out)
    return outs[0], outs[1]


# ------------------------------------------------------------------------------------------------
# the reference's function signatures (lists of VideoFeature-like objects, numpy features)
# ------------------------------------------------------------------------------------------------
_PINNED = {}      # staging buffers (page-locked host memory), reused across calls: name -> uint8 tensor
_PINNED_EV = {}   # name -> event recorded after the last upload from that buffer


def _pinned(name: str, nbytes: int) -> torch.Tensor:
    buf = _PINNED.get(name)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        _PINNED[name] = buf
    return buf


_COPY_THREADS = 4
_copy_pool = None


def _parallel(jobs):
    """Run row-range copy jobs on a few threads: numpy releases the GIL inside large contiguous copies, and one thread
    moves only ~8 GB/s of the ~40 GB/s the host memory system gives."""
    global _copy_pool
    if len(jobs) <= 1:
        for j in jobs:
            j()
        return
    if _copy_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _copy_pool = ThreadPoolExecutor(max_workers=_COPY_THREADS)
    list(_copy_pool.map(lambda j: j(), jobs))


def _cat(features: Sequence, device, name: str = "cat") -> torch.Tensor:
    """Concatenate the per-video feature arrays straight into a page-locked staging buffer (one pass over the
    host data, split over a few threads) and upload it with one asynchronous copy."""
    lens = [int(f.feature.shape[0]) for f in features]
    rows = sum(lens)
    d = int(features[0].feature.shape[1])
    stage = _pinned(name, rows * d * 4)[: rows * d * 4].view(torch.float32).view(rows, d)
    ev = _PINNED_EV.get(name)
    if ev is not None:
        ev.synchronize()                                  # the previous upload from THIS buffer has drained
    out = stage.numpy()
    offs = np.concatenate([[0], np.cumsum(lens)])
    n = len(features)
    per = max(1, -(-n // _COPY_THREADS)) if rows * d * 4 >= (8 << 20) else n

    def job(a, b):
        def run():
            for i in range(a, b):
                out[offs[i]:offs[i + 1]] = features[i].feature
        return run
    _parallel([job(a, min(n, a + per)) for a in range(0, n, per)])
    dev_t = stage.to(device, non_blocking=True)           # overlaps with the host packing of the next list
    ev = _PINNED_EV.get(name) or torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    _PINNED_EV[name] = ev
    return dev_t


def _to_host(t: torch.Tensor, name: str) -> np.ndarray:
    """Device -> page-locked host buffer -> numpy array owned by the caller."""
    stage = _pinned(name, t.numel() * 4)[: t.numel() * 4].view(torch.float32).view(t.shape)
    stage.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    src = stage.numpy()
    dst = np.empty_like(src)
    rows = src.shape[0] if src.ndim else 0
    if src.nbytes < (8 << 20) or rows < _COPY_THREADS:
        dst[...] = src
        return dst
    per = -(-rows // _COPY_THREADS)

    def job(a, b):
        def run():
            dst[a:b] = src[a:b]
        return run
    _parallel([job(a, min(rows, a + per)) for a in range(0, rows, per)])
    return dst


def _split(features: Sequence, arr: np.ndarray) -> List:
    out, i = [], 0
    for f in features:
        n = f.feature.shape[0]
        out.append(dataclasses.replace(f, feature=arr[i:i + n]))
        i += n
    return out


def _check_disjoint(refs, score_norm_refs):
    if {f.video_id for f in refs}.intersection({f.video_id for f in score_norm_refs}):
        raise Exception(
            "Normalizing on the dataset we're evaluating on is against VSC rules. "
            "An independent dataset is needed."
        )


def score_normalize(queries, refs, score_norm_refs, l2_normalize: bool = True, replace_dim: bool = True,
                    beta: float = 1.0, nk: int = 1, device="cuda"):
    """score_normalization.py:33-104."""
    _check_disjoint(refs, score_norm_refs)
    dev = torch.device(device)
    q_t, r_t, _ = score_normalize_tensors(_cat(queries, dev, "q"), _cat(refs, dev, "r"), _cat(score_norm_refs, dev, "z"),
                                          l2_normalize, replace_dim, beta, nk)
    return _split(queries, _to_host(q_t, "qo")), _split(refs, _to_host(r_t, "ro"))


def query_score_normalize(queries, score_norm_refs, video_scores=None, score_threshold: float = 0.001,
                          low_var_dim: int = 0, l2_normalize: bool = True, replace_dim: bool = True,
                          beta: float = 1.0, nk: int = 1, device="cuda"):
    """Descriptor track: score_normalization.py:107-148 ``(queries, score_norm_refs, video_scores, score_threshold,
    low_var_dim, ...)`` with the video-score gate (bias = -100 below the threshold, :142-143).
    Matching track: M/infer/vsc/baseline/score_normalization.py:107-145 ``(queries, score_norm_refs, low_var_dim, ...)``
    -- no gate; recognised by the third argument being an integer (infer_matching.py:213)."""
    if video_scores is not None and not isinstance(video_scores, dict):
        low_var_dim, video_scores = int(video_scores), None
    dev = torch.device(device)
    gated = None
    if video_scores is not None:
        gated = torch.from_numpy(np.concatenate([np.full(q.feature.shape[0], video_scores[q.video_id] < score_threshold)
                                                 for q in queries]))
    q_t, _, _ = score_normalize_tensors(_cat(queries, dev, "q"), None, _cat(score_norm_refs, dev, "z"), l2_normalize,
                                        replace_dim, beta, nk, low_var_dim_=low_var_dim, gated_rows=gated)
    return _split(queries, _to_host(q_t, "qo"))


def ref_score_normalize(refs, score_norm_refs, l2_normalize: bool = True, replace_dim: bool = True,
                        beta: float = 1.0, nk: int = 1, device="cuda"):
    """score_normalization.py:150-192."""
    _check_disjoint(refs, score_norm_refs)
    dev = torch.device(device)
    lvd = -1
    if replace_dim:
        lvd = low_var_dim(_cat(score_norm_refs, dev, "z"))
    r_t = sn_transform(_cat(refs, dev, "r"), lvd, l2_normalize, fill=1.0)
    return _split(refs, _to_host(r_t, "ro"))


def score_normalizev2(queries, refs, score_norm_refs, l2_normalize: bool = True, replace_dim: bool = True,
                      beta: float = 0.35, nk: int = 10, device="cuda"):
    """VSC22-Matching-Track-1st/vsc/baseline/score_normalization.py:115-156 (same signature; ``replace_dim`` is accepted
    and unused there too)."""
    _check_disjoint(refs, score_norm_refs)
    dev = torch.device(device)
    q_t, r_t = score_normalize_v2_tensors(_cat(queries, dev, "q"), _cat(refs, dev, "r"), _cat(score_norm_refs, dev, "z"),
                                          l2_normalize, beta, nk)
    return _split(queries, _to_host(q_t, "qo")), _split(refs, _to_host(r_t, "ro"))
