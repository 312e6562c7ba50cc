"""Drop-in for the subset of ``faiss`` the reference's hot path (B) imports.

Every symbol the reference uses is here with faiss's names and argument meaning, backed by the
hand-written sm_100a kernels of libvscb200.so through its C ABI (include/vscb200.h):

===========================  =====================================================================
symbol                       reference call sites
===========================  =====================================================================
METRIC_INNER_PRODUCT/L2      vsc/index.py:79, tests/test_index.py:42
index_factory(d,"Flat",m)    vsc/index.py:81; M/infer/infer_matching.py:219
IndexFlat / IP / L2          vsc/exhaustive_search.py:26,70,102
.add(x)                      vsc/index.py:94
.search(x,k) -> (D,I)        vsc/index.py:174; score_normalization.py:95,141; exhaustive_search.py:62
.range_search(x,t)           vsc/exhaustive_search.py:74,126,246; infer_matching.py:235
.reset() .ntotal .d          vsc/exhaustive_search.py:39,60
.metric_type                 vsc/index.py:145; exhaustive_search.py:63
get_num_gpus()               vsc/index.py:169; exhaustive_search.py:28,229
index_cpu_to_all_gpus        vsc/index.py:171; exhaustive_search.py:234; score_normalization.py:89
GpuMultipleClonerOptions     vsc/exhaustive_search.py:232-233
ResultHeap, knn              vsc/exhaustive_search.py:24 (knn_ground_truth only)
===========================  =====================================================================

Inputs and outputs are numpy host arrays exactly as with faiss; the host<->device copies happen
inside the C ABI's ``_host`` entry points.  There is no CPU implementation behind this module: if
libvscb200.so or a CUDA device is missing, calls raise.

Activation: put ``<repo>/shims`` on PYTHONPATH (it holds a ``faiss`` package re-exporting this
module); the reference tree needs no edit.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib

METRIC_INNER_PRODUCT = _lib.METRIC_INNER_PRODUCT
METRIC_L2 = _lib.METRIC_L2
_FLT_MAX = np.float32(3.4028234663852886e38)


def _f32_2d(x, d, what):
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim != 2 or x.shape[1] != d:
        # faiss: `assert d == self.d` in the SWIG wrapper
        raise AssertionError(f"{what}: expected a [n, {d}] float32 array, got shape {x.shape}")
    return x


class _Handle:
    """Ref-counted owner of a vscb200_index (shared by CPU-side index and its 'GPU clones')."""

    def __init__(self, d, metric):
        self.ptr = C.c_void_p()
        _lib.check(_lib.lib().vscb200_index_create(int(d), int(metric), C.byref(self.ptr)), "index_create")

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().vscb200_index_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class IndexFlat:
    """faiss.IndexFlat: exact brute-force inner-product / squared-L2 index, resident in HBM."""

    def __init__(self, d, metric=METRIC_L2):
        self.d = int(d)
        self.metric_type = int(metric)
        self.is_trained = True
        self._h = _Handle(self.d, self.metric_type)

    # -- faiss attributes
    @property
    def ntotal(self):
        return int(_lib.lib().vscb200_index_ntotal(self._h.ptr))

    def add(self, x):
        x = _f32_2d(x, self.d, "add")
        _lib.check(_lib.lib().vscb200_index_add_host(self._h.ptr, x.ctypes.data_as(C.c_void_p), x.shape[0]), "index.add")

    def reset(self):
        _lib.check(_lib.lib().vscb200_index_reset(self._h.ptr), "index.reset")

    def search(self, x, k):
        x = _f32_2d(x, self.d, "search")
        k = int(k)
        if k <= 0:
            raise AssertionError("search: k must be positive")
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        if nq:
            _lib.check(_lib.lib().vscb200_index_search_host(self._h.ptr, x.ctypes.data_as(C.c_void_p), nq, k,
                                                            D.ctypes.data_as(C.c_void_p),
                                                            I.ctypes.data_as(C.c_void_p)), "index.search")
        return D, I

    def range_search(self, x, thresh):
        x = _f32_2d(x, self.d, "range_search")
        nq = x.shape[0]
        lims = np.zeros(nq + 1, dtype=np.uint64)
        Dp, Ip = C.c_void_p(), C.c_void_p()
        _lib.check(_lib.lib().vscb200_index_range_search_host(self._h.ptr, x.ctypes.data_as(C.c_void_p), nq,
                                                              float(thresh), lims.ctypes.data_as(C.c_void_p),
                                                              C.byref(Dp), C.byref(Ip)), "index.range_search")
        try:
            total = int(lims[-1])
            if total:
                D = np.ctypeslib.as_array(C.cast(Dp, C.POINTER(C.c_float)), shape=(total,)).copy()
                I = np.ctypeslib.as_array(C.cast(Ip, C.POINTER(C.c_int64)), shape=(total,)).copy()
            else:
                D, I = np.zeros(0, np.float32), np.zeros(0, np.int64)
        finally:
            _lib.lib().vscb200_free(Dp)
            _lib.lib().vscb200_free(Ip)
        return lims, D, I


class IndexFlatIP(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_INNER_PRODUCT)


class IndexFlatL2(IndexFlat):
    def __init__(self, d):
        super().__init__(d, METRIC_L2)


def index_factory(d, description="Flat", metric=METRIC_L2):
    if description != "Flat":
        raise RuntimeError(f"index_factory: only the 'Flat' codec of the reference is implemented, got {description!r}")
    return IndexFlat(d, metric)


def get_num_gpus():
    return int(_lib.lib().vscb200_device_count())


class GpuMultipleClonerOptions:
    def __init__(self):
        self.shard = False
        self.useFloat16 = False


class GpuIndexView(IndexFlat):
    """What index_cpu_to_all_gpus returns on a one-GPU box.  The bank already lives in HBM, so the 'clone' is a view
    sharing the device copy (faiss would copy it, exhaustive_search.py:232-234)."""

    def __init__(self, base: IndexFlat):  # noqa: super().__init__ deliberately not called
        self.d, self.metric_type, self.is_trained = base.d, base.metric_type, True
        self._h = base._h


class MultiGpuIndex:
    """What index_cpu_to_all_gpus returns when several GPUs are visible: ONE process driving all of them, like faiss
    (vsc/index.py:171, vsc/exhaustive_search.py:229-234, score_normalization.py:88-89, infer_matching.py:225-226).

    ``co.shard = True``   the bank rows are split into contiguous shards, one per GPU, searched with global row ids; the
                          per-GPU ``[nq, k]`` results travel to GPU 0 as packed 64-bit keys (peer copies over NVLink) and
                          are merged k-way by a device kernel (csrc/merge.cu) -- faiss ``IndexShards``.
    ``co.shard = False``  (faiss's and the reference's default) every GPU holds the whole bank and takes a contiguous
                          slice of the query rows -- faiss ``IndexReplicas``; no merge.

    Either way the results are those of the single index, bit for bit: every reported score comes from the same exact
    fp32 summation (csrc/exact.cuh).  The snapshot is taken when the function is called (the reference never adds to
    the GPU index afterwards); ``add`` therefore raises."""

    def __init__(self, base: IndexFlat, devices, shard: bool):
        import torch
        from .. import search, sharding
        self._torch, self._search = torch, search
        self.d, self.metric_type, self.is_trained = base.d, base.metric_type, True
        self.shard = bool(shard)
        self.devices = [torch.device("cuda", int(i)) for i in devices]
        n = base.ntotal
        self._ntotal = n
        self.subs, self.spans = [], []
        for j, dev in enumerate(self.devices):
            a, b = sharding.shard_range(n, len(self.devices), j) if self.shard else (0, n)
            sub = search.DeviceIndex(self.d, self.metric_type, dev)
            if b > a:
                rows = torch.empty((b - a, self.d), dtype=torch.float32, device=dev)
                torch.cuda.current_stream(dev).synchronize()        # `rows` is allocated before the foreign-stream copy lands
                # the copy engine moves the rows from the base index's GPU to GPU j (unified addressing)
                _lib.check(_lib.lib().vscb200_index_reconstruct_n(base._h.ptr, a, b - a, C.c_void_p(rows.data_ptr())),
                           "index.reconstruct_n")
                sub.add(rows)
            sub.set_id_offset(a)
            self.subs.append(sub)
            self.spans.append((a, b))

    @property
    def ntotal(self):
        return self._ntotal

    def add(self, x):
        raise RuntimeError("MultiGpuIndex is a snapshot of the CPU-side index: add to that index and clone again")

    def reset(self):
        for sub in self.subs:
            sub.reset()
        self._ntotal = 0

    def search(self, x, k):
        torch = self._torch
        x = _f32_2d(x, self.d, "search")
        k = int(k)
        if k <= 0:
            raise AssertionError("search: k must be positive")
        nq = x.shape[0]
        keep_max = self.metric_type == METRIC_INNER_PRODUCT
        if nq == 0:
            return np.empty((0, k), np.float32), np.empty((0, k), np.int64)
        xt = torch.from_numpy(x)
        if not self.shard:
            # replicas: GPU j answers a contiguous slice of the queries; all slices are enqueued before the first result
            # is read back, so the GPUs work concurrently
            from .. import sharding
            parts = []
            for j, (dev, sub) in enumerate(zip(self.devices, self.subs)):
                a, b = sharding.shard_range(nq, len(self.devices), j)
                if b > a:
                    parts.append(sub.search(xt[a:b].to(dev, non_blocking=True), k))
            D = np.concatenate([p[0].cpu().numpy() for p in parts], axis=0)
            I = np.concatenate([p[1].cpu().numpy() for p in parts], axis=0)
            return D, I
        # shards: every GPU searches its rows for all queries; packed partial results -> GPU 0 -> k-way merge kernel
        dev0 = self.devices[0]
        keys = torch.empty((len(self.devices), nq, k), dtype=torch.int64, device=dev0)
        done = []
        for j, (dev, sub) in enumerate(zip(self.devices, self.subs)):
            Dj, Ij = sub.search(xt.to(dev, non_blocking=True), k)
            kj = self._search.pack_topk(Dj, Ij, keep_max)
            with torch.cuda.device(dev):
                keys[j].copy_(kj, non_blocking=True)            # peer copy, ordered on GPU j's stream
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
            done.append(ev)
        with torch.cuda.device(dev0):
            for ev in done:
                torch.cuda.current_stream(dev0).wait_event(ev)
            D, I = self._search.merge_packed_topk(keys, k, keep_max)
            return D.cpu().numpy(), I.cpu().numpy()


def index_cpu_to_all_gpus(index, co=None, ngpu=-1):
    """faiss.index_cpu_to_all_gpus: where the work is spread over the GPUs of the box."""
    avail = get_num_gpus()
    if avail == 0:
        raise RuntimeError("index_cpu_to_all_gpus: no CUDA device")
    n = avail if ngpu is None or ngpu < 0 else min(int(ngpu), avail)
    if n <= 1 or not isinstance(index, IndexFlat) or isinstance(index, MultiGpuIndex):
        return GpuIndexView(index) if isinstance(index, IndexFlat) else index
    return MultiGpuIndex(index, range(n), bool(co.shard) if co is not None else False)


def index_cpu_to_gpu(res, device, index, options=None):
    return GpuIndexView(index)


class ResultHeap:
    """faiss.ResultHeap (exhaustive_search.knn_ground_truth :24-44): running best-k merge of result
    blocks.  Host-side bookkeeping over (D, I) blocks that the GPU search produced."""

    def __init__(self, nq, k, keep_max=False):
        self.nq, self.k, self.keep_max = int(nq), int(k), bool(keep_max)
        self.D = np.full((nq, k), -_FLT_MAX if keep_max else _FLT_MAX, dtype=np.float32)
        self.I = np.full((nq, k), -1, dtype=np.int64)

    def add_result(self, D, I):
        D = np.concatenate([self.D, np.asarray(D, np.float32)], axis=1)
        I = np.concatenate([self.I, np.asarray(I, np.int64)], axis=1)
        order = np.argsort(-D if self.keep_max else D, axis=1, kind="stable")[:, :self.k]
        self.D = np.take_along_axis(D, order, axis=1)
        self.I = np.take_along_axis(I, order, axis=1)

    def finalize(self):
        pass


def knn(xq, xb, k, metric=METRIC_L2):
    index = IndexFlat(np.asarray(xb).shape[1], metric)
    index.add(xb)
    return index.search(xq, k)
