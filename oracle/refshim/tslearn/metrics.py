def dtw_path_from_metric(*a, **k):  # vcsl/vta.py DTW branch only; the eval path uses "TN"
    raise RuntimeError("tslearn shim: DTW alignment is not available")
