"""`import faiss` -> vsc22_submission_b200.faiss_compat (put <repo>/shims and <repo> on PYTHONPATH).
The reference tree is used unmodified: vsc/index.py:11, vsc/exhaustive_search.py:6,
vsc/baseline/score_normalization.py:10, M/infer/infer_matching.py:14."""
from vsc22_submission_b200.faiss_compat import *  # noqa: F401,F403
from vsc22_submission_b200.faiss_compat import (METRIC_INNER_PRODUCT, METRIC_L2, IndexFlat, IndexFlatIP,  # noqa: F401
                                                IndexFlatL2, ResultHeap, GpuMultipleClonerOptions, index_factory,
                                                get_num_gpus, index_cpu_to_all_gpus, index_cpu_to_gpu, knn)
