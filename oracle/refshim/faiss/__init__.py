from oracle.faiss_np import *  # noqa: F401,F403
from oracle.faiss_np import (METRIC_INNER_PRODUCT, METRIC_L2, IndexFlat, IndexFlatIP, IndexFlatL2,  # noqa: F401
                             ResultHeap, index_factory, get_num_gpus, index_cpu_to_all_gpus,
                             GpuMultipleClonerOptions, knn)
