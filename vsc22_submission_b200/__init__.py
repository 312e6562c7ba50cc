"""vsc22_submission_b200 -- B200-native (sm_100a) implementation of the VSC22 descriptor-encode +
similarity-search hot path, behind the reference's own seams:

* ``faiss_compat``  -- drop-in for the ``faiss`` subset the reference imports (seam B)
* ``encoder``       -- ``nn.Module`` standing in for the ``torch.jit.load``-ed frame encoder (seam A)
* ``_lib``          -- ctypes binding of the C ABI (include/vscb200.h); no CPU fallback exists

Importing this package performs no CUDA call and does not import torch.
"""
__all__ = ["faiss_compat", "encoder", "_lib"]
__version__ = "0.1.0"
