"""CPU-side checks of the C-ABI boundary: the library builds/loads and exports every symbol that
include/vscb200.h declares; argument validation works without a GPU (no compute calls here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vsc22_submission_b200 import _lib, build
    build.build(verbose=False)
    return _lib.lib()


def test_header_symbols_are_exported_and_bound(lib):
    from vsc22_submission_b200 import _lib
    header = open(os.path.join(REPO, "include", "vscb200.h")).read()
    declared = set(re.findall(r"\b(vscb200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/vscb200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_no_torch_import_and_no_cuda_at_load():
    import subprocess
    import sys
    code = ("import sys; import vsc22_submission_b200; from vsc22_submission_b200 import _lib, faiss_compat; "
            "_lib.lib(); assert 'torch' not in sys.modules; print('ok')")
    r = subprocess.run([sys.executable, "-c", code], cwd=REPO, capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_argument_validation_without_gpu(lib):
    from vsc22_submission_b200 import _lib
    h = C.c_void_p()
    assert lib.vscb200_index_create(0, 0, C.byref(h)) == 1          # d must be positive
    assert b"d must be positive" in lib.vscb200_last_error()
    assert lib.vscb200_index_create(8, 7, C.byref(h)) == 1          # unknown metric
    assert lib.vscb200_index_create(8, 0, C.byref(h)) == 0
    assert lib.vscb200_index_dim(h) == 8 and lib.vscb200_index_metric(h) == 0 and lib.vscb200_index_ntotal(h) == 0
    x = np.zeros((5, 8), np.float32)
    import torch
    rc = lib.vscb200_index_add_host(h, x.ctypes.data_as(C.c_void_p), 5)            # rows go to a page-locked staging slot
    if torch.cuda.is_available():
        assert rc == 0 and lib.vscb200_index_ntotal(h) == 5
    else:                                           # no device: the product path fails loudly, it has no host fallback
        assert rc == 2 and b"cudaHostAlloc" in lib.vscb200_last_error() and lib.vscb200_index_ntotal(h) == 0
    assert lib.vscb200_index_add_host(h, None, 5) == 1                              # null rows
    assert lib.vscb200_index_search(h, None, 3, 0, None, None, None) == 1           # null pointers / bad k
    assert lib.vscb200_index_reset(h) == 0 and lib.vscb200_index_ntotal(h) == 0
    lib.vscb200_index_destroy(h)
    spec = _lib.VitSpecC(img=224, patch=16, width=700, layers=12, heads=12)
    v = C.c_void_p()
    assert lib.vscb200_vit_create(C.byref(spec), 8, C.byref(v)) == 1                # width != heads*64
    assert b"heads*64" in lib.vscb200_last_error()


def test_faiss_compat_shape_errors(lib):
    from vsc22_submission_b200 import faiss_compat as faiss
    ix = faiss.index_factory(16, "Flat", faiss.METRIC_INNER_PRODUCT)
    with pytest.raises(AssertionError):
        ix.add(np.zeros((3, 15), np.float32))
    with pytest.raises(RuntimeError):
        faiss.index_factory(16, "IVF16,Flat", faiss.METRIC_L2)
    assert ix.ntotal == 0 and ix.d == 16 and ix.metric_type == faiss.METRIC_INNER_PRODUCT


def test_encoder_module_contract_cpu():
    import torch
    from oracle import vit_ref
    from vsc22_submission_b200.encoder import (B200ViTEncoder, VitSpec, encoder_from_state_dict, param_names)
    spec = VitSpec(img=64, patch=16, width=128, layers=2, heads=2, tail="gem_linear", out_dim=64)
    w = vit_ref.init_weights(vit_ref.VitSpec(img=64, patch=16, width=128, layers=2, heads=2, tail="gem_linear", out_dim=64))
    enc = B200ViTEncoder(spec, w, max_frames=4).eval()
    assert any(p.requires_grad for p in enc.parameters())          # DDP needs one (extract_ref_feats.py:26)
    assert len(enc.state_dict()) == len(param_names(spec)) + 1
    with pytest.raises(RuntimeError):                              # no CPU fallback
        enc(torch.zeros(1, 3, 64, 64))
    with pytest.raises(KeyError):
        B200ViTEncoder(spec, {k: v for k, v in w.items() if k != "cls"})
    assert abs(spec.flops_per_frame() - vit_ref.VitSpec(img=64, patch=16, width=128, layers=2, heads=2, tail="gem_linear", out_dim=64).flops_per_frame()) < 1
    with pytest.raises(RuntimeError):
        encoder_from_state_dict({"foo": torch.zeros(1)})
