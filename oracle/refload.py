"""Import the reference's OWN, unmodified classes from /root/reference (container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  /root/reference does not exist on the GPU box,
so nothing reachable from ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this module;
it is used by tests/golden/make_golden.py (fixture generation) and by the container-only
pin tests (tests/test_oracle_reference_pins.py, skipped when the tree is absent).

No reference source is copied: modules are loaded from where they lie, with the shims in
oracle/refshim/ standing in for absent third-party imports.
"""
from __future__ import annotations

import contextlib
import importlib
import importlib.util
import os
import sys

REF_ROOT = "/root/reference"
D = os.path.join(REF_ROOT, "VSC22-Descriptor-Track-1st")
M = os.path.join(REF_ROOT, "VSC22-Matching-Track-1st")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available() -> bool:
    return os.path.isdir(D)


def _load_file(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def _shimmed(*extra_paths):
    saved = list(sys.path)
    sys.path[:0] = [_SHIM, _REPO, *extra_paths]
    try:
        yield
    finally:
        sys.path[:] = saved


def clip_module():
    """D/train/train_vid_score/video/clip.py (CLIPModel, ResidualAttentionBlock, QuickGELU)."""
    with _shimmed():
        return _load_file("_ref_clip", os.path.join(D, "train/train_vid_score/video/clip.py"))


def swinv2_module():
    """D/train/train_v106/vsc/baseline/model_factory/backbones/swinv2.py, loaded as a file-backed
    module (it holds a @torch.jit.script function) with the BACKBONES registry neutralised."""
    path = os.path.join(D, "train/train_v106/vsc/baseline/model_factory/backbones/swinv2.py")
    import types

    pkg = types.ModuleType("_ref_backbones")
    pkg.__path__ = [os.path.dirname(path)]
    utils = types.ModuleType("_ref_backbones_utils")

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls

    utils.BACKBONES = _Reg()
    sys.modules["_ref_backbones"] = pkg
    sys.modules["_ref_pkg"] = types.ModuleType("_ref_pkg")
    sys.modules["_ref_pkg"].__path__ = []
    sys.modules["_ref_pkg.utils"] = utils
    sys.modules["_ref_pkg.backbones"] = pkg
    with _shimmed():
        spec = importlib.util.spec_from_file_location("_ref_pkg.backbones.swinv2", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["_ref_pkg.backbones.swinv2"] = mod
        spec.loader.exec_module(mod)
    return mod


def _backbones_package(path: str, pkg: str):
    """Load one file of .../model_factory/backbones/ as `<pkg>.backbones.<name>` with the BACKBONES registry of its
    `..utils` neutralised (the registry pulls in mmcv.Registry and every other backbone of the tree)."""
    import types

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls

    utils = types.ModuleType(pkg + ".utils")
    utils.BACKBONES = _Reg()
    top, sub = types.ModuleType(pkg), types.ModuleType(pkg + ".backbones")
    top.__path__, sub.__path__ = [], [os.path.dirname(path)]
    sys.modules[pkg], sys.modules[pkg + ".utils"], sys.modules[pkg + ".backbones"] = top, utils, sub
    name = pkg + ".backbones." + os.path.splitext(os.path.basename(path))[0]
    with _shimmed():
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    return mod


def sscd_gem_module():
    """D/train/train_v68/vsc/baseline/model_factory/backbones/sscd.py: GlobalGeMPool2d (:11-40), Model (:59-106) and
    SSCDModel (:109-152), unmodified.  Its `timm.create_model` call resolves to the ViT restated in
    oracle/refshim/timm (timm 0.6.12 itself is not in the tree); set `timm.TIMM_SHIM_VIT` for a toy size."""
    return _backbones_package(os.path.join(D, "train/train_v68/vsc/baseline/model_factory/backbones/sscd.py"), "_ref_v68")


def vit_hf_module():
    """D/train/train_v106/vsc/baseline/model_factory/backbones/vit.py: backbone `VIT` (:10-58) = transformers.ViTModel +
    gem + Linear, unmodified (transformers is installed here)."""
    return _backbones_package(os.path.join(D, "train/train_v106/vsc/baseline/model_factory/backbones/vit.py"), "_ref_v106")


def vsc_package(track: str = "D_infer"):
    """The reference's ``vsc`` package (Meta baseline code), with ``faiss`` -> oracle.faiss_np.
    track: 'D_infer' (D/infer/vsc), 'M' (M/vsc), 'M_infer' (M/infer/vsc), 'T106' (train copy
    that owns the unit tests)."""
    root = {
        "D_infer": os.path.join(D, "infer"),
        "M": M,
        "M_infer": os.path.join(M, "infer"),
        "T106": os.path.join(D, "train/train_v106"),
    }[track]
    for name in [n for n in sys.modules if n == "vsc" or n.startswith("vsc.") or n == "vcsl" or
                 n.startswith("vcsl.") or n == "faiss"]:
        del sys.modules[name]
    sys.path[:0] = [_SHIM, _REPO, root]
    importlib.invalidate_caches()
    import vsc  # noqa: F401
    return root


def unload_vsc():
    for name in [n for n in sys.modules if n == "vsc" or n.startswith("vsc.") or n == "vcsl" or
                 n.startswith("vcsl.") or n == "faiss"]:
        del sys.modules[name]
    sys.path[:] = [p for p in sys.path if p != _SHIM and not p.startswith(REF_ROOT)]
