def build_model(*a, **k):  # reference: sscd.py:6,68 (ResNet trunks; not on the ViT path)
    raise RuntimeError("classy_vision shim: ResNet trunks are not available in this environment")
