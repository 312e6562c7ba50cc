def load_checkpoint(*a, **k):  # reference: clip.py:5,136 (only used with pretrained weights)
    raise RuntimeError("mmcv shim: no pretrained checkpoints in this environment")
