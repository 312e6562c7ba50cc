"""timm.models.layers subset imported by the reference's swinv2.py:13 (timm 0.6.12 semantics)."""
import collections.abc
from itertools import repeat

import torch.nn as nn


class DropPath(nn.Module):
    """Stochastic depth; identity in eval mode (the only mode the inference path uses)."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.training and self.drop_prob > 0.0:
            raise RuntimeError("timm shim: DropPath only supports eval mode")
        return x


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return tuple(repeat(x, 2))


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)
