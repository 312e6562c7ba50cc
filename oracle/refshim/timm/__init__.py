"""Stand-in for the `timm` package (pinned timm==0.6.12 in the reference's dockerfile:33; absent here).

TEST INFRASTRUCTURE.  `create_model` builds the VisionTransformer restated in timm/models/vision_transformer.py so that
the reference's own SSCDModel / Model / GlobalGeMPool2d classes (D/train/train_v68/.../backbones/sscd.py) can be
instantiated, run and traced exactly as D/train/train_v68/torch2scripts.py does.  The size of the ViT is taken from
TIMM_SHIM_VIT (a dict set by the caller) because the reference hard-codes only the NAME of the architecture."""
from .models.vision_transformer import VisionTransformer

TIMM_SHIM_VIT = dict(img_size=384, patch_size=32, embed_dim=768, depth=12, num_heads=12)


def list_models(pretrained=False):
    return ["vit_base_patch32_384"]


def create_model(name, pretrained=False, global_pool="", num_classes=0, **kwargs):
    assert name in list_models() and not pretrained and global_pool == "" and num_classes == 0, (name, pretrained)
    return VisionTransformer(**TIMM_SHIM_VIT)
