"""video_depth_normal_v2_b200 — B200-native (sm_100a) inference path for the depth/normal network of
injun-baek/Video-Depth-Normal-v2: DINOv2 ViT encoder -> DPT head -> temporal motion modules -> depth (+normals).

Host code is Python; every device kernel lives in the in-tree C-ABI library ``libvdn_b200.so``
(``include/vdn_b200.h``).  There is no CPU fallback: using the ops without the built library raises.
"""
from . import ops  # noqa: F401
from . import torch_ops  # noqa: F401  (registers torch.ops.vdn.*)

_LAZY = {"VideoDepthAnything": "models", "VideoDepthRefinerV5": "models", "VideoDepthRefinerV4": "models", "ENCODER_CONFIGS": "models", "DepthAnythingV2": "da2"}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(f".{_LAZY[name]}", __name__), name)
    raise AttributeError(name)
