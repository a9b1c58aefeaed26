"""TEST INFRASTRUCTURE — import the reference's own modules.

``/root/reference`` exists only in the build container; there it is used by ``tests/golden/gen_golden.py`` to produce the
committed fixtures and by the ``not gpu`` tests that re-check the oracle against the live reference (they skip otherwise).
On the GPU box the only copy is ``oracle/_ref`` (unmodified files placed there by ``oracle/vendor_ref.py``, git-ignored), and
the only caller is the CPU arm of ``bench.py`` (``oracle/cpu_arm.py``, a subprocess without a visible GPU): nothing under
``tests -m gpu`` or ``smoke()`` reads either location.  SURVEY.md Appendix A.
"""
from __future__ import annotations

import os
import sys
import types

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("VDN_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/video_depth_anything") else _VENDORED)


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "video_depth_anything"))


def _install_shims():
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "easydict" not in sys.modules:
        class EasyDict(dict):  # stand-in for the missing `easydict` (used only at dpt_temporal.py:35-40)
            def __init__(self, d=None, **kw):
                super().__init__()
                self.update(dict(d or {}, **kw))
            __getattr__ = dict.__getitem__
            __setattr__ = dict.__setitem__
        sys.modules["easydict"] = types.SimpleNamespace(EasyDict=EasyDict)


def load_vda(encoder: str, state_dict, **kw):
    """Reference VideoDepthAnything (video_depth_anything/video_depth.py:35) with our recipe weights, strict.
    ``kw``: use_clstoken / pe constructor switches."""
    from .init_recipe import ENCODERS
    _install_shims()
    from video_depth_anything.video_depth import VideoDepthAnything
    cfg = ENCODERS[encoder]
    m = VideoDepthAnything(encoder=encoder, features=cfg["features"], out_channels=cfg["out_channels"], **kw).eval()
    m.load_state_dict(state_dict, strict=True)
    return m


def load_v5(encoder: str, state_dict):
    """Reference models/video_depth_model_v5.py:128 VideoDepthAnything with our recipe weights, strict."""
    from .init_recipe import ENCODERS
    _install_shims()
    from models.video_depth_model_v5 import VideoDepthAnything as V5
    cfg = ENCODERS[encoder]
    m = V5(encoder=encoder, features=cfg["features"], out_channels=cfg["out_channels"]).eval()
    m.load_state_dict(state_dict, strict=True)
    return m


def load_v4(encoder: str, state_dict):
    """Reference models/video_depth_model_v4.py:88 VideoDepthAnything (v5 without the 224x224 resize; same keys), strict."""
    from .init_recipe import ENCODERS
    _install_shims()
    from models.video_depth_model_v4 import VideoDepthAnything as V4
    cfg = ENCODERS[encoder]
    m = V4(encoder=encoder, features=cfg["features"], out_channels=cfg["out_channels"]).eval()
    m.load_state_dict(state_dict, strict=True)
    return m


def load_da2(encoder: str, state_dict, **kw):
    """Reference depth_anything_v2/depth_anything_v2.py:12 DepthAnythingV2 (the fork with the SAM2-style memory block), strict.
    Run it on CPU tensors via forward(): image2tensor / RoPEAttention move data to CUDA when a GPU is visible."""
    from .init_recipe import ENCODERS
    _install_shims()
    from depth_anything_v2.depth_anything_v2 import DepthAnythingV2
    cfg = ENCODERS[encoder]
    m = DepthAnythingV2(encoder=encoder, features=cfg["features"], out_channels=cfg["out_channels"], **kw).eval()
    m.load_state_dict(state_dict, strict=True)
    return m


def load_vda_stream(encoder: str, state_dict, **kw):
    """Reference streaming model (video_depth_anything/video_depth_stream.py:32), same weights / keys as load_vda, strict."""
    from .init_recipe import ENCODERS
    _install_shims()
    from video_depth_anything.video_depth_stream import VideoDepthAnything as VDAStream
    cfg = ENCODERS[encoder]
    m = VDAStream(encoder=encoder, features=cfg["features"], out_channels=cfg["out_channels"], **kw).eval()
    m.load_state_dict(state_dict, strict=True)
    return m
