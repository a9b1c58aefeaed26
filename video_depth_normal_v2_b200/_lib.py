"""ctypes binding of libvdn_b200.so (include/vdn_b200.h).  Fails loudly when the library is missing — no fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VDN_LIB_PATH") or os.path.join(HERE, "libvdn_b200.so")  # VDN_LIB_PATH: debug builds (scripts/build_fa_timeline.sh)

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmDesc(C.Structure):
    """Mirror of ``vdn_gemm_desc`` (include/vdn_b200.h)."""
    _fields_ = [
        ("a", c_void_p), ("w", c_void_p),
        ("M", c_int64), ("N", c_int64), ("K", c_int64),
        ("lda", c_int64), ("ldw", c_int64),
        ("conv", c_int), ("B", c_int), ("H", c_int), ("W", c_int),
        ("bias", c_void_p), ("gamma", c_void_p),
        ("res", c_void_p), ("res_f32", c_int), ("ld_res", c_int64),
        ("res2", c_void_p), ("ld_res2", c_int64),
        ("out", c_void_p), ("out_f32", c_int), ("ldc", c_int64),
        ("out2", c_void_p), ("out2_relu", c_int), ("ld_out2", c_int64),
        ("act", c_int), ("geglu", c_int), ("row_map", c_int),
        ("rm0", c_int), ("rm1", c_int), ("rm2", c_int), ("rm3", c_int),
        ("head_w", c_void_p), ("head_b", c_float),
        ("qkv_split", c_int), ("qkv_tokens_out", c_int), ("qkv_token_offset", c_int),
    ]


# every symbol include/vdn_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "vdn_last_error": (C.c_char_p, []),
    "vdn_version": (c_int, []),
    "vdn_set_operand_format": (c_int, [c_int]),
    "vdn_get_operand_format": (c_int, []),
    "vdn_launch_count": (c_int64, []),
    "vdn_reset_launch_count": (None, []),
    "vdn_add_launch_count": (None, [c_int64]),
    "vdn_gemm": (c_int, [C.POINTER(GemmDesc), c_void_p]),
    "vdn_flash_attn": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdn_flash_attn_ex": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vdn_stream_temporal_attn": (c_int, [C.POINTER(c_void_p), c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdn_stream_temporal_attn_ring": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdn_ring_store": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "vdn_temporal_attn_tc": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "vdn_temporal_attn": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vdn_layernorm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_int, c_int, c_void_p, c_int, c_void_p]),
    "vdn_groupnorm_stats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vdn_groupnorm_apply_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vdn_groupnorm_to_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vdn_patch_im2col": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vdn_write_cls": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdn_im2col_3x3_s2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vdn_conv_tail": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdn_conv_tail_up": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdn_bilinear_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vdn_bilinear_nhwc2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vdn_bilinear_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vdn_relu16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "vdn_cast_f32_to_16": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "vdn_preprocess_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, C.POINTER(c_float), C.POINTER(c_float), c_void_p]),
    "vdn_lsq_sums": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "vdn_lsq_solve": (c_int, [c_void_p, c_void_p, c_void_p]),
    "vdn_affine_clamp": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "vdn_crossfade": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_void_p]),
    "vdn_window_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p]),
    "vdn_window_keys": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "vdn_sobel_normals": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vdn_frame_median_scale": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_float, c_float, c_float, c_void_p]),
    "vdn_v5_net_input": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "vdn_v5_residual": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p]),
    "vdn_rope2d": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int, c_int64, c_int64, c_void_p]),
    "vdn_rope_chunks": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int, c_void_p]),
    "vdn_readout_concat": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "vdn_add_rowvec": (c_int, [c_void_p, c_int, c_void_p, c_float, c_void_p, c_int64, c_int, c_void_p]),
    "vdn_add_rowscalar": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vdn_dwconv7_ln": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vdn_mask_down1": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdn_mask_down2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library.  Raises RuntimeError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `python video_depth_normal_v2_b200/build.py`). There is no CPU / PyTorch fallback for this path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
