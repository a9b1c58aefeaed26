"""Thin torch-tensor wrappers over the C ABI (include/vdn_b200.h).  PyTorch supplies device memory and streams only;
every op here launches a hand-written sm_100a kernel on the current CUDA stream.  No fallbacks: CPU tensors raise."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import GemmDesc

ACT_NONE, ACT_GELU, ACT_RELU, ACT_SILU = 0, 1, 2, 3
ROWMAP_IDENTITY, ROWMAP_PIXEL_SHUFFLE, ROWMAP_TEMPORAL, ROWMAP_PATCH_TOKENS, ROWMAP_QKV_SPLIT = 0, 1, 2, 3, 4

_FMT_DTYPE = {0: torch.float16, 1: torch.bfloat16}


def lib():
    return _lib.load()


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what}: {lib().vdn_last_error().decode()}")


def set_operand_dtype(dtype: torch.dtype):
    """Library-wide 16-bit operand format: torch.float16 (default) or torch.bfloat16."""
    fmt = {torch.float16: 0, torch.bfloat16: 1}[dtype]
    _check(lib().vdn_set_operand_format(fmt), "vdn_set_operand_format")


def operand_dtype() -> torch.dtype:
    return _FMT_DTYPE[lib().vdn_get_operand_format()]


def launch_count() -> int:
    return int(lib().vdn_launch_count())


def reset_launch_count():
    lib().vdn_reset_launch_count()


class KernelProfiler:
    """Optional CUDA-event timing of every op launch on the current stream (used by bench.py for the live roofline).
    work = algorithmic FLOPs (kind 'tensor') or algorithmic bytes (kind 'hbm') of the launch."""

    def __init__(self, by_shape: bool = False):
        self.by_shape = by_shape
        self.records = []  # (name, kind, work, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, kind, work, e0, e1 in self.records:
            a = agg.setdefault(name, {"kind": kind, "launches": 0, "ms": 0.0, "work": 0.0})
            a["launches"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["work"] += work
        return agg


_profiler: Optional[KernelProfiler] = None


def set_profiler(p: Optional[KernelProfiler]):
    global _profiler
    _profiler = p


def _run(name: str, kind: str, work: float, fn, *args):
    if _profiler is None:
        return fn(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = fn(*args)
    e1.record()
    _profiler.records.append((name, kind, work, e0, e1))
    return rc


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor], dtype=None, name="tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (this path has no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t.data_ptr()


RIDGE_FLOP_PER_BYTE = 219.0  # MEASURED_PEAKS.json: 1417.6 TFLOP/s sustained / 6469.9 GB/s; bench.py overwrites it with the box's own figures


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int, lda: Optional[int] = None, ldw: Optional[int] = None,
         ldc: Optional[int] = None, conv: Optional[Tuple[int, int, int]] = None, bias=None, gamma=None, res=None, ld_res=None, res2=None,
         ld_res2=None, out2=None, out2_relu=False, ld_out2=None, act=ACT_NONE, geglu=False, row_map=ROWMAP_IDENTITY,
         rm: Sequence[int] = (0, 0, 0, 0), head_w=None, head_b: float = 0.0, qkv_split: int = 0, qkv_tokens_out: int = 0, qkv_token_offset: int = 0):
    """out = epilogue(A @ W^T) on tcgen05.  See vdn_gemm_desc in include/vdn_b200.h for the field meanings."""
    od = operand_dtype()
    d = GemmDesc()
    d.a = _ptr(a, od, "a")
    d.w = _ptr(w, od, "w")
    d.M, d.N, d.K = M, N, K
    d.lda = lda if lda is not None else K
    d.ldw = ldw if ldw is not None else w.shape[-1]
    if conv is not None:
        d.conv, (d.B, d.H, d.W) = 1, conv
    d.bias = _ptr(bias, torch.float32, "bias")
    d.gamma = _ptr(gamma, torch.float32, "gamma")
    n_out = N // 2 if geglu else N
    if res is not None:
        d.res = _ptr(res, None, "res")
        d.res_f32 = 1 if res.dtype == torch.float32 else 0
        if not d.res_f32 and res.dtype != od:
            raise RuntimeError("res must be fp32 or the operand dtype")
        d.ld_res = ld_res if ld_res is not None else n_out
    if res2 is not None:
        d.res2 = _ptr(res2, od, "res2")
        d.ld_res2 = ld_res2 if ld_res2 is not None else n_out
    d.out = _ptr(out, None, "out")
    d.out_f32 = 1 if out.dtype == torch.float32 else 0
    if not d.out_f32 and out.dtype != od:
        raise RuntimeError("out must be fp32 or the operand dtype")
    d.ldc = ldc if ldc is not None else n_out
    if out2 is not None:
        d.out2 = _ptr(out2, od, "out2")
        d.out2_relu = 1 if out2_relu else 0
        d.ld_out2 = ld_out2 if ld_out2 is not None else n_out
    d.act, d.geglu, d.row_map = act, 1 if geglu else 0, row_map
    d.rm0, d.rm1, d.rm2, d.rm3 = [int(v) for v in rm]
    d.qkv_split, d.qkv_tokens_out, d.qkv_token_offset = int(qkv_split), int(qkv_tokens_out), int(qkv_token_offset)
    if head_w is not None:
        d.head_w = _ptr(head_w, torch.float32, "head_w")
        d.head_b = float(head_b)
    flops = 2.0 * M * N * K * (9 if conv is not None else 1)
    name, kind, work = ("gemm_conv3x3" if conv is not None else "gemm"), "tensor", flops
    if _profiler is not None:
        # algorithmic bytes (every operand / result once); a GEMM whose arithmetic intensity is below the machine's ridge point
        # (tensor peak / HBM peak, ~219 FLOP/B on B200) is bound by HBM: it is reported against that roofline, as its own family
        nbytes = 2.0 * M * K + 2.0 * N * K * (9 if conv is not None else 1) + M * n_out * (4.0 if d.out_f32 else 2.0)
        if res is not None:
            nbytes += M * n_out * (4.0 if d.res_f32 else 2.0)
        if res2 is not None:
            nbytes += 2.0 * M * n_out
        if out2 is not None:
            nbytes += 2.0 * M * n_out
        if flops / nbytes < RIDGE_FLOP_PER_BYTE:
            name, kind, work = name + "_lowk", "hbm", nbytes
        if _profiler.by_shape:
            name = f"{name}[{M}x{N}x{K}{',f32out' if d.out_f32 else ''}{',act' + str(act) if act else ''}{',geglu' if geglu else ''}{',map' + str(row_map) if row_map else ''}]"
    _check(_run(name, kind, work, lib().vdn_gemm, C.byref(d), _stream()), "vdn_gemm")
    return out


def flash_attn(qk: torch.Tensor, vT: torch.Tensor, out: torch.Tensor, B: int, tokens: int, heads: int):
    od = operand_dtype()
    _check(_run("flash_attn", "tensor", 4.0 * B * heads * tokens * tokens * 64, lib().vdn_flash_attn, _ptr(qk, od, "qk"), qk.shape[-1],
                _ptr(vT, od, "vT"), vT.shape[-1], _ptr(out, od, "out"), B, tokens, heads, _stream()), "vdn_flash_attn")
    return out


def flash_attn_ex(q: torch.Tensor, ld_q: int, q_batch_stride: int, k: torch.Tensor, ld_k: int, k_batch_stride: int, vT: torch.Tensor, ld_vT: int,
                  out: torch.Tensor, B: int, tokens_q: int, tokens_kv: int, heads: int):
    """Cross-attention form: q [B, tokens_q, ld_q], k [B, tokens_kv, ld_k] (element strides), V^T [B*heads, 64, ld_vT]."""
    od = operand_dtype()
    _check(_run("flash_attn", "tensor", 4.0 * B * heads * tokens_q * tokens_kv * 64, lib().vdn_flash_attn_ex, _ptr(q, od, "q"), ld_q, q_batch_stride,
                _ptr(k, od, "k"), ld_k, k_batch_stride, _ptr(vT, od, "vT"), ld_vT, _ptr(out, od, "out"), B, tokens_q, tokens_kv, heads, _stream()),
           "vdn_flash_attn_ex")
    return out


def temporal_attn(qkv: torch.Tensor, out: torch.Tensor, D: int, T: int, C_: int, heads: int):
    od = operand_dtype()
    _check(_run("temporal_attn", "hbm", 8.0 * D * T * C_, lib().vdn_temporal_attn, _ptr(qkv, od, "qkv"), _ptr(out, od, "out"), D, T, C_, heads,
                _stream()), "vdn_temporal_attn")
    return out


def temporal_attn_tc(qk: torch.Tensor, vT: torch.Tensor, out: torch.Tensor, rows: int, C_: int, heads: int):
    od = operand_dtype()
    _check(_run("temporal_attn", "hbm", 8.0 * rows * C_, lib().vdn_temporal_attn_tc, _ptr(qk, od, "qk"), qk.shape[-1], _ptr(vT, od, "vT"),
                _ptr(out, od, "out"), rows, C_, heads, _stream()), "vdn_temporal_attn_tc")
    return out


def stream_temporal_attn(entries, pos: torch.Tensor, out: torch.Tensor, D: int, C_: int, heads: int):
    """entries: list of L <= 32 cached [D, 3C] 16-bit projections (oldest first, the current frame last)."""
    od = operand_dtype()
    L = len(entries)
    arr = (C.c_void_p * L)(*[_ptr(e, od, "entry") for e in entries])
    _check(_run("stream_temporal_attn", "hbm", 2.0 * L * D * 2 * C_ + 4.0 * D * C_, lib().vdn_stream_temporal_attn, arr, L, entries[0].shape[-1],
                _ptr(pos, torch.float32, "pos"), _ptr(out, od, "out"), D, C_, heads, _stream()), "vdn_stream_temporal_attn")
    return out


def stream_temporal_attn_ring(pool: torch.Tensor, staging: torch.Tensor, table: torch.Tensor, L: int, pos: torch.Tensor, out: torch.Tensor, D: int,
                              C_: int, heads: int):
    """Ring form: pool [slots, D, 3C], staging [D, 3C] (this frame), table int32 device tensor (entry j = slot of cached frame j, < 0 = staging)."""
    od = operand_dtype()
    _check(_run("stream_temporal_attn", "hbm", 2.0 * L * D * 2 * C_ + 4.0 * D * C_, lib().vdn_stream_temporal_attn_ring, _ptr(pool, od, "pool"),
                _ptr(staging, od, "staging"), _ptr(table, torch.int32, "table"), L, staging.shape[-1], _ptr(pos, torch.float32, "pos"), _ptr(out, od, "out"),
                D, C_, heads, _stream()), "vdn_stream_temporal_attn_ring")
    return out


def ring_store(staging: torch.Tensor, pool: torch.Tensor, table: torch.Tensor, slot_index: int):
    """pool[table[slot_index]] = staging (device-side slot lookup; a negative slot skips the store)."""
    od = operand_dtype()
    _check(_run("ring_store", "hbm", 4.0 * staging.numel(), lib().vdn_ring_store, _ptr(staging, od, "staging"), _ptr(pool, od, "pool"),
                _ptr(table, torch.int32, "table"), slot_index, staging.numel(), _stream()), "vdn_ring_store")
    return pool


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, out: torch.Tensor, eps: float, drop_first: bool = False, rows_per_batch: int = 0,
              pe: Optional[torch.Tensor] = None):
    rows, C_ = x.shape[0], x.shape[1]
    _check(_run("layernorm", "hbm", 6.0 * rows * C_, lib().vdn_layernorm, _ptr(x, torch.float32, "x"), _ptr(w, torch.float32, "w"),
                _ptr(b, torch.float32, "b"), _ptr(out, operand_dtype(), "out"), rows, C_, eps, 1 if drop_first else 0, rows_per_batch,
                _ptr(pe, torch.float32, "pe"), pe.shape[0] if pe is not None else 0, _stream()), "vdn_layernorm")
    return out


def groupnorm_stats(x, stats, frames, D, C_, groups, eps):
    _check(_run("groupnorm_stats", "hbm", 2.0 * frames * D * C_, lib().vdn_groupnorm_stats, _ptr(x, operand_dtype(), "x"),
                _ptr(stats, torch.float32, "stats"), frames, D, C_, groups, eps, _stream()), "vdn_groupnorm_stats")
    return stats


def groupnorm_apply_tc(x, stats, w, b, out, Bv, T, D, C_, groups):
    od = operand_dtype()
    _check(_run("groupnorm_apply_tc", "hbm", 4.0 * Bv * T * D * C_, lib().vdn_groupnorm_apply_tc, _ptr(x, od, "x"), _ptr(stats, torch.float32, "stats"),
                _ptr(w, torch.float32, "w"), _ptr(b, torch.float32, "b"), _ptr(out, od, "out"), Bv, T, D, C_, groups, _stream()),
           "vdn_groupnorm_apply_tc")
    return out


def groupnorm_to_tc(x, w, b, out, stats, Bv, T, D, C_, groups, eps):
    """GroupNorm + affine + frame-major -> pixel-major transpose in one call (motion_module.py:103-115): one cluster launch for the
    ViT-L / ViT-g head shapes, the statistics and apply kernels otherwise; ``stats`` (frames * groups * 2 fp32) receives (mean, rstd)."""
    od = operand_dtype()
    _check(_run("groupnorm_to_tc", "hbm", 4.0 * Bv * T * D * C_, lib().vdn_groupnorm_to_tc, _ptr(x, od, "x"), _ptr(w, torch.float32, "w"),
                _ptr(b, torch.float32, "b"), _ptr(out, od, "out"), _ptr(stats, torch.float32, "stats"), Bv, T, D, C_, groups, eps, _stream()),
           "vdn_groupnorm_to_tc")
    return out


def patch_im2col(img, out, B, H, W, Kp):
    _check(_run("patch_im2col", "hbm", B * 3.0 * H * W * 4 + B * (H // 14) * (W // 14) * Kp * 2.0, lib().vdn_patch_im2col, _ptr(img, torch.float32, "img"),
                _ptr(out, operand_dtype(), "out"), B, H, W, Kp, _stream()), "vdn_patch_im2col")
    return out


def write_cls(x, cls, pos, B, tokens, C_):
    _check(lib().vdn_write_cls(_ptr(x, torch.float32, "x"), _ptr(cls, torch.float32, "cls"), _ptr(pos, torch.float32, "pos"), B, tokens, C_, _stream()),
           "vdn_write_cls")
    return x


def im2col_3x3_s2(x, out, B, H, W, C_):
    od = operand_dtype()
    _check(_run("im2col_3x3_s2", "hbm", 2.0 * B * H * W * C_ + 2.0 * out.numel(), lib().vdn_im2col_3x3_s2, _ptr(x, od, "x"), _ptr(out, od, "out"), B, H, W,
                C_, _stream()), "vdn_im2col_3x3_s2")
    return out


def bilinear_nhwc(x, out, B, H, W, Ho, Wo, C_, relu_out=False):
    od = operand_dtype()
    _check(_run("bilinear_nhwc", "hbm", 2.0 * B * C_ * (H * W + Ho * Wo), lib().vdn_bilinear_nhwc, _ptr(x, od, "x"), _ptr(out, od, "out"), B, H, W, Ho, Wo,
                C_, 1 if relu_out else 0, _stream()), "vdn_bilinear_nhwc")
    return out


def conv_tail(x, wpacked, bias, head_w, head_b: float, out, B: int, H: int, W: int, src_hw=None):
    """3x3 conv 128 -> 32 + ReLU + 1x1 conv 32 -> 1 + ReLU -> fp32 [B, H, W] (dpt_temporal.py:107-111).  ``src_hw`` = (Hs, Ws): ``x`` is
    output_conv1's [B, Hs, Ws, 128] map and the align_corners bilinear resize to (H, W) happens inside the kernel (the tensor-core operand
    is produced in shared memory); without it ``x`` is the already resized [B, H, W, 128] map."""
    od = operand_dtype()
    flops = 2.0 * B * H * W * (9 * 128 * 32 + 32)
    if src_hw is None:
        rc = _run("conv_tail", "tensor", flops, lib().vdn_conv_tail, _ptr(x, od, "x"), _ptr(wpacked, od, "wpacked"), _ptr(bias, torch.float32, "bias"),
                  _ptr(head_w, torch.float32, "head_w"), float(head_b), _ptr(out, torch.float32, "out"), B, H, W, _stream())
    else:
        rc = _run("conv_tail", "tensor", flops, lib().vdn_conv_tail_up, _ptr(x, od, "x"), int(src_hw[0]), int(src_hw[1]), _ptr(wpacked, od, "wpacked"),
                  _ptr(bias, torch.float32, "bias"), _ptr(head_w, torch.float32, "head_w"), float(head_b), _ptr(out, torch.float32, "out"), B, H, W, _stream())
    _check(rc, "vdn_conv_tail")
    return out


def bilinear_nhwc2(x, out, out_relu, B, H, W, Ho, Wo, C_):
    """out = bilinear(x), out_relu = relu(out) in one pass."""
    od = operand_dtype()
    _check(_run("bilinear_nhwc", "hbm", 2.0 * B * C_ * (H * W + 2 * Ho * Wo), lib().vdn_bilinear_nhwc2, _ptr(x, od, "x"), _ptr(out, od, "out"),
                _ptr(out_relu, od, "out_relu"), B, H, W, Ho, Wo, C_, _stream()), "vdn_bilinear_nhwc2")
    return out, out_relu


def bilinear_f32(x, out, N, H, W, Ho, Wo, relu=False):
    _check(_run("bilinear_f32", "hbm", 4.0 * N * (H * W + Ho * Wo), lib().vdn_bilinear_f32, _ptr(x, torch.float32, "x"), _ptr(out, torch.float32, "out"),
                N, H, W, Ho, Wo, 1 if relu else 0, _stream()), "vdn_bilinear_f32")
    return out


def relu16(x, out):
    od = operand_dtype()
    _check(_run("relu16", "hbm", 4.0 * x.numel(), lib().vdn_relu16, _ptr(x, od, "x"), _ptr(out, od, "out"), x.numel(), _stream()), "vdn_relu16")
    return out


def cast_f32_to_16(x, out):
    _check(lib().vdn_cast_f32_to_16(_ptr(x, torch.float32, "x"), _ptr(out, operand_dtype(), "out"), x.numel(), _stream()), "vdn_cast_f32_to_16")
    return out


def preprocess_u8(frames_u8, out, h: int, w: int, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """uint8 RGB [N, H, W, 3] (device) -> fp32 [N, 3, h, w]: /255, cv2.INTER_CUBIC-equivalent resize, ImageNet normalise."""
    N, H, W = frames_u8.shape[:3]
    m, s_ = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _check(_run("preprocess_u8", "hbm", 3.0 * N * H * W + 12.0 * N * h * w, lib().vdn_preprocess_u8, _ptr(frames_u8, torch.uint8, "frames"),
                _ptr(out, torch.float32, "out"), N, H, W, h, w, m, s_, _stream()), "vdn_preprocess_u8")
    return out


def lsq_sums(pred, target, sums5):
    _check(lib().vdn_lsq_sums(_ptr(pred, torch.float32, "pred"), _ptr(target, torch.float32, "target"), pred.numel(), _ptr(sums5, torch.float64, "sums"),
                              _stream()), "vdn_lsq_sums")
    return sums5


def lsq_solve(sums5, scale_shift):
    _check(lib().vdn_lsq_solve(_ptr(sums5, torch.float64, "sums"), _ptr(scale_shift, torch.float32, "ss"), _stream()), "vdn_lsq_solve")
    return scale_shift


def affine_clamp(x, out, scale_shift):
    _check(lib().vdn_affine_clamp(_ptr(x, torch.float32, "x"), _ptr(out, torch.float32, "out"), x.numel(), _ptr(scale_shift, torch.float32, "ss"),
                                  _stream()), "vdn_affine_clamp")
    return out


def crossfade(pre, post, out, scale_shift, w: float):
    _check(lib().vdn_crossfade(_ptr(pre, torch.float32, "pre"), _ptr(post, torch.float32, "post"), _ptr(out, torch.float32, "out"), pre.numel(),
                               _ptr(scale_shift, torch.float32, "ss"), float(w), _stream()), "vdn_crossfade")
    return out


def window_finalize(cur, prev_tail, ss_cur, ss_prev, out, first_slot: int, count: int, is_first: bool):
    """out [count, H, W] = the aligned / cross-faded output frames of slots first_slot.. of one window (see vdn_window_finalize)."""
    n = cur[0].numel()
    _check(_run("window_finalize", "hbm", 8.0 * n * count + 4.0 * n * max(0, min(10, first_slot + count) - first_slot), lib().vdn_window_finalize,
                _ptr(cur, torch.float32, "cur"), _ptr(prev_tail, torch.float32, "prev_tail"), _ptr(ss_cur, torch.float32, "ss_cur"),
                _ptr(ss_prev, torch.float32, "ss_prev"), _ptr(out, torch.float32, "out"), n, first_slot, count, 1 if is_first else 0, _stream()),
           "vdn_window_finalize")
    return out


def window_keys(cur, keys):
    """keys [3, H, W] = slots (0, 1, 12) of cur [32, H, W]."""
    _check(lib().vdn_window_keys(_ptr(cur, torch.float32, "cur"), _ptr(keys, torch.float32, "keys"), cur[0].numel(), _stream()), "vdn_window_keys")
    return keys


def sobel_normals(depth, normals, N, H, W, channels_out=3):
    _check(lib().vdn_sobel_normals(_ptr(depth, torch.float32, "depth"), _ptr(normals, torch.float32, "normals"), N, H, W, channels_out, _stream()),
           "vdn_sobel_normals")
    return normals


def frame_median_scale(x, scale, n_per_frame: int, inv_max: float, w: float, b: float, median=None):
    N = x.numel() // n_per_frame
    _check(_run("frame_median_scale", "hbm", 4.0 * x.numel(), lib().vdn_frame_median_scale, _ptr(x, torch.float32, "x"), _ptr(median, torch.float32, "median"),
                _ptr(scale, torch.float32, "scale"), N, n_per_frame, float(inv_max), float(w), float(b), _stream()), "vdn_frame_median_scale")
    return scale


def v5_net_input(r, scale, x, N, h, w, inv_max: float):
    _check(_run("v5_net_input", "hbm", 16.0 * N * h * w, lib().vdn_v5_net_input, _ptr(r, torch.float32, "r"), _ptr(scale, torch.float32, "scale"),
                _ptr(x, torch.float32, "x"), N, h, w, float(inv_max), _stream()), "vdn_v5_net_input")
    return x


def v5_residual(din, o, scale, out, N, H, W, h, w, ws: float, bs: float, max_depth: float):
    _check(_run("v5_residual", "hbm", 8.0 * N * H * W + 4.0 * N * h * w, lib().vdn_v5_residual, _ptr(din, torch.float32, "din"), _ptr(o, torch.float32, "o"),
                _ptr(scale, torch.float32, "scale"), _ptr(out, torch.float32, "out"), N, H, W, h, w, float(ws), float(bs), float(max_depth), _stream()),
           "vdn_v5_residual")
    return out


def rope2d(x, rows: int, ld: int, col0: int, heads: int, cos_sin, P: int, rows_per_batch: int = 0, batch_pitch: int = 0, ptr_offset: int = 0):
    """In-place axial RoPE; ``ptr_offset`` (elements) selects a slot of a KV cache whose batches are ``batch_pitch`` rows apart."""
    base = _ptr(x, operand_dtype(), "x") + 2 * ptr_offset
    _check(_run("rope2d", "hbm", 4.0 * rows * heads * 64, lib().vdn_rope2d, base, rows, ld, col0, heads, _ptr(cos_sin, torch.float32, "cos_sin"), P,
                rows_per_batch, batch_pitch, _stream()), "vdn_rope2d")
    return x


def rope_chunks(x, rows: int, ld: int, col0: int, chunks: int, cos_sin, P: int):
    """Temporal RoPE in place on `chunks` 64-wide column chunks of 16-bit rows; row r is frame r % P; cos_sin [P, chunks, 64] fp32."""
    od = operand_dtype()
    _check(_run("rope", "hbm", 4.0 * rows * chunks * 64, lib().vdn_rope_chunks, _ptr(x, od, "x"), rows, ld, col0, chunks,
                _ptr(cos_sin, torch.float32, "cos_sin"), P, _stream()), "vdn_rope_chunks")
    return x


def readout_concat(tok, tok_row0: int, tok_frame_pitch: int, cls, cls_frame_pitch: int, out, frames: int, P: int, C_: int):
    """out [frames*P, 2C] = [patch token | cls of its frame]; token p of frame f = row tok_row0 + f*tok_frame_pitch + p of ``tok``
    ([rows, C]), cls of frame f = row f*cls_frame_pitch of ``cls``."""
    od = operand_dtype()
    esz = 2
    _check(_run("readout_concat", "hbm", 2.0 * frames * P * 3 * C_, lib().vdn_readout_concat, _ptr(tok, od, "tok") + tok_row0 * C_ * esz, tok_frame_pitch,
                _ptr(cls, od, "cls"), cls_frame_pitch, _ptr(out, od, "out"), frames, P, C_, _stream()), "vdn_readout_concat")
    return out


def add_rowvec(x, vec, alpha: float, out):
    rows, C_ = out.shape[0], out.shape[1]
    _check(_run("add_rowvec", "hbm", (4.0 if x.dtype == torch.float32 else 2.0) * rows * C_ + 4.0 * rows * C_, lib().vdn_add_rowvec, _ptr(x, None, "x"),
                1 if x.dtype == torch.float32 else 0, _ptr(vec, torch.float32, "vec"), float(alpha), _ptr(out, torch.float32, "out"), rows, C_, _stream()),
           "vdn_add_rowvec")
    return out


def add_rowscalar(x, m):
    rows, C_ = x.shape[0], x.shape[1]
    _check(_run("add_rowscalar", "hbm", 8.0 * rows * C_, lib().vdn_add_rowscalar, _ptr(x, torch.float32, "x"), _ptr(m, torch.float32, "m"), rows, C_, _stream()),
           "vdn_add_rowscalar")
    return x


def dwconv7_ln(x, w, bias, ln_w, ln_b, out, B, H, W, C_, eps: float):
    _check(_run("dwconv7_ln", "hbm", 6.0 * B * H * W * C_, lib().vdn_dwconv7_ln, _ptr(x, torch.float32, "x"), _ptr(w, torch.float32, "w"),
                _ptr(bias, torch.float32, "bias"), _ptr(ln_w, torch.float32, "ln_w"), _ptr(ln_b, torch.float32, "ln_b"), _ptr(out, operand_dtype(), "out"),
                B, H, W, C_, float(eps), _stream()), "vdn_dwconv7_ln")
    return out


def mask_down1(depth, params, out, B, H, W):
    _check(lib().vdn_mask_down1(_ptr(depth, torch.float32, "depth"), _ptr(params, torch.float32, "params"), _ptr(out, torch.float32, "out"), B, H, W, _stream()),
           "vdn_mask_down1")
    return out


def mask_down2(x, params, out, B, Hi, Wi):
    _check(lib().vdn_mask_down2(_ptr(x, torch.float32, "in"), _ptr(params, torch.float32, "params"), _ptr(out, torch.float32, "out"), B, Hi, Wi, _stream()),
           "vdn_mask_down2")
    return out
