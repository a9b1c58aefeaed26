"""The C-ABI kernels registered as ``torch.ops.vdn.*`` (SURVEY.md §8b "What the C-ABI extension exports").

Each op is an out-variant over CUDA tensors that forwards to the same ctypes launcher the host modules use (``ops.py``); they
exist so that a maintainer of the reference can call a single kernel from the reference's own ``nn.Module`` tree
 without touching ctypes.  No CPU implementation is registered: calling one with CPU tensors raises."""
from __future__ import annotations

import torch

from . import ops

_lib = torch.library.Library("vdn", "DEF")
_DEFS = {
    "linear": "(Tensor a, Tensor w, Tensor(a!) out, Tensor? bias=None, Tensor? gamma=None, Tensor? res=None, int act=0) -> Tensor(a!)",
    "geglu_linear": "(Tensor a, Tensor w, Tensor(a!) out, Tensor? bias=None) -> Tensor(a!)",
    "conv3x3": "(Tensor x, Tensor w, Tensor(a!) out, Tensor? bias=None, Tensor? res=None, int act=0) -> Tensor(a!)",
    "convT_ps": "(Tensor x, Tensor w, Tensor(a!) out, Tensor bias, int stride) -> Tensor(a!)",
    "head_tail": "(Tensor x, Tensor w, Tensor(a!) out, Tensor bias, Tensor head_w, float head_b) -> Tensor(a!)",
    "head_tail_up": "(Tensor x, Tensor wpacked, Tensor(a!) out, Tensor bias, Tensor head_w, float head_b) -> Tensor(a!)",
    "layernorm": "(Tensor x, Tensor w, Tensor b, Tensor(a!) out, float eps) -> Tensor(a!)",
    "flash_attn": "(Tensor qk, Tensor vT, Tensor(a!) out, int B, int tokens, int heads) -> Tensor(a!)",
    "temporal_attn": "(Tensor qkv, Tensor(a!) out, int D, int T, int heads) -> Tensor(a!)",
    "groupnorm_to_tc": "(Tensor x, Tensor w, Tensor b, Tensor(a!) out, int Bv, int T, int groups, float eps) -> Tensor(a!)",
    "bilinear_ac": "(Tensor x, Tensor(a!) out) -> Tensor(a!)",
    "sobel_normal": "(Tensor depth, Tensor(a!) normals) -> Tensor(a!)",
    "median_scale": "(Tensor x, Tensor(a!) scale, float inv_max, float w, float b) -> Tensor(a!)",
    "affine_align": "(Tensor x, Tensor scale_shift, Tensor(a!) out) -> Tensor(a!)",
}
for _name, _schema in _DEFS.items():
    _lib.define(_name + _schema)


def _linear(a, w, out, bias=None, gamma=None, res=None, act=0):
    M, K = a.shape[0], a.shape[-1]
    return ops.gemm(a, w, out, M=M, N=w.shape[0], K=K, bias=bias, gamma=gamma, res=res, act=act)


def _geglu_linear(a, w, out, bias=None):
    return ops.gemm(a, w, out, M=a.shape[0], N=w.shape[0], K=a.shape[-1], bias=bias, geglu=True)


def _conv3x3(x, w, out, bias=None, res=None, act=0):
    B, H, W, Ci = x.shape  # NHWC
    return ops.gemm(x, w, out, M=B * H * W, N=w.shape[0], K=Ci, conv=(B, H, W), bias=bias, res=res, act=act)


def _convT_ps(x, w, out, bias, stride):
    B, H, W, Ci = x.shape
    co = w.shape[0] // (stride * stride)
    return ops.gemm(x, w, out, M=B * H * W, N=w.shape[0], K=Ci, bias=bias, ldc=co, row_map=ops.ROWMAP_PIXEL_SHUFFLE, rm=(H, W, stride, co))


def _head_tail(x, w, out, bias, head_w, head_b):
    B, H, W, Ci = x.shape
    return ops.gemm(x, w, out, M=B * H * W, N=w.shape[0], K=Ci, conv=(B, H, W), bias=bias, head_w=head_w, head_b=head_b)


def _head_tail_up(x, wpacked, out, bias, head_w, head_b):
    """x: output_conv1's [B, Hs, Ws, 128] map; out: [B, H, W] fp32; wpacked: packing.pack_conv_tail of output_conv2.0 (dpt_temporal.py:103-111)."""
    B, Hs, Ws, _ = x.shape
    return ops.conv_tail(x, wpacked, bias, head_w, head_b, out, B, out.shape[1], out.shape[2], src_hw=(Hs, Ws))


def _layernorm(x, w, b, out, eps):
    return ops.layernorm(x, w, b, out, eps)


def _flash_attn(qk, vT, out, B, tokens, heads):
    return ops.flash_attn(qk, vT, out, B, tokens, heads)


def _temporal_attn(qkv, out, D, T, heads):
    return ops.temporal_attn(qkv, out, D, T, out.shape[-1], heads)


def _groupnorm_to_tc(x, w, b, out, Bv, T, groups, eps):
    frames, D, C = x.shape
    stats = torch.empty((frames * groups * 2,), dtype=torch.float32, device=x.device)
    return ops.groupnorm_to_tc(x, w, b, out, stats, Bv, T, D, C, groups, eps)


def _bilinear_ac(x, out):
    if x.dtype == torch.float32:
        return ops.bilinear_f32(x, out, x.shape[0], x.shape[1], x.shape[2], out.shape[1], out.shape[2])
    return ops.bilinear_nhwc(x, out, x.shape[0], x.shape[1], x.shape[2], out.shape[1], out.shape[2], x.shape[3])


def _sobel_normal(depth, normals):
    return ops.sobel_normals(depth, normals, depth.shape[0], depth.shape[-2], depth.shape[-1], normals.shape[1])


def _median_scale(x, scale, inv_max, w, b):
    return ops.frame_median_scale(x, scale, x.numel() // scale.numel(), inv_max, w, b)


def _affine_align(x, scale_shift, out):
    return ops.affine_clamp(x, out, scale_shift)


for _name in _DEFS:
    _lib.impl(_name, globals()["_" + _name], "CUDA")

REGISTERED = tuple(_DEFS)
