"""Drop-in host modules for the reference's model API, running entirely on the sm_100a kernels.

  VideoDepthAnything   <-> video_depth_anything/video_depth.py:35-156  (forward, infer_video_depth)
  VideoDepthRefinerV5  <-> models/video_depth_model_v5.py:128-192      (forward; the class is also exported under the
                                                                        reference's name ``VideoDepthAnything`` in .v5)

Same constructor kwargs, same ``state_dict`` key names (packed once by packing.py), same tensor shapes in and out.
Activations are token-major / NHWC 16-bit between kernels; the ViT residual stream and the motion-module hidden state
stay fp32; every statistic / softmax / accumulator is fp32.  No PyTorch compute op is on the path — torch only allocates.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional

import os

import numpy as np
import torch
import torch.nn as nn

from . import ops, packing

ENCODER_CONFIGS = {
    # video_depth.py:48-51 / run_video.py:28-33
    "vits": dict(embed_dim=384, depth=12, heads=6, taps=[2, 5, 8, 11]),
    "vitl": dict(embed_dim=1024, depth=24, heads=16, taps=[4, 11, 17, 23]),
}
# depth_anything_v2.py:24-29 also lists vitb (dinov2.py:353-364) and vitg (dinov2.py:381-395: 40 blocks, embed_dim 1536, 24 heads,
# SwiGLU FFN, dinov2_layers/swiglu_ffn.py)
DA2_ENCODER_CONFIGS = dict(ENCODER_CONFIGS, vitb=dict(embed_dim=768, depth=12, heads=12, taps=[2, 5, 8, 11]),
                           vitg=dict(embed_dim=1536, depth=40, heads=24, taps=[9, 19, 29, 39], ffn="swiglu"))


def swiglu_hidden(C: int) -> int:
    """SwiGLUFFNFused: hidden = round_up_8(int(4 C * 2 / 3))  (dinov2_layers/swiglu_ffn.py:52-63 with mlp_ratio 4)."""
    return (int(4 * C * 2 / 3) + 7) // 8 * 8

# video_depth.py:29-33 — "infer settings, do not change"
INFER_LEN = 32
OVERLAP = 10
KEYFRAMES = [0, 12, 24, 25, 26, 27, 28, 29, 30, 31]
INTERP_LEN = 8


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


class GraphRunner:
    """Replays a fixed-shape kernel sequence as one CUDA graph.  A forward is ~270 launches of which ~150 are short head /
    motion-module kernels: launched one by one from Python the GPU idles between them (~5 ms of a 59 ms ViT-L window); the
    sequence is static for a given input shape (tensor maps are baked into the kernel parameters), so it is captured once —
    first call eager (fills the weight / pos-embed / workspace caches), second call captured, later calls replayed.
    ``fn(*static_inputs) -> tensor or list of tensors``; results live in graph-owned buffers that the next replay overwrites."""

    enabled = os.environ.get("VDN_NO_GRAPHS", "0") != "1"

    def __init__(self):
        self.entries = {}

    def clear(self):
        self.entries.clear()

    def run(self, key, fn, inputs, static_inputs: bool = False):
        """``static_inputs``: the caller passes the same persistent tensors on every call (their addresses are baked into the graph),
        so nothing is copied; otherwise the inputs are copied into graph-owned buffers before every replay."""
        if not GraphRunner.enabled or ops._profiler is not None or torch.cuda.is_current_stream_capturing():
            return fn(*inputs)
        if static_inputs:
            key = key + tuple(t.data_ptr() for t in inputs)
        e = self.entries.get(key)
        if e is None:
            self.entries[key] = {"calls": 1}
            return fn(*inputs)
        if "graph" not in e:
            if static_inputs:
                static_in = list(inputs)
            else:
                static_in = [torch.empty_like(t) for t in inputs]
                for s_, t in zip(static_in, inputs):
                    s_.copy_(t, non_blocking=True)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                out = fn(*static_in)
            e.update(graph=g, inputs=static_in, out=out, launches=ops.launch_count() - n0)
            ops.lib().vdn_add_launch_count(-e["launches"])  # capture issued no work; every replay (below) counts them
        if not static_inputs:
            for s_, t in zip(e["inputs"], inputs):
                s_.copy_(t, non_blocking=True)
        e["graph"].replay()
        ops.lib().vdn_add_launch_count(e["launches"])
        return e["out"]


# ======================================================================================================
# encoder: DINOv2 get_intermediate_layers (dinov2.py:212-231, 271-321; block.py:82-107)
# ======================================================================================================
def encoder_forward(enc: dict, x: torch.Tensor, readout: Optional[list] = None, defer_last_readout: bool = False) -> List[torch.Tensor]:
    """x (Bf, 3, H, W) fp32 -> 4 x [Bf*ph*pw, C] 16-bit (final-norm'ed patch tokens of the tapped blocks, cls dropped).
    ``readout`` (use_clstoken=True, dpt.py:129-132): the four packed readout_projects; the per-frame [token | cls] -> Linear -> GELU
    readout has no cross-frame term, so it runs here and the features keep their shape for every consumer (head, window reuse, streaming).
    ``defer_last_readout`` (DepthAnythingV2: the last tap goes through the memory block first, depth_anything_v2.py:49-51): the last
    feature is returned without its readout and a fifth element is appended, the normed token matrix [Bf*N, C] holding its cls rows."""
    Bf, _, H, W = x.shape
    if H % 14 != 0 or W % 14 != 0:
        raise RuntimeError(f"Input image height {H} / width {W} is not a multiple of patch size 14")  # patch_embed.py:73-74
    dev, od = x.device, ops.operand_dtype()
    C, heads = enc["C"], enc["heads"]
    ph, pw = H // 14, W // 14
    P, N = ph * pw, ph * pw + 1
    rows = Bf * N
    pos = packing.encoder_pos_embed(enc, ph, pw, dev)
    patches = _empty((Bf * P, 592), od, dev)
    ops.patch_im2col(x.contiguous(), patches, Bf, H, W, 592)
    xs = _empty((rows, C), torch.float32, dev)  # fp32 residual stream
    ops.gemm(patches, enc["patch_w"], xs, M=Bf * P, N=C, K=592, bias=enc["patch_b"], res=pos, row_map=ops.ROWMAP_PATCH_TOKENS, rm=(P, 0, 0, 0))
    ops.write_cls(xs, enc["cls"], pos, Bf, N, C)
    npad = (N + 7) // 8 * 8
    xn = _empty((rows, C), od, dev)
    qk = _empty((rows, 2 * C), od, dev)
    vT = _empty((Bf * heads, 64, npad), od, dev)
    ao = _empty((rows, C), od, dev)
    hid = _empty((rows, enc.get("hidden", 4 * C)), od, dev)
    feats = []
    for i, blk in enumerate(enc["blocks"]):
        ops.layernorm(xs, blk["ln1_w"], blk["ln1_b"], xn, 1e-6)
        ops.gemm(xn, blk["qkv"]["w"], qk, M=rows, N=3 * C, K=C, bias=blk["qkv"]["b"], ldc=2 * C, out2=vT, row_map=ops.ROWMAP_QKV_SPLIT,
                 rm=(N, npad, C, 0))
        ops.flash_attn(qk, vT, ao, Bf, N, heads)
        ops.gemm(ao, blk["proj"]["w"], xs, M=rows, N=C, K=C, bias=blk["proj"]["b"], gamma=blk["ls1"], res=xs)
        ops.layernorm(xs, blk["ln2_w"], blk["ln2_b"], xn, 1e-6)
        if "w12" in blk:  # SwiGLU FFN (ViT-g): hidden = silu(x1) * x2 in the epilogue of the interleaved (x2, x1) projection
            Hd = enc["hidden"]
            ops.gemm(xn, blk["w12"]["w"], hid, M=rows, N=2 * Hd, K=C, bias=blk["w12"]["b"], geglu=True, act=ops.ACT_SILU)
            ops.gemm(hid, blk["w3"]["w"], xs, M=rows, N=C, K=Hd, bias=blk["w3"]["b"], gamma=blk["ls2"], res=xs)
        else:
            ops.gemm(xn, blk["fc1"]["w"], hid, M=rows, N=4 * C, K=C, bias=blk["fc1"]["b"], act=ops.ACT_GELU)
            ops.gemm(hid, blk["fc2"]["w"], xs, M=rows, N=C, K=4 * C, bias=blk["fc2"]["b"], gamma=blk["ls2"], res=xs)
        if i in enc["taps"]:
            f = _empty((Bf * P, C), od, dev)
            last_deferred = readout is not None and defer_last_readout and len(feats) == 3
            if readout is None or last_deferred:
                ops.layernorm(xs, enc["norm_w"], enc["norm_b"], f, 1e-6, drop_first=True, rows_per_batch=N)
                if last_deferred:
                    ops.layernorm(xs, enc["norm_w"], enc["norm_b"], xn, 1e-6)  # the last tap is the last block: xn stays untouched
            else:
                ops.layernorm(xs, enc["norm_w"], enc["norm_b"], xn, 1e-6)  # xn is free until the next block's ln1
                readout_apply(readout[len(feats)], xn, 1, N, xn, N, f, Bf, P, C)
            feats.append(f)
    if readout is not None and defer_last_readout:
        feats.append(xn)
    return feats


def readout_apply(ro: dict, tok: torch.Tensor, tok_row0: int, tok_pitch: int, cls: torch.Tensor, cls_pitch: int, out: torch.Tensor, Bf: int, P: int, C: int):
    """readout_projects[i](cat(token, cls)) = GELU(Linear(2C -> C)) (dpt.py:129-132)."""
    cat = _empty((Bf * P, 2 * C), ops.operand_dtype(), tok.device)
    ops.readout_concat(tok, tok_row0, tok_pitch, cls, cls_pitch, cat, Bf, P, C)
    ops.gemm(cat, ro["w"], out, M=Bf * P, N=C, K=2 * C, bias=ro["b"], act=ops.ACT_GELU)
    return out


# ======================================================================================================
# temporal motion module (motion_module.py:60-136, 174-192, 245-326)
# ======================================================================================================
def motion_module_forward(mm: dict, x: torch.Tensor, Bv: int, T: int, D: int, relu_copy: bool = False):
    """x: [Bv*T, D, C] 16-bit frame-major NHWC -> same layout; optional ReLU'd copy for the next RCU."""
    dev, od, C = x.device, ops.operand_dtype(), mm["C"]
    rows = Bv * D * T
    stats = _empty((Bv * T * 32 * 2,), torch.float32, dev)
    xt = _empty((rows, C), od, dev)
    ops.groupnorm_to_tc(x, mm["gn_w"], mm["gn_b"], xt, stats, Bv, T, D, C, 32, 1e-6)
    h = _empty((rows, C), torch.float32, dev)  # fp32 hidden state, pixel-major rows (b*D+d)*T+f
    ops.gemm(xt, mm["proj_in"]["w"], h, M=rows, N=C, K=C, bias=mm["proj_in"]["b"])
    n16 = xt  # reuse
    ao = _empty((rows, C), od, dev)
    # T == 32 with head_dim 32 / 64 / 128 (ViT-L) runs on the tcgen05 kernel: the fused q|k|v projection writes q|k row-major
    # and V transposed per 128-row tile (4 pixels x 32 frames); other shapes (ViT-S heads, short clips) take the CUDA-core kernel
    tc = T == 32 and (C // 8) in (32, 64, 128)
    if tc:
        qk = _empty((rows, 2 * C), od, dev)
        key = ("vT", rows, od)
        if key not in mm:  # zero beyond the last valid row of the last tile, written only inside the valid rows afterwards
            mm[key] = torch.zeros(((rows + 127) // 128 * C, 128), dtype=od, device=dev)
        vT = mm[key]
    else:
        qkv = _empty((rows, 3 * C), od, dev)
    for a in mm["attn"]:
        rope = a["pe"] is None  # pe='rope': rotate the projected q | k columns by the frame index instead of adding a table to the input
        ops.layernorm(h, a["ln_w"], a["ln_b"], n16, 1e-5, pe=None if rope else a["pe"][:T])
        if tc:
            ops.gemm(n16, a["qkv_w"], qk, M=rows, N=3 * C, K=C, ldc=2 * C, out2=vT, row_map=ops.ROWMAP_QKV_SPLIT, rm=(128, 128, C, 0))
            if rope:
                ops.rope_chunks(qk, rows, 2 * C, 0, 2 * C // 64, a["rope"], T)
            ops.temporal_attn_tc(qk, vT, ao, rows, C, 8)
        else:
            ops.gemm(n16, a["qkv_w"], qkv, M=rows, N=3 * C, K=C)
            if rope:
                ops.rope_chunks(qkv, rows, 3 * C, 0, 2 * C // 64, a["rope"], T)
            ops.temporal_attn(qkv, ao, Bv * D, T, C, 8)
        ops.gemm(ao, a["out"]["w"], h, M=rows, N=C, K=C, bias=a["out"]["b"], res=h)
    ops.layernorm(h, mm["ffn_w"], mm["ffn_b"], n16, 1e-5)
    g = _empty((rows, 4 * C), od, dev)
    ops.gemm(n16, mm["ff1_w"], g, M=rows, N=8 * C, K=C, bias=mm["ff1_b"], geglu=True)
    h16 = ao  # reuse
    ops.gemm(g, mm["ff2"]["w"], h, M=rows, N=C, K=4 * C, bias=mm["ff2"]["b"], res=h, out2=h16)
    y = _empty((Bv * T, D, C), od, dev)
    y_relu = _empty((Bv * T, D, C), od, dev) if relu_copy else None
    ops.gemm(h16, mm["proj_out"]["w"], y, M=rows, N=C, K=C, bias=mm["proj_out"]["b"], res=x, row_map=ops.ROWMAP_TEMPORAL, rm=(T, D, 0, 0),
             out2=y_relu, out2_relu=True)
    return (y, y_relu) if relu_copy else y


RING_NEW_SLOT = 32   # index in the slot table of the pool slot that receives this frame's projections
RING_SLOTS = 44      # the reference's cache list holds at most 42 frames (video_depth_stream.py:150-158) + the incoming one


def motion_module_stream(mm: dict, x: torch.Tensor, D: int, cached: Optional[list], new_cache: list, ring: Optional[dict] = None):
    """One new frame through a motion module with cached history (motion_module.py:102-136 with cached_hidden_state_list,
    TemporalAttention.forward :252-259).  x: [1, D, C] 16-bit NHWC.  ``cached``: per attention block the list of earlier frames'
    cached projections (oldest first); ``new_cache`` receives this frame's two projections.  The reference caches the normed hidden
    state n_j and re-projects all <= 32 frames on every call; to_q/k/v have no bias, so W (n_j + pe_j) = W n_j + W pe_j: the
    projection W n_j is cached instead and the positional part is the weights-only table ``pos_qkv``."""
    dev, od, C = x.device, ops.operand_dtype(), mm["C"]
    stats = _empty((32 * 2,), torch.float32, dev)
    ops.groupnorm_stats(x, stats, 1, D, C, 32, 1e-6)
    xt = _empty((D, C), od, dev)
    ops.groupnorm_apply_tc(x, stats, mm["gn_w"], mm["gn_b"], xt, 1, 1, D, C, 32)
    h = _empty((D, C), torch.float32, dev)
    ops.gemm(xt, mm["proj_in"]["w"], h, M=D, N=C, K=C, bias=mm["proj_in"]["b"])
    n16 = xt
    ao = _empty((D, C), od, dev)
    for i, a in enumerate(mm["attn"]):
        ops.layernorm(h, a["ln_w"], a["ln_b"], n16, 1e-5)  # the positional term is applied in the attention kernel
        if ring is not None:
            # graph-replayable form: this frame's projection goes to a fixed staging buffer, the cached ones are addressed through
            # the device-side slot table, and the cache insertion (staging -> pool[new slot]) is a kernel of the same graph
            proj, pool = ring["staging"][i], ring["pool"][i]
            ops.gemm(n16, a["qkv_w"], proj, M=D, N=3 * C, K=C)
            ops.stream_temporal_attn_ring(pool, proj, ring["table"], ring["L"], a["pos_qkv"], ao, D, C, 8)
            ops.ring_store(proj, pool, ring["table"], RING_NEW_SLOT)
        else:
            proj = _empty((D, 3 * C), od, dev)             # (Wq n | Wk n | Wv n) of this frame: what gets cached
            ops.gemm(n16, a["qkv_w"], proj, M=D, N=3 * C, K=C)
            new_cache.append(proj)
            entries = (cached[i] if cached is not None else []) + [proj]
            ops.stream_temporal_attn(entries, a["pos_qkv"], ao, D, C, 8)
        ops.gemm(ao, a["out"]["w"], h, M=D, N=C, K=C, bias=a["out"]["b"], res=h)
    ops.layernorm(h, mm["ffn_w"], mm["ffn_b"], n16, 1e-5)
    g = _empty((D, 4 * C), od, dev)
    ops.gemm(n16, mm["ff1_w"], g, M=D, N=8 * C, K=C, bias=mm["ff1_b"], geglu=True)
    h16 = ao
    ops.gemm(g, mm["ff2"]["w"], h, M=D, N=C, K=4 * C, bias=mm["ff2"]["b"], res=h, out2=h16)
    y = _empty((1, D, C), od, dev)
    ops.gemm(h16, mm["proj_out"]["w"], y, M=D, N=C, K=C, bias=mm["proj_out"]["b"], res=x)
    return y


# ======================================================================================================
# DPT head (dpt.py:126-159, dpt_temporal.py:53-127, util/blocks.py:68-162)
# ======================================================================================================
def _conv3(x, cw, B, H, W, out=None, **kw):
    dev, od = x.device, ops.operand_dtype()
    co = cw["co"]
    if out is None:
        out = _empty((B, H, W, co), od, dev)
    ops.gemm(x, cw["w"], out, M=B * H * W, N=co, K=cw["ci"], conv=(B, H, W), bias=cw.get("b"), **kw)
    return out


def _rcu(x, x_relu, rcu, B, H, W, res2=None, relu_copy=False):
    """ResidualConvUnit: conv2(relu(conv1(relu(x)))) + x (+ res2).  x_relu = relu(x) precomputed by the producer."""
    dev, od = x.device, ops.operand_dtype()
    c1 = _conv3(x_relu, rcu[0], B, H, W, act=ops.ACT_RELU)
    out2 = _empty(tuple(x.shape), od, dev) if relu_copy else None
    o = _conv3(c1, rcu[1], B, H, W, res=x, res2=res2, out2=out2, out2_relu=True)
    return (o, out2) if relu_copy else o


def _fusion(rf, B, H, W, Ho, Wo, x0, x0_relu=None, x1=None, x1_relu=None, relu_copy=False):
    """FeatureFusionBlock: out_conv(upsample(RCU2(x0 + RCU1(x1)))).  The 1x1 out_conv commutes exactly with the
    align_corners bilinear upsample (interpolation weights sum to 1), so it runs at the low resolution."""
    dev, od = x0.device, ops.operand_dtype()
    Fe = x0.shape[-1]
    if x1 is not None:
        s, s_relu = _rcu(x1, x1_relu, rf["rcu1"], B, H, W, res2=x0, relu_copy=True)
    else:
        s, s_relu = x0, x0_relu
    o = _rcu(s, s_relu, rf["rcu2"], B, H, W)
    lo = _empty((B, H, W, Fe), od, dev)
    ops.gemm(o, rf["out_conv"]["w"], lo, M=B * H * W, N=Fe, K=Fe, bias=rf["out_conv"]["b"])
    up = _empty((B, Ho, Wo, Fe), od, dev)
    if relu_copy:
        upr = _empty((B, Ho, Wo, Fe), od, dev)
        ops.bilinear_nhwc2(lo, up, upr, B, H, W, Ho, Wo, Fe)
        return up, upr
    ops.bilinear_nhwc(lo, up, B, H, W, Ho, Wo, Fe)
    return up


_FUSED_TAIL = os.environ.get("VDN_FUSED_TAIL", "1") != "0"  # evaluation switch: 0 = separate bilinear + implicit-GEMM output_conv2


def head_forward(head: dict, feats: List[torch.Tensor], Bf: int, ph: int, pw: int, T: Optional[int], stream: Optional[dict] = None) -> torch.Tensor:
    """-> depth fp32 [Bf, 14*ph, 14*pw] (after output_conv2's ReLUs).
    ``stream`` (Bf == 1): {"cached": per motion module, per attention block, the list of cached projections or None; "new": []}
    — the streaming path of video_depth_stream.py / dpt_temporal.py:72-96 with cached_hidden_state_list."""
    dev, od = feats[0].device, ops.operand_dtype()
    C, Fe, oc = feats[0].shape[-1], head["features"], head["oc"]
    P = ph * pw
    M = Bf * P
    # ---- reassemble (dpt_temporal.py:55-69)
    pr = []
    for i in range(4):
        t = _empty((M, oc[i]), od, dev)
        ops.gemm(feats[i], head["projects"][i]["w"], t, M=M, N=oc[i], K=C, bias=head["projects"][i]["b"])
        pr.append(t)
    H1, W1, H2, W2, H3, W3 = 4 * ph, 4 * pw, 2 * ph, 2 * pw, ph, pw
    H4, W4 = (ph - 1) // 2 + 1, (pw - 1) // 2 + 1
    layer1 = _empty((Bf, H1, W1, oc[0]), od, dev)
    ops.gemm(pr[0], head["resize0"]["w"], layer1, M=M, N=16 * oc[0], K=oc[0], bias=head["resize0"]["b"], ldc=oc[0],
             row_map=ops.ROWMAP_PIXEL_SHUFFLE, rm=(ph, pw, 4, oc[0]))
    layer2 = _empty((Bf, H2, W2, oc[1]), od, dev)
    ops.gemm(pr[1], head["resize1"]["w"], layer2, M=M, N=4 * oc[1], K=oc[1], bias=head["resize1"]["b"], ldc=oc[1],
             row_map=ops.ROWMAP_PIXEL_SHUFFLE, rm=(ph, pw, 2, oc[1]))
    layer3 = pr[2].view(Bf, P, oc[2])
    col = _empty((Bf * H4 * W4, 9 * oc[3]), od, dev)
    ops.im2col_3x3_s2(pr[3], col, Bf, ph, pw, oc[3])
    layer4 = _empty((Bf, H4 * W4, oc[3]), od, dev)
    ops.gemm(col, head["resize3"]["w"], layer4, M=Bf * H4 * W4, N=oc[3], K=9 * oc[3], bias=head["resize3"]["b"])
    # ---- temporal mixing on layer_3 / layer_4 (dpt_temporal.py:81-84)
    def mm_stream(m, x, D):
        if "ring" in stream:  # {"ring": per motion module {"pool": [2], "staging": [2]}, "table": int32[33] device, "L": entries}
            r = dict(stream["ring"][m], table=stream["table"], L=stream["L"])
            return motion_module_stream(head["mm"][m], x, D, None, [], ring=r)
        cached = stream["cached"][m] if stream["cached"] is not None else None
        return motion_module_stream(head["mm"][m], x, D, cached, stream["new"])

    if stream is not None:
        layer3 = mm_stream(0, layer3, P)
        layer4 = mm_stream(1, layer4, H4 * W4)
    elif T is not None:
        Bv = Bf // T
        layer3 = motion_module_forward(head["mm"][0], layer3, Bv, T, P)
        layer4 = motion_module_forward(head["mm"][1], layer4, Bv, T, H4 * W4)
    # ---- layerN_rn (3x3, no bias) with ReLU'd copies for the RCUs
    def rn(i, x, H, W):
        out2 = _empty((Bf, H, W, Fe), od, dev)
        o = _conv3(x, head["rn"][i], Bf, H, W, out2=out2, out2_relu=True)
        return o, out2
    l1r, l1r_relu = rn(0, layer1, H1, W1)
    l2r, l2r_relu = rn(1, layer2, H2, W2)
    l3r, l3r_relu = rn(2, layer3, H3, W3)
    l4r, l4r_relu = rn(3, layer4, H4, W4)
    # ---- refinenets (dpt_temporal.py:91-101)
    rf = head["refine"]
    path4 = _fusion(rf[4], Bf, H4, W4, H3, W3, l4r, x0_relu=l4r_relu)
    if stream is not None:
        path4 = mm_stream(2, path4.view(Bf, H3 * W3, Fe), H3 * W3)
    elif T is not None:
        path4 = motion_module_forward(head["mm"][2], path4.view(Bf, H3 * W3, Fe), Bv, T, H3 * W3)
    path3 = _fusion(rf[3], Bf, H3, W3, H2, W2, path4, x1=l3r, x1_relu=l3r_relu)
    if stream is not None:
        path3 = mm_stream(3, path3.view(Bf, H2 * W2, Fe), H2 * W2)
    elif T is not None:
        path3 = motion_module_forward(head["mm"][3], path3.view(Bf, H2 * W2, Fe), Bv, T, H2 * W2)
    path2 = _fusion(rf[2], Bf, H2, W2, H1, W1, path3, x1=l2r, x1_relu=l2r_relu)
    path1 = _fusion(rf[1], Bf, H1, W1, 2 * H1, 2 * W1, path2, x1=l1r, x1_relu=l1r_relu)
    # ---- output convs (dpt_temporal.py:103-111): 3x3 F->F/2, bilinear to 14x, fused [3x3 -> ReLU -> 1x1 -> ReLU]
    o1 = _conv3(path1, head["oc1"], Bf, 2 * H1, 2 * W1)
    Ho, Wo = 14 * ph, 14 * pw
    depth = _empty((Bf, Ho, Wo), torch.float32, dev)
    if "oc2_tail" in head and _FUSED_TAIL:  # resize + 3x3 + ReLU + 1x1 + ReLU in one kernel: the 128-channel full-resolution map is never written
        ops.conv_tail(o1, head["oc2_tail"], head["oc2"]["b"], head["oc2_head_w"], head["oc2_head_b"], depth, Bf, Ho, Wo, src_hw=(2 * H1, 2 * W1))
        return depth
    up = _empty((Bf, Ho, Wo, Fe // 2), od, dev)
    ops.bilinear_nhwc(o1, up, Bf, 2 * H1, 2 * W1, Ho, Wo, Fe // 2)
    _conv3(up, head["oc2"], Bf, Ho, Wo, out=depth, head_w=head["oc2_head_w"], head_b=head["oc2_head_b"])
    return depth


# ======================================================================================================
# modules
# ======================================================================================================
class _PackedModule(nn.Module):
    """Holds the reference-format state_dict on the host and the packed kernel weights on the device."""

    _kind = ""

    def __init__(self):
        super().__init__()
        self._sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self._packed: Optional[dict] = None
        self._packed_key = None
        self._dev = torch.device("cpu")
        self._graphs = GraphRunner()

    # --- nn.Module surface the reference callers use -----------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True):
        expected = self._expected_shapes()
        missing = [k for k in expected if k not in state_dict]
        unexpected = [k for k in state_dict if k not in expected]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing keys {missing[:8]}, unexpected keys {unexpected[:8]}")
        for k, shp in expected.items():
            if k in state_dict and tuple(state_dict[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: copying a param with shape {tuple(state_dict[k].shape)}, expected {tuple(shp)}")
        self._sd = OrderedDict((k, state_dict[k].detach().to("cpu", torch.float32).clone()) for k in expected if k in state_dict)
        self._packed = None
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def state_dict(self, *args, **kwargs):
        return OrderedDict(self._sd)

    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0))
        if probe.device != self._dev:
            self._dev = probe.device
            self._packed = None
        return super()._apply(fn, recurse)

    def _weights(self) -> dict:
        if self._dev.type != "cuda":
            raise RuntimeError("this model runs on CUDA only (sm_100a kernels, no CPU fallback): call .cuda() first")
        key = (self._dev, ops.operand_dtype())
        if self._packed is None or self._packed_key != key:
            if not self._sd:
                raise RuntimeError("no weights loaded: call load_state_dict() with a reference-format state_dict")
            self._packed = self._pack(self._sd, self._dev, ops.operand_dtype())
            self._packed_key = key
            self._graphs.clear()  # captured graphs hold pointers into the old packed weights
        return self._packed

    def _expected_shapes(self) -> Dict[str, tuple]:
        raise NotImplementedError

    def _pack(self, sd, dev, dt) -> dict:
        raise NotImplementedError


def _encoder_shapes(prefix: str, cfg: dict) -> Dict[str, tuple]:
    C = cfg["embed_dim"]
    s = OrderedDict()
    s[prefix + "cls_token"] = (1, 1, C)
    s[prefix + "pos_embed"] = (1, 1370, C)
    s[prefix + "mask_token"] = (1, C)
    s[prefix + "patch_embed.proj.weight"] = (C, 3, 14, 14)
    s[prefix + "patch_embed.proj.bias"] = (C,)
    for i in range(cfg["depth"]):
        p = f"{prefix}blocks.{i}."
        for n in ("norm1", "norm2"):
            s[p + n + ".weight"] = (C,)
            s[p + n + ".bias"] = (C,)
        s[p + "attn.qkv.weight"], s[p + "attn.qkv.bias"] = (3 * C, C), (3 * C,)
        s[p + "attn.proj.weight"], s[p + "attn.proj.bias"] = (C, C), (C,)
        s[p + "ls1.gamma"] = s[p + "ls2.gamma"] = (C,)
        if cfg.get("ffn") == "swiglu":
            Hd = swiglu_hidden(C)
            s[p + "mlp.w12.weight"], s[p + "mlp.w12.bias"] = (2 * Hd, C), (2 * Hd,)
            s[p + "mlp.w3.weight"], s[p + "mlp.w3.bias"] = (C, Hd), (C,)
        else:
            s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = (4 * C, C), (4 * C,)
            s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = (C, 4 * C), (C,)
    s[prefix + "norm.weight"] = s[prefix + "norm.bias"] = (C,)
    return s


def _head_shapes(prefix: str, C: int, Fe: int, oc: List[int], temporal: bool, pe: str = "ape", use_clstoken: bool = False) -> Dict[str, tuple]:
    s = OrderedDict()
    if use_clstoken:
        for i in range(4):
            s[f"{prefix}readout_projects.{i}.0.weight"], s[f"{prefix}readout_projects.{i}.0.bias"] = (C, 2 * C), (C,)
    for i in range(4):
        s[f"{prefix}projects.{i}.weight"], s[f"{prefix}projects.{i}.bias"] = (oc[i], C, 1, 1), (oc[i],)
    s[prefix + "resize_layers.0.weight"], s[prefix + "resize_layers.0.bias"] = (oc[0], oc[0], 4, 4), (oc[0],)
    s[prefix + "resize_layers.1.weight"], s[prefix + "resize_layers.1.bias"] = (oc[1], oc[1], 2, 2), (oc[1],)
    s[prefix + "resize_layers.3.weight"], s[prefix + "resize_layers.3.bias"] = (oc[3], oc[3], 3, 3), (oc[3],)
    sc = prefix + "scratch."
    for i in range(4):
        s[f"{sc}layer{i + 1}_rn.weight"] = (Fe, oc[i], 3, 3)
    for r in (1, 2, 3, 4):
        p = f"{sc}refinenet{r}."
        s[p + "out_conv.weight"], s[p + "out_conv.bias"] = (Fe, Fe, 1, 1), (Fe,)
        for u in (1, 2):
            for c in (1, 2):
                s[f"{p}resConfUnit{u}.conv{c}.weight"], s[f"{p}resConfUnit{u}.conv{c}.bias"] = (Fe, Fe, 3, 3), (Fe,)
    s[sc + "output_conv1.weight"], s[sc + "output_conv1.bias"] = (Fe // 2, Fe, 3, 3), (Fe // 2,)
    s[sc + "output_conv2.0.weight"], s[sc + "output_conv2.0.bias"] = (32, Fe // 2, 3, 3), (32,)
    s[sc + "output_conv2.2.weight"], s[sc + "output_conv2.2.bias"] = (1, 32, 1, 1), (1,)
    if temporal:
        for m, Cm in enumerate([oc[2], oc[3], Fe, Fe]):
            p = f"{prefix}motion_modules.{m}.temporal_transformer."
            s[p + "norm.weight"] = s[p + "norm.bias"] = (Cm,)
            s[p + "proj_in.weight"], s[p + "proj_in.bias"] = (Cm, Cm), (Cm,)
            tb = p + "transformer_blocks.0."
            for a in range(2):
                ab = f"{tb}attention_blocks.{a}."
                s[ab + "to_q.weight"] = s[ab + "to_k.weight"] = s[ab + "to_v.weight"] = (Cm, Cm)
                s[ab + "to_out.0.weight"], s[ab + "to_out.0.bias"] = (Cm, Cm), (Cm,)
                if pe == "ape":
                    s[ab + "pos_encoder.pe"] = (1, 32, Cm)
                s[f"{tb}norms.{a}.weight"] = s[f"{tb}norms.{a}.bias"] = (Cm,)
            s[tb + "ff.net.0.proj.weight"], s[tb + "ff.net.0.proj.bias"] = (8 * Cm, Cm), (8 * Cm,)
            s[tb + "ff.net.2.weight"], s[tb + "ff.net.2.bias"] = (Cm, 4 * Cm), (Cm,)
            s[tb + "ff_norm.weight"] = s[tb + "ff_norm.bias"] = (Cm,)
            s[p + "proj_out.weight"], s[p + "proj_out.bias"] = (Cm, Cm), (Cm,)
    return s


class VideoDepthAnything(_PackedModule):
    """Drop-in for video_depth_anything/video_depth.py:35 ``VideoDepthAnything`` (vits / vitl, use_bn=False; use_clstoken and
    pe in {'ape', 'rope'} as in the reference constructor)."""

    def __init__(self, encoder="vitl", features=256, out_channels=(256, 512, 1024, 1024), use_bn=False, use_clstoken=False, num_frames=32, pe="ape"):
        super().__init__()
        if encoder not in ENCODER_CONFIGS:
            raise KeyError(encoder)  # the reference indexes intermediate_layer_idx[encoder] (video_depth.py:48-51)
        if use_bn:
            raise NotImplementedError("use_bn is never exercised by the reference (SURVEY.md §8b)")
        if pe not in ("ape", "rope"):
            raise NotImplementedError(pe)  # motion_module.py:242-243
        assert num_frames > 0
        self.encoder = encoder
        self.use_clstoken, self.pe = bool(use_clstoken), pe
        self.num_frames = num_frames
        self.cfg = dict(ENCODER_CONFIGS[encoder], features=features, out_channels=list(out_channels))
        self.intermediate_layer_idx = {k: v["taps"] for k, v in ENCODER_CONFIGS.items()}

    def _expected_shapes(self):
        s = _encoder_shapes("pretrained.", self.cfg)
        s.update(_head_shapes("head.", self.cfg["embed_dim"], self.cfg["features"], self.cfg["out_channels"], True, self.pe, self.use_clstoken))
        return s

    def _pack(self, sd, dev, dt):
        return {"enc": packing.pack_encoder(sd, "pretrained.", self.cfg, dev, dt),
                "head": packing.pack_head(sd, "head.", self.cfg, dev, dt, True, self.pe, self.use_clstoken)}

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x (B, T, 3, H, W) fp32 -> depth (B, T, H, W) fp32   (video_depth.py:58-65)."""
        if x.dim() != 5 or x.shape[2] != 3:
            raise RuntimeError(f"expected (B, T, 3, H, W), got {tuple(x.shape)}")
        w = self._weights()
        B, T, _, H, W = x.shape
        if T > 32:
            raise RuntimeError("temporal attention supports at most 32 frames per window")
        ph, pw = H // 14, W // 14

        def run(xd):
            feats = encoder_forward(w["enc"], xd.reshape(B * T, 3, H, W), w["head"].get("readout"))
            # F.interpolate(depth, (H, W), align_corners=True) is the identity here (H == 14*ph) and the head's output is already >= 0
            return head_forward(w["head"], feats, B * T, ph, pw, T)

        if x.device != self._dev or x.dtype != torch.float32:
            xs = torch.empty(x.shape, dtype=torch.float32, device=self._dev)
            xs.copy_(x, non_blocking=True)  # H2D straight from (pinned) host memory
            x = xs
        depth = self._graphs.run(("forward", B, T, H, W), run, [x.contiguous()])
        return depth.view(B, T, H, W).clone()  # the graph's output buffer is overwritten by the next call

    # --- the two halves of forward(), exposed for the long-video driver: the encoder is per-frame (dinov2.py:212-321 has no
    # cross-frame op), so the features of the 10 key frames a window shares with its predecessor are computed once.
    @torch.no_grad()
    def encode_frames(self, x: torch.Tensor, clone: bool = True) -> List[torch.Tensor]:
        """x (F, 3, H, W) fp32 -> 4 x [F*ph*pw, C] tapped, final-norm'ed patch tokens (frame-major).  ``clone=False`` returns the
        replayed graph's own output buffers (valid until the next call with the same shape)."""
        w = self._weights()
        x = x.to(device=self._dev, dtype=torch.float32).contiguous()
        feats = self._graphs.run(("encode",) + tuple(x.shape), lambda xd: encoder_forward(w["enc"], xd, w["head"].get("readout")), [x])
        return [f.clone() for f in feats] if clone else list(feats)  # the long-video driver keeps per-frame views of these across windows

    def window_feature_buffers(self, ph: int, pw: int) -> List[torch.Tensor]:
        """Four persistent [32 * ph * pw, C] buffers (one per tapped layer) holding the encoder features of one window in slot order:
        the static inputs of the captured head graph.  They live as long as the packed weights, so every long-video call of this
        model replays the same graph (one pipeline per model at a time: not re-entrant, like the reference)."""
        self._weights()
        bufs = self.__dict__.setdefault("_win_bufs", {})
        key = (ph, pw, ops.operand_dtype(), self._dev)
        if key not in bufs:
            C = self.cfg["embed_dim"]
            bufs[key] = [torch.empty((INFER_LEN * ph * pw, C), dtype=ops.operand_dtype(), device=self._dev) for _ in range(4)]
        return bufs[key]

    @torch.no_grad()
    def head_from_features(self, feats: List[torch.Tensor], T: int, ph: int, pw: int, static_inputs: bool = False, clone: bool = True) -> torch.Tensor:
        """4 x [T*ph*pw, C] -> depth (T, 14*ph, 14*pw) fp32 (one video, T <= 32 frames).  ``static_inputs``: ``feats`` are persistent
        buffers handed over on every call (the long-video driver's window buffers): the captured graph reads them in place."""
        w = self._weights()
        if T > 32:
            raise RuntimeError("temporal attention supports at most 32 frames per window")
        feats = [f.contiguous() for f in feats]
        d = self._graphs.run(("head", T, ph, pw), lambda *f: head_forward(w["head"], list(f), T, ph, pw, T), feats, static_inputs=static_inputs)
        return d.clone() if clone else d

    # ------------------------------------------------------------------------------------------------
    # streaming: drop-in for video_depth_anything/video_depth_stream.py:76-160 (same class name there; one model serves both here)
    def reset_stream(self):
        self._stream = None

    def _stream_ring(self, ph: int, pw: int) -> dict:
        """Slot pools for the cached per-frame projections of the 4 motion modules x 2 attention blocks at this patch grid, the
        staging buffers of the incoming frame and the device-side slot table (persistent: the stream graph holds their addresses)."""
        rings = self.__dict__.setdefault("_rings", {})
        key = (ph, pw, ops.operand_dtype(), self._dev)
        if key not in rings:
            od, dev = ops.operand_dtype(), self._dev
            oc, Fe = self.cfg["out_channels"], self.cfg["features"]
            H4, W4 = (ph - 1) // 2 + 1, (pw - 1) // 2 + 1
            geo = [(ph * pw, oc[2]), (H4 * W4, oc[3]), (ph * pw, Fe), (4 * ph * pw, Fe)]  # (pixels, channels) of mm0..mm3
            rings[key] = {"mods": [{"pool": [torch.empty((RING_SLOTS, D, 3 * C), dtype=od, device=dev) for _ in range(2)],
                                    "staging": [torch.empty((D, 3 * C), dtype=od, device=dev) for _ in range(2)]} for D, C in geo],
                          "table": torch.full((RING_NEW_SLOT + 1,), -1, dtype=torch.int32, device=dev)}
        return rings[key]

    @torch.no_grad()
    def stream_step(self, x: torch.Tensor) -> torch.Tensor:
        """One pre-processed frame x (3, h, w) fp32 -> depth (h, w) fp32 on the device, attending to the cached history:
        frame 0, the second-oldest kept frame and the last 29 frames (video_depth_stream.py:130-158).  The cache list holds pool
        slot numbers; from the second frame on the whole step (encoder, head, cache insertion) replays as one CUDA graph whose
        only per-frame input besides the image is the 33-entry device slot table."""
        w = self._weights()
        st = getattr(self, "_stream", None)
        _, h, wd = x.shape
        ph, pw = h // 14, wd // 14
        ring = self._stream_ring(ph, pw)
        if st is None:
            st = self._stream = {"id": -1, "cache": [], "free": list(range(RING_SLOTS - 1, -1, -1)), "grid": (ph, pw)}
        elif st["grid"] != (ph, pw):
            raise RuntimeError("frame size changed mid-stream")
        st["id"] += 1
        xd = x.to(device=self._dev, dtype=torch.float32).unsqueeze(0).contiguous()
        slot = st["free"].pop()
        if st["id"] == 0:
            # the encoder of one frame has fixed shapes and touches no cache: CUDA-graph replay
            feats = self.encode_frames(xd, clone=False)
            new: list = []
            depth = head_forward(w["head"], feats, 1, ph, pw, 1, stream={"cached": None, "new": new})
            for m in range(4):
                for a in range(2):
                    ring["mods"][m]["pool"][a][slot].copy_(new[2 * m + a])
            st["cache"] = [slot] * INFER_LEN  # "copy multiple cache to simulate the windows"
            depth = depth[0]
        else:
            cl = st["cache"]
            cur = cl[0:2] + cl[-INFER_LEN + 3:]
            table = cur + [-1] * (RING_NEW_SLOT - len(cur)) + [slot]  # entry len(cur) = -1 = this frame (staging)
            ring["table"].copy_(torch.tensor(table, dtype=torch.int32))
            L = len(cur) + 1

            def run(xs):
                feats = encoder_forward(w["enc"], xs, w["head"].get("readout"))
                return head_forward(w["head"], feats, 1, ph, pw, 1, stream={"ring": ring["mods"], "table": ring["table"], "L": L})

            depth = self._graphs.run(("stream", ph, pw, L), run, [xd])[0].clone()
            st["cache"] = cl + [slot]
        gap = (INFER_LEN - OVERLAP) * 2 - 1 - (OVERLAP - INTERP_LEN)
        if st["id"] + INFER_LEN > gap + 1:
            dropped = st["cache"][1]
            st["cache"] = st["cache"][:1] + st["cache"][2:]
            if dropped not in st["cache"]:
                st["free"].append(dropped)
        return depth

    @torch.no_grad()
    def infer_video_depth_one(self, frame, input_size=518, device="cuda", fp32=False):
        """video_depth_stream.py:76: frame np.uint8 (H, W, 3) RGB -> np.float32 depth (H, W) of this frame; call once per frame."""
        from . import video as V
        if str(device).split(":")[0] != "cuda":
            raise RuntimeError("infer_video_depth_one runs on CUDA only (no CPU fallback)")
        st = getattr(self, "_stream", None)
        fh, fw = frame.shape[:2]
        if st is None:
            self._stream_size = V._resolve_input_size(fh, fw, input_size)
            self._stream_hw = (fh, fw)
        elif (fh, fw) != self._stream_hw:
            raise RuntimeError("frame size changed mid-stream")  # video_depth_stream.py:131-133 asserts
        x = torch.from_numpy(V.preprocess_frames(frame[None], self._stream_size)[0])
        d = self.stream_step(x)
        if tuple(d.shape) != (fh, fw):
            r = torch.empty((1, fh, fw), dtype=torch.float32, device=d.device)
            ops.bilinear_f32(d.unsqueeze(0).contiguous(), r, 1, d.shape[0], d.shape[1], fh, fw)
            d = r[0]
        return d.cpu().numpy()

    @torch.no_grad()
    def infer_video_depth(self, frames, target_fps, input_size=518, device="cuda", fp32=False, **kw):
        """Drop-in for video_depth.py:67-156.  frames: np.uint8 (N, H, W, 3) RGB -> (np.float32 (N, H, W), target_fps).
        Extra keyword arguments (``reuse_features``, ``group``, ``gather``, ``preprocessed``) go to video.infer_video_depth."""
        from .video import infer_video_depth
        return infer_video_depth(self, frames, target_fps, input_size=input_size, device=device, fp32=fp32, **kw)


class VideoDepthRefinerV5(_PackedModule):
    """Drop-in for models/video_depth_model_v5.py:128 ``VideoDepthAnything`` (the depth-sequence refinement model):
    forward(input_depth (B, S, H, W) in [0, max_depth]) -> refined depth (B, S, H, W).  use_residual=True, input_normal=True
    (the reference defaults) are the supported configuration.  The dense part (DINOv2 + temporal DPT head at 224x224) runs on
    the same kernels as VideoDepthAnything; the full-resolution part is three bandwidth-bound kernels (csrc/vdn_v5.cu)."""

    NET_SIZE = 224  # models/video_depth_model_v5.py:169

    def __init__(self, encoder="vitl", features=256, out_channels=(256, 512, 1024, 1024), use_bn=False, use_clstoken=False, num_frames=32,
                 max_depth=65535, pe="ape", use_residual=True, input_normal=True):
        super().__init__()
        if encoder not in ENCODER_CONFIGS:
            raise KeyError(encoder)
        if use_bn:
            raise NotImplementedError("use_bn is never exercised by the reference (SURVEY.md §8b)")
        if pe not in ("ape", "rope"):
            raise NotImplementedError(pe)
        self.use_clstoken, self.pe = bool(use_clstoken), pe
        if not use_residual or not input_normal:
            raise NotImplementedError("only use_residual=True, input_normal=True (the reference defaults) are built")
        self.encoder, self.num_frames, self.max_depth = encoder, num_frames, float(max_depth)
        self.use_residual, self.input_normal = use_residual, input_normal
        self.cfg = dict(ENCODER_CONFIGS[encoder], features=features, out_channels=list(out_channels))
        self.intermediate_layer_idx = {k: v["taps"] for k, v in ENCODER_CONFIGS.items()}

    def _expected_shapes(self):
        s = _encoder_shapes("pretrained.", self.cfg)
        s["scale_head.feat.1.weight"], s["scale_head.feat.1.bias"] = (1, 1, 1, 1), (1,)
        s.update(_head_shapes("temporal_head.", self.cfg["embed_dim"], self.cfg["features"], self.cfg["out_channels"], True, self.pe, self.use_clstoken))
        s["shift_head.0.weight"], s["shift_head.0.bias"] = (1, 1, 1, 1), (1,)
        return s

    def _pack(self, sd, dev, dt):
        return {"enc": packing.pack_encoder(sd, "pretrained.", self.cfg, dev, dt), "head": packing.pack_head(sd, "temporal_head.", self.cfg, dev, dt, True, self.pe, self.use_clstoken),
                "scale_w": float(sd["scale_head.feat.1.weight"].reshape(())), "scale_b": float(sd["scale_head.feat.1.bias"].reshape(())),
                "shift_w": float(sd["shift_head.0.weight"].reshape(())), "shift_b": float(sd["shift_head.0.bias"].reshape(()))}

    @torch.no_grad()
    def forward(self, input_depth: torch.Tensor) -> torch.Tensor:
        if input_depth.dim() != 4:
            raise RuntimeError(f"expected (B, S, H, W), got {tuple(input_depth.shape)}")
        w = self._weights()
        B, S, H, W = input_depth.shape
        if S > 32:
            raise RuntimeError("temporal attention supports at most 32 frames per window")
        dev, n = self._dev, B * S
        nh, nw = (self.NET_SIZE, self.NET_SIZE) if self.NET_SIZE else (H, W)  # v4 runs the network at the native resolution
        din = input_depth.to(device=dev, dtype=torch.float32).contiguous()
        inv_max = 1.0 / self.max_depth
        scale = _empty((n,), torch.float32, dev)
        ops.frame_median_scale(din, scale, H * W, inv_max, w["scale_w"], w["scale_b"])          # :164-167
        if (nh, nw) == (H, W):
            r = din.view(n, H, W)
        else:
            r = _empty((n, nh, nw), torch.float32, dev)
            ops.bilinear_f32(din.view(n, H, W), r, n, H, W, nh, nw)                              # :169 (scale commutes with the resize)
        x = _empty((n, 3, nh, nw), torch.float32, dev)
        ops.v5_net_input(r, scale, x, n, nh, nw, inv_max)                                        # :173-178
        feats = encoder_forward(w["enc"], x, w["head"].get("readout"))
        o = head_forward(w["head"], feats, n, nh // 14, nw // 14, S)  # temporal modules mix the S frames of each of the B sequences
        out = _empty((B, S, H, W), torch.float32, dev)
        ops.v5_residual(din, o, scale, out, n, H, W, nh, nw, w["shift_w"], w["shift_b"], self.max_depth)  # :183-192
        return out


class VideoDepthRefinerV4(VideoDepthRefinerV5):
    """Drop-in for models/video_depth_model_v4.py:88 ``VideoDepthAnything``: the v5 module tree (same state_dict keys) with the
    network evaluated at the input's own resolution (forward :120-148; H and W multiples of 14) instead of 224x224."""

    NET_SIZE = None
