"""TEST INFRASTRUCTURE — deterministic random-init recipe shared by oracle, reference and product.

The reference has no checkpoints available offline and its default init is degenerate
(every ``proj_out`` / ``ZeroConv`` is zero and ~50 % of seeds give an
all-zero depth map).  This module generates a ``state_dict`` with *exactly the reference's
key names and shapes* (checked by ``tests/golden/gen_golden.py`` with
``load_state_dict(strict=True)`` on the live reference) from a seeded CPU generator, so the
GPU box — where ``/root/reference`` does not exist — can rebuild identical weights.

Recipe (stated, applied identically to every side of a comparison):
  * ViT linears: trunc-normal(std 0.02) weights, normal(std 0.02) biases
    (reference: trunc_normal_/zeros, video_depth_anything/dinov2.py:398-403 init_weights_vit_timm).
  * LayerNorm / GroupNorm affine: weight 1 + 0.1 n, bias 0.1 n (exercises the affine path).
  * LayerScale gamma: U(0.5, 1.0) (reference init 1.0, dinov2.py:412).
  * convs / other linears: PyTorch default U(+-1/sqrt(fan_in)) for weight and bias.
  * motion-module ``proj_out``: de-zeroed with the default linear init (reference zeroes it,
    motion_module/motion_module.py:57-58).
  * last 1x1 conv (``output_conv2.2``): |w|, bias 0.05 -> strictly positive depth, so relative
    errors are well defined.
  * v5 ``ZeroConv``s: normal(0, 0.5) (reference zero, models/video_depth_model_v5.py:55-61).
  * DA2 memory block: attention / MLP linears default-linear init, embeddings trunc-normal(0.02) as the reference
    (memory_block.py:52-61), ConvNeXt layer scale ``gamma`` U(0.5, 1.0) (reference 1e-6, which would make the
    fuser an identity map, sam2/modeling/memory_encoder.py:86-90).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

ENCODERS = {
    # name: embed_dim, depth, heads, taps, features, out_channels   (run_video.py:28-33, video_depth.py:48-51)
    "vits": dict(embed_dim=384, depth=12, heads=6, taps=[2, 5, 8, 11], features=64, out_channels=[48, 96, 192, 384]),
    "vitl": dict(embed_dim=1024, depth=24, heads=16, taps=[4, 11, 17, 23], features=256, out_channels=[256, 512, 1024, 1024]),
    # DepthAnythingV2 only (depth_anything_v2.py:24-29, dinov2.py:353-364); VideoDepthAnything knows vits / vitl (video_depth.py:48-51)
    "vitb": dict(embed_dim=768, depth=12, heads=12, taps=[2, 5, 8, 11], features=128, out_channels=[96, 192, 384, 768]),
    # ViT-g (dinov2.py:381-395, run.py:32): SwiGLU FFN, hidden = round_up_8(int(4 C * 2 / 3)) = 4096
    "vitg": dict(embed_dim=1536, depth=40, heads=24, taps=[9, 19, 29, 39], features=384, out_channels=[1536, 1536, 1536, 1536], ffn="swiglu"),
}
POS_GRID = 37  # DINOv2(img_size=518, patch 14) -> 37x37 learned pos-embed (dinov2.py:407-409)
NUM_FRAMES = 32


class _Gen:
    def __init__(self, seed: int):
        self.g = torch.Generator(device="cpu").manual_seed(seed)

    def normal(self, shape, std=1.0, mean=0.0):
        return torch.randn(shape, generator=self.g, dtype=torch.float32) * std + mean

    def trunc_normal(self, shape, std=0.02):
        return (torch.randn(shape, generator=self.g, dtype=torch.float32).clamp_(-2.0, 2.0)) * std

    def uniform(self, shape, lo, hi):
        return torch.rand(shape, generator=self.g, dtype=torch.float32) * (hi - lo) + lo


def _norm_affine(sd, g, name, c):
    sd[name + ".weight"] = 1.0 + g.normal((c,), 0.1)
    sd[name + ".bias"] = g.normal((c,), 0.1)


def _vit_linear(sd, g, name, out_f, in_f):
    sd[name + ".weight"] = g.trunc_normal((out_f, in_f), 0.02)
    sd[name + ".bias"] = g.normal((out_f,), 0.02)


def _default_linear(sd, g, name, out_f, in_f, bias=True):
    b = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = g.uniform((out_f, in_f), -b, b)
    if bias:
        sd[name + ".bias"] = g.uniform((out_f,), -b, b)


def _conv(sd, g, name, out_c, in_c, k, bias=True):
    b = 1.0 / math.sqrt(in_c * k * k)
    sd[name + ".weight"] = g.uniform((out_c, in_c, k, k), -b, b)
    if bias:
        sd[name + ".bias"] = g.uniform((out_c,), -b, b)


def _convT(sd, g, name, in_c, out_c, k):
    # nn.ConvTranspose2d weight is (in, out, kh, kw); torch's fan_in uses size(1)*k*k
    b = 1.0 / math.sqrt(out_c * k * k)
    sd[name + ".weight"] = g.uniform((in_c, out_c, k, k), -b, b)
    sd[name + ".bias"] = g.uniform((out_c,), -b, b)


def sinusoid_pe(d_model: int, max_len: int = NUM_FRAMES) -> torch.Tensor:
    """motion_module/motion_module.py:195-209 (PositionalEncoding buffer ``pe``)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def _encoder(sd, g, prefix, cfg):
    C, depth = cfg["embed_dim"], cfg["depth"]
    sd[prefix + "cls_token"] = g.normal((1, 1, C), 0.02)
    sd[prefix + "pos_embed"] = g.trunc_normal((1, POS_GRID * POS_GRID + 1, C), 0.02)
    sd[prefix + "mask_token"] = torch.zeros(1, C)
    _conv(sd, g, prefix + "patch_embed.proj", C, 3, 14)
    for i in range(depth):
        p = f"{prefix}blocks.{i}."
        _norm_affine(sd, g, p + "norm1", C)
        _vit_linear(sd, g, p + "attn.qkv", 3 * C, C)
        _vit_linear(sd, g, p + "attn.proj", C, C)
        sd[p + "ls1.gamma"] = g.uniform((C,), 0.5, 1.0)
        _norm_affine(sd, g, p + "norm2", C)
        if cfg.get("ffn") == "swiglu":
            Hd = (int(4 * C * 2 / 3) + 7) // 8 * 8
            _vit_linear(sd, g, p + "mlp.w12", 2 * Hd, C)
            _vit_linear(sd, g, p + "mlp.w3", C, Hd)
        else:
            _vit_linear(sd, g, p + "mlp.fc1", 4 * C, C)
            _vit_linear(sd, g, p + "mlp.fc2", C, 4 * C)
        sd[p + "ls2.gamma"] = g.uniform((C,), 0.5, 1.0)
    _norm_affine(sd, g, prefix + "norm", C)


def _dpt_head(sd, g, prefix, cfg):
    C, F, oc = cfg["embed_dim"], cfg["features"], cfg["out_channels"]
    for i in range(4):
        _conv(sd, g, f"{prefix}projects.{i}", oc[i], C, 1)
    _convT(sd, g, prefix + "resize_layers.0", oc[0], oc[0], 4)
    _convT(sd, g, prefix + "resize_layers.1", oc[1], oc[1], 2)
    _conv(sd, g, prefix + "resize_layers.3", oc[3], oc[3], 3)
    for i in range(4):
        _conv(sd, g, f"{prefix}scratch.layer{i + 1}_rn", F, oc[i], 3, bias=False)
    for r in (1, 2, 3, 4):
        p = f"{prefix}scratch.refinenet{r}."
        _conv(sd, g, p + "out_conv", F, F, 1)
        for u in (1, 2):
            _conv(sd, g, f"{p}resConfUnit{u}.conv1", F, F, 3)
            _conv(sd, g, f"{p}resConfUnit{u}.conv2", F, F, 3)
    _conv(sd, g, prefix + "scratch.output_conv1", F // 2, F, 3)
    _conv(sd, g, prefix + "scratch.output_conv2.0", 32, F // 2, 3)
    _conv(sd, g, prefix + "scratch.output_conv2.2", 1, 32, 1)
    sd[prefix + "scratch.output_conv2.2.weight"] = sd[prefix + "scratch.output_conv2.2.weight"].abs()
    sd[prefix + "scratch.output_conv2.2.bias"] = torch.full((1,), 0.05)


def _motion_modules(sd, g, prefix, cfg, pe="ape"):
    oc, F = cfg["out_channels"], cfg["features"]
    for m, C in enumerate([oc[2], oc[3], F, F]):
        p = f"{prefix}motion_modules.{m}.temporal_transformer."
        _norm_affine(sd, g, p + "norm", C)
        _default_linear(sd, g, p + "proj_in", C, C)
        tb = p + "transformer_blocks.0."
        for a in range(2):
            ab = f"{tb}attention_blocks.{a}."
            _default_linear(sd, g, ab + "to_q", C, C, bias=False)
            _default_linear(sd, g, ab + "to_k", C, C, bias=False)
            _default_linear(sd, g, ab + "to_v", C, C, bias=False)
            _default_linear(sd, g, ab + "to_out.0", C, C)
            if pe == "ape":  # pe='rope' keeps freqs_cis as a plain attribute, not in the state_dict (motion_module.py:236-240)
                sd[ab + "pos_encoder.pe"] = sinusoid_pe(C)
        for a in range(2):
            _norm_affine(sd, g, f"{tb}norms.{a}", C)
        _default_linear(sd, g, tb + "ff.net.0.proj", 8 * C, C)
        _default_linear(sd, g, tb + "ff.net.2", C, 4 * C)
        _norm_affine(sd, g, tb + "ff_norm", C)
        _default_linear(sd, g, p + "proj_out", C, C)  # de-zeroed


def _memory_block(sd, g, prefix, cfg, max_len=6, layers=4):
    """depth_anything_v2/memory_block.py:13-82 (MemoryAttention x4, MemoryEncoder with two MaskDownSamplers and a 2-layer fuser)."""
    C = cfg["embed_dim"]
    sd[prefix + "curr_pos_enc"] = g.trunc_normal((1, 1, C), 0.02)
    sd[prefix + "maskmem_tpos_enc"] = g.trunc_normal((1, max_len, C), 0.02)
    sd[prefix + "no_mem_embed"] = g.trunc_normal((1, 1, C), 0.02)
    for l in range(layers):
        p = f"{prefix}memory_attention.layers.{l}."
        for att in ("self_attn", "cross_attn_image"):
            for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
                _default_linear(sd, g, f"{p}{att}.{proj}", C, C)
        _default_linear(sd, g, p + "linear1", 2 * C, C)
        _default_linear(sd, g, p + "linear2", C, 2 * C)
        for n in ("norm1", "norm2", "norm3"):
            _norm_affine(sd, g, p + n, C)
    _norm_affine(sd, g, prefix + "memory_attention.norm", C)
    me = prefix + "memory_encoder."
    _conv(sd, g, me + "mask_downsampler.0.encoder.0", 4, 1, 3)
    _norm_affine(sd, g, me + "mask_downsampler.0.encoder.1", 4)
    _conv(sd, g, me + "mask_downsampler.0.encoder.3", 1, 4, 1)
    _conv(sd, g, me + "mask_downsampler.1.encoder.0", 49, 1, 7)
    _norm_affine(sd, g, me + "mask_downsampler.1.encoder.1", 49)
    _conv(sd, g, me + "mask_downsampler.1.encoder.3", 1, 49, 1)
    _conv(sd, g, me + "pix_feat_proj", C, C, 1)
    for l in range(2):
        p = f"{me}fuser.layers.{l}."
        sd[p + "gamma"] = g.uniform((C,), 0.5, 1.0)
        b = 1.0 / math.sqrt(49)
        sd[p + "dwconv.weight"] = g.uniform((C, 1, 7, 7), -b, b)
        sd[p + "dwconv.bias"] = g.uniform((C,), -b, b)
        _norm_affine(sd, g, p + "norm", C)
        _default_linear(sd, g, p + "pwconv1", 4 * C, C)
        _default_linear(sd, g, p + "pwconv2", C, 4 * C)


def make_state_dict(model: str, encoder: str, seed: int = 0, use_clstoken: bool = False, pe: str = "ape") -> "OrderedDict[str, torch.Tensor]":
    """model in {'vda', 'v5', 'da2'}; returns fp32 CPU tensors under the reference's key names.

    'vda' = video_depth_anything/video_depth.py:35-56 (prefixes ``pretrained.`` / ``head.``)
    'v5'  = models/video_depth_model_v5.py:128-158 (``pretrained.`` / ``temporal_head.`` /
            ``scale_head.feat.1.`` / ``shift_head.0.``)
    'da2' = depth_anything_v2/depth_anything_v2.py:12-43 (``pretrained.`` / ``memory_block.`` / ``depth_head.``)
    """
    cfg = ENCODERS[encoder]
    g = _Gen(seed * 7919 + {"vda": 1, "v5": 2, "da2": 3}[model])
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _encoder(sd, g, "pretrained.", cfg)
    if model == "vda":
        _dpt_head(sd, g, "head.", cfg)
        _motion_modules(sd, g, "head.", cfg, pe)
        if use_clstoken:  # dpt.py:92-98; drawn last so the other tensors do not depend on the switch
            C = cfg["embed_dim"]
            for i in range(4):
                _default_linear(sd, g, f"head.readout_projects.{i}.0", C, 2 * C)
    elif model == "v5":
        sd["scale_head.feat.1.weight"] = g.normal((1, 1, 1, 1), 0.5)
        sd["scale_head.feat.1.bias"] = g.normal((1,), 0.5)
        _dpt_head(sd, g, "temporal_head.", cfg)
        _motion_modules(sd, g, "temporal_head.", cfg)
        sd["shift_head.0.weight"] = g.normal((1, 1, 1, 1), 0.5)
        sd["shift_head.0.bias"] = g.normal((1,), 0.5)
    elif model == "da2":
        _memory_block(sd, g, "memory_block.", cfg)
        _dpt_head(sd, g, "depth_head.", cfg)
        if use_clstoken:
            C = cfg["embed_dim"]
            for i in range(4):
                _default_linear(sd, g, f"depth_head.readout_projects.{i}.0", C, 2 * C)
    else:
        raise ValueError(model)
    return sd


def make_input(kind: str, shape, seed: int = 0) -> torch.Tensor:
    """Seeded synthetic inputs. kind='rgb': N(0,1) post-normalisation frames;
    kind='depth': smooth positive depth maps in [0, 65535] (low-pass noise so Sobel normals are non-trivial)."""
    g = torch.Generator(device="cpu").manual_seed(seed * 104729 + 17)
    if kind == "rgb":
        return torch.randn(shape, generator=g, dtype=torch.float32)
    if kind == "depth":
        B, S, H, W = shape
        low = torch.rand((B * S, 1, max(H // 16, 2), max(W // 16, 2)), generator=g)
        d = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=True)
        d = (d - d.amin()) / (d.amax() - d.amin() + 1e-12)
        return (d.reshape(B, S, H, W) * 0.8 + 0.1) * 65535.0
    raise ValueError(kind)
