"""One-time, host-side conversion of the reference's ``state_dict`` (SURVEY.md §8b key names) into kernel layouts:
K-major 16-bit GEMM weights, tap-major 3x3 conv filters padded to 64 input channels per tap, pixel-shuffle ConvTranspose
weights, fused q|k|v temporal projections, (value, gate)-interleaved GEGLU weights, fp32 biases / norm affines.

Pure data movement on weights (no activations touch this file); the bicubic pos-embed interpolation is the reference's
own weights-only computation (dinov2.py:179-210) evaluated once per patch grid."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _w16(t: torch.Tensor, dev, dt) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).to(dt).contiguous()


def pack_linear(sd, name, dev, dt, bias=True):
    d = {"w": _w16(sd[name + ".weight"], dev, dt)}
    if bias and (name + ".bias") in sd:
        d["b"] = _f32(sd[name + ".bias"], dev)
    return d


def pack_conv1x1(sd, name, dev, dt):
    w = sd[name + ".weight"]
    return {"w": _w16(w.reshape(w.shape[0], w.shape[1]), dev, dt), "b": _f32(sd[name + ".bias"], dev)}


def pack_conv3x3(sd, name, dev, dt, bias=True):
    """(Co, Ci, 3, 3) -> [Co, 9 * roundup(Ci, 64)], column = (r*3+s)*Cip + ci (implicit-GEMM tap-major order)."""
    w = sd[name + ".weight"].detach().float()
    co, ci = w.shape[:2]
    cip = (ci + 63) // 64 * 64
    p = torch.zeros(co, 9, cip)
    p[:, :, :ci] = w.permute(0, 2, 3, 1).reshape(co, 9, ci)
    d = {"w": _w16(p.reshape(co, 9 * cip), dev, dt), "ci": ci, "co": co}
    if bias:
        d["b"] = _f32(sd[name + ".bias"], dev)
    return d


def pack_conv_tail(sd, name, dev, dt):
    """(32, 128, 3, 3) -> the B operand of vdn_conv_tail as the tensor core reads it from shared memory: for each vertical tap dy and
    each 64-channel chunk a K-major tile [96 rows = (dx, co)][64 ci] of 128-byte rows, 16-byte chunk j of row n stored at j ^ (n & 7)
    (the 128B swizzle TMA would apply).  73728 bytes."""
    w = sd[name + ".weight"].detach().float()
    co, ci = w.shape[:2]
    if (co, ci) != (32, 128):
        raise ValueError(f"vdn_conv_tail is built for the 128 -> 32 output convolution, got {ci} -> {co}")
    t = w.permute(2, 3, 0, 1).reshape(3, 96, 2, 8, 8)          # [dy][n = dx*32 + co][chunk][j][8 ci]
    t = t.permute(0, 2, 1, 3, 4).contiguous()                   # [dy][chunk][n][j][8]
    n = torch.arange(96, device=t.device).view(1, 1, 96, 1, 1)
    j = torch.arange(8, device=t.device).view(1, 1, 1, 8, 1)
    src_j = (j ^ (n & 7)).expand(3, 2, 96, 8, 8)                # position j holds logical chunk j ^ (n & 7)
    return _w16(torch.gather(t, 3, src_j).reshape(-1), dev, dt)


def pack_conv3x3_im2col(sd, name, dev, dt):
    """(Co, Ci, 3, 3) -> [Co, 9*Ci] matching vdn_im2col_3x3_s2's column order (tap-major, unpadded)."""
    w = sd[name + ".weight"].detach().float()
    co, ci = w.shape[:2]
    return {"w": _w16(w.permute(0, 2, 3, 1).reshape(co, 9 * ci), dev, dt), "b": _f32(sd[name + ".bias"], dev)}


def pack_conv_transpose(sd, name, dev, dt, s):
    """ConvTranspose2d (Ci, Co, s, s), kernel == stride -> GEMM weight [(i*s+j)*Co + co, ci] + expanded bias."""
    w = sd[name + ".weight"].detach().float()
    ci, co = w.shape[:2]
    return {"w": _w16(w.permute(2, 3, 1, 0).reshape(s * s * co, ci), dev, dt), "b": _f32(sd[name + ".bias"].detach().float().repeat(s * s), dev), "co": co}


def pack_encoder(sd: Dict[str, torch.Tensor], prefix: str, cfg: dict, dev, dt) -> dict:
    C = cfg["embed_dim"]
    enc = {"C": C, "heads": cfg["heads"], "depth": cfg["depth"], "taps": list(cfg["taps"])}
    pw = sd[prefix + "patch_embed.proj.weight"].detach().float().reshape(C, 588)
    enc["patch_w"] = _w16(F.pad(pw, (0, 4)), dev, dt)  # K 588 -> 592 (16-byte row pitch); pad columns are zero
    enc["patch_b"] = _f32(sd[prefix + "patch_embed.proj.bias"], dev)
    enc["cls"] = _f32(sd[prefix + "cls_token"].reshape(C), dev)
    enc["pos_embed_raw"] = sd[prefix + "pos_embed"].detach().float().cpu()
    enc["pos_cache"] = {}
    blocks = []
    swiglu = cfg.get("ffn") == "swiglu"
    if swiglu:
        enc["hidden"] = sd[f"{prefix}blocks.0.mlp.w3.weight"].shape[1]
    for i in range(cfg["depth"]):
        p = f"{prefix}blocks.{i}."
        blk = {
            "ln1_w": _f32(sd[p + "norm1.weight"], dev), "ln1_b": _f32(sd[p + "norm1.bias"], dev),
            "qkv": pack_linear(sd, p + "attn.qkv", dev, dt), "proj": pack_linear(sd, p + "attn.proj", dev, dt),
            "ls1": _f32(sd[p + "ls1.gamma"], dev),
            "ln2_w": _f32(sd[p + "norm2.weight"], dev), "ln2_b": _f32(sd[p + "norm2.bias"], dev),
            "ls2": _f32(sd[p + "ls2.gamma"], dev),
        }
        if swiglu:
            # x12 = w12(x); x1, x2 = x12.chunk(2); hidden = silu(x1) * x2 (swiglu_ffn.py:29-33).  The GLU epilogue computes
            # value * act(gate) on interleaved (value, gate) output columns: row 2i <- x2_i (value), row 2i+1 <- x1_i (gate)
            w12, b12 = sd[p + "mlp.w12.weight"].detach().float(), sd[p + "mlp.w12.bias"].detach().float()
            Hd = w12.shape[0] // 2
            wi = torch.stack([w12[Hd:], w12[:Hd]], dim=1).reshape(2 * Hd, -1)
            bi = torch.stack([b12[Hd:], b12[:Hd]], dim=1).reshape(2 * Hd)
            blk["w12"] = {"w": _w16(wi, dev, dt), "b": _f32(bi, dev)}
            blk["w3"] = pack_linear(sd, p + "mlp.w3", dev, dt)
        else:
            blk["fc1"], blk["fc2"] = pack_linear(sd, p + "mlp.fc1", dev, dt), pack_linear(sd, p + "mlp.fc2", dev, dt)
        blocks.append(blk)
    enc["blocks"] = blocks
    enc["norm_w"], enc["norm_b"] = _f32(sd[prefix + "norm.weight"], dev), _f32(sd[prefix + "norm.bias"], dev)
    return enc


def encoder_pos_embed(enc: dict, ph: int, pw: int, dev) -> torch.Tensor:
    """fp32 [(1 + ph*pw), C] positional embedding for the requested grid (dinov2.py:179-210), cached."""
    key = (ph, pw)
    if key not in enc["pos_cache"]:
        pe = enc["pos_embed_raw"]
        N = pe.shape[1] - 1
        if not (ph * pw == N and ph == pw):
            s = math.sqrt(N)
            C = pe.shape[-1]
            sx, sy = float(ph + 0.1) / s, float(pw + 0.1) / s
            patch = F.interpolate(pe[:, 1:].reshape(1, int(s), int(s), C).permute(0, 3, 1, 2), scale_factor=(sx, sy), mode="bicubic", antialias=False)
            assert patch.shape[-2] == ph and patch.shape[-1] == pw
            pe = torch.cat((pe[:, :1], patch.permute(0, 2, 3, 1).reshape(1, -1, C)), dim=1)
        enc["pos_cache"][key] = pe[0].to(device=dev, dtype=torch.float32).contiguous()
    return enc["pos_cache"][key]


def temporal_rope_table(C: int, max_len: int = 32, theta: float = 10000.0) -> torch.Tensor:
    """cos|sin table of the temporal RoPE (precompute_freqs_cis(query_dim, temporal_max_len), motion_module/attention.py:403-408) laid
    out for vdn_rope_chunks over the q|k columns: [max_len, 2C/64, 64] = per (frame, 64-channel chunk) cos[32] | sin[32]; the k half
    repeats the q half (one frequency per channel pair of query_dim).  Weights-free, fp64 once."""
    if C % 64 != 0:
        raise RuntimeError("pe='rope' needs motion-module widths that are multiples of 64")
    freqs = 1.0 / (theta ** (torch.arange(0, C, 2, dtype=torch.float32)[: C // 2] / C))  # fp32 like the reference
    ang = torch.outer(torch.arange(max_len, dtype=torch.float32), freqs)                # [max_len, C/2]
    cos, sin = ang.cos().reshape(max_len, C // 64, 32), ang.sin().reshape(max_len, C // 64, 32)
    tab = torch.cat((cos, sin), dim=-1)                                                  # [max_len, C/64, 64]
    return torch.cat((tab, tab), dim=1).contiguous()                                     # q chunks then k chunks


def pack_motion_module(sd, prefix: str, C: int, dev, dt, pe_type: str = "ape") -> dict:
    p = prefix + "temporal_transformer."
    tb = p + "transformer_blocks.0."
    mm = {"C": C, "gn_w": _f32(sd[p + "norm.weight"], dev), "gn_b": _f32(sd[p + "norm.bias"], dev),
          "proj_in": pack_linear(sd, p + "proj_in", dev, dt), "proj_out": pack_linear(sd, p + "proj_out", dev, dt), "attn": []}
    for a in range(2):
        ab = f"{tb}attention_blocks.{a}."
        qkv = torch.cat([sd[ab + "to_q.weight"], sd[ab + "to_k.weight"], sd[ab + "to_v.weight"]], dim=0)
        blk = {"ln_w": _f32(sd[f"{tb}norms.{a}.weight"], dev), "ln_b": _f32(sd[f"{tb}norms.{a}.bias"], dev),
               "qkv_w": _w16(qkv, dev, dt), "out": pack_linear(sd, ab + "to_out.0", dev, dt)}
        if pe_type == "ape":
            pe = sd[ab + "pos_encoder.pe"][0].detach().double()
            blk["pe"] = _f32(sd[ab + "pos_encoder.pe"][0], dev)  # (32, C)
            # streaming path: positional part of the bias-free projections, W (n + pe_j) = W n + W pe_j (weights only, fp64 once)
            blk["pos_qkv"] = _f32((pe @ qkv.detach().double().t()).float(), dev)  # (32, 3C)
        else:
            blk["pe"] = None
            blk["rope"] = mm.setdefault("rope_table", temporal_rope_table(C).to(dev))  # (32, 2C/64, 64)
            # streaming + rope: the reference rotates by freqs_cis[:1] (angle 0) broadcast over all keys (motion_module.py:279-282
            # with a one-frame query), i.e. no positional term at all
            blk["pos_qkv"] = torch.zeros((32, 3 * C), dtype=torch.float32, device=dev)
        mm["attn"].append(blk)
    w1, b1 = sd[tb + "ff.net.0.proj.weight"].detach().float(), sd[tb + "ff.net.0.proj.bias"].detach().float()
    h = w1.shape[0] // 2  # rows [0,h) = value, [h,2h) = gate (GEGLU.chunk(2), motion_module/attention.py:382-384)
    mm["ff1_w"] = _w16(torch.stack([w1[:h], w1[h:]], dim=1).reshape(2 * h, -1), dev, dt)
    mm["ff1_b"] = _f32(torch.stack([b1[:h], b1[h:]], dim=1).reshape(2 * h), dev)
    mm["ff2"] = pack_linear(sd, tb + "ff.net.2", dev, dt)
    mm["ffn_w"], mm["ffn_b"] = _f32(sd[tb + "ff_norm.weight"], dev), _f32(sd[tb + "ff_norm.bias"], dev)
    return mm


def pack_head(sd, prefix: str, cfg: dict, dev, dt, temporal: bool, pe_type: str = "ape", use_clstoken: bool = False) -> dict:
    oc, Fe = cfg["out_channels"], cfg["features"]
    s = prefix + "scratch."
    head = {"features": Fe, "oc": list(oc)}
    head["projects"] = [pack_conv1x1(sd, f"{prefix}projects.{i}", dev, dt) for i in range(4)]
    head["resize0"] = pack_conv_transpose(sd, prefix + "resize_layers.0", dev, dt, 4)
    head["resize1"] = pack_conv_transpose(sd, prefix + "resize_layers.1", dev, dt, 2)
    head["resize3"] = pack_conv3x3_im2col(sd, prefix + "resize_layers.3", dev, dt)
    head["rn"] = [pack_conv3x3(sd, f"{s}layer{i + 1}_rn", dev, dt, bias=False) for i in range(4)]
    head["refine"] = {}
    for r in (1, 2, 3, 4):
        p = f"{s}refinenet{r}."
        head["refine"][r] = {
            "out_conv": pack_conv1x1(sd, p + "out_conv", dev, dt),
            "rcu1": (pack_conv3x3(sd, p + "resConfUnit1.conv1", dev, dt), pack_conv3x3(sd, p + "resConfUnit1.conv2", dev, dt)),
            "rcu2": (pack_conv3x3(sd, p + "resConfUnit2.conv1", dev, dt), pack_conv3x3(sd, p + "resConfUnit2.conv2", dev, dt)),
        }
    head["oc1"] = pack_conv3x3(sd, s + "output_conv1", dev, dt)
    head["oc2"] = pack_conv3x3(sd, s + "output_conv2.0", dev, dt)
    if tuple(sd[s + "output_conv2.0.weight"].shape[:2]) == (32, 128):
        head["oc2_tail"] = pack_conv_tail(sd, s + "output_conv2.0", dev, dt)
    head["oc2_head_w"] = _f32(sd[s + "output_conv2.2.weight"].reshape(32), dev)
    head["oc2_head_b"] = float(sd[s + "output_conv2.2.bias"].reshape(()).item())
    if temporal:
        head["mm"] = [pack_motion_module(sd, f"{prefix}motion_modules.{m}.", C, dev, dt, pe_type) for m, C in enumerate([oc[2], oc[3], Fe, Fe])]
    if use_clstoken:  # dpt.py:92-98: Linear(2C -> C) + GELU on [token | cls]
        head["readout"] = [pack_linear(sd, f"{prefix}readout_projects.{i}.0", dev, dt) for i in range(4)]
    return head
