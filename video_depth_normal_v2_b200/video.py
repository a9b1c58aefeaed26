"""Long-video driver: drop-in for ``VideoDepthAnything.infer_video_depth`` (video_depth_anything/video_depth.py:67-156).

Same semantics as the reference — 32-slot windows every 22 frames, first 10 slots overwritten with the previous window's
input key frames, per-window forward, least-squares scale/shift alignment on key frames, 8-frame linear cross-fade —
re-organised for the GPU:
  * every source frame is pre-processed once (the reference transforms each frame up to 32/22 times, identically);
  * the window schedule is unrolled up front (the forward of window k depends only on raw frames), so windows
    are independent units: they can be sharded across ranks (``rank`` / ``world_size``) with no data-path collective;
  * resize-to-frame-size, affine alignment, clamp and cross-fade run as device kernels on device-resident depth maps
    (the reference does them in numpy after a per-frame ``.cpu()``), with one D2H copy of the final result.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .models import INFER_LEN, INTERP_LEN, KEYFRAMES, OVERLAP

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406])
IMAGENET_STD = np.array([0.229, 0.224, 0.225])


def window_schedule(n_frames: int) -> List[List[int]]:
    """Source-frame index of every slot of every window (video_depth.py:88-102 unrolled)."""
    step = INFER_LEN - OVERLAP
    pad = (step - (n_frames % step)) % step + (INFER_LEN - step)
    padded = list(range(n_frames)) + [n_frames - 1] * pad
    windows, prev = [], None
    for f0 in range(0, n_frames, step):
        cur = [padded[f0 + i] for i in range(INFER_LEN)]
        if prev is not None:
            cur[:OVERLAP] = [prev[k] for k in KEYFRAMES]
        windows.append(cur)
        prev = cur
    return windows


def _target_size(width: int, height: int, input_size: int) -> Tuple[int, int]:
    """util/transform.py:56-105 with keep_aspect_ratio=True, ensure_multiple_of=14, resize_method='lower_bound'."""
    scale_h, scale_w = input_size / height, input_size / width
    if scale_w > scale_h:
        scale_h = scale_w
    else:
        scale_w = scale_h

    def constrain(x, min_val):
        y = int(np.round(x / 14) * 14)
        if y < min_val:
            y = int(np.ceil(x / 14) * 14)
        return y
    return constrain(scale_w * width, input_size), constrain(scale_h * height, input_size)


def preprocess_frames(frames: np.ndarray, input_size: int) -> np.ndarray:
    """uint8 RGB (N, H, W, 3) -> float32 (N, 3, h, w): /255, cubic resize to a multiple of 14, ImageNet normalise
    (video_depth.py:74-86,98-99; util/transform.py).  Host-side, like the reference (boundary)."""
    import cv2
    n, fh, fw = frames.shape[:3]
    w, h = _target_size(fw, fh, input_size)
    out = np.empty((n, 3, h, w), np.float32)
    for i in range(n):
        img = frames[i].astype(np.float32) / 255.0
        img = cv2.resize(img, (w, h), interpolation=cv2.INTER_CUBIC)
        img = (img - IMAGENET_MEAN) / IMAGENET_STD  # float64, as in NormalizeImage
        out[i] = np.transpose(img, (2, 0, 1)).astype(np.float32)
    return out


def _solve_scale_shift(sums: Sequence[float]) -> Tuple[float, float]:
    """utils/util.py:40-62 from the five sums (a00, a01, a11, b0, b1)."""
    a00, a01, a11, b0, b1 = [float(v) for v in sums]
    det = a00 * a11 - a01 * a01
    if det != 0:
        return (a11 * b0 - a01 * b1) / det, (-a01 * b0 + a00 * b1) / det
    return 1.0, 0.0


class WindowAligner:
    """Sequential affine alignment + cross-fade of window outputs on the device (video_depth.py:118-154)."""

    def __init__(self, n_windows: int, H: int, W: int, device):
        self.H, self.W = H, W
        n_out = INFER_LEN + (INFER_LEN - OVERLAP) * (n_windows - 1)
        self.aligned = torch.empty((n_out, H, W), dtype=torch.float32, device=device)
        self.pos = 0
        self.ref = None  # [2, H, W]: depth of frame 0 from window 0, aligned key-frame 12 of the previous window
        self.sums = torch.empty(5, dtype=torch.float64, device=device)
        self.ss = torch.empty(2, dtype=torch.float32, device=device)
        self.w = [0.0] + [i * (1.0 / (INTERP_LEN - 1)) for i in range(1, INTERP_LEN - 1)] + [1.0]  # utils/util.py:65-70

    def push(self, d: torch.Tensor):
        """d: [32, H, W] fp32 device tensor of one window (already resized to the output size)."""
        align_len = OVERLAP - INTERP_LEN
        kf = KEYFRAMES[:align_len]
        if self.pos == 0:
            self.aligned[:INFER_LEN].copy_(d)
            self.pos = INFER_LEN
            self.ref = torch.stack([d[k] for k in kf]).contiguous()
            return
        ops.lsq_sums(d[:align_len].contiguous(), self.ref, self.sums)
        scale, shift = _solve_scale_shift(self.sums.cpu().tolist())
        self.ss.copy_(torch.tensor([scale, shift], dtype=torch.float32))
        for i in range(INTERP_LEN):
            tgt = self.aligned[self.pos - INTERP_LEN + i]
            ops.crossfade(tgt, d[align_len + i], tgt, self.ss, self.w[i])
        n_new = INFER_LEN - OVERLAP
        ops.affine_clamp(d[OVERLAP:], self.aligned[self.pos:self.pos + n_new], self.ss)
        self.pos += n_new
        ops.affine_clamp(d[kf[1]], self.ref[1], self.ss)

    def result(self, n_frames: int) -> torch.Tensor:
        return self.aligned[:n_frames]


@torch.no_grad()
def window_depths(model, frames_t: torch.Tensor, windows: Sequence[Sequence[int]], out_hw: Tuple[int, int], device) -> List[torch.Tensor]:
    """Forward every window in ``windows`` (lists of source-frame indices into ``frames_t`` (N,3,h,w) pinned host fp32)
    and resize to ``out_hw``.  Returns device tensors [32, H, W]."""
    outs = []
    h, w = frames_t.shape[-2:]
    H, W = out_hw
    for win in windows:
        idx = torch.as_tensor(win, dtype=torch.long)
        x = frames_t.index_select(0, idx)
        x = x.pin_memory() if device != "cpu" and not x.is_pinned() else x
        xd = x.to(device, non_blocking=True).unsqueeze(0)
        d = model.forward(xd)[0]  # [32, h, w]
        if (H, W) != (h, w):
            r = torch.empty((d.shape[0], H, W), dtype=torch.float32, device=d.device)
            ops.bilinear_f32(d.contiguous(), r, d.shape[0], h, w, H, W)
            d = r
        outs.append(d)
    return outs


@torch.no_grad()
def infer_video_depth(model, frames, target_fps, input_size=518, device="cuda", fp32=False, preprocessed: Optional[torch.Tensor] = None):
    """video_depth.py:67-156.  ``fp32`` is accepted for signature compatibility: this path always accumulates in fp32 and its
    16-bit operands meet the fp32-reference tolerance.  Returns (np.float32 (N, H, W), target_fps)."""
    if str(device).split(":")[0] != "cuda":
        raise RuntimeError("infer_video_depth runs on CUDA only (no CPU fallback)")
    frames = np.asarray(frames)
    n = frames.shape[0]
    fh, fw = frames.shape[1:3]
    ratio = max(fh, fw) / min(fh, fw)
    if ratio > 1.78:  # video_depth.py:68-72
        input_size = int(input_size * 1.777 / ratio)
        input_size = round(input_size / 14) * 14
    ft = preprocessed if preprocessed is not None else torch.from_numpy(preprocess_frames(frames, input_size))
    windows = window_schedule(n)
    model_dev = model._dev if model._dev.type == "cuda" else torch.device(device)
    if model._dev.type != "cuda":
        model.to(model_dev)
    aligner = WindowAligner(len(windows), fh, fw, model_dev)
    for win in windows:
        d = window_depths(model, ft, [win], (fh, fw), model_dev)[0]
        aligner.push(d)
    return aligner.result(n).cpu().numpy(), target_fps
