"""Long-video driver: drop-in for ``VideoDepthAnything.infer_video_depth`` (video_depth_anything/video_depth.py:67-156).

Same semantics as the reference — 32-slot windows every 22 frames, first 10 slots overwritten with the previous window's
input key frames, per-window forward, least-squares scale/shift alignment on key frames, 8-frame linear cross-fade —
re-organised for the GPU as a three-stream pipeline:

  copy-in stream   raw uint8 frames of window k+1, host -> device (one async copy per contiguous run of frames)
  compute stream   cubic resize + normalise (vdn_preprocess_u8) -> ViT on the frames the window does not share with its
                   predecessor (the encoder is per-frame, so a steady-state window encodes 22 frames, not 32) -> temporal
                   head on all 32 slots -> resize to the frame size -> scale/shift fit -> ONE kernel that writes every
                   output frame the window owns (affine + clamp + cross-fade, vdn_window_finalize)
  copy-out stream  the frames window k finalised, device -> pinned host result, while window k+1 computes

Windows are independent units (the forward of window k depends only on raw frames): ``shard=True`` / ``group=``
partitions them over the ranks of a ``torch.distributed`` group in contiguous blocks.  NCCL is used for
exactly three things: the 9 overlap key-frame feature sets at each rank boundary, the three key-frame depth maps per window
that the (sequential) scale/shift chain needs, and the previous rank's last 8 raw depth maps for the boundary cross-fade.
Every rank copies the frames it owns to its own host memory (``gather='shard'``); ``'rank0'`` / ``'all'`` assemble the one
array the reference returns, through a shared host segment when the ranks share a node, else through NCCL.
Sharding is opt-in: a plain drop-in call never communicates (the reference's never does).
"""
from __future__ import annotations

import mmap
import os
import zlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .models import INFER_LEN, INTERP_LEN, KEYFRAMES, OVERLAP

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406])
IMAGENET_STD = np.array([0.229, 0.224, 0.225])
STEP = INFER_LEN - OVERLAP          # 22 new frames per window
ALIGN_LEN = OVERLAP - INTERP_LEN    # 2 key frames used for the scale/shift fit
KF_ALIGN = KEYFRAMES[:ALIGN_LEN]    # slots [0, 12] of the previous window


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


def partition_windows(n_windows: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous blocks of windows per rank, sizes differing by at most one."""
    base, extra = divmod(n_windows, world_size)
    bounds, k = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        bounds.append((k, k + n))
        k += n
    return bounds


def owned_output_range(k: int, n_windows: int, n_out: int) -> Tuple[int, int]:
    """Output frames finalised by window k: window 0 owns [0, 24); window k >= 1 owns the 8 frames it cross-fades with its
    predecessor plus its 14 untouched new frames, [22k+2, 22k+24); the last window also owns its final 8 frames."""
    lo = 0 if k == 0 else STEP * k + ALIGN_LEN
    hi = STEP * k + INFER_LEN - INTERP_LEN
    if k == n_windows - 1:
        hi = STEP * k + INFER_LEN
    return min(lo, n_out), min(hi, n_out)


def rank_output_range(bounds: Sequence[Tuple[int, int]], rank: int, n_windows: int, n_out: int) -> Tuple[int, int]:
    """Output frames owned by the windows of one rank (empty for a rank without windows)."""
    k0, k1 = bounds[rank]
    if k1 <= k0:
        return 0, 0
    return owned_output_range(k0, n_windows, n_out)[0], owned_output_range(k1 - 1, n_windows, n_out)[1]


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


def preprocess_frames(frames: np.ndarray, input_size: int, indices: Optional[Sequence[int]] = None, pinned: bool = False):
    """uint8 RGB (N, H, W, 3) -> float32 (N, 3, h, w): /255, cubic resize to a multiple of 14, ImageNet normalise
    (video_depth.py:74-86,98-99; util/transform.py).  The reference's host-side transform, kept for ``device_preprocess=False``.
    ``indices``: only these frames are transformed, into consecutive rows of the result, in the given order.
    ``pinned``: write into page-locked memory and return a torch tensor (so that the per-window H2D copies are asynchronous)."""
    import cv2
    n, fh, fw = frames.shape[:3]
    w, h = _target_size(fw, fh, input_size)
    sel = list(range(n)) if indices is None else list(indices)
    if pinned:
        out_t = torch.empty((len(sel), 3, h, w), dtype=torch.float32, pin_memory=True)
        out = out_t.numpy()
    else:
        out = np.empty((len(sel), 3, h, w), np.float32)
    for row, i in enumerate(sel):
        img = frames[i].astype(np.float32) / 255.0
        img = cv2.resize(img, (w, h), interpolation=cv2.INTER_CUBIC)
        img = (img - IMAGENET_MEAN) / IMAGENET_STD  # float64, as in NormalizeImage
        out[row] = np.transpose(img, (2, 0, 1)).astype(np.float32)
    return out_t if pinned else out


def _solve_scale_shift(sums: Sequence[float]) -> Tuple[float, float]:
    """utils/util.py:40-62 from the five sums (a00, a01, a11, b0, b1) — host mirror of the vdn_lsq_solve kernel."""
    a00, a01, a11, b0, b1 = [float(v) for v in sums]
    det = a00 * a11 - a01 * a01
    if det != 0:
        return (a11 * b0 - a01 * b1) / det, (-a01 * b0 + a00 * b1) / det
    return 1.0, 0.0


CROSSFADE_W = [0.0] + [i * (1.0 / (INTERP_LEN - 1)) for i in range(1, INTERP_LEN - 1)] + [1.0]  # utils/util.py:65-70


class DeviceAlignOps:
    """The alignment primitives on device tensors, through the C ABI (no host synchronisation)."""

    def scale_shift(self, pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        sums = torch.empty(5, dtype=torch.float64, device=pred.device)
        ss = torch.empty(2, dtype=torch.float32, device=pred.device)
        ops.lsq_sums(pred.contiguous(), target.contiguous(), sums)
        ops.lsq_solve(sums, ss)
        return ss

    def affine_clamp(self, x: torch.Tensor, ss: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        out = torch.empty_like(x) if out is None else out
        ops.affine_clamp(x.contiguous(), out, ss)
        return out

    def keys(self, d: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """out [3, H, W] = slots (0, 1, 12) of a window's raw depth d [32, H, W]."""
        return ops.window_keys(d, out)

    def finalize(self, d: torch.Tensor, prev_tail: Optional[torch.Tensor], ss: Optional[torch.Tensor], ss_prev: Optional[torch.Tensor],
                 out: torch.Tensor, first_slot: int, is_first: bool) -> torch.Tensor:
        """out [count, H, W] = output frames of slots first_slot.. of one window: cross-fade (slots 2..9) / affine + clamp."""
        return ops.window_finalize(d, prev_tail, ss, ss_prev, out, first_slot, out.shape[0], is_first)


def scale_shift_chain(keys: torch.Tensor, aops) -> torch.Tensor:
    """The sequential part of the alignment (video_depth.py:131-152).  keys: [K, 3, H, W] = raw depth of slots (0, 1, 12)
    of every window.  Returns [K, 2] (scale, shift) per window (window 0: identity)."""
    K = keys.shape[0]
    table = torch.zeros((K, 2), dtype=torch.float32, device=keys.device)
    table[0, 0] = 1.0
    ref = torch.stack([keys[0, 0], keys[0, 2]]).contiguous()  # depth of frame 0, aligned key frame 12 of the previous window
    for k in range(1, K):
        ss = aops.scale_shift(keys[k, :ALIGN_LEN], ref)
        table[k].copy_(ss)
        aops.affine_clamp(keys[k, 2], ss, out=ref[1])
    return table


class _Phases:
    """Optional per-phase timing of a call (``stats``): CUDA events on the compute stream plus host wall-clock stamps."""

    def __init__(self, device, enabled: bool):
        self.on = enabled and torch.device(device).type == "cuda"
        self.marks = []
        self.device = device

    def mark(self, name: str):
        if not self.on:
            return
        import time
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        self.marks.append((name, ev, time.perf_counter()))

    def report(self) -> dict:
        if not self.on or len(self.marks) < 2:
            return {}
        torch.cuda.synchronize(self.device)
        out = {}
        for (n0, e0, t0), (n1, e1, t1) in zip(self.marks[:-1], self.marks[1:]):
            out[n1] = {"gpu_ms": round(e0.elapsed_time(e1), 3), "host_ms": round((t1 - t0) * 1e3, 3)}
        return out


class _PinnedPool:
    """Page-locked result buffers that outlive a call.  Locking fresh pages costs ~0.6 ms per MB (measured on the B200 hosts: 120-220 ms
    for the 190 MB a rank owns of a 350-frame clip), more than the copies themselves, so a buffer goes back to this pool when the array
    handed to the caller — and every view of it — has been garbage-collected, and the next call of the same size takes it from here."""

    def __init__(self, max_bytes: int = 24 << 30):
        self.free: Dict[int, List[torch.Tensor]] = {}
        self.max_bytes = max_bytes
        self.pooled = 0

    def acquire(self, numel: int) -> torch.Tensor:
        lst = self.free.get(numel)
        if lst:
            self.pooled -= numel * 4
            return lst.pop()
        return torch.empty((numel,), dtype=torch.float32, pin_memory=True)

    def release(self, t: torch.Tensor):
        if self.pooled + t.numel() * 4 > self.max_bytes:
            return  # dropped: torch frees the pages
        self.free.setdefault(t.numel(), []).append(t)
        self.pooled += t.numel() * 4

    def export(self, t: torch.Tensor, shape) -> np.ndarray:
        """The numpy array handed to the caller; ``t`` returns to the pool when the array and all views of it are gone."""
        import weakref
        base = t.numpy()  # every later view / slice keeps `base` alive through .base
        weakref.finalize(base, self.release, t)
        return base.reshape(shape)

    def clear(self):
        self.free.clear()
        self.pooled = 0


pinned_pool = _PinnedPool()


def reserve_host_result(n_frames: int, H: int, W: int):
    """Page-lock a result buffer of this size ahead of the first call that needs it (a service does this once at start-up)."""
    pinned_pool.release(pinned_pool.acquire(n_frames * H * W))


class HostSink:
    """Device -> host copies of finalised output frames on their own stream, into one page-locked result (from ``pinned_pool``)."""

    def __init__(self, n_rows: int, H: int, W: int, device, host: Optional[torch.Tensor] = None):
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.shape = (n_rows, H, W)
        self.flat = None
        if host is not None:
            self.host = host
        elif self.cuda:
            self.flat = pinned_pool.acquire(n_rows * H * W)
            self.host = self.flat.view(n_rows, H, W)
        else:
            self.host = torch.empty((n_rows, H, W), dtype=torch.float32)
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.bytes = 0

    def result(self) -> np.ndarray:
        """Wait for the copies and hand the result over as a numpy array (the pinned block returns to the pool when it is dropped)."""
        self.finish()
        if self.flat is not None:
            return pinned_pool.export(self.flat, self.shape)
        return self.host.numpy()

    def push(self, rows: torch.Tensor, row0: int):
        if rows.shape[0] == 0:
            return
        dst = self.host[row0:row0 + rows.shape[0]]
        self.bytes += rows.numel() * 4
        if not self.cuda:
            dst.copy_(rows)
            return
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            dst.copy_(rows, non_blocking=True)

    def finish(self) -> torch.Tensor:
        if self.cuda:
            self.stream.synchronize()
        return self.host


class WindowAligner:
    """Sequential affine alignment + cross-fade of window outputs on the device (video_depth.py:118-154), streaming: window k's
    owned output frames are final as soon as window k has been pushed (its scale/shift needs only windows < k)."""

    def __init__(self, n_windows: int, H: int, W: int, device, aops=None, n_frames: Optional[int] = None, sink: Optional[HostSink] = None):
        self.K, self.H, self.W = n_windows, H, W
        self.n_out = INFER_LEN + STEP * (n_windows - 1) if n_frames is None else n_frames
        self.aligned = torch.empty((self.n_out, H, W), dtype=torch.float32, device=device)
        self.k = 0
        self.ref = None
        self.prev = None      # raw depth of the previous window (its slots 24..31 are the cross-fade operands)
        self.ss_prev = None   # None while the previous window is window 0 (never re-scaled)
        self.aops = aops or DeviceAlignOps()
        self.sink = sink

    def push(self, d: torch.Tensor):
        """d: [32, H, W] fp32 device tensor of one window (already resized to the output size)."""
        k = self.k
        ss = None
        if k == 0:
            self.ref = torch.stack([d[j] for j in KF_ALIGN]).contiguous()
        else:
            ss = self.aops.scale_shift(d[:ALIGN_LEN], self.ref)
            self.aops.affine_clamp(d[KF_ALIGN[1]], ss, out=self.ref[1])
        lo, hi = owned_output_range(k, self.K, self.n_out)
        if hi > lo:
            out = self.aligned[lo:hi]
            tail = self.prev[INFER_LEN - INTERP_LEN:] if k > 0 else None
            self.aops.finalize(d, tail, ss, self.ss_prev, out, lo - STEP * k, k == 0)
            if self.sink is not None:
                self.sink.push(out, lo)
        self.prev, self.ss_prev = d, ss
        self.k += 1

    def result(self, n_frames: int) -> torch.Tensor:
        return self.aligned[:n_frames]


class FrameSource:
    """The clip as the caller handed it over — raw uint8 RGB (N, H, W, 3) or pre-processed fp32 (N, 3, h, w); numpy or torch;
    pageable, page-locked or already on the device — and the upload of selected frames on a copy stream.  Nothing is copied
    on the host: a page-locked array is read by the DMA engine directly, a pageable one goes through the driver's staging."""

    def __init__(self, frames, device, frame_rows: Optional[Dict[int, int]] = None):
        if isinstance(frames, torch.Tensor):
            t = frames
        else:
            a = np.asarray(frames)
            if not a.flags["C_CONTIGUOUS"]:
                a = np.ascontiguousarray(a)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")  # a read-only array is fine: the frames are only read
                t = torch.from_numpy(a)
        if t.dtype not in (torch.uint8, torch.float32):
            raise RuntimeError(f"frames must be uint8 (raw RGB) or float32 (pre-processed), got {t.dtype}")
        self.t = t
        self.raw = t.dtype == torch.uint8
        self.device = torch.device(device)
        self.frame_rows = frame_rows
        self.resident = t.is_cuda
        self.stream = torch.cuda.Stream(device=self.device) if (self.device.type == "cuda" and not self.resident) else None
        self.h2d_bytes = 0

    def upload(self, idx: Sequence[int]) -> Tuple[torch.Tensor, Optional["torch.cuda.Event"]]:
        """-> (device tensor of the listed frames in order, event to wait for before reading it)."""
        rows = [self.frame_rows[f] for f in idx] if self.frame_rows is not None else list(idx)
        if self.resident:
            if rows and rows == list(range(rows[0], rows[0] + len(rows))):
                return self.t[rows[0]:rows[0] + len(rows)], None
            return self.t[torch.tensor(rows, device=self.t.device)], None
        x = torch.empty((len(rows),) + tuple(self.t.shape[1:]), dtype=self.t.dtype, device=self.device)
        self.h2d_bytes += x.numel() * x.element_size()

        def copy_runs():
            pos = 0
            while pos < len(rows):
                run = 1
                while pos + run < len(rows) and rows[pos + run] == rows[pos] + run:
                    run += 1
                x[pos:pos + run].copy_(self.t[rows[pos]:rows[pos] + run], non_blocking=True)
                pos += run

        if self.stream is None:
            copy_runs()
            return x, None
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)  # x's memory may have been in use by earlier work on the compute stream
        with torch.cuda.stream(self.stream):
            copy_runs()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        x.record_stream(self.stream)
        return x, ev


class WindowForwarder:
    """Forwards windows (lists of source-frame indices) through the model and resizes to ``out_hw``.  With ``reuse`` the tapped
    encoder features of one window live in slot order in persistent buffers (the static inputs of the captured head graph): a new
    window moves the rows of the frames it shares with its predecessor (slots 0, 12, 24..31 -> 0..9), encodes the rest and writes
    them into their slots — three device copies per tap and window in the steady state."""

    def __init__(self, model, frames, out_hw: Tuple[int, int], device, reuse: bool = True,
                 frame_rows: Optional[Dict[int, int]] = None, net_hw: Optional[Tuple[int, int]] = None):
        """``frames``: anything FrameSource takes.  Raw uint8 frames need ``net_hw`` = (h, w), the network input size: they are
        uploaded as bytes and resized / normalised on the device (``vdn_preprocess_u8``), 4x less H2D traffic and no host cv2 loop.
        ``frame_rows``: source-frame index -> row of ``frames`` when only a rank's shard of the clip was handed over."""
        self.model, self.out_hw, self.device = model, out_hw, torch.device(device)
        self.src = frames if isinstance(frames, FrameSource) else FrameSource(frames, device, frame_rows)
        self.raw = self.src.raw
        self.net_hw = tuple(net_hw) if net_hw is not None else tuple(self.src.t.shape[-2:])
        self.reuse = reuse and hasattr(model, "encode_frames")
        self.win_feats: Optional[List[torch.Tensor]] = None  # per tap [32*P, C], slot order of ``self.prev_win``
        self.prev_win: Optional[List[int]] = None
        self.seeded: Dict[int, List[torch.Tensor]] = {}      # frame -> per tap [P, C] features computed on another rank
        self.encoded_frames = 0  # bookkeeping: encoder work actually done (frames)
        self._pending = None     # (window, upload) issued ahead of its forward

    # ---- host -> device ------------------------------------------------------------------------------------------------
    def _need(self, win: Sequence[int]) -> List[int]:
        if not self.reuse:
            return list(win)
        have = set(self.prev_win or ()) | set(self.seeded)
        return [f for f in dict.fromkeys(win) if f not in have]

    def prefetch(self, win: Sequence[int], arriving: Sequence[int] = ()):
        """Start the upload of the frames ``win`` will need — called right after the kernels of the window before it were queued, so
        the copy overlaps them.  ``arriving``: frames whose features another rank is about to deliver (not uploaded)."""
        skip = set(arriving)
        need = [f for f in self._need(win) if f not in skip]
        self._pending = (list(win), need, self.src.upload(need) if need else (None, None))

    def _load(self, win: Sequence[int]) -> Tuple[List[int], Optional[torch.Tensor]]:
        need = self._need(win)
        if self._pending is not None and self._pending[0] == list(win) and self._pending[1] == need:
            x, ev = self._pending[2]
        else:
            x, ev = self.src.upload(need) if need else (None, None)
        self._pending = None
        if x is None:
            return need, None
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            x.record_stream(torch.cuda.current_stream(self.device))
        if self.raw:
            out = torch.empty((len(need), 3) + self.net_hw, dtype=torch.float32, device=self.device)
            ops.preprocess_u8(x, out, self.net_hw[0], self.net_hw[1])
            return need, out
        return need, x

    # ---- rank-boundary feature exchange (collective (1) of sharded_video_depth) -------------------------------------------
    def export_features(self, frames: Sequence[int]) -> List[torch.Tensor]:
        """Packed copies [len(frames)*P, C] per tap of frames of the last forwarded window (what a rank boundary ships)."""
        slots = [self.prev_win.index(f) for f in frames]
        out = []
        for t in range(4):
            P = self.win_feats[t].shape[0] // INFER_LEN
            v = self.win_feats[t].view(INFER_LEN, P, -1)
            out.append(torch.cat([v[s] for s in slots]).contiguous() if slots != list(range(slots[0], slots[0] + len(slots)))
                       else v[slots[0]:slots[0] + len(slots)].reshape(len(slots) * P, -1).clone())
        return out

    def feature_buffers(self, n_frames: int) -> List[torch.Tensor]:
        """Empty receive buffers shaped like ``export_features`` of n_frames frames."""
        return [torch.empty((n_frames * (t.shape[0] // INFER_LEN), t.shape[1]), dtype=t.dtype, device=t.device) for t in self.win_feats]

    def import_features(self, frames: Sequence[int], packed: List[torch.Tensor]):
        """Adopt features computed elsewhere (the next rank's copy of this rank's last key frames)."""
        P = [t.shape[0] // len(frames) for t in packed]
        for j, f in enumerate(frames):
            self.seeded[f] = [packed[t][j * P[t]:(j + 1) * P[t]] for t in range(4)]

    # ---- one window ---------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, win: Sequence[int]) -> torch.Tensor:
        win = list(win)
        h, w = self.net_hw
        H, W = self.out_hw
        if not self.reuse:
            _, x = self._load(win)
            d = self.model.forward(x.unsqueeze(0))[0]
            self.encoded_frames += len(win)
        else:
            need, x = self._load(win)
            new = self.model.encode_frames(x, clone=False) if need else None
            self.encoded_frames += len(need)
            if self.win_feats is None:
                if hasattr(self.model, "window_feature_buffers"):  # persistent per model: the captured head graph reads them in place
                    self.win_feats = self.model.window_feature_buffers(h // 14, w // 14)
                else:
                    P = [t.shape[0] // len(need) for t in new]
                    self.win_feats = [torch.empty((INFER_LEN * P[t], new[t].shape[1]), dtype=new[t].dtype, device=new[t].device) for t in range(4)]
            prev_slot = {f: s for s, f in reversed(list(enumerate(self.prev_win)))} if self.prev_win is not None else {}
            new_row = {f: i for i, f in enumerate(need)}
            # (kind, source index) of every slot; slots that already hold their frame are skipped; consecutive slots coalesce
            moves = []
            for s, f in enumerate(win):
                if f in prev_slot:
                    if prev_slot[f] != s:
                        moves.append(("prev", prev_slot[f], s))
                elif f in new_row:
                    moves.append(("new", new_row[f], s))
                else:
                    moves.append(("seed", f, s))
            # in-place moves inside the window buffers are safe when no move reads a slot that an earlier move wrote
            written, hazard = set(), False
            for kind, src, dst in moves:
                if kind == "prev" and src in written:
                    hazard = True
                written.add(dst)
            for t in range(4):
                buf = self.win_feats[t]
                P = buf.shape[0] // INFER_LEN
                old = buf.clone() if hazard else buf
                i = 0
                while i < len(moves):
                    kind, src, dst = moves[i]
                    run = 1
                    while (kind != "seed" and i + run < len(moves) and moves[i + run][0] == kind and moves[i + run][1] == src + run
                           and moves[i + run][2] == dst + run):
                        run += 1
                    if kind == "prev":
                        buf[dst * P:(dst + run) * P].copy_(old[src * P:(src + run) * P])
                    elif kind == "new":
                        buf[dst * P:(dst + run) * P].copy_(new[t][src * P:(src + run) * P])
                    else:
                        buf[dst * P:(dst + 1) * P].copy_(self.seeded[src][t])
                    i += run
            self.prev_win = win
            self.seeded = {}
            # the head graph's output buffer is overwritten by the next window: keep a copy unless the resize below makes one anyway
            d = self.model.head_from_features(self.win_feats, len(win), h // 14, w // 14, static_inputs=True, clone=(H, W) == (h, w))
        if (H, W) != (h, w):
            r = torch.empty((d.shape[0], H, W), dtype=torch.float32, device=d.device)
            ops.bilinear_f32(d.contiguous(), r, d.shape[0], h, w, H, W)
            d = r
        return d


def _resolve_input_size(fh: int, fw: int, input_size: int) -> int:
    ratio = max(fh, fw) / min(fh, fw)
    if ratio > 1.78:  # video_depth.py:68-72
        input_size = int(input_size * 1.777 / ratio)
        input_size = round(input_size / 14) * 14
    return input_size


class FrameShard(np.ndarray):
    """The output frames one rank owns (``gather='shard'``): a float32 (hi-lo, H, W) array with ``frame_range = (lo, hi)``."""
    frame_range: Tuple[int, int] = (0, 0)


def _frames_fingerprint(frames, n: int) -> int:
    """Cheap content check that every rank of a sharded call was handed the same clip: CRC of three frames."""
    crc = 0
    for i in sorted({0, n // 2, n - 1}):
        f = frames[i]
        f = f.cpu().numpy() if isinstance(f, torch.Tensor) else np.asarray(f)
        crc = zlib.crc32(np.ascontiguousarray(f).view(np.uint8).reshape(-1), crc)
    return crc


@torch.no_grad()
def infer_video_depth(model, frames, target_fps, input_size=518, device="cuda", fp32=False, preprocessed: Optional[torch.Tensor] = None,
                      reuse_features: bool = True, group=None, shard: Optional[bool] = None, gather: str = "all", device_preprocess: bool = True,
                      output: str = "numpy", stats: Optional[dict] = None):
    """video_depth.py:67-156.  ``fp32`` is accepted for signature compatibility: this path always accumulates in fp32 and its
    16-bit operands meet the fp32-reference tolerance.  Returns (np.float32 (N, H, W), target_fps).

    ``frames``: uint8 RGB (N, H, W, 3), numpy or torch, pageable, page-locked or already on the device.
    ``device_preprocess``: upload the raw uint8 frames and run the cubic resize + normalisation on the GPU (default); False keeps the
    reference's host-side cv2 transform.  ``preprocessed``: already transformed fp32 (N, 3, h, w) frames (skips both).

    Sharding over ``torch.distributed`` is opt-in (``shard=True`` for the default group, or ``group=``): every rank of the group must
    call with the same clip (checked).  ``gather='all'`` returns the full result on every rank, ``'rank0'`` only on rank 0 (None
    elsewhere), ``'shard'`` the frames the rank owns as a ``FrameShard`` (no gather at all).  Without ``shard`` / ``group`` the
    call never communicates, whatever process groups exist.  ``output='device'`` leaves the result on the GPU (a torch tensor; with
    ``shard`` only ``gather='shard'``).  ``stats`` (optional dict) receives bookkeeping of the call."""
    if str(device).split(":")[0] != "cuda":
        raise RuntimeError("infer_video_depth runs on CUDA only (no CPU fallback)")
    import torch.distributed as dist
    if not isinstance(frames, torch.Tensor):
        frames = np.asarray(frames)
    n = frames.shape[0]
    fh, fw = frames.shape[1:3]
    input_size = _resolve_input_size(fh, fw, input_size)
    model_dev = model._dev if model._dev.type == "cuda" else torch.device(device)
    if model._dev.type != "cuda":
        model.to(model_dev)
    windows = window_schedule(n)
    sharded = bool(shard) or group is not None
    if sharded and not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("shard=True needs an initialised torch.distributed process group")
    if gather not in ("all", "rank0", "shard"):
        raise ValueError(f"gather must be 'all', 'rank0' or 'shard', got {gather!r}")
    if output not in ("numpy", "device"):
        raise ValueError(f"output must be 'numpy' or 'device', got {output!r}")
    net_hw = None
    if preprocessed is not None:
        source = preprocessed
    elif device_preprocess:
        source, net_hw = frames, _target_size(fw, fh, input_size)[::-1]
    else:
        source = None  # host transform below, only for the frames this rank needs
    with torch.cuda.device(model_dev):
        if sharded and dist.get_world_size(group) > 1:
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            sig = torch.tensor([n, fh, fw, input_size, _frames_fingerprint(frames, n)], dtype=torch.int64, device=model_dev)
            sigs = torch.empty((world, 5), dtype=torch.int64, device=model_dev)
            dist.all_gather_into_tensor(sigs, sig, group=group)
            if not bool((sigs == sig).all()):
                raise RuntimeError("sharded infer_video_depth: the ranks of the group were handed different clips "
                                   f"(n, H, W, input_size, crc per rank: {sigs.tolist()}); pass shard=False for independent per-rank videos")
            k0, k1 = partition_windows(len(windows), world)[rank]
            frame_rows = None
            if source is None:
                mine = sorted({f for win in windows[k0:k1] for f in win})
                frame_rows = {f: i for i, f in enumerate(mine)}
                source = preprocess_frames(frames, input_size, indices=mine, pinned=True)
            fwd = WindowForwarder(model, source, (fh, fw), model_dev, reuse=reuse_features, frame_rows=frame_rows, net_hw=net_hw)
            out = sharded_video_depth(fwd.forward, windows, n, (fh, fw), model_dev, DeviceAlignOps(), group=group, gather=gather,
                                      forwarder=fwd if fwd.reuse else None, to_host=output == "numpy", stats=stats)
            if stats is not None:
                stats.update(encoded_frames=fwd.encoded_frames, h2d_bytes=fwd.src.h2d_bytes)
            return out, target_fps
        if source is None:
            source = preprocess_frames(frames, input_size, pinned=True)
        fwd = WindowForwarder(model, source, (fh, fw), model_dev, reuse=reuse_features, net_hw=net_hw)
        sink = HostSink(n, fh, fw, model_dev) if output == "numpy" else None
        aligner = WindowAligner(len(windows), fh, fw, model_dev, n_frames=n, sink=sink)
        fwd.prefetch(windows[0])
        for k, win in enumerate(windows):
            d = fwd.forward(win)
            if k + 1 < len(windows):
                fwd.prefetch(windows[k + 1])  # H2D of the next window's frames while this one computes
            aligner.push(d)
        if stats is not None:
            stats.update(encoded_frames=fwd.encoded_frames, h2d_bytes=fwd.src.h2d_bytes, d2h_bytes=sink.bytes if sink is not None else 0, windows=len(windows))
        if sink is None:
            return aligner.result(n), target_fps
        out = sink.result()
        if sharded and gather == "shard":
            out = out.view(FrameShard)
            out.frame_range = (0, n)
        return out, target_fps


_SHM_SEQ = [0]
_PAGE = mmap.PAGESIZE


class _SharedHostResult:
    """One host array visible to every rank of a single-node group: a POSIX shared-memory segment mapped by all of them, each
    rank page-locking only the part it will fill, so that every rank copies the frames it owns straight to their final place over
    its own PCIe link.  ``tensor`` is None when the ranks are not on one node or the segment cannot be created / registered
    (callers fall back to NCCL).  The set-up is collective; it runs while the windows are being forwarded."""

    def __init__(self, n_out: int, H: int, W: int, own_rows: Tuple[int, int], group, device):
        import socket
        import torch.distributed as dist
        self.tensor, self.mm, self._registered, self.reason = None, None, None, ""
        rank = dist.get_rank(group)
        nbytes = n_out * H * W * 4
        name, ok, fd = None, 1, None
        if rank == 0:
            _SHM_SEQ[0] += 1
            name = f"/dev/shm/vdn_{os.getpid()}_{_SHM_SEQ[0]}"
            try:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize < nbytes + (64 << 20):
                    raise OSError("not enough room in /dev/shm")
                fd = os.open(name, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, nbytes)
            except OSError as exc:
                ok, self.reason = 0, f"create: {exc}"
        info = [(socket.gethostname(), name, ok)] if rank == 0 else [None]
        dist.broadcast_object_list(info, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        host0, name, ok = info[0]
        t = None
        if ok and socket.gethostname() == host0:
            try:
                if rank != 0:
                    fd = os.open(name, os.O_RDWR)
                self.mm = mmap.mmap(fd, nbytes)
                os.close(fd)
                t = torch.frombuffer(self.mm, dtype=torch.float32).view(n_out, H, W)
            except (OSError, ValueError) as exc:
                t, self.reason = None, f"map: {exc}"
        elif ok:
            self.reason = "ranks on different hosts"
        if t is not None and device.type == "cuda" and own_rows[1] > own_rows[0]:
            lo_b = own_rows[0] * H * W * 4 // _PAGE * _PAGE
            hi_b = -(-(own_rows[1] * H * W * 4) // _PAGE) * _PAGE  # whole pages: the mapping extends to the end of its last page
            rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr() + lo_b, hi_b - lo_b, 0)
            if int(rc) != 0:
                t, self.reason = None, f"cudaHostRegister({hi_b - lo_b} bytes): {rc}"
            else:
                self._registered = t.data_ptr() + lo_b
        flag = torch.tensor([1 if t is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0 and t is not None:
            self.reason = "another rank could not map / register the segment"
        if rank == 0 and name is not None and ok:
            try:
                os.unlink(name)  # the mappings keep the segment alive; nothing is left behind in /dev/shm
            except OSError:
                pass
        if t is None or int(flag.item()) == 0:
            self.release()
            t = None
        self.tensor = t

    def release(self):
        if self._registered is not None:
            torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None


def sharded_video_depth(forward, windows: Sequence[Sequence[int]], n_frames: int, out_hw: Tuple[int, int], device, aops, group=None,
                        gather: str = "all", forwarder: Optional[WindowForwarder] = None, to_host: bool = False, stats: Optional[dict] = None):
    """Window-sharded long-video inference over ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests, which inject a
    stand-in ``forward`` / ``aops`` / ``forwarder``).  ``forward(win) -> [32, H, W]`` depth of one window at the output size.

    Collectives: (1) boundary key-frame features, rank r+1 -> rank r, so that rank r's last window does not
    re-encode the 9 frames rank r+1 needs anyway [only with a ``forwarder``]; (2) all-gather of the slots (0, 1, 12) depth
    maps of every window — every rank then runs the same sequential scale/shift chain on the same data up to its own last
    window, so all ranks apply bit-identical coefficients without a broadcast; (3) the last window's raw slots 24..31,
    rank r -> r+1, for the cross-fade at the boundary.  No output frame crosses NVLink for ``gather='shard'``; ``'rank0'`` /
    ``'all'`` assemble one array (``to_host``: in a shared host segment on one node, else on the device through NCCL).

    Returns a device tensor (``to_host=False``; the CPU tests) or a numpy array / ``FrameShard`` / None (``to_host=True``)."""
    import torch.distributed as dist
    device = torch.device(device)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    K = len(windows)
    H, W = out_hw
    bounds = partition_windows(K, world)
    k0, k1 = bounds[rank]
    mine = list(range(k0, k1))
    nxt = rank + 1 if rank + 1 < world and bounds[rank + 1][0] < bounds[rank + 1][1] else None
    prv = rank - 1 if rank > 0 and mine else None

    def _grank(r):  # group rank -> global rank for p2p
        return dist.get_global_rank(group, r) if group is not None else r

    n_out = n_frames
    ranges = [rank_output_range(bounds, r, K, n_out) for r in range(world)]
    lo, hi = ranges[rank]
    ph = _Phases(device, stats is not None)
    ph.mark("start")
    shared = _SharedHostResult(n_out, H, W, (lo, hi), group, device) if (to_host and gather in ("rank0", "all")) else None

    ph.mark("shared_host_setup")
    # ---- forward all own windows; (1) key-frame feature exchange at the rank boundaries ---------------------------------
    per_rank = max(b[1] - b[0] for b in bounds)
    keys = torch.zeros((per_rank, 3, H, W), dtype=torch.float32, device=device)
    depths: List[torch.Tensor] = []
    recv_req, recv_buf, recv_frames, send_req, sent = None, None, None, [], None
    prefetch = getattr(forwarder, "prefetch", None) if forwarder is not None else None
    if prefetch is not None and mine:
        prefetch(windows[k0])
    for i, k in enumerate(mine):
        last = i == len(mine) - 1
        if forwarder is not None and last and i > 0 and recv_req is not None:
            for rq in recv_req:
                rq.wait()
            forwarder.import_features(recv_frames, recv_buf)
            recv_req = None
        d = forward(windows[k])
        depths.append(d)
        aops.keys(d, keys[i])
        if forwarder is not None and i == 0:
            if prv is not None:  # ship the features of slots 1..9 (= window k0-1 slots 12, 24..31) to the previous rank
                sent = forwarder.export_features(windows[k][1:OVERLAP])
                send_req = [dist.isend(t, _grank(prv), group=group) for t in sent]
            if nxt is not None:
                recv_frames = list(windows[k1][1:OVERLAP])
                recv_buf = forwarder.feature_buffers(len(recv_frames))
                recv_req = [dist.irecv(t, _grank(nxt), group=group) for t in recv_buf]
        if prefetch is not None and not last:
            # H2D of the next window's frames while this one computes; the last window will not need the frames whose features
            # arrive from the next rank
            prefetch(windows[k + 1], arriving=recv_frames if (i + 1 == len(mine) - 1 and recv_req is not None) else ())
    if recv_req is not None:  # single-window rank: the features arrived too late to be useful, but the receive must complete
        for rq in recv_req:
            rq.wait()

    ph.mark("forward_windows")
    # ---- (3) boundary cross-fade operands: posted as soon as the last window is queued -----------------------------------------
    prev_tail = None
    reqs = list(send_req)
    if prv is not None and k0 > 0:
        prev_tail = torch.empty((INTERP_LEN, H, W), dtype=torch.float32, device=device)
        reqs.append(dist.irecv(prev_tail, _grank(prv), group=group))
    if nxt is not None and mine:
        tail = depths[-1][INFER_LEN - INTERP_LEN:].contiguous()
        reqs.append(dist.isend(tail, _grank(nxt), group=group))

    # ---- (2) scale/shift chain, redundantly on every rank up to its own last window -----------------------------------------------------
    all_keys = torch.empty((world * per_rank, 3, H, W), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(all_keys, keys, group=group)
    table = None
    if mine:
        parts = [all_keys[r * per_rank: r * per_rank + (min(b[1], k1) - b[0])] for r, b in enumerate(bounds) if b[0] < k1 and b[1] > b[0]]
        flat = parts[0] if len(parts) == 1 else torch.cat(parts)
        table = scale_shift_chain(flat, aops)
    for rq in reqs:
        rq.wait()
    ph.mark("exchange_and_chain")

    # ---- finalise the owned output frames: one launch per window, device -> host while the next window is finalised ------------------------
    sink, row_base = None, lo
    if to_host:
        if shared is not None and shared.tensor is not None:
            sink, row_base = HostSink(n_out, H, W, device, host=shared.tensor), 0
        elif gather == "shard":
            sink = HostSink(max(hi - lo, 0), H, W, device)
    ph.mark("host_result_alloc")
    shard = torch.empty((max(hi - lo, 0), H, W), dtype=torch.float32, device=device)
    for i, k in enumerate(mine):
        o_lo, o_hi = owned_output_range(k, K, n_out)
        if o_hi <= o_lo:
            continue
        out = shard[o_lo - lo:o_hi - lo]
        tail = None
        if k > 0:
            tail = prev_tail if i == 0 else depths[i - 1][INFER_LEN - INTERP_LEN:]
        aops.finalize(depths[i], tail, table[k] if k > 0 else None, table[k - 1] if k - 1 > 0 else None, out, o_lo - STEP * k, k == 0)
        if sink is not None:
            sink.push(out, o_lo - row_base)
    use_shared = shared is not None and shared.tensor is not None
    ph.mark("finalize")
    if sink is not None and stats is not None:
        sink.finish()
        ph.mark("d2h_tail")
    if stats is not None:
        stats.update(windows=len(mine), frame_range=(lo, hi), d2h_bytes=sink.bytes if sink is not None else 0,
                     gather_path="shard" if gather == "shard" else ("shared-host" if use_shared else "nccl"),
                     gather_note=shared.reason if shared is not None else "", phases=ph.report())

    if gather == "shard":
        if not to_host:
            return shard
        res = sink.result().view(FrameShard)
        res.frame_range = (lo, hi)
        return res
    if use_shared:
        sink.finish()
        dist.barrier(group=group)  # every rank's part has landed in the shared segment
        shared.release()
        if gather == "rank0" and rank != 0:
            return None
        return np.frombuffer(shared.mm, dtype=np.float32).reshape(n_out, H, W)  # the array keeps the mapping alive

    # ---- (4) NCCL gather (ranks on different nodes, or no shared segment) ---------------------------------------------------------------
    full = None
    if rank == 0:
        full = torch.empty((n_out, H, W), dtype=torch.float32, device=device)
        full[ranges[0][0]:ranges[0][1]].copy_(shard)
        rq = [dist.irecv(full[a:b2], _grank(r), group=group) for r, (a, b2) in enumerate(ranges) if r > 0 and b2 > a]
        for q in rq:
            q.wait()
    elif shard.shape[0] > 0:
        dist.send(shard, _grank(0), group=group)
    if gather == "all":
        if full is None:
            full = torch.empty((n_out, H, W), dtype=torch.float32, device=device)
        dist.broadcast(full, _grank(0), group=group)
    if not to_host:
        return full
    return full.cpu().numpy() if full is not None else None
