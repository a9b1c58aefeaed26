"""Long-video driver: drop-in for ``VideoDepthAnything.infer_video_depth`` (video_depth_anything/video_depth.py:67-156).

Same semantics as the reference — 32-slot windows every 22 frames, first 10 slots overwritten with the previous window's
input key frames, per-window forward, least-squares scale/shift alignment on key frames, 8-frame linear cross-fade —
re-organised for the GPU:
  * every source frame is pre-processed once (the reference transforms each frame up to 32/22 times, identically);
  * the encoder is per-frame, so the features of the 10 key frames a window shares with its predecessor are reused: a
    steady-state window runs the ViT on 22 new frames and the temporal head on all 32 slots (``reuse_features``);
  * the window schedule is unrolled up front (the forward of window k depends only on raw frames), so windows
    are independent units that shard across ranks in contiguous blocks.  NCCL is used for exactly three
    things: the 9 overlap key-frame feature sets at each rank boundary, the three key-frame depth maps per window that the
    (sequential) scale/shift chain needs, and the gather of the output shards;
  * resize-to-frame-size, the scale/shift solve, affine alignment, clamp and cross-fade are device kernels on
    device-resident depth maps (the reference does them in numpy after a per-frame ``.cpu()``): no host round trip per
    window, one D2H copy of the final result.
"""
from __future__ import annotations

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
    (video_depth.py:74-86,98-99; util/transform.py).  Host-side, like the reference (boundary).
    ``indices``: only these frames are transformed (a rank's shard), into consecutive rows of the result, in the given order.
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
    """The four alignment primitives on device tensors, through the C ABI (no host synchronisation)."""

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

    def crossfade(self, pre: torch.Tensor, post: torch.Tensor, ss: torch.Tensor, w: float, out: torch.Tensor) -> torch.Tensor:
        ops.crossfade(pre.contiguous(), post.contiguous(), out, ss, w)
        return out


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


class WindowAligner:
    """Sequential affine alignment + cross-fade of window outputs on the device (video_depth.py:118-154)."""

    def __init__(self, n_windows: int, H: int, W: int, device, aops=None):
        self.H, self.W = H, W
        n_out = INFER_LEN + STEP * (n_windows - 1)
        self.aligned = torch.empty((n_out, H, W), dtype=torch.float32, device=device)
        self.pos = 0
        self.ref = None
        self.aops = aops or DeviceAlignOps()

    def push(self, d: torch.Tensor):
        """d: [32, H, W] fp32 device tensor of one window (already resized to the output size)."""
        if self.pos == 0:
            self.aligned[:INFER_LEN].copy_(d)
            self.pos = INFER_LEN
            self.ref = torch.stack([d[k] for k in KF_ALIGN]).contiguous()
            return
        ss = self.aops.scale_shift(d[:ALIGN_LEN], self.ref)
        for i in range(INTERP_LEN):
            tgt = self.aligned[self.pos - INTERP_LEN + i]
            self.aops.crossfade(tgt, d[ALIGN_LEN + i], ss, CROSSFADE_W[i], out=tgt)
        self.aops.affine_clamp(d[OVERLAP:], ss, out=self.aligned[self.pos:self.pos + STEP])
        self.pos += STEP
        self.aops.affine_clamp(d[KF_ALIGN[1]], ss, out=self.ref[1])

    def result(self, n_frames: int) -> torch.Tensor:
        return self.aligned[:n_frames]


class WindowForwarder:
    """Forwards windows (lists of source-frame indices into ``frames_t``: (N, 3, h, w) fp32 host tensor) through the model
    and resizes to ``out_hw``.  With ``reuse`` the encoder features of frames shared with the previous window are kept."""

    def __init__(self, model, frames_t: torch.Tensor, out_hw: Tuple[int, int], device, reuse: bool = True,
                 frame_rows: Optional[Dict[int, int]] = None, net_hw: Optional[Tuple[int, int]] = None):
        """``frame_rows``: source-frame index -> row of ``frames_t`` when only a rank's shard of the clip is resident.
        ``frames_t`` is either pre-processed fp32 (N, 3, h, w) or raw uint8 RGB (N, H, W, 3) with ``net_hw`` = (h, w): raw frames are
        uploaded as bytes and resized / normalised on the device (``vdn_preprocess_u8``), 4x less H2D traffic and no host cv2 loop."""
        self.model, self.frames_t, self.out_hw, self.device = model, frames_t, out_hw, device
        self.frame_rows = frame_rows
        self.raw = frames_t.dtype == torch.uint8
        self.net_hw = tuple(net_hw) if net_hw is not None else tuple(frames_t.shape[-2:])
        self.reuse = reuse and hasattr(model, "encode_frames")
        self.cache: Dict[int, List[torch.Tensor]] = {}
        self.encoded_frames = 0  # bookkeeping: encoder work actually done (frames)

    def _load(self, idx: Sequence[int]) -> torch.Tensor:
        """Host -> device copy of the listed frames, one (asynchronous, if ``frames_t`` is pinned) copy per contiguous run."""
        idx = [self.frame_rows[f] for f in idx] if self.frame_rows is not None else list(idx)
        x = torch.empty((len(idx),) + tuple(self.frames_t.shape[1:]), dtype=self.frames_t.dtype, device=self.device)
        pos = 0
        while pos < len(idx):
            run = 1
            while pos + run < len(idx) and idx[pos + run] == idx[pos] + run:
                run += 1
            x[pos:pos + run].copy_(self.frames_t[idx[pos]:idx[pos] + run], non_blocking=True)
            pos += run
        if self.raw:
            out = torch.empty((len(idx), 3) + self.net_hw, dtype=torch.float32, device=self.device)
            ops.preprocess_u8(x, out, self.net_hw[0], self.net_hw[1])
            return out
        return x

    def seed(self, feats: Dict[int, List[torch.Tensor]]):
        """Adopt features computed elsewhere (the next rank's copy of this rank's last key frames)."""
        self.cache.update(feats)

    def features_of(self, frames: Sequence[int]) -> List[torch.Tensor]:
        """Packed [len(frames)*P, C] per tap for cached frames (what a rank boundary ships)."""
        return [torch.cat([self.cache[f][t] for f in frames]).contiguous() for t in range(4)]

    @torch.no_grad()
    def forward(self, win: Sequence[int]) -> torch.Tensor:
        h, w = self.net_hw
        H, W = self.out_hw
        if not self.reuse:
            d = self.model.forward(self._load(win).unsqueeze(0))[0]
            self.encoded_frames += len(win)
        else:
            need = [f for f in dict.fromkeys(win) if f not in self.cache]
            if need:
                new = self.model.encode_frames(self._load(need))
                P = new[0].shape[0] // len(need)
                for i, f in enumerate(need):
                    self.cache[f] = [t[i * P:(i + 1) * P] for t in new]
                self.encoded_frames += len(need)
            feats = []
            for t in range(4):
                P, C = self.cache[win[0]][t].shape
                buf = torch.empty((len(win) * P, C), dtype=self.cache[win[0]][t].dtype, device=self.cache[win[0]][t].device)
                for s, f in enumerate(win):
                    buf[s * P:(s + 1) * P].copy_(self.cache[f][t])  # device memcpy into slot order
                feats.append(buf)
            d = self.model.head_from_features(feats, len(win), h // 14, w // 14)
            keep = set(win)
            self.cache = {f: v for f, v in self.cache.items() if f in keep}
        if (H, W) != (h, w):
            r = torch.empty((d.shape[0], H, W), dtype=torch.float32, device=d.device)
            ops.bilinear_f32(d.contiguous(), r, d.shape[0], h, w, H, W)
            d = r
        return d


def _pinned_raw(frames: np.ndarray, indices: Optional[Sequence[int]]) -> torch.Tensor:
    """The (selected) raw uint8 frames in page-locked memory, rows in the given order."""
    sel = list(range(frames.shape[0])) if indices is None else list(indices)
    t = torch.empty((len(sel),) + tuple(frames.shape[1:]), dtype=torch.uint8, pin_memory=True)
    dst = t.numpy()
    for row, i in enumerate(sel):
        dst[row] = frames[i]
    return t


def _resolve_input_size(fh: int, fw: int, input_size: int) -> int:
    ratio = max(fh, fw) / min(fh, fw)
    if ratio > 1.78:  # video_depth.py:68-72
        input_size = int(input_size * 1.777 / ratio)
        input_size = round(input_size / 14) * 14
    return input_size


@torch.no_grad()
def infer_video_depth(model, frames, target_fps, input_size=518, device="cuda", fp32=False, preprocessed: Optional[torch.Tensor] = None,
                      reuse_features: bool = True, group=None, gather: str = "all", device_preprocess: bool = True):
    """video_depth.py:67-156.  ``fp32`` is accepted for signature compatibility: this path always accumulates in fp32 and its
    16-bit operands meet the fp32-reference tolerance.  Returns (np.float32 (N, H, W), target_fps).

    ``device_preprocess``: upload the raw uint8 frames and run the cubic resize + normalisation on the GPU (default); False keeps the
    reference's host-side cv2 transform.

    When ``torch.distributed`` is initialised with more than one rank (or ``group`` is given) the windows are sharded across
    the ranks; every rank must call with the same ``frames``.  ``gather='all'`` returns the full result on every rank,
    ``'rank0'`` only on rank 0 (None elsewhere)."""
    if str(device).split(":")[0] != "cuda":
        raise RuntimeError("infer_video_depth runs on CUDA only (no CPU fallback)")
    import torch.distributed as dist
    frames = np.asarray(frames)
    n = frames.shape[0]
    fh, fw = frames.shape[1:3]
    input_size = _resolve_input_size(fh, fw, input_size)
    model_dev = model._dev if model._dev.type == "cuda" else torch.device(device)
    if model._dev.type != "cuda":
        model.to(model_dev)
    windows = window_schedule(n)
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world > 1:
        rank = dist.get_rank(group)
        k0, k1 = partition_windows(len(windows), world)[rank]
        frame_rows, net_hw = None, None
        if preprocessed is None:  # only this rank's frames are kept resident (raw bytes, or transformed on the host)
            mine = sorted({f for win in windows[k0:k1] for f in win})
            frame_rows = {f: i for i, f in enumerate(mine)}
            if device_preprocess:
                preprocessed, net_hw = _pinned_raw(frames, mine), _target_size(fw, fh, input_size)[::-1]
            else:
                preprocessed = preprocess_frames(frames, input_size, indices=mine, pinned=True)
        fwd = WindowForwarder(model, preprocessed, (fh, fw), model_dev, reuse=reuse_features, frame_rows=frame_rows, net_hw=net_hw)
        out = sharded_video_depth(fwd.forward, windows, n, (fh, fw), model_dev, DeviceAlignOps(), group=group, gather=gather,
                                  forwarder=fwd if fwd.reuse else None)
        return (out.cpu().numpy() if out is not None else None), target_fps
    net_hw = None
    if preprocessed is not None:
        ft = preprocessed
    elif device_preprocess:
        ft, net_hw = _pinned_raw(frames, None), _target_size(fw, fh, input_size)[::-1]
    else:
        ft = preprocess_frames(frames, input_size, pinned=True)
    fwd = WindowForwarder(model, ft, (fh, fw), model_dev, reuse=reuse_features, net_hw=net_hw)
    aligner = WindowAligner(len(windows), fh, fw, model_dev)
    for win in windows:
        aligner.push(fwd.forward(win))
    return aligner.result(n).cpu().numpy(), target_fps


def sharded_video_depth(forward, windows: Sequence[Sequence[int]], n_frames: int, out_hw: Tuple[int, int], device, aops, group=None,
                        gather: str = "all", forwarder: Optional[WindowForwarder] = None) -> Optional[torch.Tensor]:
    """Window-sharded long-video inference over ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests, which inject a
    stand-in ``forward`` / ``aops``).  ``forward(win) -> [32, H, W]`` depth of one window at the output size.

    Collectives: (1) boundary key-frame features, rank r+1 -> rank r, so that rank r's last window does not
    re-encode the 9 frames rank r+1 needs anyway [only with a ``forwarder``]; (2) all-gather of the slots (0, 1, 12) depth
    maps, scale/shift chain on rank 0, broadcast of the [K, 2] table; (3) the last window's raw slots 24..31, rank r -> r+1,
    for the cross-fade at the boundary; (4) gather of the owned output frames."""
    import torch.distributed as dist
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

    # ---- forward all own windows; (1) key-frame feature exchange at the rank boundaries ---------------------------------
    depths: List[torch.Tensor] = []
    recv_req, recv_buf, recv_frames = None, None, None
    for i, k in enumerate(mine):
        last = i == len(mine) - 1
        if forwarder is not None and last and i > 0 and recv_req is not None:
            for rq in recv_req:
                rq.wait()
            P = recv_buf[0].shape[0] // len(recv_frames)
            forwarder.seed({f: [t[j * P:(j + 1) * P] for t in recv_buf] for j, f in enumerate(recv_frames)})
            recv_req = None
        depths.append(forward(windows[k]))
        if forwarder is not None and i == 0:
            if prv is not None:  # ship the features of slots 1..9 (= window k0-1 slots 12, 24..31) to the previous rank
                pack = forwarder.features_of(windows[k][1:OVERLAP])
                send_req = [dist.isend(t, _grank(prv), group=group) for t in pack]
            if nxt is not None:
                recv_frames = list(windows[k1][1:OVERLAP])
                ref = forwarder.cache[windows[k][0]]
                recv_buf = [torch.empty((len(recv_frames) * ref[t].shape[0], ref[t].shape[1]), dtype=ref[t].dtype, device=ref[t].device)
                            for t in range(4)]
                recv_req = [dist.irecv(t, _grank(nxt), group=group) for t in recv_buf]
            if prv is not None:
                for rq in send_req:
                    rq.wait()
    if recv_req is not None:  # single-window rank: the features arrived too late to be useful, but the receive must complete
        for rq in recv_req:
            rq.wait()

    # ---- (2) scale/shift chain ------------------------------------------------------------------------------------------
    per_rank = max(b[1] - b[0] for b in bounds)
    keys = torch.zeros((per_rank, 3, H, W), dtype=torch.float32, device=device)
    for i, d in enumerate(depths):
        keys[i, 0].copy_(d[0]); keys[i, 1].copy_(d[1]); keys[i, 2].copy_(d[KF_ALIGN[1]])
    all_keys = torch.empty((world * per_rank, 3, H, W), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(all_keys, keys, group=group)
    table = torch.zeros((K, 2), dtype=torch.float32, device=device)
    if rank == 0:
        flat = torch.cat([all_keys[r * per_rank: r * per_rank + (b[1] - b[0])] for r, b in enumerate(bounds)])
        table.copy_(scale_shift_chain(flat, aops))
    dist.broadcast(table, _grank(0), group=group)

    # ---- (3) boundary cross-fade operands -------------------------------------------------------------------------------
    prev_tail = None
    reqs = []
    if prv is not None and k0 > 0:
        prev_tail = torch.empty((INTERP_LEN, H, W), dtype=torch.float32, device=device)
        reqs.append(dist.irecv(prev_tail, _grank(prv), group=group))
    if nxt is not None and mine:
        tail = depths[-1][INFER_LEN - INTERP_LEN:].contiguous()
        reqs.append(dist.isend(tail, _grank(nxt), group=group))
    for rq in reqs:
        rq.wait()

    # ---- align the owned output frames ----------------------------------------------------------------------------------
    n_out = n_frames
    lo = owned_output_range(k0, K, n_out)[0] if mine else 0
    hi = owned_output_range(k1 - 1, K, n_out)[1] if mine else 0
    shard = torch.empty((max(hi - lo, 0), H, W), dtype=torch.float32, device=device)
    for i, k in enumerate(mine):
        d, ss = depths[i], table[k]
        o_lo, o_hi = owned_output_range(k, K, n_out)
        if o_hi <= o_lo:
            continue
        for f in range(o_lo, o_hi):
            slot = f - STEP * k
            dst = shard[f - lo]
            if k > 0 and slot < OVERLAP:  # cross-fade with the previous window's aligned slot 24 + j
                j = slot - ALIGN_LEN
                raw_prev = prev_tail[j] if i == 0 else depths[i - 1][INFER_LEN - INTERP_LEN + j]
                pre = aops.affine_clamp(raw_prev, table[k - 1]) if k - 1 > 0 else raw_prev
                aops.crossfade(pre, d[slot], ss, CROSSFADE_W[j], out=dst)
            elif k > 0:
                aops.affine_clamp(d[slot], ss, out=dst)
            else:
                dst.copy_(d[slot])

    # ---- (4) gather ------------------------------------------------------------------------------------------------------
    ranges = []
    for b in bounds:
        if b[1] > b[0]:
            ranges.append((owned_output_range(b[0], K, n_out)[0], owned_output_range(b[1] - 1, K, n_out)[1]))
        else:
            ranges.append((0, 0))
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
    return full
