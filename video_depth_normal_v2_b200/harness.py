"""Evaluation-side hand-offs of the reference, re-stated for the device path:

  evaluate_tpf        <-> scripts/evaluate_v4.py:169-233   the "TPF(msec)" loop over VideoDepthRefinerV4 / V5 (two refinement passes per
                                                             batch: ``model(model(x))``), with the device synchronised before the clock is read
  read_video_frames   <-> utils/dc_utils.py:19-67           decoded RGB frames of a video file (cv2 path: decord is not in this image), written
                                                             straight into page-locked host memory so that infer_video_depth's copy engine
                                                             reads them without a staging copy

Datasets, metrics and video writers stay out of scope; the harness takes any iterable of batches."""
from __future__ import annotations

import time
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch


def preprocess_depth_sequences(depth_batch: torch.Tensor) -> torch.Tensor:
    """scripts/evaluate_v4.py:50-99 with norm=False: clamp(min=0) and drop the channel axis: (B, S, 1, H, W) -> (B, S, H, W)."""
    if depth_batch.dim() != 5 or depth_batch.shape[2] != 1:
        raise RuntimeError(f"expected (B, S, 1, H, W), got {tuple(depth_batch.shape)}")
    return depth_batch.clamp(min=0).squeeze(2)


@torch.no_grad()
def evaluate_tpf(model, batches: Iterable[Dict[str, torch.Tensor]], device="cuda", max_eval_count: Optional[int] = None, passes: int = 2,
                 keep_outputs: bool = False, verbose: bool = True) -> dict:
    """The inference loop of scripts/evaluate_v4.py:169-233: for every batch, ``input_depths = preprocess(batch['depth_anything_v2'])``,
    ``pred = model(model(input_depths))`` (``passes`` = 2, :194-195), time per frame = inference time / frames seen.
    Returns {"tpf_ms", "frames", "infer_s", "elapsed_s", "outputs" (if keep_outputs)} and prints the reference's summary lines."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("evaluate_tpf runs on CUDA only (no CPU fallback)")
    infer, frames, outs = 0.0, 0, []
    t_epoch = time.perf_counter()
    S = None
    for i, batch in enumerate(batches):
        if max_eval_count is not None and i >= max_eval_count:
            break
        x = batch["depth_anything_v2"].to(dev)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        pred = preprocess_depth_sequences(x)
        for _ in range(passes):
            pred = model(pred)
        torch.cuda.synchronize(dev)  # the reference reads the clock without waiting for the device (:196); here the work is done
        infer += time.perf_counter() - t0
        S = pred.shape[1]
        frames += S  # evaluate_v4.py:200 counts S per batch
        if keep_outputs:
            outs.append(pred)
    elapsed = time.perf_counter() - t_epoch
    res = {"tpf_ms": infer * 1e3 / max(frames, 1), "frames": frames, "infer_s": infer, "elapsed_s": elapsed}
    if keep_outputs:
        res["outputs"] = outs
    if verbose:
        print(f"length {S}")
        print(f"Total Elapsed time: {elapsed:.2f} seconds.")
        print(f"Total frames processed: {frames}.")
        print(f"Infer Elapsed time: {infer:.2f} seconds.")
        print("=" * 80)
        print(f"TPF(msec): {res['tpf_ms']:.4f}")
    return res


def read_video_frames(video_path: str, process_length: int, target_fps: float = -1, max_res: int = -1, pinned: bool = True) -> Tuple[np.ndarray, float]:
    """utils/dc_utils.py:19-67 (the cv2 branch): every ``stride``-th frame, ``stride = max(round(original_fps / fps), 1)``, RGB, resized
    when the larger side exceeds ``max_res``, at most ``process_length`` source frames read (``> 0``).  Returns (uint8 (N, H, W, 3), fps).
    ``pinned``: the frames are decoded into one page-locked buffer (grown geometrically) and the returned array is a view of it."""
    import cv2
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise FileNotFoundError(video_path)
    original_fps = cap.get(cv2.CAP_PROP_FPS)
    oh, ow = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    resize = max_res > 0 and max(oh, ow) > max_res
    height, width = oh, ow
    if resize:
        scale = max_res / max(oh, ow)
        height, width = round(oh * scale), round(ow * scale)
    fps = original_fps if target_fps < 0 else target_fps
    stride = max(round(original_fps / fps), 1)
    use_pin = pinned and torch.cuda.is_available()

    def alloc(n):
        return torch.empty((n, height, width, 3), dtype=torch.uint8, pin_memory=use_pin)

    limit = total if total > 0 else 256
    if process_length > 0:
        limit = min(limit, process_length) if total > 0 else process_length
    buf = alloc(max(1, -(-limit // stride)))
    n, frame_count = 0, 0
    while cap.isOpened():
        ret, frame = cap.read()
        if not ret or (process_length > 0 and frame_count >= process_length):
            break
        if frame_count % stride == 0:
            if n == buf.shape[0]:  # container frame count was wrong / unknown: grow
                nb = alloc(2 * buf.shape[0])
                nb[:n].copy_(buf[:n])
                buf = nb
            dst = buf[n].numpy()
            if resize:
                cv2.cvtColor(cv2.resize(frame, (width, height)), cv2.COLOR_BGR2RGB, dst=dst)  # resize on BGR, then the channel swap: same pixels
            else:
                cv2.cvtColor(frame, cv2.COLOR_BGR2RGB, dst=dst)
            n += 1
        frame_count += 1
    cap.release()
    if n == 0:
        raise RuntimeError(f"no frame decoded from {video_path}")
    return buf.numpy()[:n], fps
