#!/usr/bin/env python
"""Headline benchmark: output frames/s of Video-Depth-Anything ViT-L at 518x518 on a long synthetic clip (BASELINE.json
configs[4]) through the drop-in ``infer_video_depth`` pipeline on B200.

    python bench.py --gpus N --steps K --warmup W          # this framework (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N ...          # the reference's own PyTorch modules on the host cores (oracle/_ref)

One "step" = one 32-slot window of the clip = 22 output frames: upload of the window's 22 new raw uint8 frames, cubic resize +
normalise on the device, the ViT on those 22 frames (the 10 key-frame slots reuse the previous window's features), the temporal
DPT head on all 32 slots, scale/shift fit, fused affine + cross-fade, download of the 22 finished frames.
K steps per rank: the clip has K * N windows (22 * K * N frames), N ranks take contiguous blocks of K windows ("weak" scaling) and
exchange the boundary key-frame features, the key-frame depth maps of the scale/shift chain and the boundary cross-fade frames over
NCCL; the time is the max over ranks.  `value` is measured with the raw frames resident in HBM and the result left there; `e2e`
goes through ``model.infer_video_depth(frames_in_pinned_host_memory, ...)`` and ends with the result in host memory.
The line also carries BASELINE configs[4] at its stated size (4096 frames, "strong" scaling: `configs4`), the sharded-vs-unsharded
parity check (`lv_parity`, N > 1), one 32-frame window forward (`window`), the encoder alone, configs[1] (`da2_batch16`) and the
streaming path.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES, SIZE, ENCODER = 32, 518, "vitl"
NEW_PER_WINDOW = 22
FEATURES, OUT_CHANNELS = 256, [256, 512, 1024, 1024]
# reference-equivalent FLOPs per slot-frame (torch flop counter on the reference modules)
GFLOP_PER_FRAME = 1404.7
ENCODER_GFLOP_PER_FRAME = 1013.6  # 24 x 42.17 + 1.65
HEAD_GFLOP_PER_FRAME = 391.1
METRIC = "frames/sec ViT-L 518x518 video"
CONFIGS4_FRAMES = 4096


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "hbm_gbs": d["hbm_gbs"], "src": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "src": "fallback"}


# Share of a kernel's algorithmic bytes that are writes, for the write-dominated bandwidth kernels of the step (2x bilinear up-sampling:
# 4 of 5 bytes; stride-2 im2col: 9 taps out for 4 pixels in at 37^2 -> 19^2; uint8 -> fp32 pre-processing: 12 of 15).  HBM3e takes writes
# alone at ~3.9 TB/s against 6.5 TB/s for a 1 : 1 copy (scripts/microbench/rw_mix.py, profiles/r02_rw_mix.txt), so these kernels are
# also reported against max(write bytes / write rate, all bytes / copy rate).
WRITE_SHARE = {"bilinear_nhwc": 0.8, "im2col_3x3_s2": 0.70, "preprocess_u8": 0.8}


def _write_peak_gbs(dev) -> float:
    """Write-only HBM rate of this box: ATen fill_ over 1 GiB of 16-bit elements, ten fills back to back between two CUDA events (as
    scripts/microbench/rw_mix.py; single timed fills read 20 % low), best of 3."""
    import torch
    buf = torch.empty(1 << 29, dtype=torch.float16, device=dev)
    for i in range(3):
        buf.fill_(float(i))
    best = float("inf")
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(10):
            buf.fill_(float(i))
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    del buf
    return (1 << 30) / (best / 1e3) / 1e9


def _round(x, sig=5):
    """Floats to `sig` significant digits, recursively: keeps the line short enough for the blocks at its end to survive a tail."""
    if isinstance(x, float):
        if x == 0 or not math.isfinite(x):
            return x
        return round(x, sig - 1 - int(math.floor(math.log10(abs(x)))))
    if isinstance(x, dict):
        return {k: _round(v, sig) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_round(v, sig) for v in x]
    return x


def synthetic_state_dict(model, seed=0):
    """Random-init weights of the reference architecture (no checkpoints offline): N(0, 0.02) matrices, unit norms, zero biases."""
    import torch
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in model._expected_shapes().items():
        if k.endswith("pos_encoder.pe"):
            C = shp[-1]
            pos = torch.arange(shp[1]).unsqueeze(1)
            div = torch.exp(torch.arange(0, C, 2) * (-math.log(10000.0) / C))
            pe = torch.zeros(shp)
            pe[0, :, 0::2], pe[0, :, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
            sd[k] = pe
        elif len(shp) >= 2 and "token" not in k:
            fan_in = 1
            for s in shp[1:]:
                fan_in *= s
            sd[k] = torch.randn(shp, generator=g) * min(0.02, fan_in ** -0.5)
        elif k.endswith("gamma") or (k.endswith(".weight") and len(shp) == 1):
            sd[k] = torch.ones(shp)
        elif k.endswith("output_conv2.2.bias"):
            sd[k] = torch.full(shp, 0.05)
        else:
            sd[k] = torch.randn(shp, generator=g) * 0.02 if len(shp) > 1 else torch.zeros(shp)
    return sd


def synthetic_clip(n: int, seed: int = 99):
    """n uint8 RGB frames (n, 518, 518, 3) in page-locked host memory: 16 random base frames, frame i = base[i % 16] shifted by i
    pixels (a moving pattern: every frame differs, generation is a memcpy per frame).  Returns (torch uint8 tensor, numpy view)."""
    import numpy as np
    import torch
    base = np.random.RandomState(seed).randint(0, 256, (16, SIZE, SIZE, 3), dtype=np.uint8)
    t = torch.empty((n, SIZE, SIZE, 3), dtype=torch.uint8, pin_memory=True)
    a = t.numpy()
    for i in range(n):
        s = i % SIZE
        b = base[i % 16]
        a[i, :, :SIZE - s] = b[:, s:]
        if s:
            a[i, :, SIZE - s:] = b[:, :s]
    return t, a


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.path = f"/tmp/vdn_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.idx)],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm), "reasons": sorted(reasons)}


def cpu_arm(mode: str, frames: int, steps: int, warmup: int, timeout: int = 1500):
    """The reference's own PyTorch modules (oracle/_ref) on the host cores, in a subprocess that sees no GPU (oracle/cpu_arm.py)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "oracle.cpu_arm", "--mode", mode, "--frames", str(frames), "--steps", str(steps), "--warmup", str(warmup),
                        "--encoder", ENCODER, "--size", str(SIZE)], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("cpu arm failed:\n" + r.stderr[-2000:])
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def run_reference(args):
    """`--impl reference`: VideoDepthAnything.infer_video_depth of the reference itself (video_depth.py:67-156, device='cpu', fp32) on a
    22-frame 518x518 clip per step = ONE 32-slot ViT-L window + alignment = one step of our arm's workload (22 output frames)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 2)), 0
    r = cpu_arm("video", NEW_PER_WINDOW, steps, warmup)
    sample = (f"{r['kind']} modules (oracle/_ref, unmodified reference files): VideoDepthAnything('{ENCODER}').infer_video_depth on a {NEW_PER_WINDOW}-frame "
              f"{SIZE}x{SIZE} uint8 clip per step (one 32-slot window incl. cv2 pre-processing and alignment, device='cpu', fp32=True), {r['steps']} step(s), "
              f"no warm-up, all host threads, torch {r['torch']}; frames/s = {NEW_PER_WINDOW} output frames / step time")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["frames_per_s"], "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"VideoDepthAnything {ENCODER} long video {SIZE}x{SIZE}: one 32-slot window (22 output frames) per step through infer_video_depth "
                               f"(CPU reference arm, bounded sample of the same per-step workload as the GPU arm)", "frames_per_step": NEW_PER_WINDOW},
        "cpu_baseline": {"value": r["frames_per_s"], "unit": "frames/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["frames_per_s"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24, help="windows per GPU of the timed long-video pass (22 output frames each)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--operands", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--da2-batch", type=int, default=16, help="batch of the DepthAnythingV2 ViT-L 518x518 measurement (BASELINE configs[1]); 0 = skip")
    ap.add_argument("--stream-frames", type=int, default=20, help="timed frames of the streaming (infer_video_depth_one) measurement; 0 = skip")
    ap.add_argument("--configs4-frames", type=int, default=CONFIGS4_FRAMES, help="length of the BASELINE configs[4] clip (strong scaling); 0 = skip")
    ap.add_argument("--window-steps", type=int, default=10, help="timed iterations of the single-window forward block; 0 = skip")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from video_depth_normal_v2_b200 import VideoDepthAnything, ops
    from video_depth_normal_v2_b200 import video as V
    ops.set_operand_dtype(torch.float16 if args.operands == "fp16" else torch.bfloat16)
    peaks = _peaks()
    ops.RIDGE_FLOP_PER_BYTE = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)

    model = VideoDepthAnything(encoder=ENCODER, features=FEATURES, out_channels=OUT_CHANNELS).to(dev).eval()
    model.load_state_dict(synthetic_state_dict(model, 0))
    shard = world > 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn):
        barrier()
        e0.record()
        r = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), r

    # ---------------- the clip: K * N windows for the headline, BASELINE configs[4]'s 4096 frames for the strong-scaling block ----------------
    K = args.steps
    n_lv = NEW_PER_WINDOW * K * world  # -> exactly K * world windows
    n_clip = max(n_lv, args.configs4_frames)
    clip_t, clip_np = synthetic_clip(n_clip)
    lv_np = clip_np[:n_lv]
    lv_dev = clip_t[:n_lv].to(dev)  # device-resident arm: the raw frames are in HBM when the timed region starts

    def lv_pass(frames, output):
        st = {}
        out, _ = model.infer_video_depth(frames, 30, input_size=SIZE, device="cuda", shard=shard, gather="shard", output=output, stats=st)
        return out, st

    # ---------------- device-resident throughput (`value`) ----------------
    # warm-up: two full passes (each K >= W windows): the first runs every shape eagerly, the second captures the CUDA graphs
    for _ in range(2):
        lv_pass(lv_dev, "device")
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    ms_total, (y, st_dev) = timed(lambda: lv_pass(lv_dev, "device"))
    launches = ops.launch_count()
    assert torch.isfinite(y).all(), "non-finite depth"
    ms_per_step = ms_total / K
    value = n_lv / (ms_total / 1e3)
    enc_frames_total = sum_over_ranks(st_dev["encoded_frames"])
    del y

    # ---------------- end to end through the public API with host buffers (`e2e`) ----------------
    lv_pass(lv_np, "numpy")  # warm-up: the pinned result block is locked here and reused from video.pinned_pool
    e2e_ms, (y_np, st_e2e) = timed(lambda: lv_pass(lv_np, "numpy"))
    clocks = sampler.stop() if rank == 0 else None
    assert np.isfinite(y_np).all()
    h2d_total, d2h_total = sum_over_ranks(st_e2e["h2d_bytes"]), sum_over_ranks(st_e2e["d2h_bytes"])
    e2e = {"value": n_lv / (e2e_ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d_total / (K * world), "d2h_bytes_per_step": d2h_total / (K * world),
           "ms_per_step": e2e_ms / K, "phases_rank0": {k: v["gpu_ms"] for k, v in (st_e2e.get("phases") or {}).items()} or None, "note": "infer_video_depth(uint8 frames in pinned host memory) -> fp32 depth in host memory; bytes per window-step per GPU"}
    del y_np, lv_dev

    # ---------------- BASELINE configs[4]: the 4096-frame clip, windows sharded over the N GPUs (strong scaling), end to end ----------------
    configs4 = None
    if args.configs4_frames > 0:
        n4 = args.configs4_frames
        c4_np = clip_np[:n4]
        wins4_ = len(V.window_schedule(n4))
        lo4, hi4 = V.rank_output_range(V.partition_windows(wins4_, world), rank, wins4_, n4)
        V.reserve_host_result(max(hi4 - lo4, 1), SIZE, SIZE)  # page-lock the result buffer ahead of the timed pass, as a long-running service would
        ms4, (y4, st4) = timed(lambda: lv_pass(c4_np, "numpy"))
        assert np.isfinite(y4).all()
        wins4 = len(V.window_schedule(n4))
        enc4 = sum_over_ranks(st4["encoded_frames"])
        configs4 = {"value": n4 / (ms4 / 1e3), "unit": "output frames/s", "frames": n4, "windows": wins4, "n_gpus": world, "scaling": "strong", "ms": ms4,
                    "windows_this_rank": st4["windows"], "encoder_frames_all_ranks": int(enc4), "slot_forwards": wins4 * FRAMES,
                    "phases_rank0": st4.get("phases"),
                    "tensor_frac_ref_equiv": wins4 * FRAMES * GFLOP_PER_FRAME * 1e9 / (ms4 / 1e3) / 1e12 / (world * peaks["tflops"]),
                    "note": "infer_video_depth(shard, gather='shard'), uint8 frames in pinned host memory -> fp32 depth in host memory of the owning ranks; one timed pass, max over ranks"}
        del y4
    del clip_t, clip_np, lv_np

    # ---------------- sharded vs unsharded result on the same clip (outside every timed region) ----------------
    lv_parity = None
    if world > 1:
        n_p = NEW_PER_WINDOW * 2 * world + 5
        _, p_np = synthetic_clip(n_p, seed=7)
        a, _ = model.infer_video_depth(p_np, 30, input_size=SIZE, device="cuda", shard=True, gather="rank0")
        if rank == 0:
            b, _ = model.infer_video_depth(p_np, 30, input_size=SIZE, device="cuda")
            lv_parity = {"max_rel": float(np.abs(a - b).max() / max(1e-12, float(np.abs(b).max()))), "frames": n_p, "windows": len(V.window_schedule(n_p)),
                         "note": "NCCL window-sharded infer_video_depth (gather='rank0') vs the single-GPU result on rank 0, max |diff| / max |ref|"}
            assert lv_parity["max_rel"] <= 1e-5, lv_parity
        dist.barrier()

    # ---------------- one 32-frame window forward (round-1 headline; kept for continuity and as the shape the kernels table refers to) ----------------
    window = None
    x_dev = torch.randn((1, FRAMES, 3, SIZE, SIZE), generator=torch.Generator().manual_seed(1234 + rank)).to(dev)
    if args.window_steps > 0:
        for _ in range(3):
            model(x_dev)
        w_ms, _ = timed(lambda: [model(x_dev) for _ in range(args.window_steps)])
        w_ms /= args.window_steps
        window = {"value": world * FRAMES / (w_ms / 1e3), "unit": "slot-frames/s", "ms_per_window": w_ms,
                  "tensor_frac": GFLOP_PER_FRAME * 1e9 * FRAMES / (w_ms / 1e3) / 1e12 / peaks["tflops"],
                  "note": "model.forward on a device-resident (1, 32, 3, 518, 518) window per rank (32 encoder frames)"}

    # ---------------- encoder alone (north_star: >= 60 % of dense tensor peak on the encoder) ----------------
    xe = x_dev[0]
    for _ in range(2):
        model.encode_frames(xe, clone=False)
    n_enc = max(3, min(args.steps, 10))
    enc_ms, _ = timed(lambda: [model.encode_frames(xe, clone=False) for _ in range(n_enc)])
    enc_ms /= n_enc
    encoder = {"ms_per_32_frames": enc_ms, "tflops": ENCODER_GFLOP_PER_FRAME * 1e9 * FRAMES / (enc_ms / 1e3) / 1e12}
    encoder["tensor_frac"] = encoder["tflops"] / peaks["tflops"]

    # ---------------- streaming: one frame per call against the cached history (video_depth_stream.py) ----------------
    stream = None
    if args.stream_frames > 0 and rank == 0:
        xs = torch.randn((3, SIZE, SIZE), generator=torch.Generator().manual_seed(3)).to(dev)
        model.reset_stream()
        for _ in range(14):  # past frame 11 the window has its steady-state 32 entries and slides
            model.stream_step(xs)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.stream_frames):
            model.stream_step(xs)
        e1.record()
        torch.cuda.synchronize()
        st_ms = e0.elapsed_time(e1) / args.stream_frames
        stream = {"value": 1e3 / st_ms, "unit": "frames/s", "ms_per_frame": st_ms}
        model.reset_stream()
    if world > 1:
        dist.barrier()

    # ---------------- live per-kernel timing of the steady-state step (CUDA events on the launching stream) ----------------
    # two more windows of a short clip with the profiler on (the profiler disables graph replay: every launch is bracketed by events)
    _, k_np = synthetic_clip(NEW_PER_WINDOW * 4, seed=11)
    src = V.FrameSource(k_np, dev)
    fwd = V.WindowForwarder(model, src, (SIZE, SIZE), dev, reuse=True, net_hw=(SIZE, SIZE))
    wins = V.window_schedule(NEW_PER_WINDOW * 4)
    al = V.WindowAligner(len(wins), SIZE, SIZE, dev, n_frames=NEW_PER_WINDOW * 4)
    al.push(fwd.forward(wins[0]))
    prof = ops.KernelProfiler()
    ops.set_profiler(prof)
    n_prof = 2
    for k in (1, 2):
        al.push(fwd.forward(wins[k]))
    ops.set_profiler(None)
    agg = prof.summary()
    del fwd, al, src
    total_prof_ms = sum(a["ms"] for a in agg.values())
    kernels = {}
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        rate = a["work"] / (a["ms"] / 1e3) if a["ms"] > 0 else 0.0
        kernels[name] = {"n": a["launches"] // n_prof, "ms": a["ms"] / n_prof, "share": a["ms"] / total_prof_ms,
                         ("tflops" if a["kind"] == "tensor" else "gbs"): rate / (1e12 if a["kind"] == "tensor" else 1e9),
                         "frac": rate / (peaks["tflops"] * 1e12 if a["kind"] == "tensor" else peaks["hbm_gbs"] * 1e9)}
    write_gbs = _write_peak_gbs(dev)
    for name, fw in WRITE_SHARE.items():
        if name in kernels and "gbs" in kernels[name]:
            mix_peak = 1.0 / max(fw / write_gbs, 1.0 / peaks["hbm_gbs"])  # GB/s of all bytes when the writes alone take fw / write rate
            kernels[name].update({"write_share": fw, "frac_write_roofline": kernels[name]["gbs"] / mix_peak})
    tname, ta = max(agg.items(), key=lambda kv: kv[1]["ms"])
    tensor = ta["kind"] == "tensor"
    achieved = ta["work"] / (ta["ms"] / 1e3) / (1e12 if tensor else 1e9)
    peak = peaks["tflops"] if tensor else peaks["hbm_gbs"]
    roofline = {"kernel": tname, "bound": "tensor" if tensor else "hbm", "achieved": achieved, "peak": peak, "unit": "TFLOP/s" if tensor else "GB/s",
                "frac": achieved / peak, "traffic": None, "avg_launch_ms": ta["ms"] / ta["launches"], "share_of_step": ta["ms"] / total_prof_ms,
                "peak_source": f"{peaks['src']} sustained bf16 (MEASURED_PEAKS.json); kind::f16 runs fp16 at the same rate" if tensor else peaks["src"]}
    traffic_file = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
    if os.path.exists(traffic_file):
        try:
            tf = json.load(open(traffic_file))
            roofline["traffic"] = tf.get(tname)
            roofline["traffic_source"] = tf.get("_source")
        except Exception:
            pass

    # ---------------- BASELINE configs[1]: DepthAnythingV2 ViT-L, single images 518x518, batch 16, stateful memory bank ----------------
    da2 = None
    if args.da2_batch > 0:
        del model
        torch.cuda.empty_cache()
        from video_depth_normal_v2_b200 import DepthAnythingV2
        m2 = DepthAnythingV2(encoder=ENCODER, features=FEATURES, out_channels=OUT_CHANNELS).to(dev).eval()
        m2.load_state_dict(synthetic_state_dict(m2, 0))
        xb = torch.randn((args.da2_batch, 3, SIZE, SIZE), generator=torch.Generator().manual_seed(7 + rank)).to(dev)
        m2(xb)                      # call 0: empty bank
        first_ms, _ = timed(lambda: m2(xb))
        for _ in range(12):         # fill the bank (6 entries) and let every ring position be captured: steady state
            m2(xb)
        n_da2 = 6
        da2_ms, _ = timed(lambda: [m2(xb) for _ in range(n_da2)])
        da2_ms /= n_da2
        da2 = {"value": world * args.da2_batch / (da2_ms / 1e3), "unit": "frames/s", "batch": args.da2_batch, "ms_per_call_full_bank": da2_ms,
               "ms_per_call_one_entry": first_ms, "tensor_frac": 1820.9e9 * args.da2_batch / (da2_ms / 1e3) / 1e12 / peaks["tflops"],
               "note": "DepthAnythingV2 ViT-L 518x518 (BASELINE configs[1]), device-resident batch, full 6-entry memory bank, 1820.9 GFLOP/frame ref-equiv"}
        del m2, xb
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_arm("forward", 8, 1, 0, timeout=600)
            fps = r["frames_per_s"] * NEW_PER_WINDOW / FRAMES
            cpu = {"value": fps, "unit": "frames/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": f"{r['kind']} modules (oracle/_ref): VideoDepthAnything('{ENCODER}').forward on (1, 8, 3, {SIZE}, {SIZE}) fp32, CPU, all host threads, one "
                             f"call, {r['sec_per_step']:.1f} s = {r['frames_per_s']:.3f} slot-frames/s; x 22/32 (a window yields 22 output frames from 32 slot "
                             f"forwards) -> output frames/s; `--impl reference` times the full 32-slot window through infer_video_depth"}
        except Exception as exc:  # the CPU arm is a reported baseline; never lose the GPU line over it
            cpu = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(exc)[-300:]}
    exec_tflop_per_step = (enc_frames_total / (K * world) * ENCODER_GFLOP_PER_FRAME + FRAMES * HEAD_GFLOP_PER_FRAME) / 1e3
    # key order: the blocks a reader of a truncated line needs first (e2e, roofline, encoder, configs4, lv_parity), the long ones last
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.operands == "fp16" else "bf16", "data": "synthetic",
        "e2e": e2e, "roofline": roofline, "gpu_launches": launches, "clocks": clocks, "encoder": encoder, "configs4": configs4, "lv_parity": lv_parity,
        "da2_batch16": da2, "window": window, "stream": stream,
        "tensor_frac_of_step": {"ref_equiv": FRAMES * GFLOP_PER_FRAME / 1e3 / (ms_per_step / 1e3) / peaks["tflops"],
                                "executed": exec_tflop_per_step / (ms_per_step / 1e3) / peaks["tflops"], "executed_tflop_per_step": exec_tflop_per_step},
        "phases_rank0": {k: v["gpu_ms"] for k, v in (st_dev.get("phases") or {}).items()} or None,
        "cpu_baseline": cpu,
        "config": {"workload": f"VideoDepthAnything {ENCODER} long synthetic video {SIZE}x{SIZE} through infer_video_depth (BASELINE configs[4] path): "
                               f"{n_lv} uint8 frames = {K * world} overlapping 32-slot windows, {K} per GPU; one step = one window = 22 output frames "
                               f"(22 ViT frames + temporal head on 32 slots + alignment); configs[4] at 4096 frames is the configs4 block",
                   "frames": n_lv, "windows": K * world, "frames_out_per_step_per_gpu": NEW_PER_WINDOW, "tokens_per_frame": 1370,
                   "parallelism": f"windows sharded x{world} in contiguous blocks; NCCL: boundary key-frame features, key-frame depth all-gather, boundary cross-fade frames"
                   if world > 1 else "single GPU",
                   "l2": "activations per step (>2 GB) exceed the 126 MB L2, no explicit flush",
                   "operands": args.operands + " (fp32 accumulate, fp32 residual stream)", "warmup_note": f"2 untimed passes over the same clip = {2 * K} window-steps per GPU (>= the requested {args.warmup})"},
        "hbm_write_gbs": write_gbs,
        "kernels": kernels,
    }
    head = {k: line[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step")}
    head.update(_round({k: v for k, v in line.items() if k not in head}))
    print(json.dumps(head), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
