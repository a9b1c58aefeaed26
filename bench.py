#!/usr/bin/env python
"""Headline benchmark: frames/s of the Video-Depth-Anything ViT-L 518x518 32-frame window forward on B200.

    python bench.py --gpus N --steps K --warmup W          # this framework (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N ...          # the reference algorithm's CPU path (oracle port) on the host cores

One "step" = one 32-slot window (1, 32, 3, 518, 518) through encoder + temporal DPT head, i.e.
32 frames.  N > 1 (torchrun, one rank per GPU): windows are independent units -> each rank forwards its own
window, no data-path collective, "weak" scaling; the time is the max over ranks.  Prints ONE JSON line on rank 0.
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
FEATURES, OUT_CHANNELS = 256, [256, 512, 1024, 1024]
# reference-equivalent FLOPs per slot-frame (torch flop counter on the reference modules)
GFLOP_PER_FRAME = 1404.7
ENCODER_GFLOP_PER_FRAME = 1013.6  # 24 x 42.17 + 1.65
METRIC = "frames/sec ViT-L 518x518 video"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d["bf16_tflops_sustained"], "tflops_burst": d["bf16_tflops"], "hbm_gbs": d["hbm_gbs"], "src": "measured"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "src": "fallback"}


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


def cpu_reference_fps(steps: int, warmup: int, sample_frames: int = 2):
    """The reference algorithm's CPU path (fp32 oracle port of the PyTorch modules), all host threads, on a bounded sample:
    a `sample_frames`-frame ViT-L 518x518 clip per step (a full 32-frame window is ~2 min of CPU work on 8 cores)."""
    import torch
    from oracle import vdn_oracle as O
    from oracle.init_recipe import make_input, make_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict("vda", ENCODER, 0)
    x = make_input("rgb", (1, sample_frames, 3, SIZE, SIZE), 0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.vda_forward(sd, x, ENCODER)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return {"value": sample_frames / (ms / 1e3), "ms_per_step": ms, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_frames}-frame ViT-L {SIZE}x{SIZE} clip per step through oracle.vda_forward (fp32 PyTorch CPU restatement of the reference "
                      f"modules, all host threads), mean of {len(times)} steps after {warmup} warm-up; frames/s = {sample_frames}/t"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(0, min(args.warmup, 1))
    r = cpu_reference_fps(steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"VideoDepthAnything {ENCODER} {SIZE}x{SIZE} clip forward (CPU reference arm, bounded sample)", "frames_per_step": 2},
        "cpu_baseline": {"value": r["value"], "unit": "frames/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--operands", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--da2-batch", type=int, default=16, help="batch of the DepthAnythingV2 ViT-L 518x518 measurement (BASELINE configs[1]); 0 = skip")
    ap.add_argument("--stream-frames", type=int, default=20, help="timed frames of the streaming (infer_video_depth_one) measurement; 0 = skip")
    ap.add_argument("--lv-windows", type=int, default=6, help="windows per GPU of the long-video (sharded infer_video_depth) measurement; 0 = skip")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

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
    ops.set_operand_dtype(torch.float16 if args.operands == "fp16" else torch.bfloat16)

    model = VideoDepthAnything(encoder=ENCODER, features=FEATURES, out_channels=OUT_CHANNELS).to(dev).eval()
    model.load_state_dict(synthetic_state_dict(model, 0))
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn((1, FRAMES, 3, SIZE, SIZE), generator=g).pin_memory()
    y_host = torch.empty((1, FRAMES, SIZE, SIZE)).pin_memory()
    x_dev = x_host.to(dev)

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

    # ---------------- device-resident throughput ----------------
    for _ in range(args.warmup):
        y = model(x_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        y = model(x_dev)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    assert torch.isfinite(y).all(), "non-finite depth"
    ms_per_step = ms_total / args.steps
    value = world * FRAMES * args.steps / (ms_total / 1e3)

    # ---------------- encoder alone (north_star: >= 60 % of dense tensor peak on the encoder) ----------------
    xe = x_dev[0]
    for _ in range(2):
        model.encode_frames(xe, clone=False)
    barrier()
    e0.record()
    for _ in range(args.steps):
        model.encode_frames(xe, clone=False)
    e1.record()
    barrier()
    enc_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    encoder = {"ms_per_window": enc_ms, "gflop_per_frame": ENCODER_GFLOP_PER_FRAME,
               "tflops": ENCODER_GFLOP_PER_FRAME * 1e9 * FRAMES / (enc_ms / 1e3) / 1e12,
               "note": "DINOv2 ViT-L, 32 frames 518x518 -> four tapped feature maps (patch embed, 24 blocks, final norms; SURVEY.md §8d: 1013.6 GFLOP/frame), "
                       "device-resident, per GPU"}

    # ---------------- end to end through the public API with host buffers ----------------
    for _ in range(2):
        y_host.copy_(model(x_host))
    barrier()
    e0.record()
    for _ in range(args.steps):
        y_host.copy_(model(x_host))  # H2D of the window inside forward(), D2H of the depth maps here
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e = {"value": world * FRAMES * args.steps / (e2e_ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": x_host.numel() * 4,
           "d2h_bytes_per_step": y_host.numel() * 4}

    # ---------------- long video: window-sharded infer_video_depth path (feature reuse, device-side alignment, NCCL exchange) ----
    long_video = None
    if args.lv_windows > 0:
        from video_depth_normal_v2_b200 import video as V
        K = args.lv_windows * world
        n_lv = V.STEP * K  # -> exactly K windows
        wins = V.window_schedule(n_lv)
        k0, k1 = V.partition_windows(K, world)[rank]
        mine = sorted({f for w_ in wins[k0:k1] for f in w_})
        g2 = torch.Generator().manual_seed(99)
        base = torch.randn((40, 3, SIZE, SIZE), generator=g2)
        lv_frames = torch.empty((len(mine), 3, SIZE, SIZE), dtype=torch.float32, pin_memory=True)
        for i, f in enumerate(mine):
            lv_frames[i].copy_(base[f % 40])
        rows = {f: i for i, f in enumerate(mine)}
        lv_out = torch.empty((n_lv, SIZE, SIZE), dtype=torch.float32, pin_memory=True) if rank == 0 else None

        def lv_run():
            fwd = V.WindowForwarder(model, lv_frames, (SIZE, SIZE), dev, reuse=True, frame_rows=rows)
            if world > 1:
                out = V.sharded_video_depth(fwd.forward, wins, n_lv, (SIZE, SIZE), dev, V.DeviceAlignOps(), gather="rank0", forwarder=fwd)
            else:
                al = V.WindowAligner(K, SIZE, SIZE, dev)
                for w_ in wins:
                    al.push(fwd.forward(w_))
                out = al.result(n_lv)
            if rank == 0:
                lv_out.copy_(out)  # D2H of every output frame
            return fwd.encoded_frames

        lv_run()
        lv_run()  # second warm-up: the encoder / head CUDA graphs are captured on their second use
        barrier()
        e0.record()
        enc = lv_run()
        e1.record()
        barrier()
        lv_ms = max_over_ranks(e0.elapsed_time(e1))
        long_video = {"value": n_lv / (lv_ms / 1e3), "unit": "output frames/s", "frames": n_lv, "windows": K, "ms": lv_ms,
                      "encoder_frames_this_rank": enc, "slot_forwards": K * FRAMES,
                      "note": "infer_video_depth path on a synthetic clip already pre-processed in pinned host memory: H2D per window, encoder "
                              "feature reuse for the 10 overlap slots, temporal head on all 32 slots, device-side scale/shift chain + cross-fade, "
                              "NCCL boundary exchange + gather when n_gpus > 1, D2H of all output frames"}

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
            d1 = model.stream_step(xs)
        e1.record()
        torch.cuda.synchronize()
        st_ms = e0.elapsed_time(e1) / args.stream_frames
        stream = {"value": 1e3 / st_ms, "unit": "frames/s", "ms_per_frame": st_ms,
                  "note": "ViT-L 518x518, one new frame per call attending to 31 cached frames (cached projections + positional table), device-resident input"}
        model.reset_stream()
    if world > 1:
        dist.barrier()

    # ---------------- BASELINE configs[1]: DepthAnythingV2 ViT-L, single images 518x518, batch 16, stateful memory bank ----------------
    da2 = None
    if args.da2_batch > 0:
        from video_depth_normal_v2_b200 import DepthAnythingV2
        m2 = DepthAnythingV2(encoder=ENCODER, features=FEATURES, out_channels=OUT_CHANNELS).to(dev).eval()
        m2.load_state_dict(synthetic_state_dict(m2, 0))
        xb = torch.randn((args.da2_batch, 3, SIZE, SIZE), generator=torch.Generator().manual_seed(7 + rank)).to(dev)
        m2(xb)                      # call 0: empty bank
        barrier()
        e0.record(); m2(xb); e1.record()
        barrier()
        first_ms = max_over_ranks(e0.elapsed_time(e1))
        for _ in range(6):          # fill the bank (6 entries): steady state
            m2(xb)
        barrier()
        n_da2 = max(3, args.steps // 4)
        e0.record()
        for _ in range(n_da2):
            m2(xb)
        e1.record()
        barrier()
        da2_ms = max_over_ranks(e0.elapsed_time(e1)) / n_da2
        da2 = {"value": world * args.da2_batch / (da2_ms / 1e3), "unit": "frames/s", "batch": args.da2_batch, "ms_per_call_full_bank": da2_ms,
               "ms_per_call_one_entry": first_ms, "gflop_per_frame_reference_equivalent": 1820.9,
               "tensor_frac": 1820.9e9 * args.da2_batch / (da2_ms / 1e3) / 1e12 / _peaks()["tflops"],
               "note": "DepthAnythingV2 (memory-block fork) ViT-L 518x518, device-resident batch, steady state with a full 6-entry memory bank"}
        del m2, xb
        torch.cuda.empty_cache()

    # ---------------- live per-kernel timing (CUDA events on the launching stream), extra instrumented steps ----------------
    prof = ops.KernelProfiler()
    ops.set_profiler(prof)
    n_prof = 2
    for _ in range(n_prof):
        model(x_dev)
    ops.set_profiler(None)
    agg = prof.summary()
    total_prof_ms = sum(a["ms"] for a in agg.values())
    peaks = _peaks()
    encoder["tensor_frac"] = encoder["tflops"] / peaks["tflops"]
    kernels = {}
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        rate = a["work"] / (a["ms"] / 1e3) if a["ms"] > 0 else 0.0
        kernels[name] = {"launches_per_step": a["launches"] // n_prof, "ms_per_step": a["ms"] / n_prof, "share": a["ms"] / total_prof_ms,
                         ("tflops" if a["kind"] == "tensor" else "gbs"): rate / (1e12 if a["kind"] == "tensor" else 1e9),
                         "frac_of_peak": rate / (peaks["tflops"] * 1e12 if a["kind"] == "tensor" else peaks["hbm_gbs"] * 1e9)}
    top = max(agg.items(), key=lambda kv: kv[1]["ms"])
    tname, ta = top
    if ta["kind"] == "tensor":
        achieved = ta["work"] / (ta["ms"] / 1e3) / 1e12
        roofline = {"kernel": tname, "bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                    "traffic": None, "peak_source": f"{peaks['src']} sustained bf16 (MEASURED_PEAKS.json); kind::f16 runs fp16 at the same rate",
                    "avg_launch_ms": ta["ms"] / ta["launches"], "share_of_step": ta["ms"] / total_prof_ms}
    else:
        achieved = ta["work"] / (ta["ms"] / 1e3) / 1e9
        roofline = {"kernel": tname, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                    "traffic": None, "peak_source": peaks["src"], "avg_launch_ms": ta["ms"] / ta["launches"], "share_of_step": ta["ms"] / total_prof_ms}
    traffic_file = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(tname)
        except Exception:
            pass

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_fps(steps=2, warmup=1)
        cpu = {"value": r["value"], "unit": "frames/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.operands == "fp16" else "bf16", "data": "synthetic",
        "config": {"workload": f"VideoDepthAnything {ENCODER} 32-frame clip {SIZE}x{SIZE} with temporal motion-module attention, one window forward per step "
                               f"per GPU (the video path BASELINE.json's metric names, at its 518x518; configs[1] DepthAnythingV2 batch 16 is the "
                               f"da2_batch16 block, configs[4] the long_video block)",
                   "frames_per_step_per_gpu": FRAMES, "tokens_per_frame": 1370, "parallelism": f"window-sharded x{world}, no data-path collective",
                   "l2": "activations per step (>2 GB) exceed the 126 MB L2, no explicit flush", "operands": args.operands + " (fp32 accumulate, fp32 residual stream)"},
        "tensor_frac_of_step": (GFLOP_PER_FRAME * 1e9 * FRAMES / (ms_per_step / 1e3)) / 1e12 / peaks["tflops"],
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "encoder": encoder, "long_video": long_video, "da2_batch16": da2, "stream": stream, "gpu_launches": launches, "clocks": clocks, "kernels": kernels,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
