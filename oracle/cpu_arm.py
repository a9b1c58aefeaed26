"""TEST INFRASTRUCTURE — the CPU arm of bench.py: the reference's own PyTorch modules (``oracle/_ref``, placed there unmodified
by ``oracle/vendor_ref.py``) timed on the host cores.  Run as a subprocess with ``CUDA_VISIBLE_DEVICES=""`` (the reference moves
tensors to CUDA whenever a GPU is visible: depth_anything_v2.py:89-90, sam/transformer.py:270-272).  Prints one JSON object.

    python -m oracle.cpu_arm --mode video --frames 22 --steps 1 --warmup 0     # VideoDepthAnything.infer_video_depth, 1 window
    python -m oracle.cpu_arm --mode forward --frames 8 --steps 1 --warmup 0    # VideoDepthAnything.forward on (1, F, 3, 518, 518)

When ``oracle/_ref`` is missing (a snapshot built without /root/reference) the oracle's restatement is timed instead and the
object says ``"kind": "port"``."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="video", choices=["video", "forward"])
    ap.add_argument("--encoder", default="vitl")
    ap.add_argument("--size", type=int, default=518)
    ap.add_argument("--frames", type=int, default=22)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    args = ap.parse_args()
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import numpy as np
    import torch
    from oracle import reference_loader as RL
    from oracle import vdn_oracle as O
    from oracle.init_recipe import make_input, make_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict("vda", args.encoder, 0)
    kind = "reference" if RL.available() else "port"
    model = RL.load_vda(args.encoder, sd) if kind == "reference" else None
    S, F = args.size, args.frames
    if args.mode == "video":
        rng = np.random.RandomState(0)
        frames = rng.randint(0, 255, (F, S, S, 3), dtype=np.uint8)

        def step():
            if model is not None:
                out, _ = model.infer_video_depth(frames, 30, input_size=S, device="cpu", fp32=True)
            else:
                from video_depth_normal_v2_b200.video import preprocess_frames
                out = O.infer_video_depth_tensor(sd, torch.from_numpy(preprocess_frames(frames, S)), args.encoder, (S, S))
            return out.shape[0]
    else:
        x = make_input("rgb", (1, F, 3, S, S), 0)

        def step():
            with torch.no_grad():
                y = model(x) if model is not None else O.vda_forward(sd, x, args.encoder)
            return y.shape[1]

    times, n_out = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        n_out = step()
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    print(json.dumps({"kind": kind, "mode": args.mode, "frames_out_per_step": n_out, "sec_per_step": sec, "steps": len(times), "warmup": args.warmup,
                      "cores": torch.get_num_threads(), "torch": torch.__version__, "frames_per_s": n_out / sec}), flush=True)


if __name__ == "__main__":
    main()
