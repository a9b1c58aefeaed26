"""Generate the committed golden fixtures by running the LIVE reference (/root/reference) in this container.

    python tests/golden/gen_golden.py            # writes tests/golden/*.npz (small, committed)

The reference has no golden vectors of its own, so these pin the oracle
(`oracle/vdn_oracle.py`) to the reference's actual outputs.  Weights come from
`oracle/init_recipe.make_state_dict` and are loaded into the reference modules with
`load_state_dict(strict=True)` — which also proves the recipe reproduces the reference's key names/shapes.
Inputs are regenerated from seeds by the tests, only outputs are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_loader as RL  # noqa: E402
from oracle.init_recipe import make_input, make_state_dict  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (name, encoder, T, H, W, seed, subsample stride for storage)
VDA_CASES = [
    ("vda_vits_t4_70x84", "vits", 4, 70, 84, 0, 1),
    ("vda_vits_t2_518x518", "vits", 2, 518, 518, 1, 4),   # native 37x37 grid -> pos_embed untouched branch
    ("vda_vitl_t2_56x70", "vitl", 2, 56, 70, 2, 1),
]
# (name, encoder, T, H, W, seed, constructor switches): SURVEY §8f rank 3 — pe='rope' motion modules and the use_clstoken readout
VDA_SWITCH_CASES = [
    ("vda_vits_t4_70x84_rope", "vits", 4, 70, 84, 9, {"pe": "rope"}),
    ("vda_vits_t4_70x84_cls", "vits", 4, 70, 84, 10, {"use_clstoken": True}),
    ("vda_vitl_t3_56x70_rope_cls", "vitl", 3, 56, 70, 11, {"pe": "rope", "use_clstoken": True}),
]
STREAM_SWITCH_CASES = [("stream_vits_n5_56x70_rope_cls", "vits", 5, 56, 70, 12, {"pe": "rope", "use_clstoken": True})]
V5_CASES = [("v5_vits_s4_60x80", "vits", 4, 60, 80, 3)]
V4_CASES = [("v4_vits_s4_56x84", "vits", 4, 56, 84, 14)]  # models/video_depth_model_v4.py: network at the native resolution
# (name, encoder, batch, H(=W), calls, seed, stride): DepthAnythingV2 is stateful -> a sequence of forward() calls on one model
DA2_CASES = [("da2_vits_b2_70_calls8", "vits", 2, 70, 8, 5, 1), ("da2_vits_b1_518_calls2", "vits", 1, 518, 2, 6, 4), ("da2_vitl_b1_70_calls3", "vitl", 1, 70, 3, 7, 1),
             ("da2_vitb_b2_70_calls3", "vitb", 2, 70, 3, 17, 1), ("da2_vitg_b1_56_calls2", "vitg", 1, 56, 2, 23, 1),
             ("da2_vits_b2_70_calls3_cls", "vits", 2, 70, 3, 19, 1)]  # name ending in _cls: use_clstoken=True
# (name, encoder, frames, H, W, seed): streaming inference, one infer_video_depth_one call per frame (window slides after frame 10)
STREAM_CASES = [("stream_vits_n16_56x70", "vits", 16, 56, 70, 8)]
VIDEO_CASES = [("video_vits_n50_56x70", "vits", 50, 56, 70, 4)]
SCHEDULE_NS = [1, 5, 21, 22, 23, 32, 44, 45, 50, 60, 100]


def video_frames(n, h, w, seed):
    """Deterministic uint8 RGB frames with slow temporal drift (so alignment has something to fit)."""
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 256, size=(h, w, 3)).astype(np.float32)
    frames = []
    for i in range(n):
        noise = rng.randint(-20, 21, size=(h, w, 3)).astype(np.float32)
        frames.append(np.clip(np.roll(base, i, axis=1) + noise, 0, 255).astype(np.uint8))
    return np.stack(frames)


def da2_inputs(B, H, calls, seed):
    """The i-th call's input: seeded N(0,1) frames (post-normalisation), a different seed per call."""
    return [make_input("rgb", (B, 1, 3, H, H), seed * 100 + i)[:, 0] for i in range(calls)]


def gen_da2(only=None):
    for name, enc, B, H, calls, seed, stride in DA2_CASES:
        if only is not None and only not in name:
            continue
        kw = {"use_clstoken": True} if name.endswith("_cls") else {}
        sd = make_state_dict("da2", enc, seed, **kw)
        m = RL.load_da2(enc, sd, **kw)
        outs = [m(x)[:, ::stride, ::stride].numpy() for x in da2_inputs(B, H, calls, seed)]
        m.clear_memory()
        again = m(da2_inputs(B, H, calls, seed)[0])[:, ::stride, ::stride].numpy()  # clear_memory() really resets the state
        assert np.array_equal(again, outs[0])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), depth=np.stack(outs), meta=np.array([B, H, calls, seed, stride]))
        print(name, np.stack(outs).shape, float(np.stack(outs).mean()), "memory effect", float(np.abs(outs[-1] - outs[0]).mean()))
        del m


def gen_stream():
    for name, enc, N, H, W, seed in STREAM_CASES:
        sd = make_state_dict("vda", enc, seed)
        m = RL.load_vda_stream(enc, sd)
        frames = video_frames(N, H, W, seed)
        outs = [m.infer_video_depth_one(frames[i], input_size=min(H, W), device="cpu", fp32=True) for i in range(N)]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), depths=np.stack(outs).astype(np.float32), meta=np.array([N, H, W, seed]))
        print(name, np.stack(outs).shape, float(np.stack(outs).mean()))
        del m


def gen_v4():
    for name, enc, S, H, W, seed in V4_CASES:
        sd = make_state_dict("v5", enc, seed)  # v4 and v5 share the module tree / key names
        m = RL.load_v4(enc, sd)
        d = make_input("depth", (1, S, H, W), seed)
        y = m(d)
        # scripts/evaluate_v4.py:186-196: the TPF loop feeds clamp(min=0)(input) through the model twice
        d5 = make_input("depth", (2, S, H, W), seed + 1).unsqueeze(2) - 500.0
        y2 = m(m(d5.clamp(min=0).squeeze(2)))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), out=y.numpy(), out_tpf=y2.numpy(), meta=np.array([S, H, W, seed]))
        print(name, tuple(y.shape), float((y - d).abs().mean()), tuple(y2.shape))
        del m


def gen_switches():
    for name, enc, T, H, W, seed, kw in VDA_SWITCH_CASES:
        sd = make_state_dict("vda", enc, seed, **kw)
        m = RL.load_vda(enc, sd, **kw)
        y = m(make_input("rgb", (1, T, 3, H, W), seed))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), depth=y.numpy(), meta=np.array([T, H, W, seed]))
        print(name, tuple(y.shape), float(y.mean()), float(y.min()))
        del m
    for name, enc, N, H, W, seed, kw in STREAM_SWITCH_CASES:
        sd = make_state_dict("vda", enc, seed, **kw)
        m = RL.load_vda_stream(enc, sd, **kw)
        frames = video_frames(N, H, W, seed)
        outs = [m.infer_video_depth_one(frames[i], input_size=min(H, W), device="cpu", fp32=True) for i in range(N)]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), depths=np.stack(outs).astype(np.float32), meta=np.array([N, H, W, seed]))
        print(name, np.stack(outs).shape, float(np.stack(outs).mean()))
        del m


def main():
    assert RL.available(), "reference not present"
    torch.set_grad_enabled(False)
    if len(sys.argv) > 1 and sys.argv[1] == "switches":
        return gen_switches()
    if len(sys.argv) > 1 and sys.argv[1] == "v4":
        return gen_v4()
    if len(sys.argv) > 1 and sys.argv[1] == "da2":
        return gen_da2(sys.argv[2] if len(sys.argv) > 2 else None)
    if len(sys.argv) > 1 and sys.argv[1] == "stream":
        return gen_stream()
    gen_da2()
    gen_stream()
    gen_switches()
    gen_v4()
    for name, enc, T, H, W, seed, stride in VDA_CASES:
        sd = make_state_dict("vda", enc, seed)
        m = RL.load_vda(enc, sd)
        x = make_input("rgb", (1, T, 3, H, W), seed)
        stages = {}
        feats = m.pretrained.get_intermediate_layers(x.flatten(0, 1), m.intermediate_layer_idx[enc], return_class_token=True)
        y = m(x)
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            depth=y[:, :, ::stride, ::stride].numpy(),
            depth_sum=np.float64(y.double().sum().item()),
            tap3=feats[3][0][:, ::7, ::5].numpy(),
            meta=np.array([T, H, W, seed, stride]),
        )
        print(name, tuple(y.shape), float(y.mean()), float(y.min()))
        del m
    for name, enc, S, H, W, seed in V5_CASES:
        sd = make_state_dict("v5", enc, seed)
        m = RL.load_v5(enc, sd)
        d = make_input("depth", (1, S, H, W), seed)
        y = m(d)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), out=y.numpy(), meta=np.array([S, H, W, seed]))
        print(name, tuple(y.shape), float((y - d).abs().mean()))
        del m
    for name, enc, N, H, W, seed in VIDEO_CASES:
        sd = make_state_dict("vda", enc, seed)
        m = RL.load_vda(enc, sd)
        frames = video_frames(N, H, W, seed)
        captured = []
        orig_forward = m.forward

        def fwd(x):
            captured.append(x.clone())
            return orig_forward(x)
        m.forward = fwd
        depths, _ = m.infer_video_depth(frames, 30, input_size=min(H, W), device="cpu", fp32=True)
        assert captured[0].shape[-2:] == (H, W), captured[0].shape  # Resize must be the identity for these sizes
        np.savez_compressed(os.path.join(OUT, name + ".npz"), depths=depths.astype(np.float32), meta=np.array([N, H, W, seed]))
        print(name, depths.shape, float(depths.mean()))
        del m
    # window schedules: run the reference driver with a stub forward and unique per-frame inputs
    sd = make_state_dict("vda", "vits", 0)
    m = RL.load_vda("vits", sd)
    sched = {}
    for N in SCHEDULE_NS:
        frames = np.zeros((N, 14, 14, 3), np.uint8)
        frames[:, 0, 0, 0] = np.arange(N) % 256
        rec = []

        def stub(x, rec=rec):
            # recover the source frame index of each slot from the red channel of pixel (0,0)
            v = x[0, :, 0, 0, 0].double() * 0.229 + 0.485
            rec.append(torch.round(v * 255).long().tolist())
            return torch.zeros(1, x.shape[1], x.shape[3], x.shape[4])
        m.forward = stub
        out, _ = m.infer_video_depth(frames, 30, input_size=14, device="cpu", fp32=True)
        assert out.shape == (N, 14, 14)
        sched[f"n{N}"] = np.array(rec, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "window_schedule.npz"), **sched)
    print("schedules", {k: v.shape for k, v in sched.items()})


if __name__ == "__main__":
    main()
