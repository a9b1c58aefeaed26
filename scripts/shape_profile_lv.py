"""Per-shape CUDA-event breakdown of one steady-state long-video step (22 encoder frames + temporal head on 32 slots + alignment),
ViT-L 518x518 — the step bench.py's headline times."""
import sys, torch
sys.path.insert(0, ".")
import bench
from video_depth_normal_v2_b200 import VideoDepthAnything, ops
from video_depth_normal_v2_b200 import video as V
dev = torch.device("cuda", 0)
model = VideoDepthAnything(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024]).to(dev).eval()
model.load_state_dict(bench.synthetic_state_dict(model, 0))
_, clip = bench.synthetic_clip(22 * 4, seed=11)
fwd = V.WindowForwarder(model, V.FrameSource(clip, dev), (518, 518), dev, reuse=True, net_hw=(518, 518))
wins = V.window_schedule(22 * 4)
al = V.WindowAligner(len(wins), 518, 518, dev, n_frames=22 * 4)
al.push(fwd.forward(wins[0]))
prof = ops.KernelProfiler(by_shape=True)
ops.set_profiler(prof)
for k in (1, 2):
    al.push(fwd.forward(wins[k]))
ops.set_profiler(None)
agg = prof.summary()
tot = sum(a["ms"] for a in agg.values())
print(f"total {tot / 2:.2f} ms per step")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:60]:
    rate = a["work"] / (a["ms"] / 1e3)
    print(f"{name:70s} n={a['launches'] // 2:3d} {a['ms'] / 2:8.3f} ms {100 * a['ms'] / tot:5.1f}%  " + (f"{rate / 1e12:7.1f} TF/s" if a["kind"] == "tensor" else f"{rate / 1e9:7.1f} GB/s"))
