"""Per-shape CUDA-event breakdown of one ViT-L 32x518x518 window (same model/weights as bench.py)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from video_depth_normal_v2_b200 import VideoDepthAnything, ops
dev = torch.device("cuda", 0)
model = VideoDepthAnything(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024]).to(dev).eval()
model.load_state_dict(bench.synthetic_state_dict(model, 0))
x = torch.randn(1, 32, 3, 518, 518, device=dev)
for _ in range(3):
    model(x)
prof = ops.KernelProfiler(by_shape=True)
ops.set_profiler(prof)
for _ in range(2):
    model(x)
ops.set_profiler(None)
agg = prof.summary()
tot = sum(a["ms"] for a in agg.values())
print(f"total {tot / 2:.2f} ms per step")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:45]:
    rate = a["work"] / (a["ms"] / 1e3)
    print(f"{name:70s} n={a['launches'] // 2:3d} {a['ms'] / 2:8.3f} ms {100 * a['ms'] / tot:5.1f}%  " + (f"{rate / 1e12:7.1f} TF/s" if a["kind"] == "tensor" else f"{rate / 1e9:7.1f} GB/s"))
