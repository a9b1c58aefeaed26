"""Per-shape GEMM timing of one ViT-L 32x518x518 window (CUDA events on the launching stream)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from video_depth_normal_v2_b200 import VideoDepthAnything, ops
m = VideoDepthAnything(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024]).cuda().eval()
m.load_state_dict(bench.synthetic_state_dict(m, 0))
x = torch.randn(1, 32, 3, 518, 518, device="cuda")
for _ in range(2): m(x)
prof = ops.KernelProfiler(by_shape=True); ops.set_profiler(prof)
for _ in range(2): m(x)
ops.set_profiler(None)
agg = prof.summary()
tot = sum(a["ms"] for a in agg.values())
print(f"total {tot/2:.2f} ms/step")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:40]:
    rate = a["work"] / (a["ms"] / 1e3) / (1e12 if a["kind"] == "tensor" else 1e9)
    print(f"{name:60s} n={a['launches']//2:3d} {a['ms']/2:8.3f} ms  {a['ms']/a['launches']*1e3:8.1f} us/launch  {rate:8.1f} {'TF' if a['kind']=='tensor' else 'GB/s'}")
