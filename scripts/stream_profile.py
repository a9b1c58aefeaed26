"""Per-shape CUDA-event breakdown of one streaming step (ViT-L 518x518, one new frame against 31 cached frames)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from video_depth_normal_v2_b200 import VideoDepthAnything, ops
dev = torch.device("cuda", 0)
model = VideoDepthAnything(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024]).to(dev).eval()
model.load_state_dict(bench.synthetic_state_dict(model, 0))
x = torch.randn(3, 518, 518, device=dev)
for _ in range(6):
    model.stream_step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    model.stream_step(x)
e1.record()
torch.cuda.synchronize()
print(f"graph replay: {e0.elapsed_time(e1) / 10:.3f} ms per frame")
prof = ops.KernelProfiler(by_shape=True)
ops.set_profiler(prof)
n = 3
for _ in range(n):
    model.stream_step(x)
ops.set_profiler(None)
agg = prof.summary()
tot = sum(a["ms"] for a in agg.values())
print(f"sum of kernel times {tot / n:.3f} ms per frame, {sum(a['launches'] for a in agg.values()) // n} launches")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:40]:
    rate = a["work"] / (a["ms"] / 1e3)
    print(f"{name:70s} n={a['launches'] // n:3d} {a['ms'] / n:8.3f} ms {100 * a['ms'] / tot:5.1f}%  " + (f"{rate / 1e12:7.1f} TF/s" if a["kind"] == "tensor" else f"{rate / 1e9:7.1f} GB/s"))
