"""Per-shape CUDA-event breakdown of one DepthAnythingV2 ViT-L 518x518 batch-16 call with a full memory bank."""
import sys, torch
sys.path.insert(0, ".")
from oracle.init_recipe import make_state_dict
from video_depth_normal_v2_b200 import DepthAnythingV2, ops
dev = torch.device("cuda", 0)
m = DepthAnythingV2(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024]).to(dev).eval()
m.load_state_dict(make_state_dict("da2", "vitl", 0))
x = torch.randn(16, 3, 518, 518, device=dev)
for _ in range(7):
    m(x)
prof = ops.KernelProfiler(by_shape=True)
ops.set_profiler(prof)
m(x)
ops.set_profiler(None)
agg = prof.summary()
tot = sum(a["ms"] for a in agg.values())
print(f"total {tot:.2f} ms per call")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:30]:
    rate = a["work"] / (a["ms"] / 1e3)
    print(f"{name:70s} n={a['launches']:3d} {a['ms']:8.3f} ms {100 * a['ms'] / tot:5.1f}%  " + (f"{rate / 1e12:7.1f} TF/s" if a["kind"] == "tensor" else f"{rate / 1e9:7.1f} GB/s"))
