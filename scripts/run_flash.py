"""Stand-alone flash-attention launch at the ViT-L bench shape (for ncu and CUDA-event timing)."""
import sys, torch
sys.path.insert(0, ".")
from video_depth_normal_v2_b200 import ops
B, N, H = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 1370, 16
C = H * 64
od = ops.operand_dtype()
g = torch.Generator(device="cuda").manual_seed(0)
qk = (torch.randn(B * N, 2 * C, device="cuda", generator=g)).to(od)
npad = (N + 7) // 8 * 8
vT = torch.randn(B * H, 64, npad, device="cuda", generator=g).to(od)
out = torch.empty(B * N, C, device="cuda", dtype=od)
for _ in range(3):
    ops.flash_attn(qk, vT, out, B, N, H)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.flash_attn(qk, vT, out, B, N, H)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"flash_attn B{B} N{N} H{H}: {ms:.4f} ms  {4.0 * B * H * N * N * 64 / ms / 1e9:.1f} TFLOP/s")
