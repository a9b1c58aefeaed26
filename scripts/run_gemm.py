"""Stand-alone timing of the ViT-L block GEMMs (M = 32 x 1370) through the C ABI, CUDA events, 20 iterations each."""
import sys, torch
sys.path.insert(0, ".")
from video_depth_normal_v2_b200 import ops
od = ops.operand_dtype()
M, C = 32 * 1370, 1024
g = torch.Generator(device="cuda").manual_seed(0)
def r16(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.05).to(od)
def f32(*s): return torch.randn(*s, device="cuda", generator=g) * 0.05
cases = []
x16, h16 = r16(M, C), r16(M, 4 * C)
xs = f32(M, C)
qk, vT = torch.empty(M, 2 * C, device="cuda", dtype=od), torch.empty(32 * 16, 64, 1376, device="cuda", dtype=od)
hid = torch.empty(M, 4 * C, device="cuda", dtype=od)
wqkv, wproj, wfc1, wfc2 = r16(3 * C, C), r16(C, C), r16(4 * C, C), r16(C, 4 * C)
b3, b1, b4 = f32(3 * C), f32(C), f32(4 * C)
gam = f32(C)
cases.append(("qkv  1024->3072 (split, V^T)", 2.0 * M * 3 * C * C, lambda: ops.gemm(x16, wqkv, qk, M=M, N=3 * C, K=C, bias=b3, ldc=2 * C, out2=vT, row_map=ops.ROWMAP_QKV_SPLIT, rm=(1370, 1376, C, 0))))
cases.append(("proj 1024->1024 (+=, fp32)", 2.0 * M * C * C, lambda: ops.gemm(x16, wproj, xs, M=M, N=C, K=C, bias=b1, gamma=gam, res=xs)))
cases.append(("fc1  1024->4096 (gelu)", 2.0 * M * 4 * C * C, lambda: ops.gemm(x16, wfc1, hid, M=M, N=4 * C, K=C, bias=b4, act=ops.ACT_GELU)))
cases.append(("fc2  4096->1024 (+=, fp32)", 2.0 * M * 4 * C * C, lambda: ops.gemm(h16, wfc2, xs, M=M, N=C, K=4 * C, bias=b1, gamma=gam, res=xs)))
tot = 0.0
for name, flops, fn in cases:
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    tot += ms
    print(f"{name:32s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s")
print(f"sum {tot * 1e3:.1f} us per ViT-L block")
