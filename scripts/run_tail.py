"""Head tail at the bench shape (32 frames, 296^2 x 128 -> 518^2): separate resize + implicit-GEMM output_conv2 against the fused kernel."""
import sys, torch
sys.path.insert(0, ".")
from video_depth_normal_v2_b200 import ops, packing
B, Hs, H, C = 32, 296, 518, 128
od = ops.operand_dtype()
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(B, Hs, Hs, C, device="cuda", generator=g).abs() * 0.5).to(od)
w = (torch.randn(32, C, 3, 3, device="cuda", generator=g) * (9 * C) ** -0.5)
sd = {"c.weight": w, "c.bias": torch.zeros(32)}
wp = packing.pack_conv_tail(sd, "c", "cuda", od)
cw = packing.pack_conv3x3(sd, "c", "cuda", od)
bias = torch.randn(32, device="cuda", generator=g) * 0.1
hw = torch.randn(32, device="cuda", generator=g).abs()
up = torch.empty(B, H, H, C, device="cuda", dtype=od)
out = torch.empty(B, H, H, device="cuda", dtype=torch.float32)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
fl = 2.0 * B * H * H * 9 * C * 32
if len(sys.argv) > 1 and sys.argv[1] == "fused":  # ncu target: the fused kernel only
    for _ in range(12):
        ops.conv_tail(x, wp, bias, hw, 0.05, out, B, H, H, src_hw=(Hs, Hs))
    torch.cuda.synchronize()
    sys.exit(0)
t_bil = timeit(lambda: ops.bilinear_nhwc(x, up, B, Hs, Hs, H, H, C))
t_old = timeit(lambda: ops.gemm(up, cw["w"], out, M=B * H * H, N=32, K=C, conv=(B, H, H), bias=bias, head_w=hw, head_b=0.05))
t_new = timeit(lambda: ops.conv_tail(up, wp, bias, hw, 0.05, out, B, H, H))
t_fused = timeit(lambda: ops.conv_tail(x, wp, bias, hw, 0.05, out, B, H, H, src_hw=(Hs, Hs)))
print(f"bilinear 296->518 x128ch: {t_bil:.3f} ms")
print(f"implicit-GEMM output_conv2 (9 taps, N=32): {t_old:.3f} ms  {fl / t_old / 1e9:.0f} TFLOP/s")
print(f"conv_tail on the resized map (row tiles, N=96): {t_new:.3f} ms  {fl / t_new / 1e9:.0f} TFLOP/s")
print(f"conv_tail_up (resize fused): {t_fused:.3f} ms  {fl / t_fused / 1e9:.0f} TFLOP/s   vs {t_bil + t_old:.3f} ms before")
