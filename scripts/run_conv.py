"""Stand-alone timing of the DPT-head 3x3 convolutions (implicit GEMM) at the ViT-L 518x518 shapes, 32 frames."""
import sys, torch
sys.path.insert(0, ".")
from video_depth_normal_v2_b200 import ops
od = ops.operand_dtype()
g = torch.Generator(device="cuda").manual_seed(0)
def r16(*s): return (torch.randn(*s, device="cuda", generator=g) * 0.05).to(od)
def f32(*s): return torch.randn(*s, device="cuda", generator=g) * 0.05
B = 32
cases = []
def conv_case(name, H, W, Ci, Co, head=False, **kw):
    x = r16(B, H, W, Ci)
    cip = (Ci + 63) // 64 * 64
    w = r16(Co, 9 * cip)
    b = f32(Co)
    if head:
        out = torch.empty(B, H, W, device="cuda", dtype=torch.float32)
        hw = f32(32)
        fn = lambda: ops.gemm(x, w, out, M=B * H * W, N=Co, K=Ci, conv=(B, H, W), bias=b, head_w=hw, head_b=0.05)
    else:
        out = torch.empty(B, H, W, Co, device="cuda", dtype=od)
        fn = lambda: ops.gemm(x, w, out, M=B * H * W, N=Co, K=Ci, conv=(B, H, W), bias=b, **kw)
    cases.append((name, 2.0 * B * H * W * Co * Ci * 9, fn))
conv_case("oc2   518x518 128->32 +head", 518, 518, 128, 32, head=True)
conv_case("oc1   296x296 256->128", 296, 296, 256, 128)
conv_case("rcu   148x148 256->256 relu", 148, 148, 256, 256, act=ops.ACT_RELU)
conv_case("rcu    74x74  256->256 relu", 74, 74, 256, 256, act=ops.ACT_RELU)
# what the residual epilogue of an RCU's second convolution costs (res = x, res2 = the other branch, out2 = relu copy for the next RCU)
_res, _res2, _o2 = r16(B, 148, 148, 256), r16(B, 148, 148, 256), torch.empty(B, 148, 148, 256, device="cuda", dtype=od)
conv_case("rcu   148x148 256->256 plain", 148, 148, 256, 256)
conv_case("rcu   148x148 256->256 +out2relu", 148, 148, 256, 256, out2=_o2, out2_relu=True)
conv_case("rcu   148x148 256->256 +res", 148, 148, 256, 256, res=_res)
conv_case("rcu   148x148 256->256 +res+out2", 148, 148, 256, 256, res=_res, out2=_o2, out2_relu=True)
conv_case("rcu   148x148 256->256 +res+res2+out2", 148, 148, 256, 256, res=_res, res2=_res2, out2=_o2, out2_relu=True)
# the same with both residuals read from ONE row (ld = 0: L1 / L2 hits, no DRAM latency): separates load latency from instruction / store cost
conv_case("rcu   148x148 256->256 +res+res2+out2 (residual rows cached)", 148, 148, 256, 256, res=_res, ld_res=0, res2=_res2, ld_res2=0, out2=_o2, out2_relu=True)
conv_case("rcu   148x148 256->256 +res (cached)", 148, 148, 256, 256, res=_res, ld_res=0)
which = sys.argv[1] if len(sys.argv) > 1 else ""
for name, flops, fn in cases:
    if which and which not in name:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:32s} {ms * 1e3:8.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s")
