"""Stand-alone timing of the DPT-head bilinear upsamples (align_corners=True, NHWC 16-bit) at the ViT-L 518x518 shapes, 32 frames.
usage: run_bilinear.py [substring of the case name]   (VDN_BILINEAR_RUN=N overrides the run length per thread)"""
import sys, torch
sys.path.insert(0, ".")
from video_depth_normal_v2_b200 import ops
od = ops.operand_dtype()
g = torch.Generator(device="cuda").manual_seed(0)
B = 32
cases = [("path4  19->37  C256", 19, 37, 256, False), ("path3  37->74  C256 relu2", 37, 74, 256, True), ("path2  74->148 C256 relu2", 74, 148, 256, True),
         ("path1 148->296 C256", 148, 296, 256, False), ("oc1   296->518 C128", 296, 518, 128, False)]
which = sys.argv[1] if len(sys.argv) > 1 else ""
for name, H, Ho, C, relu2 in cases:
    if which and which not in name:
        continue
    x = (torch.randn(B, H, H, C, device="cuda", generator=g)).to(od)
    out = torch.empty(B, Ho, Ho, C, device="cuda", dtype=od)
    out2 = torch.empty_like(out) if relu2 else None
    fn = (lambda: ops.bilinear_nhwc2(x, out, out2, B, H, H, Ho, Ho, C)) if relu2 else (lambda: ops.bilinear_nhwc(x, out, B, H, H, Ho, Ho, C))
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
    nbytes = 2.0 * B * C * (H * H + (2 if relu2 else 1) * Ho * Ho)
    print(f"{name:28s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.1f} GB/s")
