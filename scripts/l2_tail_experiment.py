"""Experiment: the head tail (bilinear 296 -> 518 of 128 channels, then output_conv2) frame by frame through ONE reused 68.7 MB buffer
(L2-resident between producer and consumer) against the batched form (2.2 GB intermediate through HBM)."""
import sys, torch
sys.path.insert(0, ".")
from video_depth_normal_v2_b200 import ops
od = ops.operand_dtype()
g = torch.Generator(device="cuda").manual_seed(0)
B, Hi, Ho, C = 32, 296, 518, 128
o1 = (torch.randn(B, Hi, Hi, C, device="cuda", generator=g) * 0.05).to(od)
w = (torch.randn(32, 9 * 128, device="cuda", generator=g) * 0.03).to(od)
b = torch.randn(32, device="cuda", generator=g) * 0.05
hw = torch.randn(32, device="cuda", generator=g).abs()
up = torch.empty(B, Ho, Ho, C, device="cuda", dtype=od)
depth = torch.empty(B, Ho, Ho, device="cuda", dtype=torch.float32)
depth2 = torch.empty_like(depth)

def batched():
    ops.bilinear_nhwc(o1, up, B, Hi, Hi, Ho, Ho, C)
    ops.gemm(up, w, depth, M=B * Ho * Ho, N=32, K=C, conv=(B, Ho, Ho), bias=b, head_w=hw, head_b=0.05)

def chunked(n):
    buf = up[:n]
    for i in range(0, B, n):
        ops.bilinear_nhwc(o1[i:i + n], buf, n, Hi, Hi, Ho, Ho, C)
        ops.gemm(buf, w, depth2[i:i + n], M=n * Ho * Ho, N=32, K=C, conv=(n, Ho, Ho), bias=b, head_w=hw, head_b=0.05)

def timeit(fn, graph=True):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if graph:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        run = gr.replay
    else:
        run = fn
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10

print(f"batched (32 frames)          {timeit(batched):8.3f} ms")
for n in (1, 2, 4):
    print(f"chunks of {n} frame(s), graph   {timeit(lambda: chunked(n)):8.3f} ms")
batched(); chunked(1); torch.cuda.synchronize()
print("identical:", torch.equal(depth, depth2))
