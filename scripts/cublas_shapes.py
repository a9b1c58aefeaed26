"""cuBLAS (torch.matmul, fp16 and bf16, fp32 accumulate) at the ViT-L block GEMM shapes, as a yardstick for vdn_gemm (scripts/run_gemm.py):
plain C = A W^T without the fused epilogues."""
import torch
M, C = 32 * 1370, 1024
for dt in (torch.float16, torch.bfloat16):
    for name, N, K in (("qkv", 3 * C, C), ("proj", C, C), ("fc1", 4 * C, C), ("fc2", C, 4 * C), ("square 8192", 8192, 8192)):
        m = 8192 if name.startswith("square") else M
        a = torch.randn(m, K, device="cuda", dtype=dt) * 0.05
        w = torch.randn(N, K, device="cuda", dtype=dt) * 0.05
        out = torch.empty(m, N, device="cuda", dtype=dt)
        for _ in range(3):
            torch.matmul(a, w.t(), out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            torch.matmul(a, w.t(), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{str(dt):16s} {name:12s} M={m} N={N} K={K}: {ms * 1e3:8.1f} us  {2.0 * m * N * K / ms / 1e9:8.1f} TFLOP/s")
