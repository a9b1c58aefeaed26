"""HBM bandwidth by read : write mix, with ATen's own elementwise kernels (fill_, copy_, broadcast copy, sum) on buffers larger than
L2: the roofline of a write-dominated kernel (the ~2x bilinear up-sampling writes 4 bytes for every byte it reads) is not the
1 : 1 copy figure of MEASURED_PEAKS.json.  Prints GB/s of algorithmic bytes (each byte of DRAM traffic once)."""
import torch

dev = torch.device("cuda", 0)
n = 1 << 29  # 512 Mi fp16 elements = 1 GiB
src = torch.randn(n // 4, device=dev, dtype=torch.float16)
dst = torch.empty(n, device=dev, dtype=torch.float16)
big = torch.randn(n, device=dev, dtype=torch.float16)


def timed(fn, bytes_moved, name, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:46s} {ms:8.3f} ms  {bytes_moved / ms / 1e6:8.1f} GB/s")


timed(lambda: dst.fill_(1.0), 2.0 * n, "write only        (fill_, 1 GiB)")
timed(lambda: dst.copy_(big), 4.0 * n, "read 1 : write 1  (copy_, 1 GiB -> 1 GiB)")
k = 2048  # a 4 KB source chunk is read once from DRAM and three more times from L1 / L2
timed(lambda: dst.view(-1, 4, k).copy_(src.view(-1, 1, k).expand(-1, 4, k)), 2.0 * n + 0.5 * n, "read 1 : write 4  (broadcast copy, 256 MiB -> 1 GiB)")
timed(lambda: big.sum(dtype=torch.float32), 2.0 * n, "read only         (sum, 1 GiB)")
h = torch.empty(n // 2, device=dev, dtype=torch.float16)
timed(lambda: torch.add(big.view(2, -1)[0], big.view(2, -1)[1], out=h), 3.0 * n, "read 2 : write 1  (add, 2 x 512 MiB -> 512 MiB)")
