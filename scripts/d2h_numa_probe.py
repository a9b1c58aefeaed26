"""Concurrent device -> host copy bandwidth of all ranks of one box, with the pinned destination allocated (a) wherever the process
happens to run, (b) after binding the process to the CPUs NVML names for its GPU, (c) with an MPOL_PREFERRED policy for the GPU's NUMA
node.  torchrun --nproc-per-node N scripts/d2h_numa_probe.py"""
import ctypes, os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
def say(*a):
    print(f"[rank {rank}]", *a, flush=True)
def gpu_numa_node():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local if "CUDA_VISIBLE_DEVICES" not in os.environ else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local]))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        p = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node"
        node = int(open(p).read()) if os.path.exists(p) else None
        try:
            aff = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        except Exception as e:
            aff = f"err {e}"
        return h, bus, node, aff
    except Exception as e:
        return None, None, None, f"nvml: {e}"
h, bus, node, aff = gpu_numa_node()
say("pci", bus, "numa_node", node, "nvml affinity mask", [hex(x) for x in aff] if isinstance(aff, list) else aff, "allowed cpus", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
if rank == 0:
    os.system("nvidia-smi topo -m 2>&1 | head -20; ls /sys/devices/system/node/ | tr '\\n' ' '; echo; cat /sys/devices/system/node/node*/cpulist 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; grep -i 'Mems_allowed_list\\|Cpus_allowed_list' /proc/self/status")
N = 512 << 20
src = torch.empty(N, dtype=torch.uint8, device=dev).random_(0, 255)
def bench(tag):
    dst = torch.empty(N, dtype=torch.uint8, pin_memory=True)
    dst.fill_(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 0.0
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, N / e0.elapsed_time(e1) / 1e6)
    # where did the pages land?
    try:
        import re
        addr = dst.data_ptr()
        nodes = {}
        for line in open("/proc/self/numa_maps"):
            f = line.split()
            if int(f[0], 16) <= addr < int(f[0], 16) + (1 << 40) and any(x.startswith("N") for x in f) and abs(int(f[0], 16) - addr) < (1 << 21):
                nodes = {x.split("=")[0]: int(x.split("=")[1]) for x in f if re.match(r"N\d+=", x)}
    except Exception as e:
        nodes = str(e)
    t = torch.tensor([best], device=dev, dtype=torch.float64)
    if world > 1:
        allb = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allb, t)
        if rank == 0:
            v = [float(x) for x in allb]
            say(f"{tag}: concurrent D2H GB/s per rank {[round(x, 1) for x in v]} sum {sum(v):.1f}")
    say(f"{tag}: {best:.1f} GB/s, pages {nodes}, cpu {os.sched_getcpu() if hasattr(os, 'sched_getcpu') else '?'}")
    del dst
bench("default")
# (b) NVML cpu affinity
try:
    import pynvml
    pynvml.nvmlDeviceSetCpuAffinity(h)
    say("after nvmlDeviceSetCpuAffinity: allowed", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4])
    bench("nvml-affinity")
except Exception as e:
    say("nvmlDeviceSetCpuAffinity failed:", e)
# (c) MPOL_PREFERRED on the GPU's node
if node is not None and node >= 0:
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    rc = libc.syscall(238, 1, ctypes.byref(mask), 64)  # set_mempolicy(MPOL_PREFERRED, &mask, maxnode)
    say("set_mempolicy(MPOL_PREFERRED, node", node, ") rc", rc, "errno", ctypes.get_errno())
    bench("mempolicy")
if world > 1:
    dist.destroy_process_group()
