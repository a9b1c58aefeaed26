"""2-GPU debug of the sharded long-video path: per-phase timings (GPU events and host clock) and the shared-host-segment set-up."""
import os, sys, time, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from video_depth_normal_v2_b200 import VideoDepthAnything, ops
from video_depth_normal_v2_b200 import video as V
model = VideoDepthAnything(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024]).to(dev).eval()
model.load_state_dict(bench.synthetic_state_dict(model, 0))
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 22 * K * world
clip_t, clip = bench.synthetic_clip(n)
dev_clip = clip_t.to(dev)
for it in range(8):
    st = {}
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    if it < 4:
        out, _ = model.infer_video_depth(clip, 30, input_size=518, device="cuda", shard=True, gather="shard", stats=st)
    else:
        out, _ = model.infer_video_depth(dev_clip, 30, input_size=518, device="cuda", shard=True, gather="shard", output="device", stats=st)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"[rank {rank}] pass {it}: {1e3 * (t1 - t0):.1f} ms for {st['windows']} windows; phases {json.dumps(st.get('phases'))}", flush=True)
dist.barrier()
# shared host segment
for g in ("rank0", "all"):
    st = {}
    t0 = time.perf_counter()
    out, _ = model.infer_video_depth(clip, 30, input_size=518, device="cuda", shard=True, gather=g, stats=st)
    t1 = time.perf_counter()
    print(f"[rank {rank}] gather={g}: {1e3 * (t1 - t0):.1f} ms path={st['gather_path']} phases {json.dumps(st.get('phases'))}", flush=True)
import mmap
nbytes = 64 << 20
name = "/dev/shm/vdn_dbg"
if rank == 0:
    fd = os.open(name, os.O_CREAT | os.O_RDWR, 0o600); os.ftruncate(fd, nbytes)
dist.barrier()
if rank != 0:
    fd = os.open(name, os.O_RDWR)
mm = mmap.mmap(fd, nbytes)
t = torch.frombuffer(mm, dtype=torch.float32)
rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)
print(f"[rank {rank}] cudaHostRegister rc={rc} int={int(rc)} is_pinned={t.is_pinned()}", flush=True)
dist.barrier()
if rank == 0:
    os.unlink(name)
dist.destroy_process_group()
