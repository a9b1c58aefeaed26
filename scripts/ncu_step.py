"""One steady-state long-video window step (22 new ViT-L frames + temporal head on 32 slots + alignment, 518x518) between
cudaProfilerStart / cudaProfilerStop: the target of the ncu passes (`--profile-from-start off`).  Launches eagerly (VDN_NO_GRAPHS=1
is set here) so that every kernel is its own launch in the list; the kernels and their order are those of the replayed graphs."""
import os, sys
os.environ.setdefault("VDN_NO_GRAPHS", "1")
import torch
sys.path.insert(0, ".")
import bench
from video_depth_normal_v2_b200 import VideoDepthAnything, ops
from video_depth_normal_v2_b200 import video as V
dev = torch.device("cuda", 0)
model = VideoDepthAnything(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024]).to(dev).eval()
model.load_state_dict(bench.synthetic_state_dict(model, 0))
_, clip = bench.synthetic_clip(22 * 4, seed=11)
fwd = V.WindowForwarder(model, V.FrameSource(clip, dev), (518, 518), dev, reuse=True, net_hw=(518, 518))
wins = V.window_schedule(22 * 4)
al = V.WindowAligner(len(wins), 518, 518, dev, n_frames=22 * 4)
for k in (0, 1):
    al.push(fwd.forward(wins[k]))
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.reset_launch_count()
al.push(fwd.forward(wins[2]))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches in the profiled step:", ops.launch_count())
