"""Where the long-video window step spends time that is not kernel time: a CUPTI trace (torch.profiler) of steady-state steps through
the shipped path (graph replay, copy streams), reduced to GPU busy time per stream, idle gaps on the compute stream and the kernels /
memcpys next to the largest gaps."""
import sys, json, collections
import torch
sys.path.insert(0, ".")
import bench
from video_depth_normal_v2_b200 import VideoDepthAnything, ops
from video_depth_normal_v2_b200 import video as V
dev = torch.device("cuda", 0)
model = VideoDepthAnything(encoder="vitl", features=256, out_channels=[256, 512, 1024, 1024]).to(dev).eval()
model.load_state_dict(bench.synthetic_state_dict(model, 0))
n = 22 * 10
clip_t, clip = bench.synthetic_clip(n, seed=11)
frames = clip_t.to(dev) if (len(sys.argv) > 1 and sys.argv[1] == "device") else clip
out = "device" if (len(sys.argv) > 1 and sys.argv[1] == "device") else "numpy"
for _ in range(2):
    model.infer_video_depth(frames, 30, input_size=518, device="cuda", output=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
model.infer_video_depth(frames, 30, input_size=518, device="cuda", output=out)
e1.record()
torch.cuda.synchronize()
print(f"untraced pass: {e0.elapsed_time(e1):.2f} ms for 10 windows = {e0.elapsed_time(e1) / 10:.2f} ms per window")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    model.infer_video_depth(frames, 30, input_size=518, device="cuda", output=out)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/lv_trace.json")
tr = json.load(open("/tmp/lv_trace.json"))
ev = [e for e in tr["traceEvents"] if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
streams = collections.defaultdict(list)
for e in ev:
    streams[e["args"].get("stream")].append(e)
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print(f"traced pass: {len(ev)} GPU activities over {(t1 - t0) / 1e3:.2f} ms on {len(streams)} streams")
for s, lst in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    busy = sum(e["dur"] for e in lst)
    print(f"  stream {s}: {len(lst)} activities, busy {busy / 1e3:.2f} ms ({100 * busy / (t1 - t0):.1f} %)")
main = max(streams.values(), key=lambda l: sum(e["dur"] for e in l))
gaps = []
for a, b in zip(main[:-1], main[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    if g > 0:
        gaps.append((g, a["name"][:60], b["name"][:60]))
tot_gap = sum(g for g, _, _ in gaps)
print(f"compute stream: idle {tot_gap / 1e3:.2f} ms in {len(gaps)} gaps ({tot_gap / 10 / 1e3:.3f} ms per window); gap histogram (us):")
hist = collections.Counter()
for g, _, _ in gaps:
    hist["<2" if g < 2 else "2-5" if g < 5 else "5-10" if g < 10 else "10-50" if g < 50 else "50-200" if g < 200 else ">200"] += 1
sumh = collections.Counter()
for g, _, _ in gaps:
    sumh["<2" if g < 2 else "2-5" if g < 5 else "5-10" if g < 10 else "10-50" if g < 50 else "50-200" if g < 200 else ">200"] += g
for k in ["<2", "2-5", "5-10", "10-50", "50-200", ">200"]:
    print(f"    {k:7s} n={hist[k]:5d}  total {sumh[k] / 1e3:7.2f} ms")
print("largest gaps (us): after kernel -> before kernel")
for g, a, b in sorted(gaps, reverse=True)[:25]:
    print(f"  {g:8.1f}  {a}  ->  {b}")
byname = collections.Counter()
for e in main:
    byname[e["name"][:50]] += e["dur"]
print("compute-stream time by activity (ms per window):")
for k, v in byname.most_common(14):
    print(f"  {v / 10 / 1e3:7.3f}  {k}")
