"""Timeline of one CTA of the flash-attention kernel (debug build: VDN_EXTRA_NVCC_FLAGS=-DVDN_FA_TIMELINE python -m
video_depth_normal_v2_b200.build --force).  Prints, per KV step in the steady state, where the two softmax groups and the
MMA issuer spend their cycles (SM clock)."""
import ctypes, sys, collections
import numpy as np
import torch
sys.path.insert(0, ".")
from video_depth_normal_v2_b200 import ops, _lib
lib = _lib.load()
B, N, H = 32, 1370, 16
C = H * 64
od = ops.operand_dtype()
g = torch.Generator(device="cuda").manual_seed(0)
qk = (torch.randn(B * N, 2 * C, device="cuda", generator=g)).to(od)
vT = torch.randn(B * H, 64, (N + 7) // 8 * 8, device="cuda", generator=g).to(od)
out = torch.empty(B * N, C, device="cuda", dtype=od)
for _ in range(3):
    ops.flash_attn(qk, vT, out, B, N, H)
torch.cuda.synchronize()
lib.vdn_debug_fa_timeline.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
lib.vdn_debug_fa_timeline(None, None)
ops.flash_attn(qk, vT, out, B, N, H)
CAP = 4096
buf = np.zeros((5, CAP), dtype=np.uint64)
cnt = np.zeros(5, dtype=np.int32)
assert lib.vdn_debug_fa_timeline(buf.ctypes.data, cnt.ctypes.data) == 0
names = {1: "wait_S", 2: "ld_S", 3: "rowmax", 4: "wait_PV", 5: "exp", 6: "st_wait", 7: "->next"}
t0 = int(min(buf[s, 0] & np.uint64(0xffffffffffff) for s in range(5) if cnt[s]))
for s in range(2):  # softmax groups: duration of each phase = time until the next event
    ev = [(int(x >> np.uint64(48)), int(x & np.uint64(0xffffffffffff)) - t0) for x in buf[s, :cnt[s]]]
    dur = collections.defaultdict(list)
    for (e, t), (e2, t2) in zip(ev[:-1], ev[1:]):
        dur[e].append(t2 - t)
    steps = [t for e, t in ev if e == 1]
    per = np.diff(steps)
    print(f"group {s}: {len(steps)} tiles, cycles per tile median {np.median(per):.0f} mean {per.mean():.0f}")
    for e in sorted(dur):
        d = np.array(dur[e][5:])
        print(f"   {names.get(e, e):8s} n={len(d):4d} median {np.median(d):7.0f} mean {d.mean():7.0f} p90 {np.percentile(d, 90):7.0f}")
mn = {10: "wait_P0", 11: "wait_P1", 12: "wait_V(0)", 13: "waitV(1)", 14: "issue_PV0", 15: "issue_PV1", 16: "wait_Sfree0", 17: "wait_Sfree1", 18: "wait_K(0)", 19: "wait_K(1)",
      20: "issue_S0", 21: "issue_S1", 22: "after_PV0", 23: "after_PV1", 24: "after_S0", 25: "after_S1"}
for slot in (2, 4):
    if not cnt[slot]:
        continue
    ev = [(int(x >> np.uint64(48)), int(x & np.uint64(0xffffffffffff)) - t0) for x in buf[slot, :cnt[slot]]]
    dur = collections.defaultdict(list)
    for (e, t), (e2, t2) in zip(ev[:-1], ev[1:]):
        dur[e].append(t2 - t)
    print(f"MMA issuer (slot {slot}):")
    for e in sorted(dur):
        d = np.array(dur[e][5:])
        print(f"   {mn.get(e, e):12s} n={len(d):4d} median {np.median(d):7.0f} mean {d.mean():7.0f} p90 {np.percentile(d, 90):7.0f}")
ev = [(int(x >> np.uint64(48)), int(x & np.uint64(0xffffffffffff)) - t0) for x in buf[3, :cnt[3]]]
dur = collections.defaultdict(list)
for (e, t), (e2, t2) in zip(ev[:-1], ev[1:]):
    dur[e].append(t2 - t)
print("producer:")
for e in sorted(dur):
    d = np.array(dur[e][5:])
    print(f"   {e:3d} n={len(d):4d} median {np.median(d):7.0f} mean {d.mean():7.0f}")
# raw interleaved trace of steady-state steps 40..44 of group 0
lo = [t for e, t in [(int(x >> np.uint64(48)), int(x & np.uint64(0xffffffffffff)) - t0) for x in buf[0, :cnt[0]]] if e == 1]
a, b = lo[40], lo[43]
allev = []
for s in range(5):
    for x in buf[s, :cnt[s]]:
        e, t = int(x >> np.uint64(48)), int(x & np.uint64(0xffffffffffff)) - t0
        if a <= t <= b:
            allev.append((t, s, e))
for t, s, e in sorted(allev):
    nm = names.get(e) if s < 2 else mn.get(e, e)
    print(f"{t - a:7d} {'  ' * s * 6}{['G0', 'G1', 'MMA', 'TMA', 'MMA1'][s]}:{nm}")
