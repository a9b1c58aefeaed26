"""Debug: per-tile event timeline of flash-attention CTA (0,0,0).  Needs a library built with -DVDN_FA_TIMELINE:
   VDN_EXTRA_NVCC_FLAGS=-DVDN_FA_TIMELINE python -m video_depth_normal_v2_b200.build --force"""
import ctypes, sys, torch
sys.path.insert(0, ".")
from video_depth_normal_v2_b200 import ops
lib = ops.lib()
B, N, H = 32, 1370, 16
C = H * 64
od = ops.operand_dtype()
qk = torch.randn(B * N, 2 * C, device="cuda").to(od)
vT = torch.randn(B * H, 64, (N + 7) // 8 * 8, device="cuda").to(od)
out = torch.empty(B * N, C, device="cuda", dtype=od)
tl = torch.zeros(16 * 2 * 64, dtype=torch.int64, device="cuda")
ops.flash_attn(qk, vT, out, B, N, H)
lib.vdn_debug_set_fa_timeline.argtypes = [ctypes.c_void_p]
lib.vdn_debug_set_fa_timeline(tl.data_ptr())
ops.flash_attn(qk, vT, out, B, N, H)
torch.cuda.synchronize()
t = tl.cpu().view(16, 2, 64)[:, :, :11]
t0 = int(t[t > 0].min())
names = ["PVissued", "S+2issued", "s_full", "Sloaded", "maxdone", "pv_done", "expstart", "expdone", "I:p_full", "I:s_free"]
for j in range(11):
    for g in range(2):
        print(f"tile {j:2d} group {g}: " + "  ".join(f"{names[e]}={int(t[e, g, j]) - t0 if t[e, g, j] > 0 else -1:6d}" for e in [2, 3, 4, 5, 6, 7, 8, 0, 9, 1]))
