"""`not gpu`: host-side logic of the product (window schedule, resize rule, weight packing layouts, state_dict contract)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import reference_loader as RL
from oracle import vdn_oracle as O
from oracle.init_recipe import ENCODERS, make_state_dict
from video_depth_normal_v2_b200 import packing
from video_depth_normal_v2_b200 import video as V
from video_depth_normal_v2_b200.models import VideoDepthAnything

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_window_schedule_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "window_schedule.npz"))
    for key in g.files:
        n = int(key[1:])
        assert (np.array(V.window_schedule(n)) % 256 == g[key]).all(), key
    # closed form stated in SURVEY.md §5: window k = [frame 0, frame 22k-10, frames 22k+2 .. 22k+31] (clamped by the padding rule)
    w = V.window_schedule(4096)
    assert len(w) == 187
    for k in (1, 7, 100):
        assert w[k][:2] == [0, 22 * k - 10] and w[k][2:] == list(range(22 * k + 2, 22 * k + 32))


@pytest.mark.parametrize("w,h,size", [(84, 70, 70), (1920, 1080, 518), (640, 480, 518), (518, 518, 518), (924, 518, 518), (300, 500, 280)])
def test_target_size_matches_reference_resize(w, h, size):
    if not RL.available():
        pytest.skip("reference not present")
    RL._install_shims()
    import cv2
    from video_depth_anything.util.transform import Resize
    r = Resize(width=size, height=size, resize_target=False, keep_aspect_ratio=True, ensure_multiple_of=14, resize_method="lower_bound",
               image_interpolation_method=cv2.INTER_CUBIC)
    assert tuple(int(v) for v in r.get_size(w, h)) == V._target_size(w, h, size)


def test_preprocess_matches_reference_transform():
    if not RL.available():
        pytest.skip("reference not present")
    RL._install_shims()
    import cv2
    from torchvision.transforms import Compose
    from video_depth_anything.util.transform import NormalizeImage, PrepareForNet, Resize
    tf = Compose([Resize(width=84, height=84, resize_target=False, keep_aspect_ratio=True, ensure_multiple_of=14, resize_method="lower_bound",
                         image_interpolation_method=cv2.INTER_CUBIC),
                  NormalizeImage(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]), PrepareForNet()])
    frames = np.random.RandomState(0).randint(0, 256, size=(3, 100, 150, 3)).astype(np.uint8)
    ours = V.preprocess_frames(frames, 84)
    for i in range(3):
        ref = tf({"image": frames[i].astype(np.float32) / 255.0})["image"]
        assert ref.shape == ours[i].shape and np.array_equal(ref, ours[i])


def test_conv_packing_layouts():
    sd = make_state_dict("vda", "vits", 0)
    dt = torch.float32  # exact check of the permutation
    c = packing.pack_conv3x3(sd, "head.scratch.layer1_rn", "cpu", dt, bias=False)
    w = sd["head.scratch.layer1_rn.weight"]  # (64, 48, 3, 3) -> Ci padded to 64 per tap
    assert c["w"].shape == (64, 9 * 64)
    for tap in (0, 4, 8):
        assert torch.equal(c["w"][:, tap * 64:tap * 64 + 48], w[:, :, tap // 3, tap % 3])
        assert (c["w"][:, tap * 64 + 48:(tap + 1) * 64] == 0).all()
    t = packing.pack_conv_transpose(sd, "head.resize_layers.0", "cpu", dt, 4)
    wt = sd["head.resize_layers.0.weight"]  # (in 48, out 48, 4, 4)
    x = torch.randn(1, 48, 3, 2)
    ref = F.conv_transpose2d(x, wt, sd["head.resize_layers.0.bias"], stride=4)
    g = (x.permute(0, 2, 3, 1).reshape(-1, 48) @ t["w"].T + t["b"]).reshape(1, 3, 2, 4, 4, 48).permute(0, 5, 1, 3, 2, 4).reshape(1, 48, 12, 8)
    assert torch.allclose(g, ref, atol=1e-5)


def test_geglu_interleave_and_qkv_fusion():
    sd = make_state_dict("vda", "vits", 0)
    mm = packing.pack_motion_module(sd, "head.motion_modules.2.", 64, "cpu", torch.float32)
    p = "head.motion_modules.2.temporal_transformer.transformer_blocks.0."
    x = torch.randn(5, 64)
    val, gate = F.linear(x, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"]).chunk(2, dim=-1)
    inter = x @ mm["ff1_w"].T + mm["ff1_b"]
    assert torch.allclose(inter[:, 0::2], val, atol=1e-6) and torch.allclose(inter[:, 1::2], gate, atol=1e-6)
    q = F.linear(x, sd[p + "attention_blocks.1.to_k.weight"])
    assert torch.allclose((x @ mm["attn"][1]["qkv_w"].T)[:, 64:128], q, atol=1e-6)


def test_pos_embed_interpolation_matches_oracle():
    sd = make_state_dict("vda", "vits", 0)
    enc = packing.pack_encoder(sd, "pretrained.", ENCODERS["vits"], "cpu", torch.float32)
    for ph, pw in ((37, 37), (5, 6), (37, 66), (16, 16)):
        ours = packing.encoder_pos_embed(enc, ph, pw, "cpu")
        ref = O.interpolate_pos_embed(sd["pretrained.pos_embed"], ph, pw)[0]
        assert torch.equal(ours, ref)


def test_state_dict_contract():
    sd = make_state_dict("vda", "vits", 0)
    m = VideoDepthAnything(encoder="vits", features=64, out_channels=[48, 96, 192, 384])
    m.load_state_dict(sd)  # strict: the reference's exact key set
    assert set(m.state_dict().keys()) == set(sd.keys())
    bad = dict(sd)
    bad["pretrained.blocks.0.attn.qkv.weight"] = torch.zeros(3, 3)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict(bad)
    with pytest.raises(KeyError):
        VideoDepthAnything(encoder="vitb")
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(1, 2, 3, 56, 70))


def test_scale_shift_solver_matches_reference_formula():
    rng = np.random.RandomState(1)
    p, t = rng.rand(2000).astype(np.float32) + 0.1, rng.rand(2000).astype(np.float32) * 3
    sums = [float(np.sum(p.astype(np.float64) ** 2)), float(p.sum(dtype=np.float64)), float(p.size), float(np.sum(p.astype(np.float64) * t)),
            float(t.sum(dtype=np.float64))]
    s, b = V._solve_scale_shift(sums)
    s_ref, b_ref = O.compute_scale_and_shift(p, t)
    assert abs(s - s_ref) < 1e-4 and abs(b - b_ref) < 1e-4
    assert V._solve_scale_shift([0, 0, 0, 0, 0]) == (1.0, 0.0)


@pytest.mark.parametrize("process_length,target_fps,max_res", [(-1, -1, -1), (7, -1, -1), (-1, 10, -1), (9, 15, 40)])
def test_read_video_frames_matches_the_reference_reader(tmp_path, process_length, target_fps, max_res):
    """harness.read_video_frames against a restatement of utils/dc_utils.py:19-67 (cv2 branch) on a small generated clip."""
    import cv2
    from video_depth_normal_v2_b200.harness import read_video_frames
    path = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (64, 48))
    if not wr.isOpened():
        pytest.skip("no MJPG writer in this cv2 build")
    rng = np.random.RandomState(0)
    for i in range(12):
        wr.write(cv2.GaussianBlur(rng.randint(0, 255, (48, 64, 3), dtype=np.uint8), (7, 7), 0))
    wr.release()

    def reference(video_path, process_length, target_fps=-1, max_res=-1):  # dc_utils.py:41-67
        cap = cv2.VideoCapture(video_path)
        original_fps = cap.get(cv2.CAP_PROP_FPS)
        oh, ow = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
        if max_res > 0 and max(oh, ow) > max_res:
            scale = max_res / max(oh, ow)
            height, width = round(oh * scale), round(ow * scale)
        fps = original_fps if target_fps < 0 else target_fps
        stride = max(round(original_fps / fps), 1)
        frames, count = [], 0
        while cap.isOpened():
            ret, frame = cap.read()
            if not ret or (process_length > 0 and count >= process_length):
                break
            if count % stride == 0:
                frame = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
                if max_res > 0 and max(oh, ow) > max_res:
                    frame = cv2.resize(frame, (width, height))
                frames.append(frame)
            count += 1
        cap.release()
        return np.stack(frames, axis=0), fps

    exp, fps_e = reference(path, process_length, target_fps, max_res)
    got, fps_g = read_video_frames(path, process_length, target_fps, max_res)
    assert fps_g == fps_e and got.shape == exp.shape and got.dtype == np.uint8
    assert np.array_equal(got, exp)
