"""`-m gpu`: end-to-end parity of the CUDA path (through the C ABI) with the reference goldens and the fp32 oracle.

Tolerances are BASELINE.json's north_star: per-pixel relative depth error <= 1e-2, AbsRel <= 1e-3, normal angle <= 0.5 deg."""
import os

import numpy as np
import pytest
import torch

from oracle import vdn_oracle as O
from oracle.init_recipe import ENCODERS, make_input, make_state_dict

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MAX_REL, ABS_REL, MAX_ANGLE = 1e-2, 1e-3, 0.5


@pytest.fixture(scope="module")
def vdn():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import video_depth_normal_v2_b200 as pkg
    pkg.ops.set_operand_dtype(torch.float16)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return pkg


def _model(vdn, enc, seed):
    cfg = ENCODERS[enc]
    m = vdn.VideoDepthAnything(encoder=enc, features=cfg["features"], out_channels=cfg["out_channels"]).cuda().eval()
    sd = make_state_dict("vda", enc, seed)
    m.load_state_dict(sd)
    return m, sd


def _check(name, pred, ref, floor_frac=1e-3):
    e = O.depth_errors(pred.float().cpu(), ref.float().cpu(), floor_frac=floor_frac)
    ang = O.normal_angle_deg(O.normals_from_depth(pred.float().cpu().flatten(0, -3)), O.normals_from_depth(ref.float().cpu().flatten(0, -3)))
    print(f"{name}: {e} normal_angle_deg={ang:.4f}")
    assert e["max_rel"] <= MAX_REL and e["abs_rel"] <= ABS_REL, (name, e)
    assert ang <= MAX_ANGLE, (name, ang)
    return e


@pytest.mark.parametrize("name,enc", [("vda_vits_t4_70x84", "vits"), ("vda_vits_t2_518x518", "vits"), ("vda_vitl_t2_56x70", "vitl")])
def test_forward_matches_reference_golden(vdn, name, enc):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    T, H, W, seed, stride = [int(v) for v in g["meta"]]
    m, sd = _model(vdn, enc, seed)
    x = make_input("rgb", (1, T, 3, H, W), seed)
    vdn.ops.reset_launch_count()
    y = m(x.cuda())
    torch.cuda.synchronize()
    assert y.shape == (1, T, H, W) and y.dtype == torch.float32 and y.is_cuda
    assert vdn.ops.launch_count() > 50
    _check(name, y[:, :, ::stride, ::stride], torch.from_numpy(g["depth"]))


def test_encoder_taps_match_oracle(vdn):
    from video_depth_normal_v2_b200.models import encoder_forward
    m, sd = _model(vdn, "vits", 5)
    x = make_input("rgb", (3, 3, 70, 84), 5)
    feats = encoder_forward(m._weights()["enc"], x.cuda())
    ref = O.dinov2_intermediate(sd, x, "vits", O._Ops())
    for i, (f, (tok, _)) in enumerate(zip(feats, ref)):
        err = (f.float().cpu().reshape(tok.shape) - tok).abs().max() / tok.abs().max()
        print(f"tap {i}: rel-to-max err {float(err):.3e}")
        assert err < 5e-3


def test_stagewise_head_matches_oracle(vdn):
    """Per-stage check of the head on oracle features (isolates head/motion-module errors from encoder errors)."""
    from video_depth_normal_v2_b200.models import head_forward
    enc = "vits"
    m, sd = _model(vdn, enc, 6)
    T, ph, pw = 4, 5, 6
    x = make_input("rgb", (1, T, 3, 14 * ph, 14 * pw), 6)
    ops_ = O._Ops()
    feats = O.dinov2_intermediate(sd, x.flatten(0, 1), enc, ops_)
    ref = O.dpt_head(sd, "head.", feats, ph, pw, ops_, T=T)
    f16 = [f[0].reshape(-1, f[0].shape[-1]).cuda().to(vdn.ops.operand_dtype()).contiguous() for f in feats]
    d = head_forward(m._weights()["head"], f16, T, ph, pw, T)
    torch.cuda.synchronize()
    _check("head on oracle features", d.reshape(1, T, 14 * ph, 14 * pw), ref.reshape(1, T, 14 * ph, 14 * pw))


def test_infer_video_depth_matches_reference_golden(vdn):
    from tests.golden.gen_golden import video_frames
    g = np.load(os.path.join(GOLD, "video_vits_n50_56x70.npz"))
    N, H, W, seed = [int(v) for v in g["meta"]]
    m, sd = _model(vdn, "vits", seed)
    frames = video_frames(N, H, W, seed)
    out, fps = m.infer_video_depth(frames, 30, input_size=min(H, W), device="cuda")
    assert fps == 30 and out.shape == (N, H, W) and out.dtype == np.float32
    _check("infer_video_depth", torch.from_numpy(out), torch.from_numpy(g["depths"]))


def test_full_size_vitl_window_matches_oracle_on_gpu(vdn):
    """BASELINE config size (ViT-L, 518x518); the oracle runs in strict fp32 on the same GPU as the checker."""
    enc, T, H, W = "vitl", 8, 518, 518
    m, sd = _model(vdn, enc, 7)
    x = make_input("rgb", (1, T, 3, H, W), 7).cuda()
    y = m(x)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    ref = O.vda_forward(sd_gpu, x, enc)
    torch.cuda.synchronize()
    _check("vitl 8x518x518", y, ref)


def test_timed_shape_vitl_32x518_matches_oracle_on_gpu(vdn):
    """The exact shape bench.py times (ViT-L, one window of 32 frames at 518x518): the pair-form GEMMs at M = 43840 and the tcgen05
    temporal-attention kernel (T == 32), end to end against the fp32 oracle on the same GPU; then the same window through the
    captured CUDA graph (third call = replay) must reproduce the eager result bit for bit."""
    enc, T, H, W = "vitl", 32, 518, 518
    m, sd = _model(vdn, enc, 17)
    x = make_input("rgb", (1, T, 3, H, W), 17).cuda()
    y = m(x).clone()
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    ref = O.vda_forward(sd_gpu, x, enc)
    torch.cuda.synchronize()
    _check("vitl 32x518x518 (timed shape)", y, ref)
    del ref, sd_gpu
    torch.cuda.empty_cache()
    y2 = m(x).clone()
    y3 = m(x).clone()
    torch.cuda.synchronize()
    assert torch.equal(y, y2) and torch.equal(y, y3)


def test_long_video_pipeline_vitl_518_matches_window_forward(vdn):
    """The long-video path bench.py times (raw uint8 frames -> device pre-processing -> encoder-feature reuse -> head -> fused
    alignment kernel -> pinned host result) on ViT-L 518x518, 3 windows: against the same clip with every window forwarded in
    full (no feature reuse) and aligned by the oracle's restatement of the reference loop."""
    from video_depth_normal_v2_b200 import video as V
    enc, n, S = "vitl", 60, 518
    m, sd = _model(vdn, enc, 19)
    rng = np.random.RandomState(5)
    base = rng.randint(0, 255, (8, S, S, 3), dtype=np.uint8)
    frames = np.stack([np.roll(base[i % 8], 3 * i, axis=1) for i in range(n)])
    st = {}
    out, _ = m.infer_video_depth(frames, 30, input_size=S, device="cuda", stats=st)
    assert out.shape == (n, S, S) and st["encoded_frames"] == n and st["windows"] == 3
    wins = V.window_schedule(n)
    x = torch.empty((n, 3, S, S), dtype=torch.float32, device="cuda")
    vdn.ops.preprocess_u8(torch.from_numpy(frames).cuda(), x, S, S)
    depth_list = []
    for w in wins:
        d = m(x[torch.tensor(w, device="cuda")].unsqueeze(0))[0]
        depth_list += [f.cpu().numpy() for f in d]
    exp = O.align_windows(depth_list, n)
    err = np.abs(out - exp).max() / max(1.0, np.abs(exp).max())
    print(f"long-video pipeline vs per-window forward + reference alignment: max err / max {err:.3e}")
    assert err < 1e-5


def test_determinism_and_batch_independence(vdn):
    """Size-independent properties: repeated runs are bit-identical; frames of different clips in a batch do not interact."""
    m, sd = _model(vdn, "vits", 8)
    x = make_input("rgb", (2, 4, 3, 56, 70), 8).cuda()
    y1 = m(x)
    y2 = m(x)
    ya = m(x[:1])
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    assert torch.equal(y1[:1], ya)


def test_api_errors(vdn):
    with pytest.raises(KeyError):  # video_depth.py:48-51 knows vits / vitl only (vitb exists for DepthAnythingV2)
        vdn.VideoDepthAnything(encoder="vitb")
    m, sd = _model(vdn, "vits", 0)
    with pytest.raises(RuntimeError, match="multiple of patch size"):
        m(torch.zeros(1, 2, 3, 60, 70).cuda())
    bad = dict(sd)
    bad.pop("head.scratch.output_conv1.bias")
    with pytest.raises(RuntimeError, match="missing keys"):
        m.load_state_dict(bad)
    cpu_model = vdn.VideoDepthAnything(encoder="vits", features=64, out_channels=[48, 96, 192, 384])
    cpu_model.load_state_dict(sd)
    with pytest.raises(RuntimeError, match="CUDA only"):
        cpu_model(torch.zeros(1, 2, 3, 56, 70))


# ------------------------------------------------------------------------------------------ a12: v5 refinement model
def test_v5_refiner_matches_reference_golden(vdn):
    """models/video_depth_model_v5.py forward (median scale head, Sobel-normal input, residual output) against the output of the
    live reference (tests/golden/gen_golden.py) and against the oracle on a second, larger non-square case."""
    g = np.load(os.path.join(GOLD, "v5_vits_s4_60x80.npz"))
    S, H, W, seed = [int(v) for v in g["meta"]]
    cfg = ENCODERS["vits"]
    sd = make_state_dict("v5", "vits", seed)
    m = vdn.VideoDepthRefinerV5(encoder="vits", features=cfg["features"], out_channels=cfg["out_channels"]).cuda().eval()
    m.load_state_dict(sd)
    d = make_input("depth", (1, S, H, W), seed)
    y = m(d.cuda())
    torch.cuda.synchronize()
    assert y.shape == (1, S, H, W) and y.dtype == torch.float32
    # normals are taken on the max_depth-normalised maps, as the model itself does (video_depth_model_v5.py:162,175)
    _check("v5_vits_s4_60x80 vs reference", y / 65535.0, torch.from_numpy(g["out"]) / 65535.0)
    # batch of two sequences, odd sizes (even pixel count -> interpolated median), against the oracle
    d2 = make_input("depth", (2, 3, 75, 98), seed + 1)
    y2 = m(d2.cuda())
    ref2 = O.v5_forward(sd, d2, "vits")
    # the refined depth is input + signed residual and comes arbitrarily close to 0 here: the per-pixel *relative* bound is
    # evaluated on pixels above 5 % of the range, the absolute error everywhere (<= 1e-3 of the range)
    e2 = _check("v5_vits 2x3x75x98 vs oracle", y2 / 65535.0, ref2 / 65535.0, floor_frac=0.05)
    assert e2["max_abs"] <= 1e-3


def test_frame_median_matches_torch_quantile(vdn):
    for n in (1, 2, 7, 1000, 4801, 518 * 924):
        x = torch.rand(3, n, device="cuda") * torch.tensor([1.0, 100.0, 65535.0], device="cuda").view(3, 1) - 0.25
        med = torch.empty(3, device="cuda")
        sc = torch.empty(3, device="cuda")
        vdn.ops.frame_median_scale(x.contiguous(), sc, n, 1.0, 0.3, -0.1, median=med)
        ref = torch.quantile(x.cpu(), 0.5, dim=-1)
        assert torch.equal(med.cpu(), ref) or float((med.cpu() - ref).abs().max()) <= 1e-6 * float(ref.abs().max()), (n, med, ref)
        assert torch.allclose(sc.cpu(), torch.exp(torch.tanh(ref * 0.3 - 0.1)), rtol=1e-5)


# ------------------------------------------------------------------------------------------ a11: DepthAnythingV2 + memory block
def _da2_inputs(B, H, calls, seed):
    return [make_input("rgb", (B, 1, 3, H, H), seed * 100 + i)[:, 0] for i in range(calls)]


@pytest.mark.parametrize("name,enc", [("da2_vits_b2_70_calls8", "vits"), ("da2_vits_b1_518_calls2", "vits"), ("da2_vitl_b1_70_calls3", "vitl"),
                                      ("da2_vitb_b2_70_calls3", "vitb"), ("da2_vits_b2_70_calls3_cls", "vits"), ("da2_vitg_b1_56_calls2", "vitg")])
def test_da2_stateful_forward_matches_reference_golden(vdn, name, enc):
    """A sequence of forward() calls on one model against the live reference's outputs: empty bank (constant cross-attention
    term), filling bank (cross-attention over 1..6 cached entries) and the ring wrap after 6 entries."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    B, H, calls, seed, stride = [int(v) for v in g["meta"]]
    cfg = ENCODERS[enc]
    cls = name.endswith("_cls")  # use_clstoken=True: readout of the last tap after the memory block
    m = vdn.DepthAnythingV2(encoder=enc, features=cfg["features"], out_channels=cfg["out_channels"], use_clstoken=cls).cuda().eval()
    m.load_state_dict(make_state_dict("da2", enc, seed, use_clstoken=cls))
    xs = _da2_inputs(B, H, calls, seed)
    for i, x in enumerate(xs):
        y = m(x.cuda())
        assert y.shape == (B, H, H) and y.dtype == torch.float32
        _check(f"{name} call {i}", y[:, ::stride, ::stride], torch.from_numpy(g["depth"][i]))
    m.clear_memory()
    y0 = m(xs[0].cuda())
    _check(f"{name} after clear_memory", y0[:, ::stride, ::stride], torch.from_numpy(g["depth"][0]))


def test_da2_batch16_vitl_518_matches_oracle(vdn):
    """BASELINE configs[1] shape (ViT-L, 518x518), batch reduced to 2 for the fp32 oracle on the GPU; two calls (empty / one entry)."""
    enc = "vitl"
    cfg = ENCODERS[enc]
    sd = make_state_dict("da2", enc, 11)
    m = vdn.DepthAnythingV2(encoder=enc, features=cfg["features"], out_channels=cfg["out_channels"]).cuda().eval()
    m.load_state_dict(sd)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    bank = []
    for i, x in enumerate(_da2_inputs(2, 518, 2, 11)):
        y = m(x.cuda())
        ref = O.da2_forward(sd_gpu, x.cuda(), enc, bank)
        _check(f"da2 vitl 2x518 call {i}", y, ref)


def test_da2_timed_shape_batch16_vitl_518_matches_oracle(vdn):
    """BASELINE configs[1] at its stated size: DepthAnythingV2 ViT-L, 518x518, batch 16, three stateful calls (empty bank, one and
    two entries) against the fp32 oracle on the same GPU."""
    enc = "vitl"
    cfg = ENCODERS[enc]
    sd = make_state_dict("da2", enc, 13)
    m = vdn.DepthAnythingV2(encoder=enc, features=cfg["features"], out_channels=cfg["out_channels"]).cuda().eval()
    m.load_state_dict(sd)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    bank = []
    for i, x in enumerate(_da2_inputs(16, 518, 3, 13)):
        y = m(x.cuda()).clone()
        ref = O.da2_forward(sd_gpu, x.cuda(), enc, bank)
        _check(f"da2 vitl 16x518 call {i}", y, ref)
        del ref
        torch.cuda.empty_cache()


def test_da2_infer_image_shape_and_state(vdn):
    cfg = ENCODERS["vits"]
    m = vdn.DepthAnythingV2(encoder="vits", features=cfg["features"], out_channels=cfg["out_channels"]).cuda().eval()
    m.load_state_dict(make_state_dict("da2", "vits", 5))
    img = np.random.RandomState(0).randint(0, 255, (90, 90, 3), dtype=np.uint8)
    d = m.infer_image(img, input_size=70)
    assert d.shape == (90, 90) and d.dtype == np.float32 and np.isfinite(d).all()
    with pytest.raises(RuntimeError, match="clear_memory"):
        m(torch.zeros(2, 3, 70, 70).cuda())  # batch size changed while the bank holds an entry
    m.clear_memory()
    m(torch.zeros(2, 3, 70, 70).cuda())


# ------------------------------------------------------------------------------------------ f1: streaming inference
def test_streaming_infer_video_depth_one_matches_reference_golden(vdn):
    """video_depth_stream.py:76-160 frame by frame: cached projections + positional table against the live reference's outputs
    (which re-project the cached hidden states on every frame)."""
    import sys
    sys.path.insert(0, GOLD)
    from gen_golden import video_frames
    g = np.load(os.path.join(GOLD, "stream_vits_n16_56x70.npz"))
    N, H, W, seed = [int(v) for v in g["meta"]]
    m, _ = _model(vdn, "vits", seed)
    frames = video_frames(N, H, W, seed)
    for i in range(N):
        d = m.infer_video_depth_one(frames[i], input_size=min(H, W), device="cuda")
        assert d.shape == (H, W) and d.dtype == np.float32
        _check(f"stream frame {i}", torch.from_numpy(d)[None], torch.from_numpy(g["depths"][i])[None])
    m.reset_stream()
    d0 = m.infer_video_depth_one(frames[0], input_size=min(H, W), device="cuda")
    _check("stream frame 0 after reset", torch.from_numpy(d0)[None], torch.from_numpy(g["depths"][0])[None])


# ------------------------------------------------------------------------------------------ f3: pe='rope' / use_clstoken=True
def _model_kw(vdn, enc, seed, **kw):
    cfg = ENCODERS[enc]
    m = vdn.VideoDepthAnything(encoder=enc, features=cfg["features"], out_channels=cfg["out_channels"], **kw).cuda().eval()
    sd = make_state_dict("vda", enc, seed, **kw)
    m.load_state_dict(sd)
    return m, sd


@pytest.mark.parametrize("name,enc,kw", [("vda_vits_t4_70x84_rope", "vits", {"pe": "rope"}), ("vda_vits_t4_70x84_cls", "vits", {"use_clstoken": True}),
                                         ("vda_vitl_t3_56x70_rope_cls", "vitl", {"pe": "rope", "use_clstoken": True})])
def test_switches_match_reference_golden(vdn, name, enc, kw):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    T, H, W, seed = [int(v) for v in g["meta"]]
    m, sd = _model_kw(vdn, enc, seed, **kw)
    y = m(make_input("rgb", (1, T, 3, H, W), seed).cuda())
    _check(name, y, torch.from_numpy(g["depth"]))
    with pytest.raises(RuntimeError):  # strict key check: a default-config state_dict does not fit a switched model
        m.load_state_dict(make_state_dict("vda", enc, seed))


def test_rope_window_of_32_on_the_tensor_core_temporal_kernel(vdn):
    """T = 32 ViT-L takes the tcgen05 temporal-attention path (q|k row-major, V transposed per tile): RoPE is applied to the q|k buffer."""
    kw = {"pe": "rope", "use_clstoken": True}
    m, sd = _model_kw(vdn, "vitl", 13, **kw)
    x = make_input("rgb", (1, 32, 3, 28, 42), 13)
    y = m(x.cuda())
    dev_sd = {k: v.cuda() for k, v in sd.items()}
    ref = O.vda_forward(dev_sd, x.cuda(), "vitl")
    _check("rope+cls vitl T=32", y, ref)


def test_streaming_rope_cls_matches_reference_golden(vdn):
    import sys
    sys.path.insert(0, GOLD)
    from gen_golden import video_frames
    g = np.load(os.path.join(GOLD, "stream_vits_n5_56x70_rope_cls.npz"))
    N, H, W, seed = [int(v) for v in g["meta"]]
    m, _ = _model_kw(vdn, "vits", seed, pe="rope", use_clstoken=True)
    frames = video_frames(N, H, W, seed)
    for i in range(N):
        d = m.infer_video_depth_one(frames[i], input_size=min(H, W), device="cuda")
        _check(f"stream rope+cls frame {i}", torch.from_numpy(d)[None], torch.from_numpy(g["depths"][i])[None])


# ------------------------------------------------------------------------------------------ f4: v4 refinement model (native resolution)
def test_v4_refiner_matches_reference_golden(vdn):
    """models/video_depth_model_v4.py:120-148 = the v5 tree without the 224x224 resize, against the live reference's output."""
    g = np.load(os.path.join(GOLD, "v4_vits_s4_56x84.npz"))
    S, H, W, seed = [int(v) for v in g["meta"]]
    cfg = ENCODERS["vits"]
    sd = make_state_dict("v5", "vits", seed)
    m = vdn.VideoDepthRefinerV4(encoder="vits", features=cfg["features"], out_channels=cfg["out_channels"]).cuda().eval()
    m.load_state_dict(sd)
    y = m(make_input("depth", (1, S, H, W), seed).cuda())
    assert y.shape == (1, S, H, W) and y.dtype == torch.float32
    e = _check("v4_vits_s4_56x84 vs reference", y / 65535.0, torch.from_numpy(g["out"]) / 65535.0, floor_frac=0.05)
    assert e["max_abs"] <= 1e-3
    with pytest.raises(RuntimeError):  # patch_embed.py:73-74 through the native-resolution path
        m(make_input("depth", (1, 2, 60, 84), seed).cuda())
    # the TPF harness of scripts/evaluate_v4.py:169-233 over this model: two refinement passes per batch, against the live reference
    from video_depth_normal_v2_b200.harness import evaluate_tpf
    d5 = make_input("depth", (2, S, H, W), seed + 1).unsqueeze(2) - 500.0
    res = evaluate_tpf(m, [{"depth_anything_v2": d5}, {"depth_anything_v2": d5}], max_eval_count=1, keep_outputs=True)
    assert res["frames"] == S and res["tpf_ms"] > 0 and len(res["outputs"]) == 1
    e2 = _check("v4 TPF harness (two passes) vs reference", res["outputs"][0] / 65535.0, torch.from_numpy(g["out_tpf"]) / 65535.0, floor_frac=0.05)
    assert e2["max_abs"] <= 2e-3


# ------------------------------------------------------------------------------------------ BASELINE configs[3]: ViT-L 518x924
def test_vitl_518x924_window_matches_oracle_on_gpu(vdn):
    """The non-square BASELINE size (37 x 66 patch grid, 2443 tokens, interpolated pos-embed), ViT-L, strict-fp32 oracle on the GPU."""
    enc, T, H, W = "vitl", 4, 518, 924
    m, sd = _model(vdn, enc, 15)
    x = make_input("rgb", (1, T, 3, H, W), 15).cuda()
    y = m(x)
    ref = O.vda_forward({k: v.cuda() for k, v in sd.items()}, x, enc)
    torch.cuda.synchronize()
    _check("vitl 4x518x924", y, ref)


def test_v5_vitl_32x518x924_matches_oracle_on_gpu(vdn):
    """BASELINE configs[3] as written: video_depth_model_v5, ViT-L, one 32-frame 518x924 depth clip (median scale head, 224x224
    Sobel-normal network input, temporal DPT head, residual output) and the surface normals of the refined depth."""
    cfg = ENCODERS["vitl"]
    sd = make_state_dict("v5", "vitl", 16)
    m = vdn.VideoDepthRefinerV5(encoder="vitl", features=cfg["features"], out_channels=cfg["out_channels"]).cuda().eval()
    m.load_state_dict(sd)
    d = make_input("depth", (1, 32, 518, 924), 16).cuda()
    y = m(d)
    ref = O.v5_forward({k: v.cuda() for k, v in sd.items()}, d, "vitl")
    torch.cuda.synchronize()
    e = _check("v5 vitl 32x518x924", y / 65535.0, ref / 65535.0, floor_frac=0.05)
    assert e["max_abs"] <= 1e-3
    n = torch.empty((32, 3, 518, 924), device="cuda")
    vdn.ops.sobel_normals((y[0] / 65535.0).contiguous(), n, 32, 518, 924)
    ang = O.normal_angle_deg(n.cpu(), O.sobel_normals((ref[0] / 65535.0)[:, None]).cpu())
    print(f"v5 vitl 32x518x924 sobel normals of the refined depth: max angle {ang:.4f} deg")
    assert ang <= MAX_ANGLE


def test_streaming_vitl_matches_oracle_on_gpu(vdn):
    """ViT-L streaming (head_dim 128 / 32 motion modules -> vectorised streaming attention, ring slots, graph replay from the third
    frame on) against the oracle's restatement of video_depth_stream.py, 6 frames."""
    m, sd = _model(vdn, "vitl", 18)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    state = {}
    for i in range(6):
        x = make_input("rgb", (1, 1, 3, 56, 70), 180 + i).cuda()
        d = m.stream_step(x[0, 0])
        ref = O.vda_stream_step(sd_gpu, x, "vitl", state)
        _check(f"vitl stream frame {i}", d[None], ref.reshape(1, 56, 70))
    m.reset_stream()
    d0 = m.stream_step(make_input("rgb", (1, 1, 3, 56, 70), 180).cuda()[0, 0])
    state = {}
    _check("vitl stream frame 0 after reset", d0[None], O.vda_stream_step(sd_gpu, make_input("rgb", (1, 1, 3, 56, 70), 180).cuda(), "vitl", state).reshape(1, 56, 70))
