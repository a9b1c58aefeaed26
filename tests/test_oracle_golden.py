"""`not gpu`: pin the oracle (oracle/vdn_oracle.py) to the live reference's outputs (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import vdn_oracle as O
from oracle.init_recipe import make_input, make_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")
# fp32 CPU, different op order (functional vs nn.Module) -> round-off only
RTOL = 2e-4


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


@pytest.mark.parametrize("name,enc", [("vda_vits_t4_70x84", "vits"), ("vda_vits_t2_518x518", "vits"), ("vda_vitl_t2_56x70", "vitl")])
def test_vda_forward_matches_reference(name, enc):
    g = _load(name)
    T, H, W, seed, stride = [int(v) for v in g["meta"]]
    sd = make_state_dict("vda", enc, seed)
    x = make_input("rgb", (1, T, 3, H, W), seed)
    stages = {}
    y = O.vda_forward(sd, x, enc, stages=stages)
    assert y.shape == (1, T, H, W)
    ref = torch.from_numpy(g["depth"])
    e = O.depth_errors(y[:, :, ::stride, ::stride], ref)
    assert e["max_rel"] < RTOL, e
    assert abs(float(y.double().sum()) - float(g["depth_sum"])) / float(g["depth_sum"]) < 1e-5
    tap = stages["taps"][3][:, ::7, ::5]
    assert torch.allclose(tap, torch.from_numpy(g["tap3"]), atol=2e-4, rtol=1e-3)
    assert float(ref.min()) > 0  # recipe guarantees a non-degenerate (strictly positive) output


def test_v5_forward_matches_reference():
    g = _load("v5_vits_s4_60x80")
    S, H, W, seed = [int(v) for v in g["meta"]]
    sd = make_state_dict("v5", "vits", seed)
    d = make_input("depth", (1, S, H, W), seed)
    y = O.v5_forward(sd, d, "vits")
    ref = torch.from_numpy(g["out"])
    assert (y - ref).abs().max() / ref.abs().max() < 1e-4
    assert (ref - d).abs().mean() > 100  # the de-zeroed recipe makes v5 a non-identity map


def test_window_schedule_matches_reference_driver():
    g = _load("window_schedule")
    for key in g.files:
        n = int(key[1:])
        wins, n_padded = O.window_schedule(n)
        # frame ids were stored mod 256 in a uint8 pixel
        assert (np.array(wins) % 256 == g[key]).all(), key
        assert len(wins) == len(range(0, n, 22))


def test_infer_video_depth_matches_reference():
    from tests.golden.gen_golden import video_frames
    g = _load("video_vits_n50_56x70")
    N, H, W, seed = [int(v) for v in g["meta"]]
    sd = make_state_dict("vda", "vits", seed)
    frames = video_frames(N, H, W, seed)
    # util/transform.py:125-158 with an identity Resize: float32/255 -> float64 normalise -> float32 CHW
    mean, std = np.array([0.485, 0.456, 0.406]), np.array([0.229, 0.224, 0.225])
    ft = ((frames.astype(np.float32) / 255.0 - mean) / std).transpose(0, 3, 1, 2).astype(np.float32)
    out = O.infer_video_depth_tensor(sd, torch.from_numpy(ft), "vits", (H, W))
    ref = g["depths"]
    assert out.shape == ref.shape
    e = O.depth_errors(torch.from_numpy(out), torch.from_numpy(ref))
    assert e["max_rel"] < 5e-4, e


def test_operand_rounding_budget():
    """Documents why operands are fp16 by default: bf16 operand rounding alone costs > the 1e-3 AbsRel budget."""
    sd = make_state_dict("vda", "vits", 0)
    x = make_input("rgb", (1, 4, 3, 70, 84), 0)
    y = O.vda_forward(sd, x, "vits")
    e16 = O.depth_errors(O.vda_forward(sd, x, "vits", operand_dtype=torch.float16), y)
    ebf = O.depth_errors(O.vda_forward(sd, x, "vits", operand_dtype=torch.bfloat16), y)
    assert e16["abs_rel"] < 1e-3 and e16["max_rel"] < 1e-2, e16
    assert ebf["abs_rel"] > e16["abs_rel"]
    print("fp16", e16, "bf16", ebf)


# ------------------------------------------------------------------------------------------ a11: DepthAnythingV2 + memory block
def _da2_inputs(B, H, calls, seed):
    return [make_input("rgb", (B, 1, 3, H, H), seed * 100 + i)[:, 0] for i in range(calls)]


@pytest.mark.parametrize("name,enc", [("da2_vits_b2_70_calls8", "vits"), ("da2_vits_b1_518_calls2", "vits"), ("da2_vitl_b1_70_calls3", "vitl"),
                                      ("da2_vitb_b2_70_calls3", "vitb"), ("da2_vits_b2_70_calls3_cls", "vits"), ("da2_vitg_b1_56_calls2", "vitg")])
def test_da2_stateful_forward_matches_reference(name, enc):
    """A sequence of forward() calls on one model (memory bank filling up, then wrapping at 6 entries)."""
    g = _load(name)
    B, H, calls, seed, stride = [int(v) for v in g["meta"]]
    sd = make_state_dict("da2", enc, seed, use_clstoken=name.endswith("_cls"))
    bank = []
    for i, x in enumerate(_da2_inputs(B, H, calls, seed)):
        y = O.da2_forward(sd, x, enc, bank)
        ref = torch.from_numpy(g["depth"][i])
        assert (y[:, ::stride, ::stride] - ref).abs().max() / ref.abs().max() < 2e-4, (name, i)
        assert len(bank) == min(i + 1, 6)


def test_da2_memory_changes_the_output():
    """Teeth for the parity tests: the same frame with an empty and with a filled bank gives measurably different depth."""
    sd = make_state_dict("da2", "vits", 5)
    xs = _da2_inputs(1, 70, 3, 5)
    bank = []
    first = O.da2_forward(sd, xs[0], "vits", bank)
    O.da2_forward(sd, xs[1], "vits", bank)
    with_mem = O.da2_forward(sd, xs[0], "vits", bank)
    assert float((with_mem - first).abs().mean() / first.abs().mean()) > 2e-3  # twice the AbsRel parity tolerance


# ------------------------------------------------------------------------------------------ f1: streaming inference
def test_stream_steps_match_reference():
    """video_depth_stream.py infer_video_depth_one, 16 frames (cache list growing to 42 entries, then sliding)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from gen_golden import video_frames
    from video_depth_normal_v2_b200 import video as V
    g = _load("stream_vits_n16_56x70")
    N, H, W, seed = [int(v) for v in g["meta"]]
    sd = make_state_dict("vda", "vits", seed)
    ft = torch.from_numpy(V.preprocess_frames(video_frames(N, H, W, seed), min(H, W)))
    state = {}
    for i in range(N):
        y = O.vda_stream_step(sd, ft[i][None, None], "vits", state)
        ref = torch.from_numpy(g["depths"][i])
        assert (y - ref).abs().max() / ref.abs().max() < 2e-4, i
    assert len(state["cache"]) == 42


# ------------------------------------------------------------------------------------------ f3: constructor switches
SWITCH_CASES = [("vda_vits_t4_70x84_rope", "vits", {"pe": "rope"}), ("vda_vits_t4_70x84_cls", "vits", {"use_clstoken": True}),
                ("vda_vitl_t3_56x70_rope_cls", "vitl", {"pe": "rope", "use_clstoken": True})]


@pytest.mark.parametrize("name,enc,kw", SWITCH_CASES)
def test_vda_switches_match_reference(name, enc, kw):
    """pe='rope' (motion_module.py:236-240,279-282; attention.py:403-429) and use_clstoken=True (dpt.py:92-98,129-132)."""
    g = _load(name)
    T, H, W, seed = [int(v) for v in g["meta"]]
    sd = make_state_dict("vda", enc, seed, **kw)
    assert any("pos_encoder.pe" in k for k in sd) == (kw.get("pe", "ape") == "ape")
    assert any("readout_projects" in k for k in sd) == bool(kw.get("use_clstoken"))
    y = O.vda_forward(sd, make_input("rgb", (1, T, 3, H, W), seed), enc)
    e = O.depth_errors(y, torch.from_numpy(g["depth"]))
    assert e["max_rel"] < RTOL, e
    # the switch changes the function (the golden is not the default model's output)
    base = {k: v for k, v in make_state_dict("vda", enc, seed).items()}
    y0 = O.vda_forward(base, make_input("rgb", (1, T, 3, H, W), seed), enc)
    assert float((y0 - y).abs().mean() / y.abs().mean()) > 5e-3


def test_stream_steps_rope_cls_match_reference():
    """Streaming with pe='rope': a one-frame query makes the reference rotate by freqs_cis[:1] = identity (see temporal_attention)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from gen_golden import video_frames
    from video_depth_normal_v2_b200 import video as V
    g = _load("stream_vits_n5_56x70_rope_cls")
    N, H, W, seed = [int(v) for v in g["meta"]]
    sd = make_state_dict("vda", "vits", seed, pe="rope", use_clstoken=True)
    ft = torch.from_numpy(V.preprocess_frames(video_frames(N, H, W, seed), min(H, W)))
    state = {}
    for i in range(N):
        y = O.vda_stream_step(sd, ft[i][None, None], "vits", state)
        ref = torch.from_numpy(g["depths"][i])
        assert (y - ref).abs().max() / ref.abs().max() < 2e-4, i


def test_v4_forward_matches_reference():
    """models/video_depth_model_v4.py:120-148 (network at the native resolution): v5_forward(net_size=None)."""
    g = _load("v4_vits_s4_56x84")
    S, H, W, seed = [int(v) for v in g["meta"]]
    sd = make_state_dict("v5", "vits", seed)
    d = make_input("depth", (1, S, H, W), seed)
    y = O.v5_forward(sd, d, "vits", net_size=None)
    ref = torch.from_numpy(g["out"])
    assert (y - ref).abs().max() / ref.abs().max() < 1e-4
    assert (O.v5_forward(sd, d, "vits") - ref).abs().max() / ref.abs().max() > 1e-3  # the 224x224 variant is a different function
    # the TPF loop of scripts/evaluate_v4.py:186-196: clamp(min=0) of a (B, S, 1, H, W) batch, two refinement passes
    d5 = make_input("depth", (2, S, H, W), seed + 1).unsqueeze(2) - 500.0
    x = d5.clamp(min=0).squeeze(2)
    y2 = O.v5_forward(sd, O.v5_forward(sd, x, "vits", net_size=None), "vits", net_size=None)
    ref2 = torch.from_numpy(g["out_tpf"])
    assert (y2 - ref2).abs().max() / ref2.abs().max() < 1e-4
