"""`not gpu`: the window-sharded long-video driver on CPU with gloo, world_size 2 and 3.

The distributed logic under test is the product's (`video.sharded_video_depth`, `partition_windows`, `owned_output_range`,
`scale_shift_chain`, `WindowAligner`); the per-window forward and the four alignment primitives are replaced by CPU stand-ins
defined here (the product versions are CUDA kernels), and the checker is the oracle's restatement of the reference's
alignment loop (`oracle.vdn_oracle.align_windows`, video_depth.py:118-154)."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vdn_oracle as O
from video_depth_normal_v2_b200 import video as V

H, W = 12, 16


class CpuAlignOps:
    """torch-CPU stand-in for video.DeviceAlignOps (same contract)."""

    def scale_shift(self, pred, target):
        p, t = pred.double().flatten(), target.double().flatten()
        s, b = V._solve_scale_shift([float((p * p).sum()), float(p.sum()), float(p.numel()), float((p * t).sum()), float(t.sum())])
        return torch.tensor([s, b], dtype=torch.float32)

    def affine_clamp(self, x, ss, out=None):
        r = torch.clamp(x * ss[0] + ss[1], min=0)
        if out is None:
            return r
        out.copy_(r)
        return out

    def crossfade(self, pre, post, ss, w, out):
        out.copy_(pre * (1.0 - w) + torch.clamp(post * ss[0] + ss[1], min=0) * w)
        return out


def _frame_depth(f: int) -> torch.Tensor:
    y, x = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    return 1.0 + 0.5 * torch.sin(0.13 * f + 0.3 * x) + 0.3 * torch.cos(0.07 * f + 0.2 * y)


def fake_forward(win):
    """Per-window depth with a window-dependent scale / shift ambiguity (and some negative values, to exercise the clamp)."""
    a = 1.0 + 0.1 * (win[-1] % 7)
    b = 0.05 * (win[-1] % 5) - 0.4
    return torch.stack([_frame_depth(f) * a + b for f in win])


def _expected(n_frames):
    wins = V.window_schedule(n_frames)
    depth_list = [m.numpy() for w in wins for m in fake_forward(w)]
    return O.align_windows(depth_list, n_frames)


def _worker(rank, world, port, n_frames, out_path, gather):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wins = V.window_schedule(n_frames)
        out = V.sharded_video_depth(fake_forward, wins, n_frames, (H, W), torch.device("cpu"), CpuAlignOps(), gather=gather)
        if gather == "all" or rank == 0:
            assert out is not None and tuple(out.shape) == (n_frames, H, W)
            np.save(f"{out_path}.{rank}.npy", out.numpy())
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n_frames,gather", [(2, 100, "all"), (3, 150, "rank0"), (2, 33, "all"), (3, 40, "all"), (2, 20, "rank0")])
def test_sharded_driver_matches_reference_alignment(world, n_frames, gather):
    exp = _expected(n_frames)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "out")
        mp.spawn(_worker, args=(world, _free_port(), n_frames, path, gather), nprocs=world, join=True)
        ranks = range(world) if gather == "all" else [0]
        for r in ranks:
            got = np.load(f"{path}.{r}.npy")
            assert np.abs(got - exp).max() < 2e-5, (r, float(np.abs(got - exp).max()))


@pytest.mark.parametrize("n_frames", [1, 22, 23, 54, 100])
def test_single_process_aligner_matches_reference_alignment(n_frames):
    wins = V.window_schedule(n_frames)
    al = V.WindowAligner(len(wins), H, W, torch.device("cpu"), aops=CpuAlignOps())
    for w in wins:
        al.push(fake_forward(w))
    assert np.abs(al.result(n_frames).numpy() - _expected(n_frames)).max() < 2e-5


def test_partition_and_ownership_cover_every_frame_once():
    for n in (1, 10, 22, 23, 100, 4096):
        K = len(V.window_schedule(n))
        seen = np.zeros(n, np.int32)
        for k in range(K):
            lo, hi = V.owned_output_range(k, K, n)
            seen[lo:hi] += 1
        assert (seen == 1).all(), n
        for world in (1, 2, 3, 8):
            b = V.partition_windows(K, world)
            assert b[0][0] == 0 and b[-1][1] == K and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [x[1] - x[0] for x in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert len(V.window_schedule(4096)) == 187 and V.partition_windows(187, 8)[0] == (0, 24)


def test_feature_reuse_schedule_encodes_each_new_frame_once():
    """With feature reuse a steady-state window encodes 22 frames; the cache never holds more than one window."""
    class FakeModel:
        def encode_frames(self, x):
            return [x.reshape(x.shape[0] * 4, -1)[:, :8].clone() for _ in range(4)]  # P = 4 "tokens" per frame

        def head_from_features(self, feats, T, ph, pw):
            return feats[0].reshape(T, 4, 8)[:, 0, 0].reshape(T, 1, 1).expand(T, 28, 28).clone()

    n = 100
    frames = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1).expand(n, 3, 28, 28).contiguous()
    fwd = V.WindowForwarder(FakeModel(), frames, (28, 28), "cpu", reuse=True)
    wins = V.window_schedule(n)
    for k, w in enumerate(wins):
        d = fwd.forward(w)
        assert torch.equal(d[:, 0, 0], torch.tensor([float(f) for f in w]))  # slot order = window order
        assert len(fwd.cache) <= 32
    uniq = len({f for w in wins for f in w})
    assert fwd.encoded_frames == uniq == n  # every source frame encoded exactly once
