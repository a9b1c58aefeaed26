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

    def keys(self, d, out):
        out.copy_(torch.stack([d[0], d[1], d[12]]))
        return out

    def finalize(self, d, prev_tail, ss, ss_prev, out, first_slot, is_first):
        """Same contract as vdn_window_finalize (include/vdn_b200.h)."""
        for i in range(out.shape[0]):
            slot = first_slot + i
            if is_first:
                out[i].copy_(d[slot])
                continue
            cur = torch.clamp(d[slot] * ss[0] + ss[1], min=0)
            if slot < V.OVERLAP:
                j = slot - V.ALIGN_LEN
                pre = prev_tail[j] if ss_prev is None else torch.clamp(prev_tail[j] * ss_prev[0] + ss_prev[1], min=0)
                w = V.CROSSFADE_W[j]
                cur = pre * (1.0 - w) + cur * w
            out[i].copy_(cur)
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


class StubForwarder:
    """Stand-in for video.WindowForwarder with the same feature-exchange protocol (export_features / feature_buffers /
    import_features / prefetch): a frame's four "tapped features" are [P, C] tensors filled with a value derived from the frame
    index, the "encoder" produces them, the "head" checks that every slot of the window holds the features of its frame — wherever
    they came from (encoded here, kept from the previous window, or shipped by the next rank) — and returns fake_forward's depth."""
    P, C = 3, 4

    def __init__(self):
        self.prev_win, self.slots, self.seeded = None, None, {}
        self.encoded, self.imported, self.prefetched = [], [], []

    def _feat(self, f):
        return [torch.full((self.P, self.C), float(f * 4 + t)) for t in range(4)]

    def prefetch(self, win, arriving=()):
        have = set(self.prev_win or ()) | set(self.seeded) | set(arriving)
        self.prefetched.append([f for f in dict.fromkeys(win) if f not in have])

    def forward(self, win):
        win = list(win)
        prev = dict(zip(self.prev_win, self.slots)) if self.prev_win is not None else {}
        slots, fresh = [], {}
        for f in win:
            if f in prev:
                slots.append(prev[f])
            elif f in self.seeded:
                slots.append(self.seeded[f])
            else:
                if f not in fresh:  # a frame repeated inside a window (padded tail) is encoded once
                    fresh[f] = self._feat(f)
                    self.encoded.append(f)
                slots.append(fresh[f])
        for f, feat in zip(win, slots):  # the head sees the right features in every slot
            for t in range(4):
                assert torch.equal(feat[t], torch.full((self.P, self.C), float(f * 4 + t))), (f, t)
        self.prev_win, self.slots, self.seeded = win, slots, {}
        return fake_forward(win)

    def export_features(self, frames):
        prev = dict(zip(self.prev_win, self.slots))
        return [torch.cat([prev[f][t] for f in frames]).contiguous() for t in range(4)]

    def feature_buffers(self, n_frames):
        return [torch.empty((n_frames * self.P, self.C)) for _ in range(4)]

    def import_features(self, frames, packed):
        self.imported += list(frames)
        for j, f in enumerate(frames):
            self.seeded[f] = [packed[t][j * self.P:(j + 1) * self.P] for t in range(4)]


def _worker(rank, world, port, n_frames, out_path, gather, with_forwarder=False):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wins = V.window_schedule(n_frames)
        fwd = StubForwarder() if with_forwarder else None
        out = V.sharded_video_depth(fwd.forward if fwd else fake_forward, wins, n_frames, (H, W), torch.device("cpu"), CpuAlignOps(), gather=gather,
                                    forwarder=fwd)
        bounds = V.partition_windows(len(wins), world)
        if gather == "shard":
            lo, hi = V.rank_output_range(bounds, rank, len(wins), n_frames)
            assert tuple(out.shape) == (hi - lo, H, W)
            np.save(f"{out_path}.{rank}.npy", out.numpy())
        elif gather == "all" or rank == 0:
            assert out is not None and tuple(out.shape) == (n_frames, H, W)
            np.save(f"{out_path}.{rank}.npy", out.numpy())
        else:
            assert out is None
        if fwd is not None:
            k0, k1 = bounds[rank]
            has_next = rank + 1 < world and bounds[rank + 1][1] > bounds[rank + 1][0]
            if has_next and k1 - k0 >= 2:
                # the 9 boundary key frames came from the next rank and were neither encoded nor uploaded for the last window
                shipped = wins[k1][1:V.OVERLAP]
                assert fwd.imported == list(shipped)
                own_last_new = [f for f in dict.fromkeys(wins[k1 - 1]) if f not in set(wins[k1 - 2])]
                assert not (set(shipped) & set(fwd.encoded)), "boundary frames were encoded although the next rank shipped them"
                assert set(fwd.prefetched[-1]) == set(own_last_new) - set(shipped)
            else:
                assert fwd.imported == []
            np.save(f"{out_path}.enc{rank}.npy", np.array(fwd.encoded))
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


@pytest.mark.parametrize("world,n_frames", [(2, 100), (3, 150), (3, 60), (2, 40), (3, 300)])
def test_sharded_driver_boundary_feature_exchange_and_shard_gather(world, n_frames):
    """Collective (1) (boundary key-frame features, rank r+1 -> r) with a stub forwarder — including a single-window rank (3 ranks,
    60 frames = 3 windows) and padded tail windows — and gather='shard': the per-rank shards tile the reference result."""
    exp = _expected(n_frames)
    wins = V.window_schedule(n_frames)
    bounds = V.partition_windows(len(wins), world)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "out")
        mp.spawn(_worker, args=(world, _free_port(), n_frames, path, "shard", True), nprocs=world, join=True)
        covered = np.zeros(n_frames, np.int32)
        encoded = []
        for r in range(world):
            lo, hi = V.rank_output_range(bounds, r, len(wins), n_frames)
            got = np.load(f"{path}.{r}.npy")
            assert got.shape[0] == hi - lo
            if hi > lo:
                assert np.abs(got - exp[lo:hi]).max() < 2e-5, (r, float(np.abs(got - exp[lo:hi]).max()))
            covered[lo:hi] += 1
            encoded.append(np.load(f"{path}.enc{r}.npy"))
        assert (covered == 1).all()
        # every source frame is encoded on some rank; only frame 0 (slot 0 of every window) and the boundary frames of ranks with a
        # single window are encoded more than once
        allenc = np.concatenate(encoded)
        assert set(allenc.tolist()) == set(range(n_frames))
        multi_window_ranks = sum(1 for b in bounds if b[1] - b[0] >= 2)
        if multi_window_ranks == sum(1 for b in bounds if b[1] > b[0]):
            dup = len(allenc) - len(set(allenc.tolist()))
            assert dup == sum(1 for b in bounds[1:] if b[1] > b[0]), dup  # frame 0 once per extra rank


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
    """With feature reuse a steady-state window encodes 22 frames; the window buffers always hold the slots' own frames."""
    class FakeModel:
        def encode_frames(self, x, clone=True):
            return [x.reshape(x.shape[0] * 4, -1)[:, :8].clone() for _ in range(4)]  # P = 4 "tokens" per frame

        def head_from_features(self, feats, T, ph, pw, static_inputs=False, clone=True):
            return feats[0].reshape(T, 4, 8)[:, 0, 0].reshape(T, 1, 1).expand(T, 28, 28).clone()

    n = 100
    frames = torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1).expand(n, 3, 28, 28).contiguous()
    fwd = V.WindowForwarder(FakeModel(), frames, (28, 28), "cpu", reuse=True)
    wins = V.window_schedule(n)
    for k, w in enumerate(wins):
        d = fwd.forward(w)
        assert torch.equal(d[:, 0, 0], torch.tensor([float(f) for f in w]))  # slot order = window order
    uniq = len({f for w in wins for f in w})
    assert fwd.encoded_frames == uniq == n  # every source frame encoded exactly once
