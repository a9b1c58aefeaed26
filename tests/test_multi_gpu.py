"""`-m gpu`, needs >= 2 GPUs (skipped otherwise): the NCCL window-sharded infer_video_depth against the single-GPU result."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    from oracle.init_recipe import make_state_dict
    from video_depth_normal_v2_b200 import VideoDepthAnything
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        m = VideoDepthAnything(encoder="vits", features=64, out_channels=[48, 96, 192, 384]).cuda().eval()
        m.load_state_dict(make_state_dict("vda", "vits", 0))
        rng = np.random.RandomState(3)
        frames = rng.randint(0, 255, (120, 56, 70, 3), dtype=np.uint8)
        st = {}
        out, _ = m.infer_video_depth(frames, 30, input_size=56, device="cuda", shard=True, gather="all", stats=st)
        np.save(f"{out_path}.{rank}.npy", out)
        sh, _ = m.infer_video_depth(frames, 30, input_size=56, device="cuda", shard=True, gather="shard")
        np.save(f"{out_path}.shard{rank}.npy", np.asarray(sh))
        np.save(f"{out_path}.range{rank}.npy", np.array(sh.frame_range))
        r0, _ = m.infer_video_depth(frames, 30, input_size=56, device="cuda", shard=True, gather="rank0")
        assert (r0 is None) == (rank != 0)
        if rank == 0:
            assert np.array_equal(r0, out)
            print("gather path:", st["gather_path"], st.get("gather_note"), flush=True)
        # a plain drop-in call inside an initialised process group stays local: no collective, full result on this rank
        if rank == 1:
            solo, _ = m.infer_video_depth(frames[:40], 30, input_size=56, device="cuda")
            assert solo.shape[0] == 40
        # ranks handed different clips are refused instead of deadlocking / mixing data
        bad = frames.copy()
        if rank == 1:
            bad[0, 0, 0, 0] ^= 1
        try:
            m.infer_video_depth(bad, 30, input_size=56, device="cuda", shard=True)
            raise AssertionError("mismatched clips were accepted")
        except RuntimeError as e:
            assert "different clips" in str(e)
    finally:
        dist.destroy_process_group()


def test_sharded_infer_video_depth_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle.init_recipe import make_state_dict
    from video_depth_normal_v2_b200 import VideoDepthAnything
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "o")
        mp.spawn(_worker, args=(2, port, path), nprocs=2, join=True)
        a, b = np.load(f"{path}.0.npy"), np.load(f"{path}.1.npy")
        shards = [(np.load(f"{path}.range{r}.npy"), np.load(f"{path}.shard{r}.npy")) for r in range(2)]
    assert np.array_equal(a, b)
    assert shards[0][0][0] == 0 and shards[0][0][1] == shards[1][0][0] and shards[1][0][1] == 120
    assert np.array_equal(np.concatenate([s_[1] for s_ in shards]), a)
    m = VideoDepthAnything(encoder="vits", features=64, out_channels=[48, 96, 192, 384]).cuda().eval()
    m.load_state_dict(make_state_dict("vda", "vits", 0))
    frames = np.random.RandomState(3).randint(0, 255, (120, 56, 70, 3), dtype=np.uint8)
    ref, _ = m.infer_video_depth(frames, 30, input_size=56, device="cuda")
    no_reuse, _ = m.infer_video_depth(frames, 30, input_size=56, device="cuda", reuse_features=False)
    # same kernels on the same inputs: only the fp64 atomics of the five LSQ sums may reorder
    assert np.abs(a - ref).max() <= 1e-5 * max(1.0, float(np.abs(ref).max())), float(np.abs(a - ref).max())
    assert np.abs(no_reuse - ref).max() <= 1e-5 * max(1.0, float(np.abs(ref).max()))
