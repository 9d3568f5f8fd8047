"""N > 1 host logic on CPU: world_size 2, gloo backend (views shard one per rank, the cloud is
broadcast once, per-rank timings are max-reduced).  The CUDA path itself is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest

from splatapult_b200 import scenes, sharding


def test_views_round_robin():
    assert sharding.views_for_rank(8, 0, 8) == [0]
    assert sharding.views_for_rank(8, 7, 8) == [7]
    assert sharding.views_for_rank(8, 1, 2) == [1, 3, 5, 7]
    assert sharding.views_for_rank(3, 2, 4) == [2] and sharding.views_for_rank(3, 3, 4) == []
    covered = sorted(v for r in range(3) for v in sharding.views_for_rank(8, r, 3))
    assert covered == list(range(8))
    with pytest.raises(ValueError):
        sharding.views_for_rank(8, 2, 2)


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, ply: str, out_dir: str):
    import torch.distributed as dist

    import splatapult_b200 as sp

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        aos, full = None, None
        if rank == 0:  # only the source rank touches the file: same .ply load path as a single-GPU run
            gc = sp.GaussianCloud(import_full_sh=True, export_full_sh=True)
            assert gc.ImportPly(ply)
            aos, full = gc.GetRawData(), gc.HasFullSH()
        buf, got_full, ms = sharding.broadcast_cloud(aos, full, src=0)
        assert ms is None  # CPU tensors: no device timing
        cams = scenes.ring_cameras(8, 18.0, 3.0)
        mine = sharding.views_for_rank(len(cams), rank, world)
        # a rank-dependent "frame time": the reduction must return the slowest rank's
        (slowest,) = sharding.max_over_ranks([10.0 + rank])
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), buf.numpy())
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
            f.write(f"{int(got_full)} {slowest} {' '.join(map(str, mine))}")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_broadcast_and_shard_world2(splat, scratch):
    import torch.multiprocessing as mp

    ply = str(scratch / "shard.ply")
    scenes.write_synthetic_ply(ply, 4000, 3, seed=3)
    out = str(scratch / "shard_out")
    os.makedirs(out, exist_ok=True)
    mp.spawn(_worker, args=(2, _free_port(), ply, out), nprocs=2, join=True)
    a0, a1 = np.load(os.path.join(out, "rank0.npy")), np.load(os.path.join(out, "rank1.npy"))
    assert a0.shape == (4000, 61)
    np.testing.assert_array_equal(a0, a1)  # the replica is bit-identical
    gc = splat.GaussianCloud()
    assert gc.ImportPly(ply)
    np.testing.assert_array_equal(a0, gc.GetRawData())
    r0 = open(os.path.join(out, "rank0.txt")).read().split()
    r1 = open(os.path.join(out, "rank1.txt")).read().split()
    assert r0[0] == r1[0] == "1"
    assert float(r0[1]) == float(r1[1]) == 11.0  # max over ranks
    assert list(map(int, r0[2:])) == [0, 2, 4, 6] and list(map(int, r1[2:])) == [1, 3, 5, 7]


def test_ring_cameras_look_at_origin():
    for c in scenes.ring_cameras(8, 18.0, 3.0):
        m = c.reshape(4, 4)  # m[col,row]
        eye, fwd = m[3, :3], -m[2, :3]
        np.testing.assert_allclose(np.linalg.norm(eye), np.hypot(18.0, 3.0), rtol=1e-5)
        np.testing.assert_allclose(fwd, -eye / np.linalg.norm(eye), atol=1e-5)
        np.testing.assert_allclose(m[:3, :3] @ m[:3, :3].T, np.eye(3), atol=1e-5)
