"""The C++ multi-GPU entry point: a C++ host with one thread per GPU replicates the cloud with ONE
ncclBroadcast inside libsplat_b200.so (SplatComm + SplatRenderer::InitBroadcast, NCCL bound at run time) and renders one
cameras.json view per GPU -- the multi-GPU form of src/app.cpp:535-559.  Every rank's frame is compared with the frame the
single-GPU path (Init from the host cloud) renders for the same camera: bit-identical."""
import os
import subprocess

import numpy as np
import pytest

from splatapult_b200 import api, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 640, 480


@pytest.fixture(scope="module")
def consumer(splat, scratch):
    exe = os.path.join(str(scratch), "multigpu_consumer")
    libdir = os.path.dirname(splat.library_path())
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "multigpu_consumer.cpp"),
           "-L", libdir, "-lsplat_b200", f"-Wl,-rpath,{libdir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _run(consumer, splat, scratch, nranks, deg):
    if splat.SplatComm.NcclVersion() == 0:
        pytest.skip("libnccl.so.2 not loadable")
    d = str(scratch)
    ply = scenes.write_synthetic_ply(os.path.join(d, f"mg_{deg}.ply"), 20000, deg, seed=11)
    cams = scenes.ring_cameras(8, 10.0, 2.0)
    js = scenes.write_cameras_json(os.path.join(d, "mg_cameras.json"), cams, W, H, 500.0)
    prefix = os.path.join(d, f"mg_{deg}_{nranks}_rank")
    proj = scenes.perspective(scenes.FOVY, W / H)
    projfile = os.path.join(d, "mg_proj.bin")
    proj.astype(np.float32).tofile(projfile)
    r = subprocess.run([consumer, ply, js, prefix, str(nranks), str(W), str(H), projfile], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert any(l.startswith(f"multigpu ok {nranks} ranks") for l in r.stdout.splitlines()), r.stdout  # NCCL may print its version first
    # the same views through the single-GPU path of the C ABI
    gc = splat.GaussianCloud()
    assert gc.ImportPly(ply)
    rr = splat.SplatRenderer()
    assert rr.Init(gc, False, False)
    cfg = splat.CamerasConfig()
    assert cfg.ImportJson(js)
    for rank in range(nranks):
        raw = np.fromfile(prefix + f"{rank}.raw", np.uint8)
        assert tuple(raw[:8].view(np.int32)) == (W, H)
        got = raw[8:].view(np.float32).reshape(H, W, 4)
        cam = cfg.GetCamera(rank % 8)[0]
        rr.Sort(cam, proj, [0, 0, W, H], [0.1, 1000.0])
        rr.Render(cam, proj, [0, 0, W, H], [0.1, 1000.0])
        want = rr.ReadFramebuffer()
        assert want[..., :3].max() > 0
        np.testing.assert_array_equal(got, want, err_msg=f"rank {rank}")


def test_broadcast_init_single_rank(consumer, splat, scratch):
    """nranks = 1: the NCCL path end to end on any box with one GPU"""
    _run(consumer, splat, scratch, 1, 3)


def test_broadcast_init_all_gpus(consumer, splat, scratch):
    n = api.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    _run(consumer, splat, scratch, min(n, 8), 0)
