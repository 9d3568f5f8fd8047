"""The C++ drop-in boundary: a translation unit that includes only
include/splatapult_b200/*.h -- the reference's class names and call protocol -- compiles with g++ and links
against libsplat_b200.so (the classes are exported next to the C ABI)."""
import os
import subprocess

import numpy as np
import pytest

from splatapult_b200 import api, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def consumer(splat, scratch):
    exe = os.path.join(str(scratch), "dropin_consumer")
    libdir = os.path.dirname(splat.library_path())
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dropin_consumer.cpp"),
           "-L", libdir, "-lsplat_b200", f"-Wl,-rpath,{libdir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cloud_classes_link_and_round_trip(consumer, scratch):
    out = os.path.join(str(scratch), "debug_cloud.ply")
    r = subprocess.run([consumer, "cloud", out], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip() == "cloud ok 16 stride 244"
    gc = api.GaussianCloud()
    assert gc.ImportPly(out) and gc.GetNumGaussians() == 16


def test_renderer_class_fails_loudly_without_a_gpu(consumer):
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([consumer, "nogpu"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 3
    assert "init failed" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_api_renders_the_same_frame_as_the_c_abi(consumer, splat, golden_dir, scratch):
    ply = os.path.join(golden_dir, "reference_test.ply")
    w, h = 640, 480
    cam, proj = scenes.test_ply_camera(), scenes.perspective(scenes.FOVY, w / h)
    args = np.concatenate([cam, proj, np.array([0, 0, w, h, scenes.Z_NEAR, scenes.Z_FAR], np.float32)]).astype(np.float32)
    assert args.size == 38
    argfile, outfile = os.path.join(str(scratch), "args.bin"), os.path.join(str(scratch), "cpp_frame.raw")
    args.tofile(argfile)
    r = subprocess.run([consumer, "render", ply, argfile, outfile], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = np.fromfile(outfile, np.uint8)
    dims = raw[:8].view(np.int32)
    assert tuple(dims) == (w, h)
    got = raw[8:].view(np.float32).reshape(h, w, 4)

    gc = splat.GaussianCloud()
    assert gc.ImportPly(ply)
    rr = splat.SplatRenderer()
    assert rr.Init(gc, False, False)
    rr.Sort(cam, proj, [0, 0, w, h], [scenes.Z_NEAR, scenes.Z_FAR])
    rr.Render(cam, proj, [0, 0, w, h], [scenes.Z_NEAR, scenes.Z_FAR])
    np.testing.assert_array_equal(got, rr.ReadFramebuffer())
