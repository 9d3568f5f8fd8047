"""Mutation fuzz of the host-side PLY and cameras.json readers under AddressSanitizer + UBSan: malformed files must
be rejected (or accepted with zeros, like the reference reader), never crash, overflow or leak out an exception."""
import os
import shutil
import subprocess

import pytest

from splatapult_b200 import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "splatapult_b200", "csrc")


def test_readers_survive_mutated_inputs(scratch):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    d = os.path.join(str(scratch), "fuzz")
    os.makedirs(d, exist_ok=True)
    exe = os.path.join(d, "fuzz_host_readers")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", CSRC, os.path.join(ROOT, "tests", "cpp", "fuzz_host_readers.cpp"),
           os.path.join(CSRC, "ply.cpp"), os.path.join(CSRC, "gaussiancloud.cpp"), os.path.join(CSRC, "camerasconfig.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr or "").lower():
        pytest.skip("sanitizer runtime not available")
    assert r.returncode == 0, r.stderr
    ply = scenes.write_synthetic_ply(os.path.join(d, "seed.ply"), 40, 3, seed=1)
    js = scenes.write_cameras_json(os.path.join(d, "seed.json"), scenes.ring_cameras(3, 10.0, 2.0), 640, 480, 500.0)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([exe, ply, js, d, "4000"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert "fuzz done" in r.stdout
