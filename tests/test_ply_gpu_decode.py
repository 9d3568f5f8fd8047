"""GPU decode of the PLY body (csrc/plydecode.cu).

CPU part: the expf the kernel uses is glibc's algorithm restated in double precision; this test lifts the table and
the constants out of the .cu source, replays the algorithm with numpy float64 (IEEE, like the device's doubles) and
compares with the host libm expf the CPU decode (GaussianCloud::ImportPly) calls.
GPU part: ImportPlyGpu == ImportPly bit for bit; InitFromPly renders the same image as ImportPly + Init."""
import ctypes
import os
import re

import numpy as np
import pytest

from splatapult_b200 import api, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NF = [scenes.Z_NEAR, scenes.Z_FAR]


def _kernel_expf_constants():
    src = open(os.path.join(ROOT, "splatapult_b200", "csrc", "plydecode.cu")).read()
    tab = re.search(r"kExp2Tab\[32\] = \{(.*?)\};", src, re.S).group(1)
    table = np.array([int(h, 16) for h in re.findall(r"0x([0-9a-f]{16})ull", tab)], np.uint64)
    assert len(table) == 32
    hexf = lambda name: float.fromhex(re.search(name + r" = (0x[0-9a-fp.+-]+)", src).group(1))
    return table, hexf("kInvLn2N"), hexf("kShift"), hexf("c0"), hexf("c1"), hexf("c2")


def _expf_replica(x32: np.ndarray) -> np.ndarray:
    table, inv_ln2n, shift, c0, c1, c2 = _kernel_expf_constants()
    x = x32.astype(np.float64)
    z = inv_ln2n * x
    kd = z + shift
    ki = kd.view(np.uint64)
    kd = kd - shift
    r = z - kd
    t = table[(ki & np.uint64(31)).astype(np.int64)] + (ki << np.uint64(47))
    s = t.view(np.float64)
    zz = c0 * r + c1
    r2 = r * r
    y = c2 * r + 1.0
    y = zz * r2 + y
    y = y * s
    return y.astype(np.float32)


def test_kernel_expf_is_the_hosts_libm_expf():
    libm = ctypes.CDLL("libm.so.6")
    libm.expf.restype = ctypes.c_float
    libm.expf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(5)
    # the decode domain: log-scales and opacity logits, plus a wide band; and values around rounding-sensitive spots
    xs = np.concatenate([rng.uniform(-20, 20, 150_000), rng.uniform(-79, 79, 50_000), rng.normal(0, 1e-3, 5_000),
                         np.array([0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 79.9, -79.9, np.log(0.005), np.log(0.05)])]).astype(np.float32)
    want = np.array([libm.expf(float(v)) for v in xs], np.float32)
    got = _expf_replica(xs)
    bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))[0]
    # checked exhaustively offline for |x| <= 40: a single float (0x1.04845ep+5) differs, by one ulp
    assert len(bad) == 0, (xs[bad][:5], got[bad][:5], want[bad][:5])


def test_import_ply_gpu_fails_loudly_without_a_device(golden_dir):
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    gc = api.GaussianCloud()
    assert not gc.ImportPlyGpu(os.path.join(golden_dir, "reference_test.ply"))
    assert "no CUDA device" in api.lib().spl_last_error().decode()


def _write_odd_ply(path, n, seed):
    """vertex rows with a leading uchar (unaligned floats), a double-typed scale_1 (reads as 0), no f_dc_2, no rot_3,
    duplicate property name, float properties in a shuffled order."""
    rng = np.random.default_rng(seed)
    names = ["flag:uchar", "rot_2", "x", "scale_0", "opacity", "y", "f_dc_0", "scale_1:double", "z", "rot_0", "rot_1", "f_dc_1", "scale_2", "x"]
    dt = []
    for k, nm in enumerate(names):
        base, _, typ = nm.partition(":")
        dt.append((f"{base}__{k}", {"": "<f4", "uchar": "u1", "double": "<f8"}[typ]))
    rows = np.zeros(n, np.dtype(dt))
    for f, (nm, t) in zip(rows.dtype.names, dt):
        rows[f] = rng.uniform(-2, 2, n) if t != "u1" else rng.integers(0, 255, n)
    header = "ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n
    for nm in names:
        base, _, typ = nm.partition(":")
        header += "property %s %s\n" % (typ or "float", base)
    header += "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode())
        f.write(rows.tobytes())
    return path


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["reference", "deg0", "deg3", "deg3_base_import", "odd", "empty"])
def test_gpu_decode_equals_host_decode(splat, golden_dir, scratch, case):
    if case == "reference":
        path, full = os.path.join(golden_dir, "reference_test.ply"), True
    elif case == "deg0":
        path, full = scenes.write_synthetic_ply(os.path.join(str(scratch), "dec0.ply"), 150_001, 0, seed=3), True
    elif case in ("deg3", "deg3_base_import"):
        path, full = scenes.write_synthetic_ply(os.path.join(str(scratch), "dec3.ply"), 60_007, 3, seed=4), case == "deg3"
    elif case == "odd":
        path, full = _write_odd_ply(os.path.join(str(scratch), "odd.ply"), 4_099, 9), True
    else:
        path, full = scenes.write_synthetic_ply(os.path.join(str(scratch), "dec_empty.ply"), 0, 0, seed=1), True
    host, dev = splat.GaussianCloud(import_full_sh=full), splat.GaussianCloud(import_full_sh=full)
    assert host.ImportPly(path)
    assert dev.ImportPlyGpu(path), splat.lib().spl_last_error()
    assert dev.GetNumGaussians() == host.GetNumGaussians() and dev.GetStride() == host.GetStride()
    assert dev.HasFullSH() == host.HasFullSH()
    a, b = host.GetRawData().view(np.uint32), dev.GetRawData().view(np.uint32)
    np.testing.assert_array_equal(b, a)  # bit for bit, signed zeros included


@pytest.mark.gpu
def test_init_from_ply_renders_the_same_frame(splat, golden_dir):
    path = os.path.join(golden_dir, "reference_test.ply")
    gc = splat.GaussianCloud()
    assert gc.ImportPly(path)
    proj = scenes.perspective(scenes.FOVY, 640 / 480)
    cam = scenes.test_ply_camera()
    ra = splat.SplatRenderer()
    assert ra.Init(gc, False, False)
    ra.Render(cam, proj, [0, 0, 640, 480], NF)
    want = ra.ReadFramebuffer()
    rb = splat.SplatRenderer()
    rb.InitFromPly(path, True, False)
    rb.Render(cam, proj, [0, 0, 640, 480], NF)
    np.testing.assert_array_equal(rb.ReadFramebuffer(), want)
    assert rb.GetStats().num_gaussians == gc.GetNumGaussians()
    with pytest.raises(api.SplatError):
        splat.SplatRenderer().InitFromPly(os.path.join(golden_dir, "does_not_exist.ply"))


@pytest.mark.gpu
def test_gpu_decode_equals_host_decode_on_mutated_headers(splat, scratch):
    """400 header mutations of a valid INRIA file (the family tests/test_ref_pin.py runs against the reference's own reader):
    the CUDA decode takes the host decode's decision and produces its bytes -- repeated / typed / missing properties, vertex
    sizes that are no multiple of four (byte loads), stale-token lines, claimed counts beyond the body (zero rows)."""
    base_path = str(scratch / "gpufuzz_base.ply")
    scenes.write_synthetic_ply(base_path, 37, 3, seed=13)
    with open(base_path, "rb") as fh:
        base = fh.read()
    hdr_end = base.index(b"end_header\n") + 11
    tail = bytes(64 * 1024)
    tokens = [b"property float x\n", b"property double x\n", b"property uchar red\n", b"property list uchar int vi\n", b"element face 0\n", b"element vertex 3\n",
              b"element vertex 41\n", b"comment hi\n", b"format ascii 1.0\n", b"end_header\n", b"property float\n", b"property int8 a\n", b"property uint16 b\n", b"property int c\n",
              b"property float32 f_dc_0\n", b"\r\n", b" ", b"\n", b"\t", b"0", b"1", b"ply\n", b"property float64 opacity\n", b"property short scale_0\n", b"property\n",
              b"property double\n", b"\n\n", b"property float y\n", b"property uchar q\nproperty uchar r\nproperty uchar s\n"]
    rng = np.random.default_rng(17)
    path = str(scratch / "gpufuzz_case.ply")
    accepted = odd = 0
    for it in range(400):
        hdr = bytearray(base[:hdr_end])
        for _ in range(int(rng.integers(1, 3))):
            kind, pos = int(rng.integers(0, 5)), int(rng.integers(0, len(hdr)))
            if kind == 0:
                hdr[pos] = int(rng.integers(0, 256))
            elif kind == 1:
                hdr[pos:pos] = tokens[int(rng.integers(0, len(tokens)))]
            elif kind == 2:
                del hdr[pos:pos + int(rng.integers(1, 20))]
            elif kind == 3:
                ls, le = bytes(hdr).rfind(b"\n", 0, pos) + 1, bytes(hdr).find(b"\n", pos) + 1
                if le > ls:
                    del hdr[ls:le]
            else:
                ls = bytes(hdr).rfind(b"\n", 0, pos) + 1
                hdr[ls:ls] = tokens[int(rng.integers(0, len(tokens)))]
        m = re.search(rb"element vertex (\d+)", bytes(hdr))
        if m and int(m.group(1)) > 4096:
            continue  # keep the scratch files small; huge counts are test_host_cloud's business
        with open(path, "wb") as fh:
            fh.write(bytes(hdr) + base[hdr_end:] + tail)
        full = bool(rng.integers(0, 2))
        host, dev = splat.GaussianCloud(import_full_sh=full), splat.GaussianCloud(import_full_sh=full)
        ok_h, ok_d = host.ImportPly(path), dev.ImportPlyGpu(path)
        assert ok_h == ok_d, (it, bytes(hdr)[-200:])
        if not ok_h:
            continue
        accepted += 1
        assert host.HasFullSH() == dev.HasFullSH() and host.GetNumGaussians() == dev.GetNumGaussians()
        a, b = host.GetRawData(), dev.GetRawData()
        # bit for bit, except that a NaN is a NaN: garbage read through a mutated layout overflows expf, and inf * 0 gives the
        # x86 "indefinite" 0xFFC00000 on the host and CUDA's canonical 0x7FFFFFFF on the device
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        np.testing.assert_array_equal(nan_a, nan_b, err_msg=f"case {it}")
        np.testing.assert_array_equal(a.view(np.uint32)[~nan_a], b.view(np.uint32)[~nan_b], err_msg=f"case {it}")
        odd += int(b"uchar" in bytes(hdr) or b"int8" in bytes(hdr) or b"uint16" in bytes(hdr) or b"short" in bytes(hdr))
    assert accepted > 80 and odd > 10, (accepted, odd)


@pytest.mark.gpu
def test_gpu_decode_of_poisoned_bodies(splat, scratch):
    """bodies full of infinities, NaN, denormals, zero quaternions and overflowing log-scales (the family the reference's
    source, the oracle and the host decode agree on bit for bit, tests/test_ref_pin.py): the CUDA decode produces the host
    decode's bits; where the host has a NaN the device has a NaN (the payloads are the platforms')"""
    rng = np.random.default_rng(4)
    path = str(scratch / "gpu_poison.ply")
    for it in range(100):
        n, deg = int(rng.integers(1, 300)), int(rng.choice([0, 3]))
        a = scenes.synthetic_cloud_arrays(n, deg, seed=int(rng.integers(0, 1 << 30)))
        scenes.poison_arrays(a, rng)
        scenes.write_ply_from_arrays(path, a)
        for full in (True, False):
            host, dev = splat.GaussianCloud(import_full_sh=full), splat.GaussianCloud(import_full_sh=full)
            assert host.ImportPly(path) and dev.ImportPlyGpu(path) and host.HasFullSH() == dev.HasFullSH()
            x, y = host.GetRawData(), dev.GetRawData()
            nx, ny = np.isnan(x), np.isnan(y)
            np.testing.assert_array_equal(nx, ny, err_msg=f"case {it}")
            np.testing.assert_array_equal(x.view(np.uint32)[~nx], y.view(np.uint32)[~ny], err_msg=f"case {it}")
