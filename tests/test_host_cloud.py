"""Host-side parity (CPU only): the product's C++ GaussianCloud / Ply (through the C ABI) against the oracle's
restatement of GaussianCloud::ImportPly (gaussiancloud.cpp:138-365) and Ply::Parse (ply.cpp:72-254)."""
import os

import numpy as np
import pytest

from splatapult_b200 import scenes


def _import(splat, path, full=True):
    gc = splat.GaussianCloud(import_full_sh=full, export_full_sh=True)
    ok = gc.ImportPly(path)
    return gc, ok


def test_import_reference_fixture_bit_exact(splat, oracle, golden_dir):
    path = os.path.join(golden_dir, "reference_test.ply")
    gc, ok = _import(splat, path)
    assert ok and gc.GetNumGaussians() == 16 and gc.GetStride() == 244 and gc.HasFullSH()
    assert gc.GetTotalSize() == 16 * 244
    aos, full = oracle.import_ply(path, True)
    assert full
    np.testing.assert_array_equal(gc.GetRawData(), aos)
    np.testing.assert_array_equal(gc.GetRawData(), np.load(os.path.join(golden_dir, "test_ply_decode.npy")))


def test_import_nosh(splat, oracle, golden_dir):
    path = os.path.join(golden_dir, "reference_test.ply")
    gc, ok = _import(splat, path, full=False)
    assert ok and gc.GetStride() == 100 and not gc.HasFullSH()
    aos, full = oracle.import_ply(path, False)
    np.testing.assert_array_equal(gc.GetRawData(), aos)


@pytest.mark.parametrize("deg", [0, 3])
def test_import_synthetic_bit_exact(splat, oracle, scratch, deg):
    """random quaternions / scales / opacities: sigmoid, exp, R S S^T R^T in the reference's operation order."""
    path = str(scratch / f"synth_{deg}.ply")
    scenes.write_synthetic_ply(path, 5000, deg, seed=11)
    gc, ok = _import(splat, path)
    assert ok
    aos, full = oracle.import_ply(path, True)
    assert gc.HasFullSH() == full == (deg == 3)
    assert gc.GetStride() == (244 if deg == 3 else 100)
    np.testing.assert_array_equal(gc.GetRawData(), aos)
    # sanity of the decode itself, independent of either implementation
    a = scenes.synthetic_cloud_arrays(5000, deg, seed=11)
    np.testing.assert_array_equal(aos[:, 0:3], a["pos"])
    np.testing.assert_allclose(aos[:, 3], 1 / (1 + np.exp(-a["opacity"][:, 0].astype(np.float64))), rtol=2e-7)
    cov = aos[:, 16:25].reshape(-1, 3, 3).astype(np.float64)
    w, x, y, z = (a["rot"][:, i].astype(np.float64) for i in range(4))
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                  np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                  np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)
    S2 = np.exp(a["log_scale"].astype(np.float64)) ** 2
    ref = np.einsum("nij,nj,nkj->nik", R, S2, R)
    np.testing.assert_allclose(cov, ref, rtol=0, atol=2e-6 * np.abs(ref).max())
    if deg == 3:
        # f_rest_{15c+k-1} -> {r,g,b}_sh{k>>2}[k&3]
        blocks = [[4, 25, 29, 33], [8, 37, 41, 45], [12, 49, 53, 57]]
        for c in range(3):
            for k in (1, 5, 15):
                np.testing.assert_array_equal(aos[:, blocks[c][k >> 2] + (k & 3)], a["f_rest"][:, 15 * c + k - 1])


def test_attribute_table(splat, golden_dir):
    gc, _ = _import(splat, os.path.join(golden_dir, "reference_test.ply"))
    expect = {"posWithAlpha": 0, "r_sh0": 16, "g_sh0": 32, "b_sh0": 48, "cov3_col0": 64, "cov3_col1": 76, "cov3_col2": 88,
              "r_sh1": 100, "r_sh2": 116, "r_sh3": 132, "g_sh1": 148, "g_sh2": 164, "g_sh3": 180,
              "b_sh1": 196, "b_sh2": 212, "b_sh3": 228}
    for name, off in expect.items():
        t, s, o = gc.GetAttrib(name)
        assert (t, s, o) == (7, 4, off), name  # 7 = BinaryAttribute::Type::Float
    base, _ = _import(splat, os.path.join(golden_dir, "reference_test.ply"), full=False)
    assert base.GetAttrib("r_sh1") == (0, 0, 0)  # Unknown when not full SH (gaussiancloud.cpp:644-656)


def test_debug_cloud(splat, oracle):
    gc = splat.GaussianCloud()
    gc.InitDebugCloud()
    assert gc.GetNumGaussians() == 16 and gc.GetStride() == 244 and not gc.HasFullSH()
    np.testing.assert_array_equal(gc.GetRawData(), oracle.debug_cloud())


def test_export_import_round_trip(splat, scratch):
    src = str(scratch / "rt_src.ply")
    scenes.write_synthetic_ply(src, 3000, 3, seed=5)
    gc, ok = _import(splat, src)
    assert ok
    out = str(scratch / "rt_out.ply")
    assert gc.ExportPly(out)
    gc2, ok2 = _import(splat, out)
    assert ok2 and gc2.GetNumGaussians() == 3000 and gc2.HasFullSH()
    a, b = gc.GetRawData(), gc2.GetRawData()
    np.testing.assert_array_equal(a[:, 0:3], b[:, 0:3])
    np.testing.assert_allclose(a[:, 3], b[:, 3], rtol=2e-6)
    # every SH coefficient survives (the reference's exporter scrambles f_rest, SURVEY q10 -- fixed here)
    sh = list(range(4, 16)) + list(range(25, 61))
    np.testing.assert_array_equal(a[:, sh], b[:, sh])
    # covariance survives eigen-decomposition -> (quat, log-scale) -> recomposition
    np.testing.assert_allclose(a[:, 16:25], b[:, 16:25], rtol=0, atol=2e-5 * np.abs(a[:, 16:25]).max())


def test_prune_keeps_nearest(splat, scratch):
    src = str(scratch / "prune.ply")
    scenes.write_synthetic_ply(src, 2000, 0, seed=9)
    gc, _ = _import(splat, src)
    before = gc.GetRawData().copy()
    origin = np.array([0.5, -1.0, 2.0], np.float32)
    gc.PruneSplats(origin, 100)
    assert gc.GetNumGaussians() == 100
    d = np.linalg.norm(before[:, 0:3] - origin, axis=1)
    keep = np.argsort(d, kind="stable")[:100]
    np.testing.assert_array_equal(gc.GetRawData(), before[keep])
    gc.PruneSplats(origin, 1000)  # more than we have: no-op
    assert gc.GetNumGaussians() == 100


def _write(path, header: bytes, body: bytes = b""):
    with open(path, "wb") as f:
        f.write(header + body)
    return str(path)


HEADER_OK = b"ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\nend_header\n"


@pytest.mark.parametrize("name,header,accept", [
    ("ok", HEADER_OK, True),
    ("comments", HEADER_OK.replace(b"element", b"comment made by nobody\nelement").replace(b"end_header", b"comment x\nend_header"), True),
    ("aliases", HEADER_OK.replace(b"float x", b"float32 x").replace(b"float y", b"float32 y"), True),
    ("extra_typed_props", HEADER_OK.replace(b"end_header", b"property uchar red\nproperty double w\nproperty int16 s\nend_header"), True),
    ("crlf", HEADER_OK.replace(b"\n", b"\r\n"), False),
    ("big_endian", HEADER_OK.replace(b"little", b"big"), False),
    ("ascii", HEADER_OK.replace(b"binary_little_endian", b"ascii"), False),
    ("not_ply", b"plx\n" + HEADER_OK[4:], False),
    ("second_element", HEADER_OK.replace(b"end_header", b"element face 0\nend_header"), False),
    ("list_property", HEADER_OK.replace(b"end_header", b"property list uchar int vertex_indices\nend_header"), False),
    ("unknown_type", HEADER_OK.replace(b"float z", b"half z"), False),
    ("no_end_header", HEADER_OK.replace(b"end_header\n", b""), False),
    ("element_not_vertex", HEADER_OK.replace(b"element vertex", b"element point"), False),
])
def test_ply_header_accept_reject(splat, oracle, scratch, name, header, accept):
    """same accept/reject decisions as the reference reader (ply.cpp:140-254), in both implementations."""
    body = np.arange(2 * 64, dtype=np.float32).tobytes()
    path = _write(scratch / f"hdr_{name}.ply", header, body)
    gc, ok = _import(splat, path)
    assert ok == accept
    if accept:
        aos, _ = oracle.import_ply(path, True)
        np.testing.assert_array_equal(gc.GetRawData(), aos)
    else:
        with pytest.raises(IOError):
            oracle.import_ply(path, True)
        assert gc.LastError()


def test_missing_required_property_reads_zero(splat, oracle, scratch):
    """a missing property only logs; the value reads as 0 (binaryattribute.h:57-65): alpha = sigmoid(0) = 0.5,
    scale = exp(0) = 1, rot = 0 -> glm::normalize(quat) yields identity."""
    path = _write(scratch / "xyz_only.ply", HEADER_OK, np.array([1, 2, 3, 4, 5, 6], np.float32).tobytes())
    gc, ok = _import(splat, path)
    assert ok and gc.GetStride() == 100
    aos, _ = oracle.import_ply(path, True)
    np.testing.assert_array_equal(gc.GetRawData(), aos)
    np.testing.assert_array_equal(aos[:, 0:4], [[1, 2, 3, 0.5], [4, 5, 6, 0.5]])
    np.testing.assert_array_equal(aos[0, 16:25], np.eye(3, dtype=np.float32).reshape(-1))


def test_missing_file_and_truncated_body(splat, oracle, scratch):
    gc, ok = _import(splat, str(scratch / "does_not_exist.ply"))
    assert not ok and gc.GetNumGaussians() == 0
    # short body: not an error in the reference (ply.cpp:83 does not check the read) -> trailing vertices are zero
    path = _write(scratch / "short.ply", HEADER_OK, np.array([1, 2, 3], np.float32).tobytes())
    gc, ok = _import(splat, path)
    assert ok and gc.GetNumGaussians() == 2
    aos, _ = oracle.import_ply(path, True)
    np.testing.assert_array_equal(gc.GetRawData(), aos)


def test_empty_cloud(splat, oracle, scratch):
    path = _write(scratch / "empty.ply", HEADER_OK.replace(b"vertex 2", b"vertex 0"))
    gc, ok = _import(splat, path)
    assert ok and gc.GetNumGaussians() == 0 and gc.GetTotalSize() == 0


@pytest.mark.parametrize("count", ["999999999999", "-5", "18446744073709551615", "4611686018427387904"])
def test_absurd_vertex_count_is_an_error_not_a_crash(splat, scratch, count):
    """The reference allocates whatever the header claims (`new uint8_t[size * count]`, ply.cpp:83) and would die on
    std::bad_alloc; behind a C ABI that must come back as a status code: no exception crosses `extern "C"`."""
    path = os.path.join(str(scratch), f"absurd_{abs(hash(count)) % 10**6}.ply")
    header = f"ply\nformat binary_little_endian 1.0\nelement vertex {count}\nproperty float x\nproperty float y\nproperty float z\nend_header\n"
    with open(path, "wb") as f:
        f.write(header.encode() + b"\0" * 64)
    gc = splat.GaussianCloud()
    assert not gc.ImportPly(path)
    assert splat.lib().spl_last_error()  # a message, and the process is still alive
    assert gc.ImportPly(os.path.join(os.path.dirname(__file__), "golden", "reference_test.ply"))  # the object stays usable


@pytest.mark.parametrize("count", ["4611686018427387904", "2147483648", "75489476129874"])
def test_huge_vertex_count_without_properties(splat, scratch, count):
    """with NO property lines the vertex size is 0, so the size guard in Ply::Parse did not apply and
    GaussianCloud::Allocate's count * stride wrapped (one float allocated, `count` records written).  Counts beyond what
    the renderer accepts (2^31 - 1) are rejected in the reader; Allocate's product is checked."""
    path = os.path.join(str(scratch), f"noprops_{count}.ply")
    with open(path, "wb") as f:
        f.write(f"ply\nformat binary_little_endian 1.0\nelement vertex {count}\nend_header\n".encode())
    gc = splat.GaussianCloud()
    assert not gc.ImportPly(path)
    assert b"out of range" in splat.lib().spl_last_error()
    assert gc.ImportPly(os.path.join(os.path.dirname(__file__), "golden", "reference_test.ply"))
