"""CamerasConfig: INRIA cameras.json -> the cameraMat values SplatRenderer::Sort/Render take.
Checked against a numpy restatement of src/camerasconfig.cpp:31-56; tests/test_ref_pin.py additionally holds the class
bit for bit to the reference's own camerasconfig.cpp (compiled into oracle/_ref against a stand-in for nlohmann-json)."""
import json
import math

import numpy as np
import pytest

from splatapult_b200 import scenes


def _restatement(entries):
    """camerasconfig.cpp:31-56: mat columns = (R[:,0], -R[:,1], -R[:,2], position); fov = 2 atan(size / (2 fx))."""
    out = []
    for o in entries:
        R = np.asarray(o["rotation"], np.float32)  # rows as in the JSON
        pos = np.asarray(o["position"], np.float32)
        m = np.eye(4, dtype=np.float32)  # m[col, row]
        m[0, :3], m[1, :3], m[2, :3], m[3, :3] = R[:, 0], -R[:, 1], -R[:, 2], pos
        fx = np.float32(o["fx"])
        fov = np.array([2 * math.atan(np.float32(o["width"]) / (2 * fx)), 2 * math.atan(np.float32(o["height"]) / (2 * fx))], np.float32)
        out.append((m.reshape(-1), fov))
    return out


def test_round_trip_of_ring_cameras(splat, scratch):
    cams = scenes.ring_cameras(8, 18.0, 3.0)
    path = scenes.write_cameras_json(str(scratch / "cameras.json"), cams, 1920, 1080, 1303.7)
    cfg = splat.CamerasConfig()
    assert cfg.ImportJson(path) and cfg.GetNumCameras() == 8
    want = _restatement(json.load(open(path)))
    for i, c in enumerate(cams):
        m, fov = cfg.GetCamera(i)
        np.testing.assert_array_equal(m, want[i][0])
        np.testing.assert_allclose(fov, want[i][1], rtol=1e-6)
        np.testing.assert_allclose(m, c, atol=1e-6)  # writer and reader are inverse conventions
    n, p = cfg.EstimateFloorPlane()
    ups = np.stack([c.reshape(4, 4)[1, :3] for c in cams]).mean(0)
    np.testing.assert_allclose(n, ups / np.linalg.norm(ups), atol=1e-5)
    np.testing.assert_allclose(p, n * np.mean([c.reshape(4, 4)[3, :3] @ n for c in cams]), atol=1e-4)


def test_inria_style_entry(splat, scratch):
    """an entry shaped like real INRIA output (extra keys, float/exponent formats, nested arrays)"""
    entry = {"id": 0, "img_name": "DSC_0001", "width": 1959, "height": 1090,
             "position": [-3.0089893469241797, -0.11086489695181866, -3.7527640949141428],
             "rotation": [[0.876134201218856, 0.06925962026449776, 0.47706599800804744],
                          [-0.04747421839895102, 0.9972110940209488, -0.057586739349882114],
                          [-0.4797239414934443, 0.027805376500959853, 0.8769787916452908]],
             "fy": 1164.6601287484507, "fx": 1.1593880237791562e3}
    path = str(scratch / "inria.json")
    with open(path, "w") as f:
        json.dump([entry], f, indent=2)
    cfg = splat.CamerasConfig()
    assert cfg.ImportJson(path) and cfg.GetNumCameras() == 1
    m, fov = cfg.GetCamera(0)
    (wm, wf), = _restatement([entry])
    np.testing.assert_array_equal(m, wm)
    np.testing.assert_allclose(fov, wf, rtol=1e-6)
    assert fov[0] == pytest.approx(2 * math.atan(1959 / (2 * 1159.3880237791562)), rel=1e-6)
    with pytest.raises(splat.SplatError):
        cfg.GetCamera(1)


@pytest.mark.parametrize("text", ["", "{", "[{\"id\": 0}]", "[{\"id\":0,\"position\":[1,2],\"rotation\":[[1,0,0],[0,1,0],[0,0,1]],\"width\":1,\"height\":1,\"fx\":1,\"fy\":1}]",
                                  "{\"id\": 0}", "[1, 2, 3]", "[{\"id\":0,\"position\":[1,2,3],\"rotation\":\"x\",\"width\":1,\"height\":1,\"fx\":1,\"fy\":1}] trailing"])
def test_malformed_json_returns_false(splat, scratch, text):
    """nlohmann would throw -> ImportJson returns false (camerasconfig.cpp:58-63); a missing file too (:23-26)."""
    path = str(scratch / "bad.json")
    with open(path, "w") as f:
        f.write(text)
    cfg = splat.CamerasConfig()
    assert not cfg.ImportJson(path) and cfg.GetNumCameras() == 0 and cfg.LastError()
    assert not cfg.ImportJson(str(scratch / "missing.json"))


def test_empty_array_and_floor_default(splat, scratch):
    path = str(scratch / "empty.json")
    with open(path, "w") as f:
        f.write("[]")
    cfg = splat.CamerasConfig()
    assert cfg.ImportJson(path) and cfg.GetNumCameras() == 0
    n, p = cfg.EstimateFloorPlane()
    np.testing.assert_array_equal(n, [0, 1, 0])
    np.testing.assert_array_equal(p, [0, 0, 0])


def test_deeply_nested_json_is_refused_not_a_stack_overflow(splat, scratch):
    """the parser recurses per nesting level; a file of 200 000 '[' must come back as `false`."""
    for name, text in (("arrays", "[" * 200_000), ("objects", '{"a":' * 100_000), ("mixed", '[{"a":' * 50_000)):
        p = scratch / f"deep_{name}.json"
        p.write_text(text)
        cfg = splat.CamerasConfig()
        assert not cfg.ImportJson(str(p))
        assert "nesting" in cfg.LastError()
