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


def test_entries_before_a_bad_one_stay(splat, scratch):
    """camerasconfig.cpp:55 emplaces per entry inside the try block: a type error in entry 2 returns false with entry 1 kept"""
    good = '{"id":0,"position":[1,2,3],"rotation":[[1,0,0],[0,1,0],[0,0,1]],"width":4,"height":3,"fx":2,"fy":2}'
    bad = '{"id":1,"position":[1,2,3],"rotation":[[1,0,0],[0,1,0],[0,0,1]],"width":4,"height":3,"fy":2}'
    path = str(scratch / "partial.json")
    with open(path, "w") as f:
        f.write(f"[{good},{bad},{good}]")
    cfg = splat.CamerasConfig()
    assert not cfg.ImportJson(path) and cfg.GetNumCameras() == 1 and "fx" in cfg.LastError()
    # a SYNTAX error anywhere leaves nothing: the document is parsed before the first entry is read
    with open(path, "w") as f:
        f.write(f"[{good},{good}")
    cfg = splat.CamerasConfig()
    assert not cfg.ImportJson(path) and cfg.GetNumCameras() == 0


@pytest.mark.parametrize("text,count", [("null", 0), ('﻿[]', 0),
                                        ('{"b":{"id":1,"position":[9,9,9],"rotation":[[1,0,0],[0,1,0],[0,0,1]],"width":4,"height":3,"fx":2,"fy":2},'
                                         '"a":{"id":0,"position":[1,2,3],"rotation":[[1,0,0],[0,1,0],[0,0,1]],"width":4,"height":3,"fx":true,"fy":2}}', 2)])
def test_documents_nlohmann_iterates(splat, scratch, text, count):
    """`for (auto&& o : data)`: null iterates nothing, an object iterates its values in key order; a byte-order mark is skipped;
    a boolean converts like a number (get<float>() of true is 1.0f)"""
    path = str(scratch / "iter.json")
    with open(path, "w", encoding="utf-8") as f:
        f.write(text)
    cfg = splat.CamerasConfig()
    assert cfg.ImportJson(path), cfg.LastError()
    assert cfg.GetNumCameras() == count
    if count == 2:
        m0, f0 = cfg.GetCamera(0)
        np.testing.assert_array_equal(m0[12:15], [1, 2, 3])  # key "a" first
        assert f0[0] == np.float32(2 * math.atan(np.float32(4) / np.float32(2.0)))  # fx = true -> 1.0f: 2 atan(4 / (2 * 1))
