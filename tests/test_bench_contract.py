"""bench.py's CPU-checkable parts: the reference arm end to end on a tiny scene (JSON contract of the prompt), the
byte model, the per-rank cameras, and the loud failure of the product arm without a GPU."""
import json
import sys

import numpy as np
import pytest

import bench
from splatapult_b200 import api

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "e2e", "gpu_launches", "cpu_baseline")


def _tiny(monkeypatch, tmp_path):
    monkeypatch.setitem(bench.WORKLOADS, "tiny", dict(n=3000, deg=0, seed=5, half=4.0, cam_z=10.0, fp16=False, cfg="test scene"))
    monkeypatch.setattr(bench, "SCENE_DIR", str(tmp_path))


def test_reference_arm_prints_one_contract_line(monkeypatch, capsys, tmp_path):
    _tiny(monkeypatch, tmp_path)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1"])
    bench.main()
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "Msplats/sec at 1920x1080" and d["unit"] == "Msplats/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["steps"] == 2 and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "3000 of 3000" in cb["sample"]
    assert cb["omp"]["OMP_PROC_BIND"] == "spread" and set(cb["stage_ms"]) == {"project", "sort", "composite"}
    assert d["config"]["workload"] == bench.workload_string("tiny")
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_print_nothing(monkeypatch, capsys, tmp_path):
    _tiny(monkeypatch, tmp_path)
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1", "--gpus", "2"])
    bench.main()
    assert capsys.readouterr().out.strip() == ""


def test_byte_model():
    n, v, d, tiles, px = 1_000_000, 998_167, 4_221_697, 8160, 1920 * 1080
    ab = bench.algorithmic_bytes(n, v, d, 0, 6, tiles, px, False, False, False)
    assert ab["composite"] == 4 * d + 48 * v + 8 * tiles + 16 * px
    assert ab["sort_pass"] == 24 * d and ab["tile_ranges"] == 8 * tiles
    assert ab["frame"] == ab["project"] + ab["sort_histogram"] + 6 * ab["sort_pass"] + ab["tile_ranges"] + ab["composite"]
    ab3 = bench.algorithmic_bytes(n, v, d, 3, 6, tiles, px, True, False, False)
    assert ab3["project"] - ab["project"] == (232 - 52) * v and ab["composite"] - ab3["composite"] == 8 * px
    binned = bench.algorithmic_bytes(n, v, d, 0, 0, tiles, px, False, True, False)
    assert set(binned) == {"project", "bin_scatter", "segment_sort", "composite", "frame"}
    # default path: depth digits sorted on the N gaussians (8-byte pairs), tile digits on the D instances (8-byte instances)
    sp = bench.algorithmic_bytes(n, v, d, 0, 6, tiles, px, False, False, True)
    assert sp["sort_pass"] == 16 * d and sp["depth_sort_emit"] == 4 * 16 * n + 12 * n + 8 * d
    assert sp["frame"] == sp["project"] + sp["depth_sort_emit"] + 2 * sp["sort_pass"] + sp["tile_ranges"] + sp["composite"]


def test_both_arms_name_the_workload_identically():
    """the driver compares config.workload of the two arms byte for byte (same_config was false)"""
    import inspect

    src = inspect.getsource(bench)
    assert src.count('"workload": workload_string(') == 2
    s = bench.workload_string("1m_deg3")
    assert "configs[2]" in s and "SH degree 3" in s and "1920x1080" in s and "fp32" in s


def test_headline_is_the_full_sh_config():
    assert bench.HEADLINE == "1m_deg3" and bench.WORKLOADS[bench.HEADLINE]["deg"] == 3


def test_kernel_sources_sha_guards_traffic_json():
    """profiles/traffic.json is keyed by a hash of the kernel sources: a capture of other sources must not be reported"""
    import os

    sha = bench.kernel_sources_sha()
    assert len(sha) == 16
    path = os.path.join(bench.ROOT, "profiles", "traffic.json")
    prof = json.load(open(path))
    assert "captures" in prof
    for name, cap in prof["captures"].items():
        assert name in bench.WORKLOADS and "kernel_sources_sha" in cap and "dram_bytes_per_launch" in cap


def test_cameras_per_rank():
    c0 = bench.camera_for_rank("1m_deg0", 0, 1)
    assert c0.shape == (16,) and c0[14] == 10.0  # translate(0, 0, cam_z): column 3
    cams = [bench.camera_for_rank("1m_deg0", r, 8) for r in range(8)]
    assert len({tuple(np.round(c, 5)) for c in cams}) == 8
    for c in cams:  # all on the ring of radius cam_z around the origin
        assert abs(float(np.linalg.norm(c[12:15])) - 10.0) < 1e-4


def test_product_arm_needs_a_gpu(monkeypatch, tmp_path):
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    _tiny(monkeypatch, tmp_path)
    monkeypatch.delenv("RANK", raising=False)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "tiny", "--steps", "1", "--warmup", "1"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bench.main()
