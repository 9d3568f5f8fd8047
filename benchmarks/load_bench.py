"""Load-path timing: host PLY decode (GaussianCloud::ImportPly + Init) against the GPU decode
(SplatRenderer::InitFromPly and GaussianCloud::ImportPlyGpu), same file."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import splatapult_b200 as sp  # noqa: E402


def main():
    for wl in sys.argv[1:] or ["1m_deg0", "6m_deg3_fp16"]:
        path = bench.scene_path(wl)
        out = {"workload": wl, "ply_bytes": os.path.getsize(path)}
        open(path, "rb").read()  # page cache warm for every variant
        t = time.perf_counter(); gc = sp.GaussianCloud(); assert gc.ImportPly(path); out["import_ply_host_s"] = time.perf_counter() - t
        t = time.perf_counter(); r = sp.SplatRenderer(); assert r.Init(gc, False, False); r.Synchronize(); out["init_from_host_cloud_s"] = time.perf_counter() - t
        del r
        t = time.perf_counter(); g2 = sp.GaussianCloud(); assert g2.ImportPlyGpu(path); out["import_ply_gpu_s"] = time.perf_counter() - t
        out["gpu_decode_bit_identical"] = bool((g2.GetRawData().view("u4") == gc.GetRawData().view("u4")).all())
        del g2
        t = time.perf_counter(); r = sp.SplatRenderer(); r.InitFromPly(path, True, False); r.Synchronize(); out["init_from_ply_s"] = time.perf_counter() - t
        out["n"] = gc.GetNumGaussians()
        print(json.dumps(out))


if __name__ == "__main__":
    main()
