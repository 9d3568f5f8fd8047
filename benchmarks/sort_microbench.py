#!/usr/bin/env python
"""Sort micro-benchmark (BASELINE.json "sort Gkeys/sec").

    python benchmarks/sort_microbench.py [--sizes 1,4,16,64] [--reps 20]

Stable (u64 key, u32 payload) sort on bits [0,45) -- the width of the 1080p (tile << 32 | depthKey) key.  Inputs:
  uniform   keys uniform in [0, 2^45), numpy default_rng(7)
  frame     the real key stream of workload 1m_deg0 (emitted by the projection kernel in gaussian order; taken with the
            renderer's tap after a flat-LSD Sort), tiled up to the requested size
Rows: this repo's onesweep sorter (spl_sorter_sort, includes its first-digit histogram kernel) and, as the same-box
library baseline, cub::DeviceRadixSort::SortPairs.  Timing: CUDA events around `reps` back-to-back sorts after 3
warm-ups; inputs are restored by a device-to-device copy outside the timed events... (the sort clobbers its input).
Prints one JSON line per (input, size).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def build_cub() -> C.CDLL:
    here = os.path.dirname(os.path.abspath(__file__))
    so, src = os.path.join(here, "libcub_baseline.so"), os.path.join(here, "cub_baseline.cu")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o", so, src], check=True)
    L = C.CDLL(so)
    L.cub_sort_temp_bytes.restype = C.c_size_t
    L.cub_sort_temp_bytes.argtypes = [C.c_uint64, C.c_int, C.c_int]
    L.cub_sort_pairs.restype = C.c_int
    L.cub_sort_pairs.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
    return L


def frame_keys():
    import splatapult_b200 as sp
    from splatapult_b200 import api
    import bench

    gc = sp.GaussianCloud()
    assert gc.ImportPly(bench.scene_path("1m_deg0"))
    r = sp.SplatRenderer()
    r.SetOption(api.OPT_SORT_MODE, api.SORT_LSD_FLAT)
    assert r.Init(gc, False, False)
    cam, proj = sp.translate(0, 0, 10.0), sp.perspective(sp.FOVY, 1920 / 1080)
    r.Sort(cam, proj, [0, 0, 1920, 1080], [sp.Z_NEAR, sp.Z_FAR])
    keys, vals = r.Tap(api.TAP_KEYS), r.Tap(api.TAP_VALS)
    # the tap returns the SORTED stream; undo the order with the payload to recover emission (gaussian) order
    order = np.argsort(vals, kind="stable")
    return keys[order].copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,4,16,64")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch

    import splatapult_b200 as sp

    L = sp.lib()
    cub = build_cub()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    real = None
    for kind in ("uniform", "frame"):
        for m in [int(s) for s in args.sizes.split(",")]:
            n = m * 1_000_000
            if kind == "uniform":
                keys = np.random.default_rng(7).integers(0, 1 << 45, n, dtype=np.uint64)
            else:
                if real is None:
                    real = frame_keys()
                keys = np.tile(real, n // len(real) + 1)[:n]
            pad = 4096
            k_src = torch.from_numpy(keys.view(np.int64)).to(dev)
            v_src = torch.arange(n, dtype=torch.int32, device=dev)
            k_in = torch.zeros(n + pad, dtype=torch.int64, device=dev)
            v_in = torch.zeros(n + pad, dtype=torch.int32, device=dev)
            k_out, v_out = torch.zeros_like(k_in), torch.zeros_like(v_in)
            sorter = L.spl_sorter_create(0, n)
            tmp_bytes = cub.cub_sort_temp_bytes(n, 0, 45)
            tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=dev)
            res = {}
            for name in ("splat_b200_onesweep", "cub_DeviceRadixSort"):
                times = []
                for rep in range(args.reps + 3):
                    k_in[:n].copy_(k_src)
                    v_in[:n].copy_(v_src)
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    if name.startswith("splat"):
                        rc = L.spl_sorter_sort(sorter, k_in.data_ptr(), v_in.data_ptr(), k_out.data_ptr(), v_out.data_ptr(), n, 0, 45, C.c_void_p(stream))
                    else:
                        rc = cub.cub_sort_pairs(tmp.data_ptr(), tmp_bytes, k_in.data_ptr(), k_out.data_ptr(), v_in.data_ptr(), v_out.data_ptr(), n, 0, 45, C.c_void_p(stream))
                    b.record()
                    assert rc == 0, rc
                    torch.cuda.synchronize()
                    if rep >= 3:
                        times.append(a.elapsed_time(b))
                if name.startswith("splat"):
                    mine = (k_out[:n].clone(), v_out[:n].clone())
                else:
                    assert torch.equal(mine[0], k_out[:n]) and torch.equal(mine[1], v_out[:n]), "sorters disagree"
                ms = float(np.median(times))
                res[name] = {"ms": ms, "gkeys_per_s": n / ms / 1e6, "gbs_ideal_24B": 24 * n / ms / 1e6}
            L.spl_sorter_destroy(C.c_void_p(sorter))
            print(json.dumps({"input": kind, "n": n, "bits": 45, "reps": args.reps, **res,
                              "speedup_vs_cub": res["cub_DeviceRadixSort"]["ms"] / res["splat_b200_onesweep"]["ms"]}), flush=True)


if __name__ == "__main__":
    main()
