"""Small frames through every kernel of the library, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python benchmarks/sanitize_frame.py
    compute-sanitizer --tool racecheck python benchmarks/sanitize_frame.py
    compute-sanitizer --tool synccheck python benchmarks/sanitize_frame.py

Test-size scenes (the tools slow kernels down 10-100x): all three frame-sort modes, both target formats, degree 0 / 3,
the shared-sort second eye, the presented 8-bit frame, the GPU PLY decode, and the stand-alone (u64, u32) sorter at sizes
around the tile boundaries (multi-tile look-back, generation-tagged status words).  Every mode's image must equal the
default mode's bit for bit, so a race that changes a result also fails the run; the exit code is non-zero on any mismatch.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import splatapult_b200 as sp  # noqa: E402
from splatapult_b200 import api, scenes  # noqa: E402


def main():
    d = tempfile.mkdtemp(prefix="sanitize_")
    w, h = 333, 211
    vp, nf = [0, 0, w, h], [sp.Z_NEAR, sp.Z_FAR]
    proj = sp.perspective(sp.FOVY, w / h)
    cam = sp.translate(0, 0, 6.0)
    eye2 = cam.copy()
    eye2.reshape(4, 4)[3, 0] += 0.064
    frames = 0
    for deg in (0, 3):
        ply = scenes.write_synthetic_ply(os.path.join(d, f"s{deg}.ply"), 9000, deg, seed=3)
        for fmt in (api.FB_FP32, api.FB_FP16):
            ref = None
            for mode in (api.SORT_LSD, api.SORT_LSD_FLAT, api.SORT_BINNED):
                r = sp.SplatRenderer(fb_format=fmt)
                r.InitFromPly(ply, True, False)  # GPU PLY decode + pack
                r.SetOption(api.OPT_SORT_MODE, mode)
                for _ in range(2):  # second frame: generation-tagged look-back words of the first are stale
                    r.Sort(cam, proj, vp, nf)
                    r.Render(cam, proj, vp, nf)
                fb = r.ReadFramebuffer()
                p8 = r.ReadPresented8()
                frames += 2
                assert p8.shape == (h, w, 4)
                if ref is None:
                    ref = fb
                elif not np.array_equal(ref, fb):
                    print(f"MISMATCH deg {deg} fmt {fmt} mode {mode}: max diff {np.abs(ref - fb).max()}")
                    return 1
                if mode == api.SORT_LSD:
                    r.SetOption(api.OPT_SHARED_SORT, 1)
                    r.Sort(cam, proj, vp, nf)
                    r.Render(cam, proj, vp, nf)
                    r.Render(eye2, proj, vp, nf)  # re-projection in the first eye's order
                    r.ReadFramebuffer()
                    frames += 2
                r.close()
    rng = np.random.default_rng(1)
    for n in (1, 4095, 4097, 12289, 70001):
        keys = rng.integers(0, 1 << 45, n, dtype=np.uint64)
        vals = np.arange(n, dtype=np.uint32)
        ko, vo = sp.sort_pairs_host(keys, vals, 0, 45)
        order = np.argsort(keys, kind="stable")
        if not (np.array_equal(ko, keys[order]) and np.array_equal(vo, vals[order])):
            print(f"SORT MISMATCH n={n}")
            return 1
    print(f"sanitize_frame ok: {frames} frames, 5 sorts")
    return 0


if __name__ == "__main__":
    sys.exit(main())
