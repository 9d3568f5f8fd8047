"""Regenerates the fixtures under tests/golden/.  Run in the authoring container (needs /root/reference):

    python tests/golden/make_golden.py

  reference_test.ply        byte copy of the reference's only data fixture, data/test.ply (16 gaussians; it is
                            GaussianCloud::InitDebugCloud exported, gaussiancloud.cpp:505-578) -- data, not source
  test_ply_decode.npy       the oracle's ImportPly of that file (16 x 61 float32)
  test_ply_640x480.npz      oracle records + back-to-front order + framebuffer for SURVEY 8(d) config 1
  synth_2k_deg3_320x180.npz oracle records/order/framebuffer for a 2000-gaussian seeded deg-3 cloud
  sort_keys_45bit.npz       20k (u64,u32) pairs + the oracle's stable sort of them on bits [0,45)
The GPU box has no /root/reference: tests read only these files.
"""
import math
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import splat_oracle as O  # noqa: E402
from splatapult_b200 import scenes  # noqa: E402


def main():
    ref = "/root/reference/data/test.ply"
    dst = os.path.join(HERE, "reference_test.ply")
    shutil.copyfile(ref, dst)
    aos, full = O.import_ply(dst, True)
    assert full and aos.shape == (16, 61)
    np.save(os.path.join(HERE, "test_ply_decode.npy"), aos)

    cam = scenes.test_ply_camera()
    proj = scenes.perspective(scenes.FOVY, 640 / 480)
    fb, rec, order, _ = O.render(aos, full, cam, proj, [0, 0, 640, 480], [scenes.Z_NEAR, scenes.Z_FAR], want_records=True)
    np.savez_compressed(os.path.join(HERE, "test_ply_640x480.npz"), camera=cam, proj=proj, records=rec, order=order, fb=fb)

    tmp = os.path.join(HERE, "_tmp_synth.ply")
    scenes.write_synthetic_ply(tmp, 2000, 3, seed=7)
    aos2, full2 = O.import_ply(tmp, True)
    os.remove(tmp)
    cam2 = scenes.translate(0, 0, 10)
    proj2 = scenes.perspective(scenes.FOVY, 320 / 180)
    fb2, rec2, order2, _ = O.render(aos2, full2, cam2, proj2, [0, 0, 320, 180], [scenes.Z_NEAR, scenes.Z_FAR], want_records=True)
    np.savez_compressed(os.path.join(HERE, "synth_2k_deg3_320x180.npz"), camera=cam2, proj=proj2, aos=aos2, records=rec2, order=order2, fb=fb2)

    rng = np.random.default_rng(7)
    keys = rng.integers(0, 1 << 45, 20000, dtype=np.uint64)
    keys[::7] = keys[3]  # plenty of duplicates: exercises stability
    vals = np.arange(20000, dtype=np.uint32)
    ko, vo = O.sort_pairs_u64(keys, vals, 0, 45)
    np.savez_compressed(os.path.join(HERE, "sort_keys_45bit.npz"), keys=keys, vals=vals, keys_sorted=ko, vals_sorted=vo)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
