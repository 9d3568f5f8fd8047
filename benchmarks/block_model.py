"""CPU model (oracle records, bench scene): how many (list entry, pixel block) pairs hold a contributing pixel, for different
block shapes of the compositor's warps.  Round 2 used it to choose 8x8 blocks with two pixels per thread: 0.61x the
(entry, block) hits of round 1's 8x4 blocks (profiles/r2_block_model.txt).

    python benchmarks/block_model.py [1m_deg0]        (a 25 % sample of the gaussians; no transmittance early-out)
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import splat_oracle as O
from splatapult_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "1m_deg0"
w = bench.WORKLOADS[name]
aos, full = O.import_ply(bench.scene_path(name), True)
W, H = 1920, 1080
fr = O.Frame(scenes.translate(0, 0, w["cam_z"]), scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H], [scenes.Z_NEAR, scenes.Z_FAR])
rec = O.project(aos, full, fr)
order = O.sort(rec)[::-1]
r = rec[order]
px, py = r["px"].astype(np.float64), r["py"].astype(np.float64)
A, B, C = r["i00"].astype(np.float64), 0.5 * (r["i01"].astype(np.float64) + r["i10"]), r["i11"].astype(np.float64)
alpha = r["alpha"].astype(np.float64)
r2 = 2.0 * np.log(np.maximum(256.0 * alpha, 1e-30))
ok = r2 > 0
det = A * C - B * B
ex = np.sqrt(np.maximum(r2 * C / det, 0)); ey = np.sqrt(np.maximum(r2 * A / det, 0))
tx0 = np.clip(np.floor((px - ex - 0.5) / 16), 0, 119).astype(np.int64); tx1 = np.clip(np.floor((px + ex - 0.5) / 16), 0, 119).astype(np.int64)
ty0 = np.clip(np.floor((py - ey - 0.5) / 16), 0, 67).astype(np.int64); ty1 = np.clip(np.floor((py + ey - 0.5) / 16), 0, 67).astype(np.int64)
on = ok & (px + ex >= 0) & (px - ex <= W) & (py + ey >= 0) & (py - ey <= H)
cnt = np.where(on, (tx1 - tx0 + 1) * (ty1 - ty0 + 1), 0)
D = int(cnt.sum())
# sample a subset of gaussians to keep it light
rng = np.random.default_rng(0)
sel = rng.random(len(cnt)) < 0.25
cnt_s = np.where(sel, cnt, 0)
Ds = int(cnt_s.sum())
k = np.repeat(np.arange(len(order)), cnt_s)
first = np.cumsum(cnt_s) - cnt_s
j = np.arange(Ds) - np.repeat(first, cnt_s)
wd = (tx1 - tx0 + 1)[k]
tx = tx0[k] + j % wd; ty = ty0[k] + j // wd
uu = np.arange(16) + 0.5
shapes = [(4, 8), (8, 8), (4, 16), (16, 4), (8, 16), (16, 16), (2,8),(4,4)]
tot = {s: 0 for s in shapes}
pix = 0
CH = 200000
for s0 in range(0, Ds, CH):
    e = min(s0 + CH, Ds); kk = k[s0:e]
    dx = (tx[s0:e, None] * 16 + uu[None, :]) - px[kk, None]
    dy = (ty[s0:e, None] * 16 + uu[None, :]) - py[kk, None]
    q = (A[kk, None, None] * dx[:, None, :] ** 2 + 2 * B[kk, None, None] * dx[:, None, :] * dy[:, :, None] + C[kk, None, None] * dy[:, :, None] ** 2)
    hit = q < r2[kk, None, None]
    pix += int(hit.sum())
    for (bh, bw) in shapes:
        b = hit.reshape(e - s0, 16 // bh, bh, 16 // bw, bw).any(axis=(2, 4))
        tot[(bh, bw)] += int(b.sum())
print(name, "D", D, "sampled", Ds, "pixel hits", pix)
base = tot[(4, 8)]
for (bh, bw) in shapes:
    t = tot[(bh, bw)]
    print(f"block {bw}x{bh} (w x h): block-hits {t}  per-entry {t/Ds:.3f}  vs 8x4 {t/base:.3f}  lane eff {pix/(t*bh*bw):.3f}")
