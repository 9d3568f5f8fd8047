"""CPU model (oracle records, bench scene): which share of the (gaussian, tile) instances of the conservative tile rectangle
holds no contributing pixel at all -- what an exact ellipse-tile test at emission would remove.  32x16 tiles: 7.6 %
(profiles/r2_cull_model.txt; 15.6 % with round 1's 16x16 tiles), against ~6 us more projection: not built.

    python benchmarks/cull_model.py [1m_deg0]         (a 10 % sample of the visible gaussians)
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import splat_oracle as O
from splatapult_b200 import scenes
name = sys.argv[1] if len(sys.argv) > 1 else "1m_deg0"
TW, TH = 32, 16
w = bench.WORKLOADS[name]
aos, full = O.import_ply(bench.scene_path(name), True)
W, H = 1920, 1080
fr = O.Frame(scenes.translate(0, 0, w["cam_z"]), scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H], [scenes.Z_NEAR, scenes.Z_FAR])
rec = O.project(aos, full, fr)
r = rec[rec["visible"] == 1]
rng = np.random.default_rng(0)
r = r[rng.random(len(r)) < 0.1]
px, py = r["px"].astype(np.float64), r["py"].astype(np.float64)
A, B, C = r["i00"].astype(np.float64), 0.5 * (r["i01"].astype(np.float64) + r["i10"]), r["i11"].astype(np.float64)
alpha = r["alpha"].astype(np.float64)
r2 = 2.0 * np.log(np.maximum(256.0 * alpha, 1e-30))
det = A * C - B * B
ex = np.sqrt(np.maximum(r2 * C / det, 0)) * 1.001 + 0.01; ey = np.sqrt(np.maximum(r2 * A / det, 0)) * 1.001 + 0.01
ok = r2 > 0
fx0 = np.maximum(np.ceil(px - ex - 0.5), 0); fx1 = np.minimum(np.floor(px + ex - 0.5), W - 1)
fy0 = np.maximum(np.ceil(py - ey - 0.5), 0); fy1 = np.minimum(np.floor(py + ey - 0.5), H - 1)
on = ok & (fx1 >= fx0) & (fy1 >= fy0)
tx0 = (fx0 // TW).astype(int); tx1 = (fx1 // TW).astype(int); ty0 = (fy0 // TH).astype(int); ty1 = (fy1 // TH).astype(int)
tot_pairs = 0; live_pairs = 0; small = 0; hist = {}
for i in np.nonzero(on)[0]:
    wt, ht = tx1[i] - tx0[i] + 1, ty1[i] - ty0[i] + 1
    tot_pairs += wt * ht
    hist[(wt, ht)] = hist.get((wt, ht), 0) + 1
    if wt <= 4 and ht <= 4: small += 1
    xs = np.arange(tx0[i] * TW, (tx1[i] + 1) * TW) + 0.5 - px[i]
    ys = np.arange(ty0[i] * TH, (ty1[i] + 1) * TH) + 0.5 - py[i]
    q = A[i] * xs[None, :] ** 2 + 2 * B[i] * xs[None, :] * ys[:, None] + C[i] * ys[:, None] ** 2
    hit = q < r2[i]
    # clip to screen
    hit = hit.reshape(ht, TH, wt, TW).any(axis=(1, 3))
    live_pairs += int(hit.sum())
print(name, "sampled gaussians", int(on.sum()), "rect pairs", tot_pairs, "live pairs", live_pairs, f"dead {100*(1-live_pairs/tot_pairs):.1f}%", "rects<=4x4", small / on.sum())
top = sorted(hist.items(), key=lambda kv: -kv[1])[:10]
print("rect shapes (w,h):count", top)
