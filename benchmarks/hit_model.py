"""CPU model of the compositor's hit-loop trip counts for different sub-block shapes (no GPU).

The compositor walks, per warp and per group of 32 list entries, the entries whose mask bit is set; the trip count of
that walk is what the kernel's issue time scales with.  If the lanes of a warp are split into independent sub-groups
(each owning a smaller pixel block and walking its OWN entries), the trip count per group becomes
max over sub-groups(count) instead of |union|.  This script measures both on the bench scene from oracle records:

    python benchmarks/hit_model.py [--n 1000000] [--workload 1m_deg0]

Output: total (entry, warp) hit iterations for  1 x (8x4),  2 x (4x4),  4 x (4x2),  4 x (2x4),  8 x (2x2)  lane splits, and the
share of the walk's lanes that hold a contributing pixel.  The model has no transmittance early-out (the kernel stops a
pixel at T < 1e-4, which removes about half of these trips on the dense bench scene); the ratios are what matters.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="1m_deg0")
    ap.add_argument("--n", type=int, default=0, help="use only the first n gaussians (0 = all)")
    args = ap.parse_args()
    import bench
    from oracle import splat_oracle as O
    from splatapult_b200 import scenes

    w = bench.WORKLOADS[args.workload]
    aos, full = O.import_ply(bench.scene_path(args.workload), True)
    if args.n:
        aos = aos[: args.n]
    W, H = bench.WIDTH, bench.HEIGHT
    fr = O.Frame(scenes.translate(0, 0, w["cam_z"]), scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H], [scenes.Z_NEAR, scenes.Z_FAR])
    rec = O.project(aos, full, fr)
    order = O.sort(rec)[::-1]  # near -> far
    r = rec[order]
    px, py = r["px"].astype(np.float64), r["py"].astype(np.float64)
    A, B, C = r["i00"].astype(np.float64), 0.5 * (r["i01"].astype(np.float64) + r["i10"]), r["i11"].astype(np.float64)
    alpha = r["alpha"].astype(np.float64)
    r2 = 2.0 * np.log(np.maximum(256.0 * alpha, 1e-30))
    ok = r2 > 0
    det = A * C - B * B
    ex = np.sqrt(np.maximum(r2 * C / det, 0))
    ey = np.sqrt(np.maximum(r2 * A / det, 0))
    # tiles whose pixel-centre range meets the footprint box
    tx0 = np.clip(np.floor((px - ex - 0.5) / 16), 0, (W + 15) // 16 - 1).astype(np.int64)
    tx1 = np.clip(np.floor((px + ex - 0.5) / 16), 0, (W + 15) // 16 - 1).astype(np.int64)
    ty0 = np.clip(np.floor((py - ey - 0.5) / 16), 0, (H + 15) // 16 - 1).astype(np.int64)
    ty1 = np.clip(np.floor((py + ey - 0.5) / 16), 0, (H + 15) // 16 - 1).astype(np.int64)
    on = ok & (px + ex >= 0) & (px - ex <= W) & (py + ey >= 0) & (py - ey <= H)
    cnt = np.where(on, (tx1 - tx0 + 1) * (ty1 - ty0 + 1), 0)
    D = int(cnt.sum())
    print(f"visible {len(order)}  tile instances D = {D}  ({D / max(len(aos), 1):.2f} per gaussian)")
    # expand (splat rank k, tile)
    k = np.repeat(np.arange(len(order)), cnt)
    first = np.cumsum(cnt) - cnt
    j = np.arange(D) - np.repeat(first, cnt)
    wd = (tx1 - tx0 + 1)[k]
    tx = tx0[k] + j % wd
    ty = ty0[k] + j // wd
    tilesX = (W + 15) // 16
    tile = ty * tilesX + tx
    # stable sort by tile keeps near->far inside a tile
    perm = np.argsort(tile, kind="stable")
    k, tx, ty, tile = k[perm], tx[perm], ty[perm], tile[perm]
    # pixel test per 2x2 block (64 blocks of a tile): any pixel centre with Q < r2
    t0 = time.time()
    bits = np.zeros((D, 8, 8), bool)  # [pair, by2, bx2]
    pixel_hits = 0
    uu = np.arange(16) + 0.5
    CH = 200_000
    for s in range(0, D, CH):
        e = min(s + CH, D)
        kk = k[s:e]
        dx = (tx[s:e, None] * 16 + uu[None, :]) - px[kk, None]  # [p, 16]
        dy = (ty[s:e, None] * 16 + uu[None, :]) - py[kk, None]
        q = (A[kk, None, None] * dx[:, None, :] ** 2 + 2 * B[kk, None, None] * dx[:, None, :] * dy[:, :, None]
             + C[kk, None, None] * dy[:, :, None] ** 2)  # [p, y, x]
        hit = q < r2[kk, None, None]
        bits[s:e] = hit.reshape(e - s, 8, 2, 8, 2).any(axis=(2, 4))
        pixel_hits += int(hit.sum())
    print(f"pixel tests: {time.time() - t0:.1f} s; pairs with no pixel at all: {(~bits.any(axis=(1, 2))).mean() * 100:.1f} %")
    # groups of 32 entries inside each tile's list (chunks of 256 start at the list head, so groups are aligned to it)
    start = np.r_[True, tile[1:] != tile[:-1]]
    idx = np.arange(D)
    head = np.maximum.accumulate(np.where(start, idx, 0))
    pos = idx - head
    gid = np.cumsum(pos % 32 == 0) - 1
    G = int(gid[-1]) + 1

    def trips(bh, bw):
        """lanes split into sub-groups that each own a (bw x bh)-pixel block; a warp owns 8x4 pixels."""
        # block grid of the tile at this granularity, from the 2x2 base blocks
        ny, nx = 16 // bh, 16 // bw
        b = bits.reshape(D, ny, bh // 2, nx, bw // 2).any(axis=(2, 4))  # [pair, ny, nx]
        counts = np.zeros((G, ny, nx), np.int32)
        for yy in range(ny):
            for xx in range(nx):
                counts[:, yy, xx] = np.bincount(gid, weights=b[:, yy, xx], minlength=G)
        # warps: 8x4 pixel blocks -> (4 // bh) x (8 // bw) sub-groups each
        sy, sx = 4 // bh, 8 // bw
        cw = counts.reshape(G, 4, sy, 2, sx).transpose(0, 1, 3, 2, 4).reshape(G, 8, sy * sx)
        return int(cw.max(axis=2).sum()), int(b.sum())

    base, _ = trips(4, 8)
    print(f"contributing (pixel, entry) pairs: {pixel_hits}  = {pixel_hits / (base * 32) * 100:.1f} % of the lanes of the 8x4 walk")
    print(f"{'split':>12} {'trips':>12} {'vs 8x4':>8} {'lane-hits':>12}")
    for name, bh, bw in (("1 x 8x4", 4, 8), ("2 x 4x4", 4, 4), ("2 x 8x2", 2, 8), ("4 x 4x2", 2, 4), ("4 x 2x4", 4, 2), ("8 x 2x2", 2, 2)):
        t, lanehits = trips(bh, bw)
        print(f"{name:>12} {t:12d} {t / base:8.3f} {lanehits:12d}")


if __name__ == "__main__":
    main()
