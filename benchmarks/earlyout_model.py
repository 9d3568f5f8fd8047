"""CPU model of the compositor's transmittance early-out at its three granularities (no GPU).

The kernel zeroes a pixel's T once it falls below earlyOutT (checked after every group of 32 list entries); a warp skips
a whole 256-entry chunk when all its 32 pixels are finished (checked at the chunk's start), and the tile stops when all
256 pixels are.  A warp that finishes in the middle of a chunk still walks the rest of that chunk's hits with T = 0.
This script counts, on a sample of tiles of the bench scene (oracle records):
    trips_now      hit-loop trips the kernel executes (warp alive at the chunk start => whole chunk walked)
    trips_group    trips if the warp also left the walk at the first GROUP boundary where all its pixels are finished
    trips_ideal    trips while at least one pixel of the warp is alive, at entry granularity
    trips_full     trips without any early-out

    python benchmarks/earlyout_model.py [--tiles 400] [--workload 1m_deg0] [--early-out 1e-4]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="1m_deg0")
    ap.add_argument("--tiles", type=int, default=400)
    ap.add_argument("--early-out", type=float, default=1e-4)
    args = ap.parse_args()
    import bench
    from oracle import splat_oracle as O
    from splatapult_b200 import scenes

    w = bench.WORKLOADS[args.workload]
    aos, full = O.import_ply(bench.scene_path(args.workload), True)
    W, H = bench.WIDTH, bench.HEIGHT
    fr = O.Frame(scenes.translate(0, 0, w["cam_z"]), scenes.perspective(scenes.FOVY, W / H), [0, 0, W, H], [scenes.Z_NEAR, scenes.Z_FAR])
    rec = O.project(aos, full, fr)
    order = O.sort(rec)[::-1]  # near -> far
    r = rec[order]
    px, py = r["px"].astype(np.float64), r["py"].astype(np.float64)
    A, B, C = r["i00"].astype(np.float64), 0.5 * (r["i01"].astype(np.float64) + r["i10"]), r["i11"].astype(np.float64)
    alpha = r["alpha"].astype(np.float64)
    r2 = 2.0 * np.log(np.maximum(256.0 * alpha, 1e-30))
    det = A * C - B * B
    ex, ey = np.sqrt(np.maximum(r2 * C / det, 0)), np.sqrt(np.maximum(r2 * A / det, 0))
    tilesX, tilesY = (W + 15) // 16, (H + 15) // 16
    rng = np.random.default_rng(1)
    sample = rng.choice(tilesX * tilesY, size=min(args.tiles, tilesX * tilesY), replace=False)
    tot = dict(now=0, group=0, ideal=0, full=0, entries=0, staged_now=0, staged_all=0)
    uu = np.arange(16) + 0.5
    for t in sample:
        tx, ty = int(t % tilesX), int(t // tilesX)
        x0, y0 = tx * 16, ty * 16
        # entries: footprint box meets the tile's pixel centres (near -> far order is the index order)
        sel = np.nonzero((r2 > 0) & (px + ex >= x0 + 0.5) & (px - ex <= x0 + 15.5) & (py + ey >= y0 + 0.5) & (py - ey <= y0 + 15.5))[0]
        E = len(sel)
        if E == 0:
            continue
        dx = (x0 + uu)[None, None, :] - px[sel, None, None]
        dy = (y0 + uu)[None, :, None] - py[sel, None, None]
        q = A[sel, None, None] * dx * dx + 2 * B[sel, None, None] * dx * dy + C[sel, None, None] * dy * dy
        wgt = alpha[sel, None, None] * np.exp(-0.5 * q)
        wgt = np.where(wgt > 1.0 / 256.0, wgt, 0.0)  # [E, y, x]
        # transmittance BEFORE entry k (exclusive cumulative product)
        T = np.cumprod(1.0 - wgt, axis=0)
        # a pixel is finished from the first group boundary (multiple of 32 entries) at which T < early_out
        G = (E + 31) // 32
        gend = np.minimum((np.arange(G) + 1) * 32, E) - 1  # last entry of each group
        done_at_group = T[gend] < args.early_out  # [G, y, x]
        # first finished group per pixel (G = never)
        first = np.where(done_at_group.any(axis=0), done_at_group.argmax(axis=0), G)
        # warps: 8x4 pixel blocks
        first_w = first.reshape(4, 4, 2, 8).transpose(0, 2, 1, 3).reshape(8, 32).max(axis=1)  # group after which the warp is finished
        hit_w = (wgt > 0).reshape(E, 4, 4, 2, 8).transpose(0, 1, 3, 2, 4).reshape(E, 8, 32).any(axis=2)  # [E, warp]
        # per-pixel exact finish entry for the ideal count
        pix_done_entry = np.where((T < args.early_out).any(axis=0), (T < args.early_out).argmax(axis=0) + 1, E)
        ideal_w = pix_done_entry.reshape(4, 4, 2, 8).transpose(0, 2, 1, 3).reshape(8, 32).max(axis=1)
        tile_fin_entry = min(E, (int(first_w.max()) + 1) * 32) if first_w.max() < G else E
        # finer: entry masks per 4x4 half (or 4x2 quarter) of the warp's block, and the walk skips entries that touch only
        # halves whose pixels are all finished (liveness refreshed at group boundaries)
        for name, bh, bw in (("half", 4, 4), ("quarter", 2, 4)):
            sy, sx = 4 // bh, 8 // bw
            # [E, warpRow, sy, bh, warpCol, sx, bw]
            hs = (wgt > 0).reshape(E, 4, sy, bh, 2, sx, bw).any(axis=(3, 6)).transpose(0, 1, 3, 2, 4).reshape(E, 8, sy * sx)
            fs = first.reshape(4, sy, bh, 2, sx, bw).max(axis=(2, 5)).transpose(0, 2, 1, 3).reshape(8, sy * sx)  # finish group per sub-block
            gidx = np.arange(E) // 32
            alive = gidx[:, None, None] <= fs[None, :, :]  # sub-block still has a live pixel while group g <= its finish group
            tot.setdefault("sub_" + name, 0)
            tot["sub_" + name] += int((hs & alive).any(axis=2).sum())
        for wi in range(8):
            fin_entry = min(E, (int(first_w[wi]) + 1) * 32) if first_w[wi] < G else E  # entries walked with a live pixel (group granularity)
            now_entry = min(E, ((fin_entry + 255) // 256) * 256)  # the chunk that holds the finishing group is walked to its end
            h = hit_w[:, wi]
            tot["now"] += int(h[:now_entry].sum())
            tot["group"] += int(h[:fin_entry].sum())
            tot["ideal"] += int(h[: int(ideal_w[wi])].sum())
            tot["full"] += int(h.sum())
        tot["entries"] += E
        tot["staged_now"] += min(E, ((tile_fin_entry + 255) // 256) * 256)
        tot["staged_all"] += E
    print(f"tiles sampled {len(sample)}  list entries {tot['entries']}  (early-out T < {args.early_out})")
    print(f"entries staged before the tile stops: {tot['staged_now']} of {tot['staged_all']} ({tot['staged_now'] / tot['staged_all'] * 100:.1f} %)")
    for k in ("full", "now", "group", "ideal", "sub_half", "sub_quarter"):
        print(f"trips_{k:6s} {tot[k]:10d}  {tot[k] / tot['now']:.3f} x now")


if __name__ == "__main__":
    main()
