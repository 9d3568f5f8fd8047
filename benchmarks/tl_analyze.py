"""Per-tile phase timeline of k_sort_onesweep (diagnostic builds only: SPL_SORT_TIMELINE)."""
import sys, numpy as np
for path in sys.argv[1:]:
    raw = open(path, 'rb').read()
    passes, tiles, n, grid = np.frombuffer(raw[:16], np.uint32)
    a = np.frombuffer(raw[16:], np.uint64).reshape(passes, tiles, 8).astype(np.int64)
    print(f'== {path}: passes {passes} tiles {tiles} n {n} grid {grid}')
    for p in range(passes):
        t = a[p][:, [0, 1, 6, 2, 3, 4, 5]]
        t0 = t[:, 0].min()
        ph = np.diff(t, axis=1) / 1000.0  # us
        names = ['load', 'rank', 'nexthist', 'lookback', 'reorder', 'scatter']
        span = (t[:, -1].max() - t0) / 1000.0
        s = ' '.join(f'{nm} {ph[:, i].mean():.2f}/{np.percentile(ph[:, i], 95):.2f}' for i, nm in enumerate(names))
        print(f' pass {p}: span {span:.1f} us | mean/p95: {s} | tile {((t[:,-1]-t[:,0])/1000).mean():.2f}')
