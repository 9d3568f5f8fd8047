"""The device-wide (u64 key, u32 payload) radix sort against the oracle's stable LSD sort: bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pairs(n, bits, seed, dup_every=0):
    rng = np.random.default_rng(seed)
    keys = rng.integers(0, 1 << bits, n, dtype=np.uint64) if n else np.zeros(0, np.uint64)
    if dup_every and n:
        keys[::dup_every] = keys[0]
    return keys, np.arange(n, dtype=np.uint32)


def test_golden_vector(splat, golden_dir):
    g = np.load(os.path.join(golden_dir, "sort_keys_45bit.npz"))
    ko, vo = splat.sort_pairs_host(g["keys"], g["vals"], 0, 45)
    np.testing.assert_array_equal(ko, g["keys_sorted"])
    np.testing.assert_array_equal(vo, g["vals_sorted"])


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 4095, 4096, 4097, 8191, 8192, 12289, 100003, 1 << 20])
def test_sizes_45bit(splat, oracle, n):
    keys, vals = _pairs(n, 45, n + 1, dup_every=5)
    ko, vo = splat.sort_pairs_host(keys, vals, 0, 45)
    kw, vw = oracle.sort_pairs_u64(keys, vals, 0, 45)
    np.testing.assert_array_equal(ko, kw)
    np.testing.assert_array_equal(vo, vw)  # payload order on ties proves stability


@pytest.mark.parametrize("begin,end", [(0, 8), (0, 13), (0, 32), (0, 64), (8, 40), (3, 20), (32, 45), (5, 6)])
def test_bit_ranges(splat, oracle, begin, end):
    """sort on bits [begin,end) only: other bits must not influence the order (stable partial-key sort)."""
    keys, vals = _pairs(50000, 64, 17)
    ko, vo = splat.sort_pairs_host(keys, vals, begin, end)
    kw, vw = oracle.sort_pairs_u64(keys, vals, begin, end)
    np.testing.assert_array_equal(ko, kw)
    np.testing.assert_array_equal(vo, vw)


def test_degenerate_distributions(splat, oracle):
    n = 300000
    vals = np.arange(n, dtype=np.uint32)
    for keys in (np.zeros(n, np.uint64),                                   # one bin holds everything
                 np.full(n, (1 << 45) - 1, np.uint64),
                 np.arange(n, dtype=np.uint64),                            # already sorted
                 np.arange(n, dtype=np.uint64)[::-1].copy(),               # reversed
                 (np.arange(n, dtype=np.uint64) % 3) << np.uint64(40)):    # three huge runs in the top digit
        ko, vo = splat.sort_pairs_host(keys, vals, 0, 45)
        kw, vw = oracle.sort_pairs_u64(keys, vals, 0, 45)
        np.testing.assert_array_equal(ko, kw)
        np.testing.assert_array_equal(vo, vw)


def test_full_size_properties(splat):
    """BASELINE-scale stream (4.5 M keys, the size of config 2/3's instance list): sortedness, permutation
    (checksum of checksums), stability -- properties that need no oracle run."""
    n = 4_500_000
    rng = np.random.default_rng(7)
    tile = rng.integers(0, 8160, n, dtype=np.uint64)
    depth = rng.integers(0, 1 << 20, n, dtype=np.uint64) << np.uint64(6)  # few distinct depths -> many ties
    keys = (tile << np.uint64(32)) | depth
    vals = np.arange(n, dtype=np.uint32)
    ko, vo = splat.sort_pairs_host(keys, vals, 0, 45)
    assert np.all(ko[1:] >= ko[:-1])
    np.testing.assert_array_equal(keys[vo], ko)                      # payload follows its key
    assert int(vo.astype(np.uint64).sum()) == n * (n - 1) // 2       # permutation checksum
    ties = ko[1:] == ko[:-1]
    assert ties.any() and np.all(vo[1:][ties] > vo[:-1][ties])       # stable


def test_prefix_beyond_24_bits(splat):
    """Look-back words keep 24 count bits; a digit value with more than 2^24 keys in front of a tile takes the
    hi-word path.  20 M keys whose low byte is constant put every prefix of that digit beyond 2^24 from tile 4096 on."""
    n = 20_000_000
    rng = np.random.default_rng(11)
    keys = (rng.integers(0, 1 << 16, n, dtype=np.uint64) << np.uint64(8)) | np.uint64(0x5A)
    vals = np.arange(n, dtype=np.uint32)
    ko, vo = splat.sort_pairs_host(keys, vals, 0, 24)
    assert np.all(ko[1:] >= ko[:-1])
    np.testing.assert_array_equal(keys[vo], ko)
    ties = ko[1:] == ko[:-1]
    assert ties.any() and np.all(vo[1:][ties] > vo[:-1][ties])
    assert int(vo.astype(np.uint64).sum()) == n * (n - 1) // 2



def test_status_generation_wrap_over_many_frames(splat, golden_dir):
    """The renderer never clears the look-back words between frames; their 5-bit generation wraps every 31 frames
    (the status array is zeroed at the wrap).  70 frames over two cameras must keep giving the same two images."""
    import os

    from splatapult_b200 import scenes

    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    gc = splat.GaussianCloud()
    gc.SetRawData(g["aos"], True)
    r = splat.SplatRenderer()
    assert r.Init(gc, False, False)
    vp, nf = [0, 0, 320, 180], [scenes.Z_NEAR, scenes.Z_FAR]
    cams = [g["camera"], scenes.translate(0.7, -0.3, 9.0)]
    want = []
    for k in range(70):
        c = cams[k & 1]
        r.Sort(c, g["proj"], vp, nf)
        r.Render(c, g["proj"], vp, nf)
        fb = r.ReadFramebuffer()
        if k < 2:
            want.append(fb.copy())
        else:
            np.testing.assert_array_equal(fb, want[k & 1])
