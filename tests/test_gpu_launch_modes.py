"""How the frame's kernels are launched does not change a bit of the result: plain stream order, programmatic dependent
launch (SPL_OPT_PDL: a kernel may start while its predecessor drains; griddepcontrol.wait guards every dependent read) and
CUDA-graph replay of a Sort (SPL_OPT_GRAPH: capture + in-place update of the executable graph every frame)."""
import os

import numpy as np
import pytest

from splatapult_b200 import api, scenes

pytestmark = pytest.mark.gpu
NF = [scenes.Z_NEAR, scenes.Z_FAR]


def _frames(splat, ply, opts, cams, w, h, fmt=api.FB_FP32, sort_mode=api.SORT_LSD):
    r = splat.SplatRenderer(fb_format=fmt)
    r.InitFromPly(ply, True, False)
    r.SetOption(api.OPT_SORT_MODE, sort_mode)
    for k, v in opts.items():
        r.SetOption(k, v)
    proj = scenes.perspective(scenes.FOVY, w / h)
    out = []
    for cam in cams:
        for _ in range(2):  # the same view twice: replay / update of an unchanged graph, stale look-back words
            r.Sort(cam, proj, [0, 0, w, h], NF)
            r.Render(cam, proj, [0, 0, w, h], NF)
        out.append((r.ReadFramebuffer(), r.Tap(api.TAP_VALS).copy(), r.Tap(api.TAP_TILE_RANGES).copy()))
    st = r.GetStats()
    assert not st.overflow
    return out


@pytest.mark.parametrize("deg,n", [(0, 60000), (3, 20000)])
def test_pdl_and_graph_are_bit_identical(splat, scratch, deg, n):
    ply = scenes.write_synthetic_ply(os.path.join(str(scratch), f"modes_{deg}.ply"), n, deg, seed=21)
    w, h = 800, 450
    cams = [scenes.translate(0, 0, 10.0), scenes.look_at_camera((6.0, 2.0, 7.0), (0, 0, 0)), scenes.translate(0.5, -0.5, 4.0)]
    base = _frames(splat, ply, {api.OPT_PDL: 0, api.OPT_GRAPH: 0}, cams, w, h)
    assert base[0][0][..., :3].max() > 0
    for opts in ({api.OPT_PDL: 15, api.OPT_GRAPH: 0}, {api.OPT_PDL: 0, api.OPT_GRAPH: 1}, {api.OPT_PDL: 15, api.OPT_GRAPH: 1}, {api.OPT_PDL: 5, api.OPT_GRAPH: 1}):
        got = _frames(splat, ply, opts, cams, w, h)
        for (fb0, v0, r0), (fb1, v1, r1) in zip(base, got):
            np.testing.assert_array_equal(v0, v1, err_msg=str(opts))
            np.testing.assert_array_equal(r0, r1, err_msg=str(opts))
            np.testing.assert_array_equal(fb0, fb1, err_msg=str(opts))


def test_pdl_with_async_readback_and_other_paths(splat, scratch):
    """the overlapped read-back (event waits between the kernels), the fp16 target, the flat sort path and the shared sort"""
    ply = scenes.write_synthetic_ply(os.path.join(str(scratch), "modes_async.ply"), 30000, 0, seed=22)
    w, h = 640, 360
    cams = [scenes.translate(0, 0, 9.0), scenes.translate(0.3, 0.1, 8.0)]
    for fmt in (api.FB_FP32, api.FB_FP16):
        for mode in (api.SORT_LSD, api.SORT_LSD_FLAT):
            base = _frames(splat, ply, {api.OPT_PDL: 0, api.OPT_GRAPH: 0}, cams, w, h, fmt, mode)
            got = _frames(splat, ply, {api.OPT_PDL: 15, api.OPT_GRAPH: 1}, cams, w, h, fmt, mode)
            for a, b in zip(base, got):
                np.testing.assert_array_equal(a[0], b[0])
    proj = scenes.perspective(scenes.FOVY, w / h)
    ref = None
    for pdl in (0, 1):
        r = splat.SplatRenderer()
        r.InitFromPly(ply, True, False)
        r.SetOption(api.OPT_PDL, 15 * pdl)
        r.SetOption(api.OPT_GRAPH, pdl)
        r.SetOption(api.OPT_SHARED_SORT, 1)
        pin = [splat.host_alloc(w * h * 16), splat.host_alloc(w * h * 16)]
        for i in range(6):
            cam = cams[i & 1]
            r.Sort(cam, proj, [0, 0, w, h], NF)
            r.Render(cam, proj, [0, 0, w, h], NF)
            r.ReadFramebufferAsync(pin[i & 1].ptr, w * h * 16)
        r.Render(cams[0], proj, [0, 0, w, h], NF)  # shared sort: re-projection in the last Sort's order
        last = r.ReadFramebuffer()
        r.WaitCopies(0)
        frames = [p.array(np.float32, (h, w, 4)).copy() for p in pin] + [last]
        if ref is None:
            ref = frames
        else:
            for a, b in zip(ref, frames):
                np.testing.assert_array_equal(a, b)


def test_graph_follows_configuration_changes(splat, scratch):
    """one renderer, graph on, the frame's SHAPE changing under it: viewport sizes with other tile-bit counts (other pass
    counts), sort modes, sort tile sizes, shared sort.  Every frame equals a fresh stream-order renderer's frame: an
    executable graph of another node sequence is dropped and instantiated again, never replayed."""
    ply = scenes.write_synthetic_ply(os.path.join(str(scratch), "modes_shape.ply"), 25000, 0, seed=23)
    cam = scenes.translate(0, 0, 8.0)
    steps = [((320, 200), api.SORT_LSD, 0, 0), ((320, 200), api.SORT_LSD, 0, 0), ((1280, 720), api.SORT_LSD, 0, 0), ((64, 48), api.SORT_LSD, 0, 0),
             ((640, 360), api.SORT_LSD_FLAT, 0, 0), ((640, 360), api.SORT_BINNED, 0, 0), ((640, 360), api.SORT_LSD, 2048, 0),
             ((640, 360), api.SORT_LSD, 4096, 1), ((640, 360), api.SORT_LSD, 0, 0), ((1920, 1080), api.SORT_LSD, 0, 0)]
    live = splat.SplatRenderer()
    live.InitFromPly(ply, True, False)
    live.SetOption(api.OPT_GRAPH, 1)
    for (w, h), mode, tile, shared in steps:
        proj = scenes.perspective(scenes.FOVY, w / h)
        fresh = splat.SplatRenderer()
        fresh.InitFromPly(ply, True, False)
        fresh.SetOption(api.OPT_GRAPH, 0)
        fresh.SetOption(api.OPT_PDL, 0)
        for r in (live, fresh):
            r.SetOption(api.OPT_SORT_MODE, mode)
            r.SetOption(api.OPT_SORT_TILE, tile)
            r.SetOption(api.OPT_SHARED_SORT, shared)
        got = None
        for _ in range(3):  # direct frame, first capture, update
            live.Sort(cam, proj, [0, 0, w, h], NF)
            live.Render(cam, proj, [0, 0, w, h], NF)
            got = live.ReadFramebuffer()
        fresh.Sort(cam, proj, [0, 0, w, h], NF)
        fresh.Render(cam, proj, [0, 0, w, h], NF)
        np.testing.assert_array_equal(got, fresh.ReadFramebuffer(), err_msg=str(((w, h), mode, tile, shared)))
        fresh.close()
