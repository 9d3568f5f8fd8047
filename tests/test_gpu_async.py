"""End-to-end path: host matrices in, framebuffer out to pinned host memory on the copy stream while the next
frame renders (two alternating render targets)."""
import os

import numpy as np
import pytest

from splatapult_b200 import api, scenes

pytestmark = pytest.mark.gpu
NF = [scenes.Z_NEAR, scenes.Z_FAR]


def test_overlapped_readback_matches_blocking_reads(splat, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    gc = splat.GaussianCloud()
    gc.SetRawData(g["aos"], True)
    r = splat.SplatRenderer()
    assert r.Init(gc, False, False)
    vp = [0, 0, 320, 180]
    cams = [g["camera"], scenes.translate(0.7, -0.3, 9.0), scenes.translate(-1.0, 0.4, 11.0)]
    want = []
    for c in cams:  # blocking reference reads
        r.Sort(c, g["proj"], vp, NF)
        r.Render(c, g["proj"], vp, NF)
        want.append(r.ReadFramebuffer().copy())
    assert not np.array_equal(want[0], want[1])

    nbytes = 320 * 180 * 16
    pinned = [splat.host_alloc(nbytes) for _ in range(6)]
    for k in range(6):  # six frames back to back, no host wait in between
        c = cams[k % 3]
        r.Sort(c, g["proj"], vp, NF)
        r.Render(c, g["proj"], vp, NF)
        r.ReadFramebufferAsync(pinned[k].ptr, nbytes)
    r.FenceCopies()
    r.Synchronize()
    for k in range(6):
        got = pinned[k].array(np.float32, (180, 320, 4))
        np.testing.assert_array_equal(got, want[k % 3])


def test_wait_copies_streams_previous_frame_to_the_host(splat, golden_dir):
    """Consumer pattern of INTEGRATION.md: after frame k's async copy is enqueued, WaitCopies(1) guarantees frame
    k-1's pixels are in host memory (host-side wait, no stream synchronise)."""
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    gc = splat.GaussianCloud()
    gc.SetRawData(g["aos"], True)
    r = splat.SplatRenderer()
    assert r.Init(gc, False, False)
    vp = [0, 0, 320, 180]
    cams = [g["camera"], scenes.translate(0.7, -0.3, 9.0), scenes.translate(-1.0, 0.4, 11.0)]
    want = []
    for c in cams:
        r.Render(c, g["proj"], vp, NF)
        want.append(r.ReadFramebuffer().copy())
    nbytes = 320 * 180 * 16
    pinned = [splat.host_alloc(nbytes) for _ in range(2)]
    for p in pinned:
        p.array(np.float32, (180, 320, 4))[:] = -1.0
    for k in range(7):
        r.Sort(cams[k % 3], g["proj"], vp, NF)
        r.Render(cams[k % 3], g["proj"], vp, NF)
        r.ReadFramebufferAsync(pinned[k & 1].ptr, nbytes)
        r.WaitCopies(1)
        if k > 0:
            np.testing.assert_array_equal(pinned[(k - 1) & 1].array(np.float32, (180, 320, 4)), want[(k - 1) % 3])
    r.WaitCopies(0)
    np.testing.assert_array_equal(pinned[0].array(np.float32, (180, 320, 4)), want[0])
    with pytest.raises(api.SplatError):
        r.WaitCopies(2)


def test_fp16_async_and_resolution_change(splat, golden_dir):
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    gc = splat.GaussianCloud()
    gc.SetRawData(g["aos"], True)
    r = splat.SplatRenderer(fb_format=api.FB_FP16)
    assert r.Init(gc, False, False)
    a = splat.host_alloc(320 * 180 * 8)
    r.Render(g["camera"], g["proj"], [0, 0, 320, 180], NF)
    r.ReadFramebufferAsync(a.ptr, a.nbytes)
    b = splat.host_alloc(400 * 300 * 8)
    proj2 = scenes.perspective(scenes.FOVY, 400 / 300)
    r.Render(g["camera"], proj2, [0, 0, 400, 300], NF)  # reallocates both targets: must wait for the copy in flight
    r.ReadFramebufferAsync(b.ptr, b.nbytes)
    r.Synchronize()
    fa = a.array(np.float16, (180, 320, 4)).astype(np.float32)
    assert np.abs(fa[..., :3] - g["fb"][..., :3]).max() <= 2.0 ** -7
    fb = b.array(np.float16, (300, 400, 4)).astype(np.float32)
    assert fb[..., 3].min() == 1.0 and fb[..., :3].max() > 0.1
    with pytest.raises(splat.SplatError):
        r.ReadFramebufferAsync(a.ptr, 16)  # buffer too small


@pytest.mark.parametrize("fb_format", [api.FB_FP32, api.FB_FP16])
def test_presented_rgba8_is_the_desktop_blit(splat, golden_dir, fb_format):
    """ReadPresented8 = the reference's RenderDesktop blit (src/app.cpp:1076-1082, desktop_frag.glsl:32-37, color = 1)
    into an 8-bit window: clamp, x255, round -- bit-exact against the host conversion of the float target."""
    g = np.load(os.path.join(golden_dir, "synth_2k_deg3_320x180.npz"))
    gc = splat.GaussianCloud()
    gc.SetRawData(g["aos"], True)
    r = splat.SplatRenderer(fb_format=fb_format)
    assert r.Init(gc, False, False)
    vp = [0, 0, 320, 180]
    r.Render(g["camera"], g["proj"], vp, NF)
    fb = r.ReadFramebuffer()
    want = (np.clip(fb, 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    got = r.ReadPresented8()
    assert got.shape == (180, 320, 4) and got.dtype == np.uint8
    np.testing.assert_array_equal(got, want)
    assert got[..., 3].min() == 255 and got[..., :3].max() > 100

    # streamed: three frames, 8-bit copies overlapping the next frame
    cams = [g["camera"], scenes.translate(0.7, -0.3, 9.0), scenes.translate(-1.0, 0.4, 11.0)]
    wants, pinned = [], [splat.host_alloc(320 * 180 * 4) for _ in cams]
    for c in cams:
        r.Render(c, g["proj"], vp, NF)
        wants.append(r.ReadPresented8().copy())
    for k, c in enumerate(cams):
        r.Sort(c, g["proj"], vp, NF)
        r.Render(c, g["proj"], vp, NF)
        r.ReadPresented8Async(pinned[k].ptr, pinned[k].nbytes)
    r.WaitCopies(0)
    for k in range(3):
        np.testing.assert_array_equal(pinned[k].array(np.uint8, (180, 320, 4)), wants[k])
    with pytest.raises(api.SplatError):
        r.ReadPresented8Async(pinned[0].ptr, 16)


def test_overflow_on_the_async_path_is_reported_and_repaired(splat, oracle, scratch):
    """with D > capacity the entries that do not fit are the NEAREST splats, and on the asynchronous
    paths (Render + GetFramebuffer, ReadFramebufferAsync) nothing used to say so.  Now the compositor posts the frame's
    instance count into mapped host memory; the next host-side wait / Sort raises SPL_ERR_OVERFLOW as a notification,
    the capacity is grown before the next Sort, and the frames after that are complete."""
    a = scenes.synthetic_cloud_arrays(3000, 0, seed=5, half_extent=1.0)
    a["log_scale"] += 2.0
    path = scenes.write_ply_from_arrays(str(scratch / "async_overflow.ply"), a)
    gc = splat.GaussianCloud()
    assert gc.ImportPly(path)
    r = splat.SplatRenderer()
    r.SetOption(api.OPT_INSTANCES_PER_GAUSSIAN, 1.0)
    assert r.Init(gc, False, False)
    w, h = 640, 480
    cam, proj = scenes.translate(0, 0, 4), scenes.perspective(scenes.FOVY, w / h)
    vp, nf = [0, 0, w, h], [scenes.Z_NEAR, scenes.Z_FAR]
    pin = [splat.host_alloc(w * h * 16), splat.host_alloc(w * h * 16)]
    r.Sort(cam, proj, vp, nf)
    r.Render(cam, proj, vp, nf)
    # the notification comes from the first call that finds the frame's report: the enqueue of the copy if the (tiny) frame
    # has already finished, at the latest the wait for its pixels
    notes = []
    for call in (lambda: r.ReadFramebufferAsync(pin[0].ptr, w * h * 16), lambda: r.WaitCopies(0)):
        try:
            call()
        except splat.SplatError as e:
            notes.append(e)
    assert len(notes) == 1 and notes[0].code == -4 and "incomplete" in str(notes[0])
    r.WaitCopies(0)
    cap0 = r.GetStats().instance_capacity
    # the next frame runs with the raised capacity; no further notification
    r.Sort(cam, proj, vp, nf)
    r.Render(cam, proj, vp, nf)
    r.ReadFramebufferAsync(pin[1].ptr, w * h * 16)
    r.WaitCopies(0)
    st = r.GetStats()
    assert not st.overflow and st.instance_capacity > cap0 and st.instance_capacity >= st.num_instances
    aos, full = oracle.import_ply(path, True)
    want = oracle.render(aos, full, cam, proj, vp, nf)
    got = pin[1].array(np.float32, (h, w, 4))
    d = np.abs(got[..., :3] - want[..., :3])
    assert d.max() <= 5e-3 and d.mean() <= 1e-4
    bad = pin[0].array(np.float32, (h, w, 4))
    assert np.abs(bad[..., :3] - want[..., :3]).max() > 5e-3  # the reported frame really was incomplete
    # a Sort right after an unreported overflow carries the notification too (Render + GetFramebuffer users)
    r2 = splat.SplatRenderer()
    r2.SetOption(api.OPT_INSTANCES_PER_GAUSSIAN, 1.0)
    assert r2.Init(gc, False, False)
    r2.Sort(cam, proj, vp, nf)
    r2.Render(cam, proj, vp, nf)
    r2.Framebuffer()
    import time

    time.sleep(0.2)  # the frame finishes on its own; nobody synchronises
    with pytest.raises(splat.SplatError) as e2:
        r2.Sort(cam, proj, vp, nf)
    assert e2.value.code == -4
    r2.Render(cam, proj, vp, nf)  # that Sort WAS carried out, with the capacity raised
    fb = r2.ReadFramebuffer()
    assert np.abs(fb[..., :3] - want[..., :3]).max() <= 5e-3
