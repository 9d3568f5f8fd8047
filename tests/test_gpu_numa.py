"""NUMA helpers of the end-to-end path (spl_host_alloc_near / spl_device_numa_node / spl_bind_thread_near): the
pinned destination lands on the GPU's node, is usable as the target of ReadFramebufferAsync, and the thread binding is a
subset of the CPUs the process was allowed before."""
import os

import numpy as np
import pytest

from splatapult_b200 import api, scenes

pytestmark = pytest.mark.gpu
NF = [scenes.Z_NEAR, scenes.Z_FAR]


def _node_of_page(addr: int) -> int | None:
    """the NUMA node that holds the page at `addr`, from /proc/self/numa_maps (None when the kernel does not say)"""
    try:
        with open("/proc/self/numa_maps") as f:
            best = None
            for line in f:
                parts = line.split()
                start = int(parts[0], 16)
                if start <= addr:
                    nodes = [p for p in parts if p.startswith("N") and "=" in p and p[1:p.index("=")].isdigit()]
                    best = (start, nodes)
            if best and best[1]:
                return int(max(best[1], key=lambda p: int(p.split("=")[1]))[1:].split("=")[0])
    except OSError:
        pass
    return None


def test_near_allocation_and_binding(splat):
    node = splat.device_numa_node(0)
    assert node >= -1
    allowed = os.sched_getaffinity(0)
    try:
        buf = splat.host_alloc(8 << 20, near_device=0)
        a = buf.array(np.uint8, (8 << 20,))
        a[:] = 7  # touch every page
        assert int(a.sum()) == 7 * (8 << 20)
        if node >= 0:
            where = _node_of_page(buf.ptr)
            assert where is None or where == node, (where, node)
        # the buffer is pinned: an asynchronous read-back into it works and equals the blocking read
        gc = splat.GaussianCloud()
        gc.InitDebugCloud()
        r = splat.SplatRenderer()
        assert r.Init(gc, False, False)
        w, h = 320, 240
        cam, proj = scenes.translate(0.5, 0.5, 3), scenes.perspective(scenes.FOVY, w / h)
        r.Sort(cam, proj, [0, 0, w, h], NF)
        r.Render(cam, proj, [0, 0, w, h], NF)
        r.ReadFramebufferAsync(buf.ptr, w * h * 16)
        r.WaitCopies(0)
        np.testing.assert_array_equal(buf.array(np.float32, (-1,))[: w * h * 4].reshape(h, w, 4), r.ReadFramebuffer())
        bound = splat.bind_thread_near(0)
        assert bound == node or bound == -1
        now = os.sched_getaffinity(0)
        assert now and now <= allowed
    finally:
        os.sched_setaffinity(0, allowed)
