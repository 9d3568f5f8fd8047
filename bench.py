#!/usr/bin/env python
"""bench.py -- headline benchmark of the splat hot path (BASELINE.json: Msplats/s at 1920x1080).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one frame: SplatRenderer::Sort + SplatRenderer::Render of the whole cloud from one camera into a
cleared 1920x1080 float framebuffer.  Workloads:
    1m_deg3       BASELINE configs[2]: 1M synthetic gaussians, full SH (48 coeffs), fp32 target   (headline)
    1m_deg0       BASELINE configs[1]: 1M synthetic gaussians, SH degree 0, fp32 target           (info block)
    6m_deg3_fp16  BASELINE configs[3]: 6M gaussians, full SH, fp16 target                          (info block)
    config5       BASELINE configs[4]: the 6M cloud, 8 cameras from a cameras.json, one view per GPU (info block, N > 1)
ONE run reports all of them in ONE JSON line: the contract keys describe the headline workload, `workloads` holds the
others.  Every workload carries a `parity` block: the frame the end-to-end loop read back (and the projected records)
compared with the CPU oracle's frame of the same camera -- on every rank at N > 1.
`value` = gaussians in the cloud x frames x ranks / device time (inputs resident in HBM, L2 flushed between
frames); `e2e` = the same through the C ABI with host camera matrices in and the framebuffer copied to pinned
host memory every frame.  N > 1 (torchrun, one rank per GPU): views shard one per GPU, the cloud is replicated by ONE
ncclBroadcast inside libsplat_b200.so (SplatRenderer::InitBroadcast; torch.distributed only carries the 128-byte NCCL
id, the barriers and the max-over-ranks of the timings), no per-frame collective ("scaling": "weak").
--impl reference times the reference path's CPU stand-in (the oracle: the reference's GLSL cannot run here,
there is no OpenGL) on the host cores, rank 0 only, in a child process with pinned OpenMP threads.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    #  name            n          deg seed      half  cam_z  fb      baseline config
    "1m_deg0": dict(n=1_000_000, deg=0, seed=20240916, half=4.0, cam_z=10.0, fp16=False, cfg="configs[1]"),
    "1m_deg3": dict(n=1_000_000, deg=3, seed=20240916, half=4.0, cam_z=10.0, fp16=False, cfg="configs[2]"),
    "6m_deg3_fp16": dict(n=6_000_000, deg=3, seed=20240917, half=7.3, cam_z=18.0, fp16=True, cfg="configs[3]"),
}
HEADLINE = "1m_deg3"
WIDTH, HEIGHT = 1920, 1080
SCENE_DIR = os.environ.get("SPLAT_SCENE_DIR", "/tmp/splat_b200_scenes")
EARLY_OUT_T = 1e-4
# image bars of the CUDA compositor against the oracle
BARS_FP32 = {"max_abs": 5e-3, "p999": 1e-3, "mean": 1e-4}
BARS_FP32_EXACT = {"max_abs": 5e-3, "p999": 1e-3, "mean": 2e-5}
BAR_FP16_REL = 2.0 ** -7  # vs the oracle's round-after-every-blend half target (src/app.cpp:1010-1014), relative to max(1, |v|)


def workload_string(name: str) -> str:
    """config.workload, byte-identical in both arms (the driver compares them)."""
    w = WORKLOADS[name]
    return (f"{name}: {w['cfg']} of BASELINE.json -- {w['n']} synthetic gaussians (seed {w['seed']}), SH degree {w['deg']}, "
            f"{WIDTH}x{HEIGHT}, {'fp16' if w['fp16'] else 'fp32'} framebuffer, camera translate(0,0,{w['cam_z']:g})")


def scene_path(name: str) -> str:
    from splatapult_b200 import scenes

    w = WORKLOADS[name]
    os.makedirs(SCENE_DIR, exist_ok=True)
    path = os.path.join(SCENE_DIR, f"{name}_{w['n']}.ply")
    if not os.path.exists(path):
        scenes.write_synthetic_ply(path, w["n"], w["deg"], seed=w["seed"], half_extent=w["half"])
    return path


def cameras_json_path(name: str) -> str:
    """BASELINE configs[4]: 8 cameras on a ring, INRIA cameras.json schema (src/camerasconfig.cpp:31-56)."""
    from splatapult_b200 import scenes

    os.makedirs(SCENE_DIR, exist_ok=True)
    path = os.path.join(SCENE_DIR, f"cameras_ring8_{name}.json")
    if not os.path.exists(path):
        tmp = path + f".{os.getpid()}.tmp"
        scenes.write_cameras_json(tmp, scenes.ring_cameras(8, WORKLOADS[name]["cam_z"], 3.0), WIDTH, HEIGHT, 1303.7)
        os.replace(tmp, path)
    return path


def camera_for_rank(name: str, rank: int, world: int) -> np.ndarray:
    """rank 0 renders the BASELINE camera (translate(0,0,z)); other ranks their own view on a ring of the same radius."""
    from splatapult_b200 import scenes

    z = WORKLOADS[name]["cam_z"]
    if rank == 0:
        return scenes.translate(0, 0, z)
    ang = 2.0 * math.pi * rank / max(world, 1)
    return scenes.look_at_camera((z * math.sin(ang), 0.0, z * math.cos(ang)), (0.0, 0.0, 0.0))


def config5_camera(name: str, rank: int, world: int):
    """one view per GPU from the cameras.json, read by the C++ CamerasConfig (round-robin if there are fewer GPUs than views)"""
    import splatapult_b200 as sp
    from splatapult_b200 import sharding

    cfg = sp.CamerasConfig()
    if not cfg.ImportJson(cameras_json_path(name)):
        raise RuntimeError(cfg.LastError())
    mine = sharding.views_for_rank(cfg.GetNumCameras(), rank, world)
    view = mine[0] if mine else 0
    return cfg.GetCamera(view)[0], view, cfg.GetNumCameras()


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_sources_sha() -> str:
    """identifies the kernels an ncu capture in profiles/traffic.json belongs to: a capture of other sources is stale"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "splatapult_b200", "csrc")
    for name in ("composite.cu", "project.cu", "radix_sort.cu", "device_common.cuh"):
        with open(os.path.join(d, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.thread = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(n, v, d, deg_mode, passes, tiles, pixels, fp16, binned, split):
    """information bytes per stage (not padded strides), for the sort path actually run.
    rec = 16 (pos+alpha) + 24 (symmetric Sigma) + SH {12 | 48 | 192}."""
    rec = {0: 52, 1: 88, 3: 232}[deg_mode]
    fbb = pixels * (8 if fp16 else 16)
    if binned:
        b_p = 16 * n + (rec - 16) * v + 48 * v + 16 * n
        b_bin = 16 * n + 8 * d + 16 * tiles
        b_seg = 16 * d
        b_c = 8 * d + 48 * v + 8 * tiles + fbb
        return {"project": b_p, "bin_scatter": b_bin, "segment_sort": b_seg, "composite": b_c, "frame": b_p + b_bin + b_seg + b_c}
    b_p = 16 * n + (rec - 16) * v + 48 * v + 12 * d
    b_hist = 8 * d
    b_pass = 24 * d
    b_ranges = 8 * tiles  # written by the last sort pass's scatter (no separate pass over the keys)
    b_c = 4 * d + 48 * v + 8 * tiles + fbb
    ab = {"project": b_p, "sort_histogram": b_hist, "sort_pass": b_pass, "tile_ranges": b_ranges, "composite": b_c,
          "frame": b_p + b_hist + passes * b_pass + b_ranges + b_c}
    if split:
        # project_keys writes a u32 key + u32 id + 8 B rectangle per gaussian instead of 12 B per instance;
        # stage 1 = 4 depth-digit passes on N (u32, u32) pairs (16 N each) + expansion (4 N ids + 8 N rectangles in,
        # (u32 tile, u32 id) = 8 D out); the tile-digit passes move 8-byte instances (16 D per pass)
        d_passes = passes - 4
        ab["project"] = ab["project"] - 12 * d + 16 * n
        ab["depth_sort_emit"] = 4 * 16 * n + 12 * n + 8 * d
        ab["sort_pass"] = 16 * d
        ab["frame"] = ab["project"] + ab["depth_sort_emit"] + d_passes * ab["sort_pass"] + ab["tile_ranges"] + ab["composite"]
    return ab


# ----------------------------------------------------------------------------------------------------------------------
# parity against the CPU oracle (checker only: never on the timed path)
# ----------------------------------------------------------------------------------------------------------------------
_ORACLE_CLOUDS = {}


def oracle_cloud(name: str):
    from oracle import splat_oracle as O

    if name not in _ORACLE_CLOUDS:
        _ORACLE_CLOUDS.clear()  # one cloud at a time: the 6M AoS is 1.5 GB
        _ORACLE_CLOUDS[name] = O.import_ply(scene_path(name), True)
    return _ORACLE_CLOUDS[name]


def parity_block(name: str, cam, fb: np.ndarray, fb_exact: np.ndarray | None, rec: np.ndarray, threads: int) -> dict:
    """Compare this rank's frame (as read back by the e2e loop) and its projected records with the oracle's."""
    from oracle import splat_oracle as O
    from splatapult_b200 import scenes

    w = WORKLOADS[name]
    O.set_num_threads(max(1, threads))
    aos, full = oracle_cloud(name)
    proj = scenes.perspective(scenes.FOVY, WIDTH / HEIGHT)
    vp, nf = [0, 0, WIDTH, HEIGHT], [scenes.Z_NEAR, scenes.Z_FAR]
    t0 = time.perf_counter()
    want, wrec, order, _ = O.render(aos, full, cam, proj, vp, nf, fp16=False, want_records=True)
    fields = ("px", "py", "i00", "i01", "i10", "i11", "r", "g", "b", "alpha", "key", "visible")
    vis = wrec["visible"] != 0
    # records: key / visible / alpha of every gaussian, every field of the visible ones (the others are never read)
    rec_ok = bool(np.array_equal(rec["visible"], wrec["visible"]) and np.array_equal(rec["key"], wrec["key"]) and
                  all(np.array_equal(rec[f][vis].view(np.uint32), wrec[f][vis].view(np.uint32)) for f in fields))

    def stats(a, b):
        d = np.abs(a[..., :3].astype(np.float32) - b[..., :3])
        return {"max_abs": float(d.max()), "p999": float(np.quantile(d.reshape(-1), 0.999)), "mean": float(d.mean())}

    out = {"records_bit_exact": rec_ok, "visible": int(vis.sum()), "alpha_is_one": bool(np.all(fb[..., 3] == 1.0))}
    img = stats(fb, want)
    out.update(img)
    ok = rec_ok and out["alpha_is_one"]
    if w["fp16"]:
        # the product rounds to half once, the reference's RGBA16F target after every blend (src/app.cpp:1010-1014)
        want16 = O.composite(wrec, order, WIDTH, HEIGHT, fp16=True)
        rel = np.abs(fb[..., :3].astype(np.float32) - want16[..., :3]) / np.maximum(1.0, np.abs(want16[..., :3]))
        out["fp16_max_rel_vs_round_every_blend"] = float(rel.max())
        out["fp16_bar"] = BAR_FP16_REL
        ok = ok and out["fp16_max_rel_vs_round_every_blend"] <= BAR_FP16_REL and img["mean"] <= 1e-3
    else:
        ok = ok and all(img[k] <= BARS_FP32[k] for k in BARS_FP32)
        out["bars"] = BARS_FP32
    if fb_exact is not None:
        ex = stats(fb_exact, want)
        out["exact_mode"] = ex
        if not w["fp16"]:
            ok = ok and all(ex[k] <= BARS_FP32_EXACT[k] for k in BARS_FP32_EXACT)
    out["pass"] = bool(ok)
    out["oracle_s"] = time.perf_counter() - t0
    out["oracle_threads"] = O.num_threads()
    return out


# ----------------------------------------------------------------------------------------------------------------------
# one workload on this rank's GPU
# ----------------------------------------------------------------------------------------------------------------------
class Rig:
    """per-process state shared by the workloads of one run"""

    def __init__(self, args, rank, local_rank, world):
        import torch

        import splatapult_b200 as sp

        self.args, self.rank, self.local_rank, self.world = args, rank, local_rank, world
        self.torch, self.sp = torch, sp
        self.all_cpus = sorted(os.sched_getaffinity(0))
        # NUMA: run this rank on the CPUs of its GPU's node and keep the pinned framebuffer destination there
        self.numa_node = sp.bind_thread_near(local_rank)
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.comm = None
        if world > 1:
            import torch.distributed as dist

            uid = [sp.SplatComm.GetUniqueId() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)  # control plane: 128 bytes
            self.comm = sp.SplatComm(uid[0], rank, world, local_rank)
        self.vp, self.nf = [0, 0, WIDTH, HEIGHT], [sp.Z_NEAR, sp.Z_FAR]
        self.proj = sp.perspective(sp.FOVY, WIDTH / HEIGHT)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist

            dist.barrier()

    def max_over_ranks(self, values):
        if self.world == 1:
            return [float(v) for v in values]
        import torch.distributed as dist

        t = self.torch.tensor(list(values), device=self.dev, dtype=self.torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        import torch.distributed as dist

        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def load(self, name: str):
        """the same .ply path a splatapult user has; N > 1: ONE ncclBroadcast of the AoS inside the library"""
        from splatapult_b200 import api

        sp, w = self.sp, WORKLOADS[name]
        t0 = time.perf_counter()
        r = sp.SplatRenderer(device=self.local_rank, fb_format=api.FB_FP16 if w["fp16"] else api.FB_FP32, cuda_stream=self.stream.cuda_stream)
        broadcast_ms = None
        if self.world == 1:
            # PLY header parsed on the host, body decoded on the device (same records as GaussianCloud::ImportPly, tested)
            r.InitFromPly(scene_path(name), True, False)
        else:
            gc = None
            if self.rank == 0:
                gc = sp.GaussianCloud(import_full_sh=True, export_full_sh=True)
                if not gc.ImportPly(scene_path(name)):
                    raise RuntimeError(gc.LastError())
            broadcast_ms = r.InitBroadcast(self.comm, gc, root=0)
            del gc
        a = self.args
        r.SetOption(api.OPT_SORT_MODE, {"lsd": api.SORT_LSD, "lsd_flat": api.SORT_LSD_FLAT, "binned": api.SORT_BINNED}[a.sort_mode])
        r.SetOption(api.OPT_SORT_TILE, a.sort_tile)
        r.SetOption(api.OPT_SUBBLOCK_CULL, a.subblock_cull)
        if a.pdl >= 0:
            r.SetOption(api.OPT_PDL, a.pdl)
        if a.graph >= 0:
            r.SetOption(api.OPT_GRAPH, a.graph)
        return r, time.perf_counter() - t0, broadcast_ms

    def timed(self, fn, steps):
        """device ms per call of fn, L2 flushed before each call (same procedure as the headline loop)"""
        torch, stream = self.torch, self.stream
        for _ in range(3):
            fn()
        stream.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a_, b_ in evs:
            self.flush.zero_()
            a_.record(stream)
            fn()
            b_.record(stream)
        stream.synchronize()
        return float(sum(a_.elapsed_time(b_) for a_, b_ in evs)) / steps


def run_workload(rig: Rig, name: str, cam, steps: int, warmup: int, *, headline: bool, label: str | None = None) -> dict:
    """Load `name`, time K frames on the device and end to end, profile the stages, check parity.  Returns this rank's
    result with max-over-ranks timings (collective at N > 1: every rank must call it with the same arguments)."""
    from splatapult_b200 import api

    torch, sp, stream, world = rig.torch, rig.sp, rig.stream, rig.world
    w = WORKLOADS[name]
    r, load_s, broadcast_ms = rig.load(name)
    vp, nf, proj = rig.vp, rig.nf, rig.proj
    fb_bytes = WIDTH * HEIGHT * (8 if w["fp16"] else 16)
    fb_dtype = np.float16 if w["fp16"] else np.float32
    res = {"workload": label or name}

    def frame():
        r.Sort(cam, proj, vp, nf)
        r.Render(cam, proj, vp, nf)

    with torch.cuda.stream(stream):
        for _ in range(max(warmup, 3)):
            rig.flush.zero_()
            frame()
        stream.synchronize()
        st = r.GetStats()
        if st.overflow:  # grow once, outside the timed region
            r.Reserve(int(st.num_instances * 1.25))
            frame()
            stream.synchronize()
            st = r.GetStats()
            assert not st.overflow

        # ---- timed: K frames, device time per frame (CUDA events on the launching stream), L2 flushed in between
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        rig.barrier()
        # clocks / throttle reasons are sampled from here to the end of the e2e loop: the GPU is under the same load
        # throughout (the K timed frames alone can be shorter than nvidia-smi's start-up)
        clocks = ClockSampler(rig.local_rank) if headline else None
        if clocks:
            clocks.__enter__()
        wall0 = time.perf_counter()
        for a, b in ev:
            rig.flush.zero_()
            a.record(stream)
            frame()
            b.record(stream)
        stream.synchronize()
        wall = time.perf_counter() - wall0
        torch.cuda.synchronize()
        rig.barrier()
        total_ms = float(sum(a.elapsed_time(b) for a, b in ev))

        # ---- end to end through the C ABI: host matrices in, framebuffer out to pinned host memory, every frame
        pinned = [sp.host_alloc(fb_bytes, near_device=rig.local_rank), sp.host_alloc(fb_bytes, near_device=rig.local_rank)]
        for i in range(2):
            frame()
            r.ReadFramebufferAsync(pinned[i].ptr, fb_bytes)
        r.Synchronize()
        rig.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(steps):
            frame()  # host matrices -> kernels
            r.ReadFramebufferAsync(pinned[i & 1].ptr, fb_bytes)  # D->H on the copy stream, overlaps the next frame
        r.FenceCopies()  # the end event is recorded only after the last copy has landed
        b.record(stream)
        r.Synchronize()
        e2e_ms = a.elapsed_time(b)
        if clocks:  # keep the load up long enough for a handful of clock samples
            t_end = time.perf_counter() + 0.5
            while time.perf_counter() < t_end:
                frame()
                stream.synchronize()
            clocks.__exit__(None, None, None)
        # the frame the e2e loop delivered: kept for the parity check below
        fb_e2e = pinned[(steps - 1) & 1].array(fb_dtype, (HEIGHT, WIDTH, 4)).astype(np.float32)

        e2e8_ms = None
        if headline:
            # ---- informational: the same loop reading the PRESENTED 8-bit frame (the reference's RenderDesktop blit)
            p8 = [sp.host_alloc(WIDTH * HEIGHT * 4, near_device=rig.local_rank), sp.host_alloc(WIDTH * HEIGHT * 4, near_device=rig.local_rank)]
            for i in range(2):
                frame()
                r.ReadPresented8Async(p8[i].ptr, p8[i].nbytes)
            r.Synchronize()
            a8, b8 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a8.record(stream)
            for i in range(steps):
                frame()
                r.ReadPresented8Async(p8[i & 1].ptr, p8[i & 1].nbytes)
            r.FenceCopies()
            b8.record(stream)
            r.Synchronize()
            e2e8_ms = a8.elapsed_time(b8)
            del p8

        # ---- per-stage profile (separate, untimed pass): events between the kernels
        r.SetOption(api.OPT_PROFILE, 1)
        stages = np.zeros(8)
        reps = 10
        for _ in range(reps):
            rig.flush.zero_()
            frame()
            stream.synchronize()
            stages += np.array(r.GetStats().stage_ms)
        stages /= reps
        r.SetOption(api.OPT_PROFILE, 0)
        st = r.GetStats()
        rec = r.Tap(api.TAP_RECORDS)

        # ---- exact mode -- no transmittance early-out, the reference's full blend chain per pixel: timed + kept for parity
        r.SetOption(api.OPT_EARLY_OUT_T, 0.0)
        exact_ms = rig.timed(frame, steps)
        frame()
        fb_exact = r.ReadFramebuffer()
        r.SetOption(api.OPT_EARLY_OUT_T, EARLY_OUT_T)

        launch_ms = None
        if headline and rig.args.pdl < 0 and rig.args.graph < 0:
            # ---- informational: what the launch mechanics are worth (same kernels, same results): plain stream order, the
            # Sort replayed as one CUDA graph, and the default = graph + programmatic dependent launch of the sort passes
            launch_ms = {}
            for key, g, p in (("stream_order", 0, 0), ("graph", 1, 0), ("graph_pdl_default", 1, 5)):
                r.SetOption(api.OPT_GRAPH, g)
                r.SetOption(api.OPT_PDL, p)
                launch_ms[key] = rig.timed(frame, steps)
            launch_ms["saved_us_graph"] = (launch_ms["stream_order"] - launch_ms["graph"]) * 1e3
            launch_ms["saved_us_graph_pdl"] = (launch_ms["stream_order"] - launch_ms["graph_pdl_default"]) * 1e3

        stereo_ms = None
        if headline:
            # ---- informational: the VR call pattern (src/app.cpp:603-607): Sort for the left eye, Render for both.  Default:
            # the second Render sorts again (exact order for each eye); SPL_OPT_SHARED_SORT: the reference's behaviour, the
            # second eye is re-projected in the left eye's order (no depth passes).  6.4 cm apart, parallel axes.
            eye_l, eye_r = cam.copy(), cam.copy()
            right_axis = cam.reshape(4, 4)[0, :3]
            eye_l.reshape(4, 4)[3, :3] -= 0.032 * right_axis
            eye_r.reshape(4, 4)[3, :3] += 0.032 * right_axis

            def stereo_pair():
                r.Sort(eye_l, proj, vp, nf)
                r.Render(eye_l, proj, vp, nf)
                r.Render(eye_r, proj, vp, nf)

            stereo_ms = {}
            for key, opt in (("resort", 0), ("shared_sort", 1)):
                r.SetOption(api.OPT_SHARED_SORT, opt)
                stereo_ms[key] = rig.timed(stereo_pair, steps)
            r.SetOption(api.OPT_SHARED_SORT, 0)
        launches = int(st.kernel_launches)
    r.close()
    del r, pinned
    torch.cuda.empty_cache()

    # ---- parity of THIS rank's view against the oracle (untimed; rank-local OpenMP), worst over ranks
    par = None
    if not rig.args.no_parity:
        par = parity_block(name, cam, fb_e2e, fb_exact, rec, len(rig.all_cpus) // world)
        pars = rig.gather(par)
        worst = dict(par)
        for k in ("max_abs", "p999", "mean", "fp16_max_rel_vs_round_every_blend"):
            if k in par:
                worst[k] = max(p[k] for p in pars)
        worst["records_bit_exact"] = all(p["records_bit_exact"] for p in pars)
        worst["pass"] = all(p["pass"] for p in pars)
        worst["ranks_checked"] = len(pars)
        if world > 1:
            worst["per_rank_max_abs"] = [p["max_abs"] for p in pars]
        par = worst
    del rec, fb_e2e, fb_exact

    d2h_gbs = fb_bytes * steps / (e2e_ms * 1e-3) / 1e9
    per_rank_d2h = rig.gather(round(d2h_gbs, 2))
    total_ms, e2e_ms, exact_ms = rig.max_over_ranks([total_ms, e2e_ms, exact_ms])
    if e2e8_ms is not None:
        (e2e8_ms,) = rig.max_over_ranks([e2e8_ms])

    n, v, d = int(st.num_gaussians), int(st.num_visible), int(st.num_instances)
    tiles = int(st.tiles_x * st.tiles_y)
    binned = int(st.sort_mode) == api.SORT_BINNED
    split = int(st.sort_mode) == api.SORT_LSD
    ab = algorithmic_bytes(n, v, d, int(st.sh_mode), int(st.sort_passes), tiles, WIDTH * HEIGHT, w["fp16"], binned, split)
    if binned:
        stage = {"project": stages[0], "bin_scatter": stages[1], "segment_sort": stages[2], "composite": stages[4], "frame": stages[7]}
        cand = {"composite": (stages[4], ab["composite"], stages[4]), "project": (stages[0], ab["project"], stages[0]),
                "segment_sort": (stages[2], ab["segment_sort"], stages[2]), "bin_scatter": (stages[1], ab["bin_scatter"], stages[1])}
        names = ("project", "bin_scatter", "segment_sort", "composite")
        pass_ms = None
    else:
        d_passes = int(st.sort_passes) - 4 if split else int(st.sort_passes)
        if split:
            stage = {"project": stages[0], "depth_sort_emit": stages[1], "tile_sort_passes_total": stages[2], "composite": stages[4], "frame": stages[7]}
            names = ("project", "depth_sort_emit", "composite")
        else:
            stage = {"project": stages[0], "sort_passes_total": stages[2], "composite": stages[4], "frame": stages[7]}
            names = ("project", "composite")
        pass_ms = stages[2] / max(d_passes, 1)
        cand = {"composite": (stages[4], ab["composite"], stages[4]), "sort_onesweep": (stages[2], ab["sort_pass"], pass_ms),
                "project": (stages[0], ab["project"], stages[0])}
    stage_gbs = {k: (ab[k] / (stage[k] * 1e-3) / 1e9 if stage[k] > 0 else None) for k in names}
    if pass_ms:
        stage_gbs["sort_pass"] = ab["sort_pass"] / (pass_ms * 1e-3) / 1e9

    secs = total_ms * 1e-3
    res.update({
        "value": n * steps * world / secs / 1e6, "unit": "Msplats/s", "ms_per_step": total_ms / steps,
        "frames_per_s": steps * world / secs,
        "e2e": {"value": n * steps * world / (e2e_ms * 1e-3) / 1e6, "unit": "Msplats/s", "h2d_bytes_per_step": 38 * 4,
                "d2h_bytes_per_step": fb_bytes, "ms_per_step": e2e_ms / steps, "d2h_gbs_per_rank": per_rank_d2h},
        "exact_mode": {"ms_per_step": exact_ms, "value": n * world / (exact_ms * 1e-3) / 1e6, "unit": "Msplats/s",
                       "note": "early_out_T = 0: every list entry of every pixel is blended, as the reference does"},
        "parity": par,
        "visible": v, "tile_instances": d, "tile_px": f"{int(st.tile_w)}x{int(st.tile_h)}", "sort_bits": int(st.sort_bits), "sort_passes": int(st.sort_passes),
        "stage_ms": {k: float(x) for k, x in stage.items()}, "stage_gbs": stage_gbs,
        "sort_gkeys_per_s": (d / ((stages[1] + stages[2]) * 1e-3) / 1e9) if (stages[1] + stages[2]) > 0 else None,
        "load_s": load_s, "broadcast_ms": broadcast_ms, "fp16_target": w["fp16"],
    })
    res["_internal"] = {"stages": stages, "ab": ab, "cand": cand, "launches": launches, "wall": wall, "clocks": clocks.summary() if clocks else None,
                        "e2e8_ms": e2e8_ms, "stereo_ms": stereo_ms, "launch_ms": launch_ms, "n": n}
    return res


def roofline_block(name: str, res: dict, dev_props, clocks: dict) -> dict:
    """dominant kernel of the workload: algorithmic bytes per launch / live launch time, against the measured HBM peak"""
    it = res["_internal"]
    stages, ab, cand = it["stages"], it["ab"], it["cand"]
    peak, peak_src = measured_peak_gbs()
    dom = max(cand, key=lambda k: cand[k][0])
    dom_total, dom_bytes, dom_launch_ms = cand[dom]
    achieved = dom_bytes / (dom_launch_ms * 1e-3) / 1e9 if dom_launch_ms > 0 else 0.0
    traffic, issue, traffic_note = None, None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            prof = json.load(f)
        cap = prof.get("captures", {}).get(name, {})
        if cap and cap.get("kernel_sources_sha") == kernel_sources_sha():
            traffic = cap.get("dram_bytes_per_launch", {}).get(dom)
            wi = cap.get("warp_instructions_per_launch", {}).get(dom)
            clk = (clocks or {}).get("sm_mhz")
            # this path's dominant kernel is instruction-issue bound, not HBM bound: also report how close it is to that
            # bound -- warp instructions per launch (ncu smsp__inst_executed.sum) over the live launch time, per SMSP
            if wi and clk and dom_launch_ms > 0:
                smsps = dev_props.multi_processor_count * 4
                issue = {"warp_instructions_per_launch": wi, "ipc_per_smsp": wi / (dom_launch_ms * 1e-3 * clk * 1e6 * smsps), "peak": 1.0,
                         "source": f"ncu smsp__inst_executed.sum ({cap.get('source')}) / live launch time / (SMSPs x sampled SM clock)"}
            traffic_note = f"ncu --set full capture {cap.get('source')} of these kernel sources (sha {kernel_sources_sha()})"
        elif cap:
            traffic_note = "profiles/traffic.json holds a capture of OTHER kernel sources (stale): not reported"
        else:
            traffic_note = "no ncu capture of this workload in profiles/traffic.json"
    except Exception as e:  # noqa: BLE001
        traffic_note = f"profiles/traffic.json unreadable: {e}"
    return {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_note": traffic_note, "issue": issue, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
            "launch_ms": dom_launch_ms, "share_of_step": dom_total / stages[7] if stages[7] > 0 else None,
            "frame": {"algorithmic_bytes": ab["frame"], "achieved": ab["frame"] / (stages[7] * 1e-3) / 1e9 if stages[7] > 0 else None,
                      "frac": ab["frame"] / (stages[7] * 1e-3) / 1e9 / peak if stages[7] > 0 else None}}


def run_ours(args, rank, local_rank, world):
    import splatapult_b200 as sp

    if sp.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: splatapult_b200 has no CPU fallback")
    rig = Rig(args, rank, local_rank, world)
    head = args.workload
    steps, warmup = args.steps, max(args.warmup, 3)
    sort_desc = {"lsd": ": onesweep LSD, depth digits on N gaussians + tile digits on D instances", "lsd_flat": ": onesweep LSD, all digits on D instances",
                 "binned": ": MSD by tile id + per-tile smem sort"}[args.sort_mode]

    main = run_workload(rig, head, camera_for_rank(head, rank, world), steps, warmup, headline=True)
    it = main.pop("_internal")
    props = rig.torch.cuda.get_device_properties(rig.dev)
    main_roof = roofline_block(head, dict(main, _internal=it), props, it["clocks"])

    # ---- the other BASELINE configurations, same procedure, as info blocks of the same line
    others = {}
    if not args.headline_only:
        for name in [k for k in ("1m_deg0", "1m_deg3", "6m_deg3_fp16") if k != head]:
            blk = run_workload(rig, name, camera_for_rank(name, rank, world), steps, warmup, headline=False)
            bit = blk.pop("_internal")
            blk["roofline"] = roofline_block(name, dict(blk, _internal=bit), props, it["clocks"])
            blk["config"] = workload_string(name)
            others[name] = blk
        if world > 1:
            # BASELINE configs[4]: the 6M cloud, 8 cameras read from a cameras.json by CamerasConfig::ImportJson, one view per GPU
            name = "6m_deg3_fp16"
            cam5, view, ncams = config5_camera(name, rank, world)
            blk = run_workload(rig, name, cam5, steps, warmup, headline=False, label="config5")
            blk.pop("_internal")
            blk["config"] = (f"config5: configs[4] of BASELINE.json -- the {WORKLOADS[name]['n']}-gaussian cloud of {name}, {ncams} cameras from cameras.json "
                             f"(CamerasConfig::ImportJson, INRIA schema), view r (mod {ncams}) on rank r, {world} ranks, cloud replicated by one ncclBroadcast")
            blk["views"] = rig.gather(view)
            others["config5"] = blk

    if rig.comm is not None:
        rig.comm.close()
    if rank != 0:
        return None

    n = it["n"]
    e2e8_ms, stereo_ms = it["e2e8_ms"], it["stereo_ms"]
    out = {
        "metric": "Msplats/sec at 1920x1080", "value": main["value"], "unit": "Msplats/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(head),
                   "l2": "flushed between timed frames (256 MiB memset, untimed); e2e loop: per-frame working set > L2",
                   "tile": main["tile_px"], "early_out_T": EARLY_OUT_T, "sort": args.sort_mode + sort_desc, "sort_bits": main["sort_bits"], "sort_passes": main["sort_passes"],
                   "visible": main["visible"], "tile_instances": main["tile_instances"], "parallelism": f"views x{world}",
                   "views": "rank 0: the BASELINE camera; rank r > 0: look-at camera on a ring of the same radius" if world > 1 else "the BASELINE camera",
                   "numa_node": rig.numa_node, "nccl": sp.SplatComm.NcclVersion() if world > 1 else None},
        "frames_per_s": main["frames_per_s"], "sort_gkeys_per_s": main["sort_gkeys_per_s"],
        "clocks": it["clocks"],
        "e2e": main["e2e"],
        "parity": main["parity"],
        "exact_mode": main["exact_mode"],
        "stereo_pair": {"ms_resort": stereo_ms["resort"], "ms_shared_sort": stereo_ms["shared_sort"],
                        "note": "Sort(left) + Render(left) + Render(right), device time per pair (rank 0's view); shared_sort = the reference's VR behaviour"},
        "e2e_presented_rgba8": {"value": n * steps * world / (e2e8_ms * 1e-3) / 1e6, "unit": "Msplats/s", "d2h_bytes_per_step": WIDTH * HEIGHT * 4,
                                "ms_per_step": e2e8_ms / steps,
                                "note": "informational: same loop, reading the 8-bit presented frame (RenderDesktop blit) instead of the float target"},
        "launch_modes": it["launch_ms"],
        "gpu_launches": it["launches"] * steps,
        "roofline": main_roof,
        "stage_ms": main["stage_ms"], "stage_gbs": main["stage_gbs"],
        "wall_s_timed_loop": it["wall"], "load_s": main["load_s"], "broadcast_ms": main["broadcast_ms"],
        "workloads": others,
    }
    if world == 1 and not args.no_cpu_baseline:
        base = cpu_arm(head, steps=2, warmup=1, all_cpus=rig.all_cpus)
        out["cpu_baseline"] = base["cpu_baseline"]
    return out


# ----------------------------------------------------------------------------------------------------------------------
# the CPU arm: the reference path's stand-in (oracle/splat_oracle.c, "port") on the box's host cores
# ----------------------------------------------------------------------------------------------------------------------
def cpu_arm_child(workload: str, steps: int, warmup: int, budget_s: float) -> dict:
    """Runs in a child process whose OpenMP runtime started with OMP_PROC_BIND/OMP_PLACES set (cpu_arm): one thread per
    hardware thread, pinned, so the number does not depend on where the scheduler happens to put 128 threads."""
    from oracle import splat_oracle as O
    from splatapult_b200 import scenes

    w = WORKLOADS[workload]
    aos, full = O.import_ply(scene_path(workload), True)
    cam = scenes.translate(0, 0, w["cam_z"])
    proj = scenes.perspective(scenes.FOVY, WIDTH / HEIGHT)
    vp, nf = [0, 0, WIDTH, HEIGHT], [scenes.Z_NEAR, scenes.Z_FAR]
    cores = O.num_threads()
    n = len(aos)
    # first (warm-up) frame sizes the sample so that warmup+steps stay inside the budget
    t0 = time.perf_counter()
    O.render(aos, full, cam, proj, vp, nf, fp16=w["fp16"])
    t_full = time.perf_counter() - t0
    frac = min(1.0, budget_s / max(t_full * (steps + warmup), 1e-9))
    m = n if frac >= 1.0 else max(int(n * frac), 1000)
    sub = aos[:m]
    for _ in range(max(warmup - 1, 0)):
        O.render(sub, full, cam, proj, vp, nf, fp16=w["fp16"])
    times, stage = [], np.zeros(3)
    for _ in range(steps):
        t0 = time.perf_counter()
        _, _, _, ms = O.render(sub, full, cam, proj, vp, nf, fp16=w["fp16"], want_records=True)
        times.append(time.perf_counter() - t0)
        stage += np.array(ms)
    total = float(sum(times))
    sample = (f"{steps} whole frame(s) of the same workload ({m} of {n} gaussians, 1920x1080) after {warmup} warm-up"
              if m == n else f"{steps} frame(s) of the first {m} of {n} gaussians (bounded sample), 1920x1080, after {warmup} warm-up")
    base = {"value": m * steps / total / 1e6, "unit": "Msplats/s", "cores": cores, "kind": "port", "sample": sample,
            "ms_per_frame": total / steps * 1e3, "ms_per_frame_min": min(times) * 1e3, "ms_per_frame_max": max(times) * 1e3,
            "stage_ms": {"project": stage[0] / steps, "sort": stage[1] / steps, "composite": stage[2] / steps},
            "omp": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OMP_PROC_BIND", "OMP_PLACES")}}
    return {"cpu_baseline": base, "total_s": total, "m": m}


def cpu_arm(workload: str, steps: int, warmup: int, all_cpus=None, budget_s: float = 150.0) -> dict:
    """spawn cpu_arm_child with a clean OpenMP environment: all host threads, pinned one per hardware thread"""
    cpus = sorted(all_cpus) if all_cpus else sorted(os.sched_getaffinity(0))
    env = dict(os.environ)
    env.update({"OMP_NUM_THREADS": str(len(cpus)), "OMP_PROC_BIND": "spread", "OMP_PLACES": "threads", "OMP_DYNAMIC": "false",
                "SPLAT_SCENE_DIR": SCENE_DIR})

    def unbind():  # the parent may be bound to one NUMA node (Rig): the CPU arm uses the whole box
        try:
            os.sched_setaffinity(0, cpus)
        except OSError:
            pass

    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-arm-child", "--workload", workload, "--workload-json", json.dumps(WORKLOADS[workload]),
           "--steps", str(steps), "--warmup", str(warmup), "--cpu-budget", str(budget_s)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, preexec_fn=unbind)
    if p.returncode != 0:
        raise RuntimeError(f"CPU arm failed ({p.returncode}): {p.stderr[-2000:]}")
    return json.loads(p.stdout.strip().splitlines()[-1])


def run_reference(args, rank, world):
    if rank != 0:
        return None
    warmup = max(args.warmup, 1)
    res = cpu_arm(args.workload, steps=args.steps, warmup=warmup)
    base, total = res["cpu_baseline"], res["total_s"]
    return {
        "impl": "reference", "metric": "Msplats/sec at 1920x1080", "value": base["value"], "unit": "Msplats/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload),
                   "arm": "reference arm = CPU oracle (the reference's GLSL/OpenGL path cannot execute on this box), all host threads pinned, rank 0 only"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "Msplats/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stage_ms": base["stage_ms"],
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, help=f"the headline workload, one of {sorted(WORKLOADS)} (the others become info blocks)")
    ap.add_argument("--headline-only", action="store_true", help="skip the info blocks (other workloads, config5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison (profiling runs)")
    ap.add_argument("--sort-mode", default="lsd", choices=["binned", "lsd", "lsd_flat"])
    ap.add_argument("--sort-tile", type=int, default=0, choices=[0, 2048, 4096], help="onesweep tile: 0 = the library's automatic choice")
    ap.add_argument("--subblock-cull", type=int, default=2, choices=[0, 1, 2], help="compositor sub-block mask: 0 none, 1 box, 2 exact ellipse")
    ap.add_argument("--graph", type=int, default=-1, choices=[-1, 0, 1], help="replay a Sort's device work as one CUDA graph (-1: library default = on)")
    ap.add_argument("--pdl", type=int, default=-1, help="programmatic dependent launch, mask: 1 depth passes, 2 emission, 4 tile passes, 8 compositor (-1: library default = 5)")
    ap.add_argument("--cpu-arm-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-budget", type=float, default=150.0, help=argparse.SUPPRESS)
    ap.add_argument("--workload-json", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.workload_json:  # the parent's definition of the workload (cpu_arm)
        WORKLOADS[args.workload] = json.loads(args.workload_json)
    if args.workload not in WORKLOADS:
        ap.error(f"unknown workload {args.workload!r}: choose from {sorted(WORKLOADS)}")

    if args.cpu_arm_child:
        print(json.dumps(cpu_arm_child(args.workload, args.steps, args.warmup, args.cpu_budget)), flush=True)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if args.steps > 5 and "--steps" not in " ".join(sys.argv):
            args.steps = 3
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return

    import splatapult_b200 as sp

    if sp.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: splatapult_b200 has no CPU fallback")
    if world > 1:
        import torch
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        # generate before anybody needs them
        for name in ([args.workload] if args.headline_only else sorted(WORKLOADS)):
            scene_path(name)
        if world > 1 and not args.headline_only:
            cameras_json_path("6m_deg3_fp16")
    if world > 1:
        import torch.distributed as dist

        dist.barrier()  # scenes and cameras.json exist before any other rank looks for them
    out = run_ours(args, rank, local_rank, world)
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
