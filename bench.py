#!/usr/bin/env python
"""bench.py -- headline benchmark of the splat hot path (BASELINE.json: Msplats/s at 1920x1080).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one frame: SplatRenderer::Sort + SplatRenderer::Render of the whole cloud from one camera into a
cleared 1920x1080 float framebuffer.  Workloads:
    1m_deg0       BASELINE configs[1]: 1M synthetic gaussians, SH degree 0, fp32 target      (default)
    1m_deg3       BASELINE configs[2]: 1M synthetic gaussians, full SH (48 coeffs), fp32 target
    6m_deg3_fp16  BASELINE configs[3]: 6M gaussians, full SH, fp16 target
`value` = gaussians in the cloud x frames x ranks / device time (inputs resident in HBM, L2 flushed between
frames); `e2e` = the same through the C ABI with host camera matrices in and the framebuffer copied to pinned
host memory every frame.  N > 1 (torchrun, one rank per GPU): views shard one per GPU, the cloud is
NCCL-broadcast once at start-up (not timed), no per-frame collective ("scaling": "weak").
--impl reference times the reference path's CPU stand-in (the oracle: the reference's GLSL cannot run here,
there is no OpenGL) on the host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
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
WIDTH, HEIGHT = 1920, 1080
SCENE_DIR = os.environ.get("SPLAT_SCENE_DIR", "/tmp/splat_b200_scenes")


def scene_path(name: str) -> str:
    from splatapult_b200 import scenes

    w = WORKLOADS[name]
    os.makedirs(SCENE_DIR, exist_ok=True)
    path = os.path.join(SCENE_DIR, f"{name}_{w['n']}.ply")
    if not os.path.exists(path):
        scenes.write_synthetic_ply(path, w["n"], w["deg"], seed=w["seed"], half_extent=w["half"])
    return path


def camera_for_rank(name: str, rank: int, world: int) -> np.ndarray:
    """rank 0 renders the BASELINE camera (translate(0,0,z)); other ranks their own view on a ring of the same radius."""
    from splatapult_b200 import scenes

    z = WORKLOADS[name]["cam_z"]
    if world > 1 and name.startswith("6m"):
        # BASELINE configs[4]: 8 cameras from a cameras.json (INRIA schema, read by the C++ CamerasConfig), one view per GPU
        import splatapult_b200 as sp
        from splatapult_b200 import sharding

        path = os.path.join(SCENE_DIR, "cameras_ring8.json")
        if not os.path.exists(path):  # written by scene_path() on rank 0 before the other ranks get here
            scenes.write_cameras_json(path, scenes.ring_cameras(8, z, 3.0), WIDTH, HEIGHT, 1303.7)
        cfg = sp.CamerasConfig()
        if not cfg.ImportJson(path):
            raise RuntimeError(cfg.LastError())
        mine = sharding.views_for_rank(cfg.GetNumCameras(), rank, world)
        return cfg.GetCamera(mine[0] if mine else 0)[0]
    if rank == 0:
        return scenes.translate(0, 0, z)
    ang = 2.0 * math.pi * rank / max(world, 1)
    return scenes.look_at_camera((z * math.sin(ang), 0.0, z * math.cos(ang)), (0.0, 0.0, 0.0))


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


def algorithmic_bytes(n, v, d, deg_mode, passes, tiles, pixels, fp16, binned):
    """information bytes per stage (not padded strides), for the sort path actually run.
    rec = 16 (pos+alpha) + 24 (symmetric Sigma) + SH {12 | 48 | 192}."""
    rec = {0: 52, 1: 88, 3: 232}[deg_mode]
    fbb = pixels * (8 if fp16 else 16)
    if binned:
        # P: read cloud, write 48 B records + 16 B side records (rect, key, count); per-instance REDs not counted
        b_p = 16 * n + (rec - 16) * v + 48 * v + 16 * n
        b_bin = 16 * n + 8 * d + 16 * tiles          # scan + scatter: side records in, (key,id) pairs out
        b_seg = 16 * d                               # every segment read once and written once
        b_c = 8 * d + 48 * v + 8 * tiles + fbb       # pairs once, each record once (L2 re-gathers not credited)
        return {"project": b_p, "bin_scatter": b_bin, "segment_sort": b_seg, "composite": b_c, "frame": b_p + b_bin + b_seg + b_c}
    b_p = 16 * n + (rec - 16) * v + 48 * v + 12 * d
    b_hist = 8 * d
    b_pass = 24 * d
    b_ranges = 8 * tiles  # written by the last sort pass's scatter (no separate pass over the keys)
    b_c = 4 * d + 48 * v + 8 * tiles + fbb
    return {"project": b_p, "sort_histogram": b_hist, "sort_pass": b_pass, "tile_ranges": b_ranges, "composite": b_c,
            "frame": b_p + b_hist + passes * b_pass + b_ranges + b_c}


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import splatapult_b200 as sp
    from splatapult_b200 import api

    w = WORKLOADS[args.workload]
    if sp.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: splatapult_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)
    vp, nf = [0, 0, WIDTH, HEIGHT], [sp.Z_NEAR, sp.Z_FAR]
    proj = sp.perspective(sp.FOVY, WIDTH / HEIGHT)
    cam = camera_for_rank(args.workload, rank, world)

    # ---- load through the same .ply path a splatapult user has; broadcast once for N > 1
    t0 = time.perf_counter()
    broadcast_ms = None
    r = sp.SplatRenderer(device=local_rank, fb_format=api.FB_FP16 if w["fp16"] else api.FB_FP32, cuda_stream=stream.cuda_stream)
    if world == 1:
        # PLY header parsed on the host, body decoded on the device (same records as GaussianCloud::ImportPly, tested)
        r.InitFromPly(scene_path(args.workload), True, False)
        n_gauss = int(r.GetStats().num_gaussians)
    else:
        from splatapult_b200 import sharding

        host, host_full = None, None
        if rank == 0:
            gc = sp.GaussianCloud(import_full_sh=True, export_full_sh=True)
            if not gc.ImportPly(scene_path(args.workload)):
                raise RuntimeError(gc.LastError())
            host, host_full = gc.GetRawData(), gc.HasFullSH()
        # one ncclBroadcast of the AoS over NVLink/NVSwitch, at start-up only (not in the timed region)
        buf, full, broadcast_ms = sharding.broadcast_cloud(host, host_full, src=0, device=dev)
        n_gauss, stride = int(buf.shape[0]), int(buf.shape[1]) * 4
        r.InitFromDevice(buf.data_ptr(), n_gauss, stride, full)
        del buf
    load_s = time.perf_counter() - t0

    r.SetOption(api.OPT_SORT_MODE, {"lsd": api.SORT_LSD, "lsd_flat": api.SORT_LSD_FLAT, "binned": api.SORT_BINNED}[args.sort_mode])
    r.SetOption(api.OPT_SORT_TILE, args.sort_tile)
    r.SetOption(api.OPT_SUBBLOCK_CULL, args.subblock_cull)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def frame():
        r.Sort(cam, proj, vp, nf)
        r.Render(cam, proj, vp, nf)

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            flush.zero_()
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
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        # clocks / throttle reasons are sampled from here to the end of the e2e loop: the GPU is under the same load
        # throughout (the K timed frames alone can be shorter than nvidia-smi's start-up)
        clocks = ClockSampler(local_rank)
        clocks.__enter__()
        e2e_wall_floor = 0.5  # seconds of extra (untimed) frames after the e2e loop, for the clock samples only
        wall0 = time.perf_counter()
        for a, b in ev:
            flush.zero_()
            a.record(stream)
            frame()
            b.record(stream)
        stream.synchronize()
        wall = time.perf_counter() - wall0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        step_ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = float(sum(step_ms))

        # ---- end to end through the C ABI: host matrices in, framebuffer out to pinned host memory, every frame
        fb_bytes = WIDTH * HEIGHT * (8 if w["fp16"] else 16)
        pinned = [sp.host_alloc(fb_bytes), sp.host_alloc(fb_bytes)]
        for i in range(2):
            frame()
            r.ReadFramebufferAsync(pinned[i].ptr, fb_bytes)
        r.Synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(args.steps):
            frame()  # host matrices -> kernels
            r.ReadFramebufferAsync(pinned[i & 1].ptr, fb_bytes)  # D->H on the copy stream, overlaps the next frame
        r.FenceCopies()  # the end event is recorded only after the last copy has landed
        b.record(stream)
        r.Synchronize()
        if e2e_wall_floor > 0:  # keep the load up long enough for a handful of clock samples
            t_end = time.perf_counter() + e2e_wall_floor
            while time.perf_counter() < t_end:
                frame()
                stream.synchronize()
        clocks.__exit__(None, None, None)
        e2e_ms = a.elapsed_time(b)
        checksum = float(pinned[(args.steps - 1) & 1].array(np.float16 if w["fp16"] else np.float32, (HEIGHT, WIDTH, 4))[..., :3].mean())

        # ---- informational: the same loop reading the PRESENTED 8-bit frame (the reference's RenderDesktop blit) instead
        # of the float target -- 4 bytes per pixel over PCIe instead of 16 / 8
        p8 = [sp.host_alloc(WIDTH * HEIGHT * 4), sp.host_alloc(WIDTH * HEIGHT * 4)]
        for i in range(2):
            frame()
            r.ReadPresented8Async(p8[i].ptr, p8[i].nbytes)
        r.Synchronize()
        a8, b8 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a8.record(stream)
        for i in range(args.steps):
            frame()
            r.ReadPresented8Async(p8[i & 1].ptr, p8[i & 1].nbytes)
        r.FenceCopies()
        b8.record(stream)
        r.Synchronize()
        e2e8_ms = a8.elapsed_time(b8)

        # ---- per-stage profile (separate, untimed pass): events between the kernels
        r.SetOption(api.OPT_PROFILE, 1)
        stages = np.zeros(8)
        reps = 10
        for _ in range(reps):
            flush.zero_()
            frame()
            stream.synchronize()
            stages += np.array(r.GetStats().stage_ms)
        stages /= reps
        r.SetOption(api.OPT_PROFILE, 0)
        st = r.GetStats()

        def timed(fn, steps):
            """device ms per call of fn, L2 flushed before each call (same procedure as the headline loop)"""
            for _ in range(3):
                fn()
            stream.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a_, b_ in evs:
                flush.zero_()
                a_.record(stream)
                fn()
                b_.record(stream)
            stream.synchronize()
            return float(sum(a_.elapsed_time(b_) for a_, b_ in evs)) / steps

        # ---- informational: exact mode -- no transmittance early-out, the reference's full blend chain per pixel
        r.SetOption(api.OPT_EARLY_OUT_T, 0.0)
        exact_ms = timed(frame, args.steps)
        r.SetOption(api.OPT_EARLY_OUT_T, 1e-4)

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
        for name, opt in (("resort", 0), ("shared_sort", 1)):
            r.SetOption(api.OPT_SHARED_SORT, opt)
            stereo_ms[name] = timed(stereo_pair, args.steps)
        r.SetOption(api.OPT_SHARED_SORT, 0)

    if world > 1:
        t = torch.tensor([total_ms, e2e_ms, e2e8_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, e2e8_ms = float(t[0]), float(t[1]), float(t[2])

    if rank != 0:
        return None

    n, v, d = int(st.num_gaussians), int(st.num_visible), int(st.num_instances)
    tiles = int(st.tiles_x * st.tiles_y)
    binned = int(st.sort_mode) == api.SORT_BINNED
    ab = algorithmic_bytes(n, v, d, int(st.sh_mode), int(st.sort_passes), tiles, WIDTH * HEIGHT, w["fp16"], binned)
    peak, peak_src = measured_peak_gbs()
    # cand: kernel -> (time it takes in a step, algorithmic bytes per launch, average launch duration)
    if binned:
        stage = {"project": stages[0], "bin_scatter": stages[1], "segment_sort": stages[2], "composite": stages[4], "frame": stages[7]}
        cand = {"composite": (stages[4], ab["composite"], stages[4]), "project": (stages[0], ab["project"], stages[0]),
                "segment_sort": (stages[2], ab["segment_sort"], stages[2]), "bin_scatter": (stages[1], ab["bin_scatter"], stages[1])}
        stage_gbs = {k: (ab[k] / (stage[k] * 1e-3) / 1e9 if stage[k] > 0 else None) for k in ("project", "bin_scatter", "segment_sort", "composite")}
    else:
        split = int(st.sort_mode) == api.SORT_LSD  # depth digits sorted on the N gaussians, tile digits on the D instances
        d_passes = int(st.sort_passes) - 4 if split else int(st.sort_passes)  # onesweep passes that run on the D instances
        if split:
            # project_keys writes a u32 key + u32 id + 8 B rectangle per gaussian instead of 12 B per instance;
            # stage 1 = 4 depth-digit passes on N (u32, u32) pairs (16 N each) + expansion (4 N ids + 8 N rectangles in,
            # (u32 tile, u32 id) = 8 D out); the tile-digit passes move 8-byte instances (16 D per pass)
            ab["project"] = ab["project"] - 12 * d + 16 * n
            ab["depth_sort_emit"] = 4 * 16 * n + 12 * n + 8 * d
            ab["sort_pass"] = 16 * d
            ab["frame"] = ab["project"] + ab["depth_sort_emit"] + d_passes * ab["sort_pass"] + ab["tile_ranges"] + ab["composite"]
            stage = {"project": stages[0], "depth_sort_emit": stages[1], "tile_sort_passes_total": stages[2],
                     "composite": stages[4], "frame": stages[7]}
            names = ("project", "depth_sort_emit", "composite")
        else:
            stage = {"project": stages[0], "sort_passes_total": stages[2], "composite": stages[4], "frame": stages[7]}
            names = ("project", "composite")
        pass_ms = stages[2] / max(d_passes, 1)
        cand = {"composite": (stages[4], ab["composite"], stages[4]), "sort_onesweep": (stages[2], ab["sort_pass"], pass_ms),
                "project": (stages[0], ab["project"], stages[0])}
        stage_gbs = {k: (ab[k] / (stage[k] * 1e-3) / 1e9 if stage[k] > 0 else None) for k in names}
        stage_gbs["sort_pass"] = ab["sort_pass"] / (pass_ms * 1e-3) / 1e9 if pass_ms > 0 else None
    dom = max(cand, key=lambda k: cand[k][0])
    dom_total, dom_bytes, dom_launch_ms = cand[dom]
    achieved = dom_bytes / (dom_launch_ms * 1e-3) / 1e9 if dom_launch_ms > 0 else 0.0
    traffic, issue = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            prof = json.load(f)
        traffic = prof.get(args.workload, {}).get(dom)
        # the dominant kernel of this path is instruction-issue bound, not HBM bound: also report how close it is to
        # that bound -- warp instructions per launch (ncu smsp__inst_executed.sum, a property of the workload) over the
        # live launch time, per SM sub-partition (peak 1 warp instruction per cycle)
        wi = prof.get("warp_instructions", {}).get(args.workload, {}).get(dom)
        clk = clocks.summary().get("sm_mhz")
        if wi and clk and dom_launch_ms > 0:
            smsps = torch.cuda.get_device_properties(dev).multi_processor_count * 4
            issue = {"warp_instructions_per_launch": wi, "ipc_per_smsp": wi / (dom_launch_ms * 1e-3 * clk * 1e6 * smsps), "peak": 1.0,
                     "source": "ncu smsp__inst_executed.sum (profiles/r1j_ncu_current.md) / live launch time / (SMSPs x sampled SM clock)"}
    except Exception:
        pass

    secs = total_ms * 1e-3
    value = n * args.steps * world / secs / 1e6
    out = {
        "metric": "Msplats/sec at 1920x1080", "value": value, "unit": "Msplats/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['cfg']} of BASELINE.json -- {n} synthetic gaussians (seed {w['seed']}), SH degree {w['deg']}, "
                               f"{WIDTH}x{HEIGHT}, {'fp16' if w['fp16'] else 'fp32'} framebuffer, one view per GPU",
                   "l2": "flushed between timed frames (256 MiB memset, untimed); e2e loop: per-frame working set > L2",
                   "tile": 16, "early_out_T": 1e-4, "sort": args.sort_mode + {"lsd": ": onesweep LSD, depth digits on N gaussians + tile digits on D instances", "lsd_flat": ": onesweep LSD, all digits on D instances", "binned": ": MSD by tile id + per-tile smem sort"}[args.sort_mode],"sort_bits": int(st.sort_bits), "sort_passes": int(st.sort_passes),
                   "visible": v, "tile_instances": d, "parallelism": f"views x{world}"},
        "frames_per_s": args.steps * world / secs,
        "sort_gkeys_per_s": (d / ((stages[1] + stages[2]) * 1e-3) / 1e9) if (stages[1] + stages[2]) > 0 else None,
        "clocks": clocks.summary(),
        "e2e": {"value": n * args.steps * world / (e2e_ms * 1e-3) / 1e6, "unit": "Msplats/s", "h2d_bytes_per_step": 38 * 4,
                "d2h_bytes_per_step": fb_bytes, "ms_per_step": e2e_ms / args.steps, "framebuffer_mean": checksum},
        "exact_mode": {"ms_per_step": exact_ms, "value": n * world / (exact_ms * 1e-3) / 1e6, "unit": "Msplats/s",
                       "note": "early_out_T = 0: every list entry of every pixel is blended, as the reference does (rank 0's view)"},
        "stereo_pair": {"ms_resort": stereo_ms["resort"], "ms_shared_sort": stereo_ms["shared_sort"],
                        "note": "Sort(left) + Render(left) + Render(right), device time per pair (rank 0's view); shared_sort = the reference's VR behaviour"},
        "e2e_presented_rgba8": {"value": n * args.steps * world / (e2e8_ms * 1e-3) / 1e6, "unit": "Msplats/s", "d2h_bytes_per_step": WIDTH * HEIGHT * 4,
                                "ms_per_step": e2e8_ms / args.steps,
                                "note": "informational: same loop, reading the 8-bit presented frame (RenderDesktop blit) instead of the float target"},
        "gpu_launches": int(st.kernel_launches) * args.steps,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "issue": issue, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                     "launch_ms": dom_launch_ms, "share_of_step": dom_total / stages[7] if stages[7] > 0 else None,
                     "frame": {"algorithmic_bytes": ab["frame"], "achieved": ab["frame"] / (stages[7] * 1e-3) / 1e9 if stages[7] > 0 else None,
                               "frac": ab["frame"] / (stages[7] * 1e-3) / 1e9 / peak if stages[7] > 0 else None}},
        "stage_ms": {k: float(x) for k, x in stage.items()},
        "stage_gbs": stage_gbs,
        "wall_s_timed_loop": wall, "load_s": load_s, "broadcast_ms": broadcast_ms,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.workload, steps=2, warmup=1)
    return out


def cpu_baseline(workload: str, steps: int, warmup: int, budget_s: float = 150.0, impl_line: bool = False):
    """The reference path's CPU stand-in (oracle/splat_oracle.c, "port") on the box's host cores."""
    from oracle import splat_oracle as O
    from splatapult_b200 import scenes

    w = WORKLOADS[workload]
    # all the host threads the box has (torchrun exports OMP_NUM_THREADS=1, which would cripple the CPU arm)
    try:
        O.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        O.set_num_threads(os.cpu_count() or 1)
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
            "ms_per_frame": total / steps * 1e3, "stage_ms": {"project": stage[0] / steps, "sort": stage[1] / steps, "composite": stage[2] / steps}}
    if impl_line:
        return base, total, m
    return base


def run_reference(args, rank, world):
    if rank != 0:
        return None
    w = WORKLOADS[args.workload]
    base, total, m = cpu_baseline(args.workload, steps=args.steps, warmup=max(args.warmup, 1), impl_line=True)
    return {
        "impl": "reference", "metric": "Msplats/sec at 1920x1080", "value": base["value"], "unit": "Msplats/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": total / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['cfg']} of BASELINE.json, {WIDTH}x{HEIGHT}; reference arm = CPU oracle (the reference's GLSL/OpenGL "
                               "path cannot execute on this box), all host threads, rank 0 only"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "Msplats/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1m_deg0", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sort-mode", default="lsd", choices=["binned", "lsd", "lsd_flat"])
    ap.add_argument("--sort-tile", type=int, default=4096, choices=[2048, 4096])
    ap.add_argument("--subblock-cull", type=int, default=2, choices=[0, 1, 2], help="compositor sub-block mask: 0 none, 1 box, 2 exact ellipse")
    args = ap.parse_args()

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

    if world > 1:
        import torch
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        scene_path(args.workload)  # generate before anybody needs it
        if world > 1 and args.workload.startswith("6m"):
            from splatapult_b200 import scenes

            scenes.write_cameras_json(os.path.join(SCENE_DIR, "cameras_ring8.json"), scenes.ring_cameras(8, WORKLOADS[args.workload]["cam_z"], 3.0),
                                      WIDTH, HEIGHT, 1303.7)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()  # scene and cameras.json exist before any other rank looks for them
    out = run_ours(args, rank, local_rank, world)
    if out is not None:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
