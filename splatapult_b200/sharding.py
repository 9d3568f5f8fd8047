"""View sharding across ranks: one process per GPU, views dealt round-robin, the gaussian cloud
replicated by ONE broadcast at start-up, nothing collective on the frame path.

torch.distributed is the plumbing (backend "nccl" over NVLink/NVSwitch on the GPU box, "gloo" in the CPU tests);
the broadcast destination is handed to SplatRenderer.InitFromDevice, which packs it on the device.
"""
from __future__ import annotations

import numpy as np


def views_for_rank(num_views: int, rank: int, world: int) -> list[int]:
    """round-robin: rank r renders views r, r + world, r + 2 world, ... (8 cameras on 8 GPUs -> one each)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, num_views, world))


def broadcast_cloud(aos_host: np.ndarray | None, has_full_sh: bool | None, src: int = 0, device=None):
    """Replicate the reference-layout AoS (float32 [N, 25|61]) from rank `src` to every rank.

    Returns (tensor [N, floats] on `device` (or CPU), has_full_sh, broadcast_ms or None).  Only rank `src` passes
    data; the others pass None.  With CUDA tensors and the nccl backend this is one ncclBroadcast of N*stride bytes.
    """
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        if aos_host is None:
            raise ValueError("the source rank must provide the cloud")
        meta = [(int(aos_host.shape[0]), int(aos_host.shape[1]), bool(has_full_sh))]
    dist.broadcast_object_list(meta, src=src)
    n, floats, full = meta[0]
    dev = device if device is not None else torch.device("cpu")
    buf = torch.empty((n, floats), dtype=torch.float32, device=dev)
    if rank == src:
        buf.copy_(torch.from_numpy(np.ascontiguousarray(aos_host, np.float32)))
    ms = None
    if buf.is_cuda:
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(buf, src=src)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
    else:
        dist.broadcast(buf, src=src)
    return buf, full, ms


def max_over_ranks(values: list[float], device=None) -> list[float]:
    """multi-GPU numbers are the max over ranks of per-rank device times."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(values, dtype=torch.float64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]
