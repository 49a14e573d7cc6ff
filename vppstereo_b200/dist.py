"""Multi-GPU plumbing: the hot path shards by FRAME.  One process per GPU, contiguous frame
ranges per rank, no communication while computing, one collective at the end to gather the disparities.  An oversized
single frame splits by ROW BANDS where the stage allows it exactly (vpp_rnd_banded, band_with_halo).

Only torch.distributed is used for communication (NCCL on GPUs, gloo in the CPU tests of this host logic).
"""
import numpy as np


def shard_range(n_frames, rank, world_size):
    """Contiguous, balanced frame range [lo, hi) of `rank`: the first n % world ranks own one extra frame."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, extra = divmod(int(n_frames), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_frames, world_size):
    return [shard_range(n_frames, r, world_size)[1] - shard_range(n_frames, r, world_size)[0] for r in range(world_size)]


def chunks(lo, hi, chunk):
    """Sub-ranges of at most `chunk` frames (bounds the workspace: 368 MB of cost volumes per KITTI frame)."""
    out = []
    while lo < hi:
        out.append((lo, min(hi, lo + chunk)))
        lo += chunk
    return out


def gather_frames(local, n_frames, group=None, dst=None):
    """All ranks pass their shard [n_local, ...] (contiguous frame range in rank order); returns the full
    [n_frames, ...] tensor on every rank (dst=None, all_gather) or on rank `dst` only (gather; others get None).
    Shards may be uneven: they are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_frames, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} frames, expected {sizes[rank]}")
    m = max(sizes)
    padded = local
    if local.shape[0] < m:
        pad = torch.zeros((m - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], 0)
    padded = padded.contiguous()
    if dst is None:
        out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, padded, group=group)
        parts = [out[r * m: r * m + sizes[r]] for r in range(world)]
        return torch.cat(parts, 0) if any(s != m for s in sizes) else out
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][: sizes[r]] for r in range(world)], 0)


def run_sharded(n_frames, make_inputs, process, chunk=16, group=None, gather=True):
    """Frame-sharded driver.  make_inputs(lo, hi) -> inputs of frames [lo, hi); process(inputs) -> tensor [hi-lo, ...].
    Returns the gathered [n_frames, ...] result (every rank) or the local shard when gather=False."""
    import torch
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)
    lo, hi = shard_range(n_frames, rank, world)
    outs = [process(make_inputs(a, b)) for a, b in chunks(lo, hi, chunk)]
    local = torch.cat(outs, 0) if outs else None
    if local is None:
        raise ValueError("a rank received no frames: use n_frames >= world_size")
    return gather_frames(local, n_frames, group=group) if gather else local


def band_with_halo(n_rows, rank, world_size, halo):
    """Row band of `rank` for an oversized frame: (lo, hi) = the rows the rank owns, (rlo, rhi) = the rows it
    has to read (owned rows plus `halo` rows on either side, clipped).  census needs halo 3, median 2, the rnd projection
    reads hint rows within its patch radius."""
    lo, hi = shard_range(n_rows, rank, world_size)
    return (lo, hi), (max(lo - halo, 0), min(hi + halo, n_rows))


def vpp_rnd_banded(left, right, hints, group=None, pattern=None, seed=0, **vpp_kwargs):
    """Random-pattern VPP of ONE oversized frame split by row bands across the ranks of `group`: every
    rank holds the full CUDA inputs, projects only its band of image rows (vppb200_vpp_scan_rnd_rows: exact, the stream
    position of every draw is a closed form of the full hint map) and the bands are all-gathered.  Returns (lc, rc) uint8
    [H,W,C] on every rank.  Keyword arguments are those of vpp() for method="rnd" without adaptive patches; all ranks must
    pass the same `pattern` or `seed`."""
    import torch
    import torch.distributed as dist
    from . import _lib
    from . import vpp_core_opt as core
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_available() and dist.is_initialized() else (0, 1)
    kw = dict(wsize=3, left2right=True, blending=0.4, uniform_color=False, c_occ=0.0, g_occ=None, discard_occ=False, interpolate=True)
    unknown = set(vpp_kwargs) - set(kw)
    if unknown:
        raise TypeError(f"vpp_rnd_banded: unsupported arguments {sorted(unknown)}")
    kw.update(vpp_kwargs)
    lc = _lib.as_device(left, torch.uint8).clone().contiguous()
    rc = _lib.as_device(right, torch.uint8).clone().contiguous()
    g = _lib.as_device(hints, torch.float32).contiguous()
    if lc.dim() == 2:
        lc, rc = lc.unsqueeze(-1).contiguous(), rc.unsqueeze(-1).contiguous()
    H, W = g.shape
    occ = torch.zeros((H, W), dtype=torch.uint8, device=g.device) if kw["g_occ"] is None else \
        (_lib.as_device(kw["g_occ"], torch.float32) != 0).to(torch.uint8)
    lo, hi = shard_range(H, rank, world)
    rng_seed = None if pattern is not None else (int(seed) | (1 << 63))
    core._scan("rnd", lc, rc, g, W, H, lc.shape[-1], kw["uniform_color"], kw["wsize"], None, 1 if kw["left2right"] else 0,
               kw["blending"], kw["c_occ"], occ, kw["discard_occ"], kw["interpolate"], pattern, 1, device_rng_seed=rng_seed,
               want_counts=False, rows=(lo, hi))
    if world == 1:
        return lc, rc
    return gather_frames(lc[lo:hi].contiguous(), H, group=group), gather_frames(rc[lo:hi].contiguous(), H, group=group)


def bind_to_gpu_numa(device_index):
    """Pin this process (and therefore the pinned host buffers it allocates next: first-touch placement) to the CPU cores
    that are local to GPU `device_index` (NVML's ideal CPU affinity).  With one process per GPU the host<->device copies of
    the end-to-end path then stay on the GPU's own socket instead of crossing the inter-socket link.  Returns the cores bound
    to, or None when NVML / the cpuset does not allow it (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = device_index
        if vis:
            ent = vis.split(",")[device_index].strip()
            phys = int(ent) if ent.isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(os.cpu_count() or 1, 1) + 63) // 64)
        ideal = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus = sorted(ideal & set(os.sched_getaffinity(0)))
        if len(cpus) < 2:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
