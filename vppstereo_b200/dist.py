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


class PeerGather:
    """All-gather of per-step results between the ranks of ONE node on the COPY ENGINES: no SM kernel, nothing on the compute
    streams.  The frame-sharded path needs exactly one exchange -- every rank's disparities to every rank -- and the sweeps that
    produce them are cooperative launches that want every SM, so an SM-resident collective kernel on the same device either
    waits for them or makes them wait.  Here every rank owns `depth` gather buffers [world, *shape]; after step k rank r copies
    its shard into slot r of every peer's buffer (cudaMemcpyAsync onto peer-mapped memory: DMA over NVLink), then a 4-byte step
    counter into the peer's flag word r; a consumer stream waits on its own flag words with stream memory operations
    (cuStreamWaitValue32, again no kernel).  Peer memory is mapped once with CUDA IPC handles exchanged through the process
    group (torch.distributed carries only that hand-shake and the barriers).

        pg = PeerGather((B, H, W), torch.float32, device)            # collective: all ranks
        pg.push(k, disp)              # after step k (disp produced on the current stream)
        full = pg.wait(k)             # [world, B, H, W] of step k, valid on the current stream
        pg.release(k)                 # when the reads of `full` have been queued on the current stream

    `available` is False (and push / wait fall back to one NCCL all_gather per step) when peer mapping is not possible."""

    MAX_STEPS = 1 << 20

    def __init__(self, shape, dtype, device, group=None, depth=2):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.device = torch.device(device)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.depth = int(depth)
        self.shape = tuple(shape)
        self.dtype = dtype
        self._raw, self._opened = [], []
        self.available, self.why = False, "single rank"
        with torch.cuda.device(self.device):
            self.ticks = torch.arange(self.MAX_STEPS, dtype=torch.int32, device=self.device)
            self.copy_stream = torch.cuda.Stream(self.device)
            self.credit_stream = torch.cuda.Stream(self.device)
        if self.world > 1:
            self._map_peers()
        if not self.available:                        # plain torch buffers (NCCL fallback / single rank)
            with torch.cuda.device(self.device):
                self.bufs = torch.empty((self.depth, self.world) + self.shape, dtype=dtype, device=self.device)
                self.words = torch.zeros((2, max(self.world, 1)), dtype=torch.int32, device=self.device)
            self.peer_bufs, self.peer_words = [self.bufs], [self.words]
        torch.cuda.current_stream(self.device).synchronize()

    class _RawView:
        """a raw device allocation as something torch.as_tensor understands"""
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}

    def _views(self, ptr):
        """(gather buffers, flag / credit words) over one raw allocation: [depth * world * shape | 2 * world int32]"""
        torch = self.torch
        nb = self.depth * self.world * int(np.prod(self.shape)) * torch.empty((), dtype=self.dtype).element_size()
        nb_al = (nb + 255) & ~255
        raw = torch.as_tensor(self._RawView(ptr, nb_al + 8 * self.world), device=self.device)
        bufs = raw[:nb].view(self.dtype).view((self.depth, self.world) + self.shape)
        words = raw[nb_al:nb_al + 8 * self.world].view(torch.int32).view(2, self.world)
        return bufs, words, nb_al + 8 * self.world

    def _map_peers(self):
        """Every rank cudaMallocs its buffers, exports them with cudaIpcGetMemHandle and opens the peers' handles IN THE CONTEXT
        OF ITS OWN DEVICE (cudaIpcMemLazyEnablePeerAccess): the peer memory then is ordinary device-addressable memory for this
        GPU's DMA engines and kernels, reached over NVLink -- the way NCCL maps its buffers.  (torch's own CUDA-IPC rebuild opens
        the handle in a context on the PEER device instead; copies into it from here were staged through the host: 25 GB/s.)"""
        torch, dist = self.torch, self.dist
        ok, why, handle = True, "", None
        try:
            from cuda.bindings import driver as cu
            from cuda.bindings import runtime as rt
            self._cu, self._rt = cu, rt
            with torch.cuda.device(self.device):
                elem = torch.empty((), dtype=self.dtype).element_size()
                nb = self.depth * self.world * int(np.prod(self.shape)) * elem
                total = ((nb + 255) & ~255) + 8 * self.world
                err, ptr = rt.cudaMalloc(total)
                if int(err) != 0:
                    raise RuntimeError(f"cudaMalloc({total}): {err}")
                self._raw.append(int(ptr))
                err, = rt.cudaMemset(ptr, 0, total)
                err2, h = rt.cudaIpcGetMemHandle(ptr)
                if int(err) != 0 or int(err2) != 0:
                    raise RuntimeError(f"cudaIpcGetMemHandle: {err} {err2}")
                handle = (bytes(h.reserved), self.device.index)
                rt.cudaDeviceSynchronize()
        except Exception as e:                          # no cuda-python / no IPC: every rank must take the same branch
            ok, why = False, f"{type(e).__name__}: {e}"
        everyone = [None] * self.world
        dist.all_gather_object(everyone, (ok, why, handle), group=self.group)
        bad = [w for o, w, _ in everyone if not o]
        if not bad:
            try:
                rt = self._rt
                self.peer_bufs, self.peer_words = [], []
                with torch.cuda.device(self.device):
                    for r, (_, _, (hb, pdev)) in enumerate(everyone):
                        if r == self.rank:
                            pptr = self._raw[0]
                        else:
                            if pdev != self.device.index and not torch.cuda.can_device_access_peer(self.device.index, pdev):
                                raise RuntimeError(f"device {self.device.index} cannot access peer device {pdev}")
                            h = rt.cudaIpcMemHandle_t()
                            h.reserved = hb
                            err, pptr = rt.cudaIpcOpenMemHandle(h, rt.cudaIpcMemLazyEnablePeerAccess)
                            if int(err) != 0:
                                raise RuntimeError(f"cudaIpcOpenMemHandle(rank {r}): {err}")
                            self._opened.append(int(pptr))
                        b, w, _ = self._views(int(pptr))
                        self.peer_bufs.append(b); self.peer_words.append(w)
                self.bufs, self.words = self.peer_bufs[self.rank], self.peer_words[self.rank]
            except Exception as e:
                ok, why = False, f"{type(e).__name__}: {e}"
        else:
            ok, why = False, bad[0]
        verdicts = [None] * self.world
        dist.all_gather_object(verdicts, (ok, why), group=self.group)
        bad = [w for o, w in verdicts if not o]
        self.available, self.why = (not bad), (bad[0] if bad else "peer-mapped (CUDA IPC), copy engines")
        if not self.available:
            self._unmap()

    def _unmap(self):
        rt = getattr(self, "_rt", None)
        if rt is None:
            return
        self.peer_bufs, self.peer_words = [], []
        with self.torch.cuda.device(self.device):
            for p in self._opened:
                rt.cudaIpcCloseMemHandle(p)
            self._opened = []
            for p in self._raw:
                rt.cudaFree(p)
            self._raw = []

    def _wait_word(self, stream, tensor, index, value):
        """stream waits until (int32)tensor[index] >= value (stream memory operation, no kernel)"""
        cu = self._cu
        err, = cu.cuStreamWaitValue32(cu.CUstream(stream.cuda_stream), cu.CUdeviceptr(tensor.data_ptr() + 4 * index), int(value),
                                      cu.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_GEQ)
        if int(err) != 0:
            raise RuntimeError(f"cuStreamWaitValue32 failed: {err}")

    def _copy(self, stream, dst, src):
        """dst <- src (same byte size, both contiguous) as ONE asynchronous driver-level copy on `stream` of this rank's own
        device.  dst may be peer-mapped memory: with unified addressing the DMA engine of this GPU writes it over NVLink.
        (torch's own cross-device copy_ would also queue events on a stream of the PEER device in this process' context on
        that GPU, and that second context then time-slices with the peer's kernels: measured 1.9 ms per step.)"""
        cu = self._cu
        nbytes = src.numel() * src.element_size()
        assert dst.numel() * dst.element_size() == nbytes and src.is_contiguous() and dst.is_contiguous()
        err, = cu.cuMemcpyAsync(cu.CUdeviceptr(dst.data_ptr()), cu.CUdeviceptr(src.data_ptr()), nbytes, cu.CUstream(stream.cuda_stream))
        if int(err) != 0:
            raise RuntimeError(f"cuMemcpyAsync failed: {err}")

    def push(self, k, local):
        """Step k's shard (produced on the current stream) goes to every rank's buffer k % depth, slot `rank`."""
        torch = self.torch
        if k + 1 >= self.MAX_STEPS:
            raise ValueError("PeerGather: step counter exhausted")
        cur = torch.cuda.current_stream(self.device)
        if not self.available:
            if self.world > 1:
                self.dist.all_gather_into_tensor(self.bufs[k % self.depth], local.contiguous(), group=self.group)
            else:
                self.bufs[k % self.depth, 0].copy_(local, non_blocking=True)
            return
        local = local.contiguous()
        cs = self.copy_stream
        cs.wait_stream(cur)
        local.record_stream(cs)
        with torch.cuda.device(self.device):
            for j in range(self.world):
                p = (self.rank + j) % self.world                         # own slot first, then the peers in ring order
                if k >= self.depth and p != self.rank:
                    self._wait_word(cs, self.words[1], p, k - self.depth + 1)       # peer p has released step k - depth
                self._copy(cs, self.peer_bufs[p][k % self.depth, self.rank], local)
                self._copy(cs, self.peer_words[p][0, self.rank:self.rank + 1], self.ticks[k + 1:k + 2])

    # ---- point-to-point use of the same buffers (a mailbox per (slot, source rank)): the row state of a banded sweep travels to
    # ONE neighbour.  Messages of one sender to one receiver carry increasing k; message k lands in bufs[k % depth, sender].
    def send(self, k, local, dst):
        """message k (produced on the current stream) to rank `dst` only"""
        torch = self.torch
        if not self.available:
            raise RuntimeError(f"PeerGather.send needs peer-mapped buffers ({self.why})")
        local = local.contiguous()
        cs = self.copy_stream
        cs.wait_stream(torch.cuda.current_stream(self.device))
        local.record_stream(cs)
        with torch.cuda.device(self.device):
            if k >= self.depth:
                self._wait_word(cs, self.words[1], dst, k - self.depth + 1)          # dst has released this sender's message k - depth
            self._copy(cs, self.peer_bufs[dst][k % self.depth, self.rank].view(-1)[:local.numel()], local.view(-1))
            self._copy(cs, self.peer_words[dst][0, self.rank:self.rank + 1], self.ticks[k + 1:k + 2])

    def wait_from(self, k, src):
        """the current stream waits until message k of rank `src` has landed in bufs[k % depth, src]"""
        with self.torch.cuda.device(self.device):
            self._wait_word(self.torch.cuda.current_stream(self.device), self.words[0], src, k + 1)
        return self.bufs[k % self.depth, src]

    def release_to(self, k, src):
        """the reads of message k of rank `src` have been queued on the current stream: `src` may overwrite that mailbox"""
        torch = self.torch
        cs = self.credit_stream
        cs.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.device(self.device):
            self._copy(cs, self.peer_words[src][1, self.rank:self.rank + 1], self.ticks[k + 1:k + 2])

    def wait(self, k):
        """[world, *shape] of step k; the current stream waits (on the device) until every rank's shard has landed."""
        torch = self.torch
        cur = torch.cuda.current_stream(self.device)
        if self.available:
            with torch.cuda.device(self.device):
                for r in range(self.world):
                    self._wait_word(cur, self.words[0], r, k + 1)
        elif self.world == 1:
            pass                                     # same stream as push
        return self.bufs[k % self.depth]

    def release(self, k):
        """The reads of step k's gathered buffer have been queued on the current stream: tell every peer it may be overwritten."""
        torch = self.torch
        if not self.available:
            return
        cs = self.credit_stream                      # not the push stream: a late peer must not hold back this rank's next push
        cs.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.device(self.device):
            for p in range(self.world):
                if p != self.rank:
                    self._copy(cs, self.peer_words[p][1, self.rank:self.rank + 1], self.ticks[k + 1:k + 2])

    def close(self):
        """Collective: nobody unmaps while a peer may still write."""
        try:
            self.torch.cuda.synchronize(self.device)
            if self.world > 1:
                self.dist.barrier(group=self.group)
        except Exception:
            pass
        if self._raw:                                # raw IPC-exported allocation and the peers' mappings
            self.bufs = self.words = None
            self._unmap()
            self.available = False


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
