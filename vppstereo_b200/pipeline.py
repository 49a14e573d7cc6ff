"""The whole hot path as one object: VPP (rnd pattern, numba arithmetic as test.py uses it) followed by compute_rsgm on a
batch of frames, with the images staying on the device between the two (test.py:158-225 without the host round-trips).

    pipe = VppRsgmPipeline(H, W, batch=64, dmax=192)
    disp = pipe.run_device(left_u8, right_u8, hints_f32)        # CUDA tensors in, CUDA float32 [N,H,W] out
    disp = pipe.run_host(left_pinned, right_pinned, hints_pinned)  # host tensors in, pinned host float32 out

run_host is the end-to-end path: host->device copies of the three inputs and the device->host copy of the disparities
are part of the call.  submit_host / collect is the same path for streams of batches: the copies of batch k+1 and of
batch k-1 ride their own CUDA streams while batch k computes (two staging sets, events between the streams).
"""
import ctypes as C

from . import _lib
from . import vpp_core_opt as _core


class VppRsgmPipeline:
    def __init__(self, H, W, channels=3, batch=64, dmax=192, wsize=3, blending=0.4, c_occ=0.0, left2right=True,
                 interpolate=True, subpixel=True, device=None, seed=1234):
        torch = _lib.require_cuda()
        self.torch = torch
        self.H, self.W, self.C, self.N, self.D = int(H), int(W), int(channels), int(batch), int(dmax)
        self.wsize, self.blending, self.c_occ = int(wsize), float(blending), float(c_occ)
        self.direction = 1 if left2right else 0
        self.interpolate, self.subpixel = bool(interpolate), bool(subpixel)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.seed = int(seed)
        self.step = 0
        L = _lib.lib()
        self.lib = L
        with torch.cuda.device(self.device):
            self.ws_rsgm = torch.empty(L.vppb200_rsgm_workspace_bytes(self.H, self.W, self.C, self.D, self.N), dtype=torch.uint8,
                                       device=self.device)
            self.ws_vpp = torch.empty(L.vppb200_vpp_workspace_bytes(self.H, self.W, self.C, self.N), dtype=torch.uint8,
                                      device=self.device)
            self.occ = torch.zeros((self.N, self.H, self.W), dtype=torch.uint8, device=self.device)
            self.lv = torch.empty((self.N, self.H, self.W, self.C), dtype=torch.uint8, device=self.device)
            self.rv = torch.empty_like(self.lv)
            # two projected-image sets + VPP stream: see run_device
            self.lv2, self.rv2 = [self.lv, torch.empty_like(self.lv)], [self.rv, torch.empty_like(self.lv)]
            self.vpp_stream = torch.cuda.Stream(self.device)
            self.vpp_done = [torch.cuda.Event(), torch.cuda.Event()]
            self.rsgm_done = [None, None]
            self.disp = torch.empty((self.N, self.H, self.W), dtype=torch.float32, device=self.device)
            # staging for run_host
            self.d_left = torch.empty_like(self.lv)
            self.d_right = torch.empty_like(self.lv)
            self.d_hints = torch.empty((self.N, self.H, self.W), dtype=torch.float32, device=self.device)
            self.h_disp = torch.empty((self.N, self.H, self.W), dtype=torch.float32).pin_memory()
            self._stream_sets = None                  # created by the first submit_host

    def workspace_bytes(self):
        return self.ws_rsgm.numel() + self.ws_vpp.numel()

    def run_device(self, left, right, hints, out=None, inputs_ready=None):
        """VPP (in copies) + compute_rsgm; all operands CUDA tensors [N,H,W,C] uint8 / [N,H,W] float32.

        VPP runs on the pipeline's own stream into one of two projected-image sets, compute_rsgm on the caller's current
        stream: the projection of call k+1 overlaps the matcher of call k (VPP is latency bound and fits beside the SGM
        sweeps).  `inputs_ready`: None = the inputs were produced on the current stream (VPP waits for everything queued
        on it so far: no overlap); True = the inputs are complete; a torch.cuda.Event = wait for that event."""
        torch, L = self.torch, self.lib
        N = left.shape[0]
        assert N <= self.N and left.shape[1:] == (self.H, self.W, self.C)
        out = self.disp[:N] if out is None else out
        main = torch.cuda.current_stream(self.device)
        side = self.vpp_stream
        b = self.step & 1
        lv, rv = self.lv2[b][:N], self.rv2[b][:N]
        self.step += 1
        seed = (self.seed * 0x9E3779B97F4A7C15 + self.step) & (2**64 - 1)
        if inputs_ready is None:
            side.wait_stream(main)
        elif inputs_ready is not True:
            side.wait_event(inputs_ready)
        if self.rsgm_done[b] is not None:
            side.wait_event(self.rsgm_done[b])           # this set was last read by the matcher two calls ago
        with torch.cuda.stream(side):
            lv.copy_(left); rv.copy_(right)              # vpp() returns copies (vpp_standalone.py:397)
            rc = L.vppb200_vpp_scan_rnd(_lib.ptr(lv), _lib.ptr(rv), _lib.ptr(hints), self.W, self.H, self.C, 0, self.wsize,
                                        self.direction, C.c_double(self.blending), C.c_double(self.c_occ), _lib.ptr(self.occ),
                                        0, int(self.interpolate), 1, None, None, C.c_uint64(seed), None, _lib.ptr(self.ws_vpp),
                                        C.c_size_t(self.ws_vpp.numel()), N, C.c_void_p(side.cuda_stream))
            _lib.check(rc, "vpp_scan_rnd")
            self.vpp_done[b].record(side)
        main.wait_event(self.vpp_done[b])
        rc = L.vppb200_compute_rsgm(_lib.ptr(left), _lib.ptr(lv), _lib.ptr(rv), None, None, _lib.ptr(out), self.H, self.W,
                                    self.C, self.D, 1 if self.subpixel else 0, None, _lib.ptr(self.ws_rsgm),
                                    C.c_size_t(self.ws_rsgm.numel()), N, C.c_void_p(main.cuda_stream))
        _lib.check(rc, "compute_rsgm")
        ev = torch.cuda.Event()
        ev.record(main)
        self.rsgm_done[b] = ev
        self.lv, self.rv = self.lv2[b], self.rv2[b]     # the projected pair of the latest call
        return out

    def run_host(self, left, right, hints):
        """Host tensors (ideally pinned) in, pinned host float32 [N,H,W] out; the copies are part of the call."""
        torch = self.torch
        N = left.shape[0]
        with torch.cuda.device(self.device):
            self.d_left[:N].copy_(left, non_blocking=True)
            self.d_right[:N].copy_(right, non_blocking=True)
            self.d_hints[:N].copy_(hints, non_blocking=True)
            out = self.run_device(self.d_left[:N], self.d_right[:N], self.d_hints[:N])
            self.h_disp[:N].copy_(out, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        return self.h_disp[:N]

    # ---- streaming front-end: overlapped host<->device copies ---------------------------------------------------
    def _streaming(self):
        torch = self.torch
        if self._stream_sets is None:
            with torch.cuda.device(self.device):
                sets = []
                for _ in range(2):
                    sets.append(dict(left=torch.empty_like(self.lv), right=torch.empty_like(self.lv),
                                     hints=torch.empty_like(self.d_hints), disp=torch.empty_like(self.disp),
                                     h_disp=torch.empty((self.N, self.H, self.W), dtype=torch.float32).pin_memory(),
                                     copied_in=torch.cuda.Event(), computed=torch.cuda.Event(), copied_out=torch.cuda.Event(),
                                     busy=False))
                self._stream_sets = dict(sets=sets, h2d=torch.cuda.Stream(self.device), d2h=torch.cuda.Stream(self.device), turn=0)
        return self._stream_sets

    def submit_host(self, left, right, hints):
        """Queue one batch of host tensors (pinned for real overlap); returns a ticket for collect().  At most two
        batches are in flight: submitting a third one first requires collecting the oldest."""
        torch = self.torch
        ss = self._streaming()
        slot = ss["turn"] & 1
        st = ss["sets"][slot]
        if st["busy"]:
            raise RuntimeError("submit_host: collect() the batch submitted two calls ago first")
        N = left.shape[0]
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.device(self.device):
            with torch.cuda.stream(ss["h2d"]):
                # the staging set was last read by the compute of two submits ago, which collect() has waited for
                st["left"][:N].copy_(left, non_blocking=True)
                st["right"][:N].copy_(right, non_blocking=True)
                st["hints"][:N].copy_(hints, non_blocking=True)
                st["copied_in"].record(ss["h2d"])
            compute.wait_event(st["copied_in"])
            self.run_device(st["left"][:N], st["right"][:N], st["hints"][:N], out=st["disp"][:N], inputs_ready=st["copied_in"])
            st["computed"].record(compute)
            with torch.cuda.stream(ss["d2h"]):
                ss["d2h"].wait_event(st["computed"])
                st["h_disp"][:N].copy_(st["disp"][:N], non_blocking=True)
                st["copied_out"].record(ss["d2h"])
        st["busy"], st["n"] = True, N
        ss["turn"] += 1
        return slot

    def collect(self, ticket):
        """Wait for a submitted batch; returns its pinned host float32 [N,H,W] disparities (valid until that staging set is
        reused by the second submit_host after this call)."""
        st = self._streaming()["sets"][ticket]
        if not st["busy"]:
            raise RuntimeError("collect: nothing in flight for this ticket")
        st["copied_out"].synchronize()
        st["busy"] = False
        return st["h_disp"][:st["n"]]
