"""Frame-path inference driver: host frames in, stylised uint8 BGR frames out.

This is the call a user of the reference's `Inference` iterator (RC/utilities.py:212-235) ends up
in; bench.py's `e2e` number times exactly this with pinned host buffers.
"""
from __future__ import annotations

import os

import torch

from . import _lib, ops


class FrameStylizer:
    """`lanes` > 1 splits every batch into that many independent sub-batches, each with its own plan (arena) on its own
    CUDA stream.  Frames are independent, and the forward alternates tensor-core-bound tap-GEMMs with HBM-bound
    InstanceNorm applies: with two lanes in flight the GPU co-schedules one lane's apply kernel (no shared memory, few
    registers) next to the other lane's persistent tap-GEMM CTAs instead of leaving either resource idle."""

    def __init__(self, model, H: int, W: int, batch: int = 1, lanes: int = 1):
        p = next(model.parameters())
        if not p.is_cuda:
            raise _lib.VstError("FrameStylizer needs the model on a CUDA device (no CPU fallback)")
        self.model, self.H, self.W, self.N = model, H, W, batch
        self.device = p.device
        self.in_ch = 3 * model.input_frame_num
        self.x_dev = torch.empty((batch, self.in_ch, H, W), dtype=torch.float32, device=self.device)
        self.u8_dev = torch.empty((batch, H, W, 3), dtype=torch.uint8, device=self.device)
        self.x_pin = torch.empty((batch, self.in_ch, H, W), dtype=torch.float32).pin_memory()
        self.u8_pin = torch.empty((batch, H, W, 3), dtype=torch.uint8).pin_memory()
        self.plan = None
        self.lanes = 1
        self.paired = False
        if model.precision in ("bf16", "fp16"):
            if lanes > 1 and batch % lanes == 0:
                self.lanes = lanes
                self.plans = [model.plan(batch // lanes, H, W, slot=i) for i in range(lanes)]
                self.lane_streams = [torch.cuda.Stream(self.device) for _ in range(lanes)]
                self.lane_done = [torch.cuda.Event() for _ in range(lanes)]
                self.plan = self.plans[0]
                # VST_PAIR=1 (opt-in, measured slower than two streams - DESIGN.md 6c): ONE stream, the two half-batch plans in
                # lock step, every tap-GEMM launch of one carrying the other's pending InstanceNorm apply on four extra warps
                self.paired = lanes == 2 and model.precision == "bf16" and os.environ.get("VST_PAIR", "0") == "1"
            else:
                self.plan = model.plan(batch, H, W)

    def run_device(self, x_dev: torch.Tensor) -> torch.Tensor:
        """x_dev fp32 NCHW on device -> uint8 BGR [N,H,W,3] on device (no host traffic)."""
        self._forward_u8(x_dev, self.u8_dev)
        return self.u8_dev

    def _forward_u8(self, x_dev: torch.Tensor, u8_dev: torch.Tensor) -> None:
        """Stylise x_dev into u8_dev on the current stream (fans out over the lanes and joins them again)."""
        if x_dev.dtype == torch.uint8:
            # decoder frames (uint8 BGR HWC): cvframe_to_tensor is folded into the first kernel (vst_plan_forward_bgr8)
            if self.plan is None or self.in_ch != 3:
                raise _lib.VstError("uint8 BGR frames need a tensor-core plan (bf16 / fp16) of a single-frame network")
            if self.lanes > 1:
                cur = torch.cuda.current_stream(self.device)
                ready = torch.cuda.Event()
                ready.record(cur)
                n = self.N // self.lanes
                for i, (pl, st) in enumerate(zip(self.plans, self.lane_streams)):
                    st.wait_event(ready)
                    with torch.cuda.stream(st):
                        pl.forward_bgr8(x_dev[i * n:(i + 1) * n], u8_out=u8_dev[i * n:(i + 1) * n])
                        self.lane_done[i].record(st)
                for ev in self.lane_done:
                    cur.wait_event(ev)
            else:
                self.plan.forward_bgr8(x_dev, u8_out=u8_dev)
        elif self.paired:
            n = self.N // 2
            xa, xb, ua, ub = x_dev[:n], x_dev[n:], u8_dev[:n], u8_dev[n:]
            _lib.check(_lib.lib().vst_plan_forward_pair(self.plans[0]._h, self.plans[1]._h, xa.data_ptr(), xb.data_ptr(),
                                                        ua.data_ptr(), ub.data_ptr(), None, None,
                                                        torch.cuda.current_stream(self.device).cuda_stream),
                       "vst_plan_forward_pair")
        elif self.lanes > 1:
            cur = torch.cuda.current_stream(self.device)
            ready = torch.cuda.Event()
            ready.record(cur)
            n = self.N // self.lanes
            for i, (pl, st) in enumerate(zip(self.plans, self.lane_streams)):
                st.wait_event(ready)
                with torch.cuda.stream(st):
                    pl.forward(x_dev[i * n:(i + 1) * n], want_img=False, u8_out=u8_dev[i * n:(i + 1) * n])
                    self.lane_done[i].record(st)
            for ev in self.lane_done:
                cur.wait_event(ev)
        elif self.plan is not None:
            self.plan.forward(x_dev, want_img=False, u8_out=u8_dev)
        else:
            # fp32 module path: clamp(0,255) -> HWC -> BGR -> uint8 truncation (RC/utilities.py:219-224) in one pack kernel
            ops.pack_bgr_u8(self.model(x_dev)[-1], out=u8_dev)

    def stylize_u8(self, x_host: torch.Tensor):
        """Host fp32 frames [N,in_ch,H,W] -> numpy uint8 BGR [N,H,W,3] (H2D + forward + D2H)."""
        src = x_host if x_host.is_pinned() else self.x_pin.copy_(x_host)
        self.x_dev.copy_(src, non_blocking=True)
        self.run_device(self.x_dev)
        self.u8_pin.copy_(self.u8_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.u8_pin.numpy()

    def stylize_frames(self, frames_host: torch.Tensor):
        """Host decoder frames, uint8 BGR [N,H,W,3] (pinned or not) -> numpy uint8 BGR [N,H,W,3]: upload (a quarter of the
        float tensor's bytes), forward with the BGR->RGB / float conversion inside the first kernel, download."""
        if not hasattr(self, "f8_dev"):
            self.f8_dev = torch.empty((self.N, self.H, self.W, 3), dtype=torch.uint8, device=self.device)
            self.f8_pin = torch.empty((self.N, self.H, self.W, 3), dtype=torch.uint8).pin_memory()
        src = frames_host if frames_host.is_pinned() else self.f8_pin.copy_(frames_host)
        self.f8_dev.copy_(src, non_blocking=True)
        self._forward_u8(self.f8_dev, self.u8_dev)
        self.u8_pin.copy_(self.u8_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.u8_pin.numpy()

    def stylize_stream(self, batches):
        """Pipelined video path: for each host batch (pinned fp32 [N,in_ch,H,W], or decoder frames uint8 BGR [N,H,W,3]) yield the uint8 BGR
        frames [N,H,W,3] (a pinned tensor, valid until the next-but-one yield).  The H2D copy of batch
        i+1 and the D2H copy of batch i-1 run on side streams under the kernels of batch i, removing the
        per-frame `.to(device)` / `.cpu()` serialisation of RC/utilities.py:217-222."""
        dev = self.device
        if not hasattr(self, "_slots"):
            self._slots = []
            for _ in range(2):
                self._slots.append({
                    "x": torch.empty_like(self.x_dev), "u8": torch.empty_like(self.u8_dev),
                    "pin": torch.empty_like(self.u8_pin).pin_memory(),
                    "xpin": None,   # per-slot pinned staging for NON-pinned inputs (allocated on first use)
                    "in_done": torch.cuda.Event(), "comp_done": torch.cuda.Event(), "out_done": torch.cuda.Event(),
                })
            self._s_in, self._s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        comp = torch.cuda.current_stream(dev)
        pending = []  # slots whose D2H is in flight, oldest first
        for i, xh in enumerate(batches):
            sl = self._slots[i & 1]
            if len(pending) == 2:  # this slot's previous result must be consumed before reuse
                old = pending.pop(0)
                old["out_done"].synchronize()
                yield old["pin"]
            is8 = xh.dtype == torch.uint8       # decoder frames (uint8 BGR [N,H,W,3]) instead of float tensors
            if is8 and "x8" not in sl:
                sl["x8"] = torch.empty((self.N, self.H, self.W, 3), dtype=torch.uint8, device=dev)
            xbuf = sl["x8"] if is8 else sl["x"]
            if xh.is_pinned():
                src = xh
            elif is8:
                if sl.get("x8pin") is None:
                    sl["x8pin"] = torch.empty((self.N, self.H, self.W, 3), dtype=torch.uint8).pin_memory()
                sl["in_done"].synchronize()
                src = sl["x8pin"].copy_(xh)
            else:
                # non-pinned batch: stage it through THIS slot's own pinned buffer, and only after the slot's previous
                # upload has left it (the H2D copy is asynchronous; a single shared staging buffer could be overwritten
                # with batch i+1 while batch i was still crossing PCIe)
                if sl["xpin"] is None:
                    sl["xpin"] = torch.empty_like(self.x_pin).pin_memory()
                sl["in_done"].synchronize()
                src = sl["xpin"].copy_(xh)
            with torch.cuda.stream(self._s_in):
                self._s_in.wait_event(sl["comp_done"])      # previous kernels reading sl["x"] are done
                xbuf.copy_(src, non_blocking=True)
                sl["in_done"].record(self._s_in)
            comp.wait_event(sl["in_done"])
            comp.wait_event(sl["out_done"])                 # previous D2H of sl["u8"] is done
            self._forward_u8(xbuf, sl["u8"])
            sl["comp_done"].record(comp)
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(sl["comp_done"])
                sl["pin"].copy_(sl["u8"], non_blocking=True)
                sl["out_done"].record(self._s_out)
            pending.append(sl)
        for old in pending:
            old["out_done"].synchronize()
            yield old["pin"]


class RtnstvStylizer:
    """RTNSTV frame path as ONE call per batch (RT/utilities.py:296-332 runs the module layer by layer and converts the frame
    with four eager torch ops): the tensor-core forward of `StylizingNetwork` (tc_graph.RtnstvTC: 15 tap-GEMMs + InstanceNorm
    applies, the 3-channel output stage on the fp32 kernels) and the uint8 BGR pack are captured once into a CUDA graph over
    static buffers; `stylize_u8` = H2D copy, one graph launch, D2H copy.  fp32 precision runs the module path + pack kernel."""

    def __init__(self, model, H: int, W: int, batch: int = 1):
        p = next(model.parameters())
        if not p.is_cuda:
            raise _lib.VstError("RtnstvStylizer needs the model on a CUDA device (no CPU fallback)")
        self.model, self.H, self.W, self.N, self.device = model, H, W, batch, p.device
        self.x_dev = torch.empty((batch, 3, H, W), dtype=torch.float32, device=self.device)
        self.u8_dev = torch.empty((batch, H, W, 3), dtype=torch.uint8, device=self.device)
        self.x_pin = torch.empty((batch, 3, H, W), dtype=torch.float32).pin_memory()
        self.u8_pin = torch.empty((batch, H, W, 3), dtype=torch.uint8).pin_memory()
        self.graph = None
        if model.precision == "bf16":
            from .tc_graph import RtnstvTC

            self.net = RtnstvTC(model, batch, H, W)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                      # warm-up outside the capture (lazy buffers, function attributes)
                self.x_dev.zero_()
                ops.pack_bgr_u8(self.net.forward(self.x_dev)[1], out=self.u8_dev)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            # the weights of a stylising run are frozen: they were packed by the warm-up forward and stay out of the captured
            # graph (31 small pack launches = a tenth of a 4-frame 640x360 replay); refresh_weights() re-packs in place
            self.net.repack = False
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                ops.pack_bgr_u8(self.net.forward(self.x_dev)[1], out=self.u8_dev)

    def refresh_weights(self):
        """Re-pack the (possibly updated) module weights into the operand buffers the captured forward reads; call after the
        model's parameters changed (training in the same process, `load_state_dict`)."""
        if self.graph is not None:
            self.net._packer.run()
        return self

    def run_device(self, x_dev: torch.Tensor) -> torch.Tensor:
        if self.graph is not None:
            if x_dev.data_ptr() != self.x_dev.data_ptr():
                self.x_dev.copy_(x_dev, non_blocking=True)
            self.graph.replay()
        else:
            ops.pack_bgr_u8(self.model(x_dev), out=self.u8_dev)
        return self.u8_dev

    def stylize_stream(self, batches):
        """Pipelined video path (the RTNSTV twin of `FrameStylizer.stylize_stream`): for each host batch (pinned fp32
        [N,3,H,W]) yield the uint8 BGR frames [N,H,W,3] (a pinned tensor, valid until the next-but-one yield).  The upload of
        batch i+1 and the download of batch i-1 run on side streams under the captured forward of batch i; the graph works on
        its static buffers, a slot's frames enter / leave them through device-to-device copies (14 MB per 4 frames)."""
        dev = self.device
        if not hasattr(self, "_slots"):
            self._slots = [{"x": torch.empty_like(self.x_dev), "u8": torch.empty_like(self.u8_dev),
                            "pin": torch.empty_like(self.u8_pin).pin_memory(), "xpin": None,
                            "in_done": torch.cuda.Event(), "comp_done": torch.cuda.Event(), "out_done": torch.cuda.Event()}
                           for _ in range(2)]
            self._s_in, self._s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        comp = torch.cuda.current_stream(dev)
        pending = []
        for i, xh in enumerate(batches):
            sl = self._slots[i & 1]
            if len(pending) == 2:
                old = pending.pop(0)
                old["out_done"].synchronize()
                yield old["pin"]
            if xh.is_pinned():
                src = xh
            else:
                if sl["xpin"] is None:
                    sl["xpin"] = torch.empty_like(self.x_pin).pin_memory()
                sl["in_done"].synchronize()
                src = sl["xpin"].copy_(xh)
            with torch.cuda.stream(self._s_in):
                self._s_in.wait_event(sl["comp_done"])
                sl["x"].copy_(src, non_blocking=True)
                sl["in_done"].record(self._s_in)
            comp.wait_event(sl["in_done"])
            comp.wait_event(sl["out_done"])
            self.run_device(sl["x"])
            sl["u8"].copy_(self.u8_dev, non_blocking=True)
            sl["comp_done"].record(comp)
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(sl["comp_done"])
                sl["pin"].copy_(sl["u8"], non_blocking=True)
                sl["out_done"].record(self._s_out)
            pending.append(sl)
        for old in pending:
            old["out_done"].synchronize()
            yield old["pin"]

    def stylize_u8(self, x_host: torch.Tensor):
        """Host fp32 frames [N,3,H,W] -> numpy uint8 BGR [N,H,W,3] (H2D + one graph launch + D2H)."""
        src = x_host if x_host.is_pinned() else self.x_pin.copy_(x_host)
        self.x_dev.copy_(src, non_blocking=True)
        self.run_device(self.x_dev)
        self.u8_pin.copy_(self.u8_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.u8_pin.numpy()
