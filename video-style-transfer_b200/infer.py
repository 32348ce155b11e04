"""Frame-path inference driver: host frames in, stylised uint8 BGR frames out.

This is the call a user of the reference's `Inference` iterator (RC/utilities.py:212-235) ends up
in; bench.py's `e2e` number times exactly this with pinned host buffers.
"""
from __future__ import annotations

import torch

from . import _lib, ops


class FrameStylizer:
    def __init__(self, model, H: int, W: int, batch: int = 1):
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
        self.plan = model.plan(batch, H, W) if model.precision == "bf16" else None

    def run_device(self, x_dev: torch.Tensor) -> torch.Tensor:
        """x_dev fp32 NCHW on device -> uint8 BGR [N,H,W,3] on device (no host traffic)."""
        if self.plan is not None:
            self.plan.forward(x_dev, want_img=False, u8_out=self.u8_dev)
        else:
            with torch.no_grad():
                img = self.model(x_dev)[-1]
            # clamp(0,255) -> HWC -> BGR -> uint8 truncation (RC/utilities.py:219-224)
            self.u8_dev.copy_(img.clamp(0, 255).permute(0, 2, 3, 1).flip(-1).to(torch.uint8))
        return self.u8_dev

    def stylize_u8(self, x_host: torch.Tensor):
        """Host fp32 frames [N,in_ch,H,W] -> numpy uint8 BGR [N,H,W,3] (H2D + forward + D2H)."""
        self.x_pin.copy_(x_host)
        self.x_dev.copy_(self.x_pin, non_blocking=True)
        self.run_device(self.x_dev)
        self.u8_pin.copy_(self.u8_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.u8_pin.numpy()
