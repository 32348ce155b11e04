"""Python handle on the tensor-core plan (vst_plan_* in include/vst_b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _lib
from ._lib import NetDesc, check


class ReCoNetPlan:
    """One ReCoNet-family network at one input shape on the tcgen05/TMA path.

    Owns a device arena (activations in padded NHWC bf16, packed bf16 weights, IN statistics).
    `tensors` are the 62 state_dict tensors in the reference's registration order.
    """

    PLAN_FP16 = 1   # include/vst_b200.h VST_PLAN_FP16

    def __init__(self, tensors: List[torch.Tensor], in_ch, c1, c2, c3, d1, d2, N, H, W, device, fp16: bool = False):
        L = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.VstError("ReCoNetPlan needs a CUDA device (no CPU fallback)")
        self.desc = NetDesc(0, in_ch, c1, c2, c3, d1, d2, N, H, W, self.PLAN_FP16 if fp16 else 0)
        self.N, self.H, self.W, self.c3, self.in_ch = N, H, W, c3, in_ch
        self._widths = (c1, c2, c3, d1, d2)
        nbytes = L.vst_plan_arena_bytes(C.byref(self.desc))
        self.arena = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        host = [t.detach().to("cpu", torch.float32).contiguous() for t in tensors]
        ptrs = (C.c_void_p * len(host))(*[t.data_ptr() for t in host])
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            check(L.vst_plan_create(C.byref(self.desc), ptrs, len(host), self.arena.data_ptr(), nbytes, stream,
                                    C.byref(handle)), "vst_plan_create")
        self._h = handle
        self.launches = L.vst_plan_launches(self._h)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().vst_plan_destroy(h)
            except Exception:
                pass

    def forward(self, x: torch.Tensor, want_img=True, want_features=False, u8_out: Optional[torch.Tensor] = None,
                img_out: Optional[torch.Tensor] = None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """x: fp32 NCHW [N,in_ch,H,W] in 0..255 on the plan's device.  Returns (img fp32 NCHW | None,
        features fp32 NCHW | None); `u8_out` ([N,H,W,3] uint8) receives the BGR byte frames if given."""
        if x.dtype != torch.float32 or not x.is_cuda or not x.is_contiguous():
            raise _lib.VstError("plan.forward: x must be a contiguous float32 CUDA tensor")
        if tuple(x.shape) != (self.N, self.in_ch, self.H, self.W):
            raise _lib.VstError(f"plan.forward: expected {(self.N, self.in_ch, self.H, self.W)}, got {tuple(x.shape)}")
        img = img_out
        if img is None and want_img:
            img = torch.empty((self.N, 3, self.H, self.W), dtype=torch.float32, device=x.device)
        feat = torch.empty((self.N, self.c3, self.H // 4, self.W // 4), dtype=torch.float32, device=x.device) \
            if want_features else None
        check(_lib.lib().vst_plan_forward(self._h, x.data_ptr(), None if img is None else img.data_ptr(),
                                          None if u8_out is None else u8_out.data_ptr(),
                                          None if feat is None else feat.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "vst_plan_forward")
        return img, feat

    def forward_bgr8(self, frames: torch.Tensor, u8_out: Optional[torch.Tensor] = None, img_out: Optional[torch.Tensor] = None,
                     want_img: bool = False) -> Optional[torch.Tensor]:
        """frames: uint8 [N,H,W,3] BGR on the plan's device - decoder frames as `cv2.VideoCapture.read` returns them; the
        host-side `cvframe_to_tensor` (RC/utilities.py:119-123) is folded into the first kernel.  Same results as `forward`."""
        if frames.dtype != torch.uint8 or not frames.is_cuda or not frames.is_contiguous():
            raise _lib.VstError("plan.forward_bgr8: frames must be a contiguous uint8 CUDA tensor")
        if tuple(frames.shape) != (self.N, self.H, self.W, 3) or self.in_ch != 3:
            raise _lib.VstError(f"plan.forward_bgr8: expected {(self.N, self.H, self.W, 3)} for a single-frame network, "
                                f"got {tuple(frames.shape)} (in_ch {self.in_ch})")
        img = img_out
        if img is None and want_img:
            img = torch.empty((self.N, 3, self.H, self.W), dtype=torch.float32, device=frames.device)
        check(_lib.lib().vst_plan_forward_bgr8(self._h, frames.data_ptr(), None if img is None else img.data_ptr(),
                                               None if u8_out is None else u8_out.data_ptr(), None,
                                               torch.cuda.current_stream().cuda_stream), "vst_plan_forward_bgr8")
        return img

    def forward_upto(self, x: torch.Tensor, layer: int) -> torch.Tensor:
        """Run stages 0..layer only and return that stage's activation (test hook; activation buffers
        are recycled, so a stage must be read before later stages overwrite it)."""
        L = _lib.lib()
        check(L.vst_plan_set_stop_after(self._h, layer), "vst_plan_set_stop_after")
        try:
            self.forward(x, want_img=True)
            return self.activation(layer)
        finally:
            check(L.vst_plan_set_stop_after(self._h, -1), "vst_plan_set_stop_after")

    def activation(self, layer: int) -> torch.Tensor:
        """Post-IN activation of stage `layer` (0 conv1 ... 14 deconv2) as fp32 NCHW (test hook)."""
        c1, c2, c3, d1, d2 = self._widths
        C_ = [c1, c2] + [c3] * 11 + [d1, d2]
        div = [1, 2] + [4] * 11 + [2, 1]
        out = torch.empty((self.N, C_[layer], self.H // div[layer], self.W // div[layer]), dtype=torch.float32,
                          device=self.device)
        check(_lib.lib().vst_plan_debug_activation(self._h, layer, out.data_ptr(), out.numel(),
                                                   torch.cuda.current_stream().cuda_stream), "vst_plan_debug_activation")
        return out

    STAGE_NAMES = ["conv1", "conv2", "conv3"] + [f"res{i}.conv{j}" for i in range(1, 6) for j in (1, 2)] + \
                  ["deconv1", "deconv2", "deconv3"]

    def set_timing(self, enable: bool) -> None:
        check(_lib.lib().vst_plan_set_timing(self._h, int(enable)), "vst_plan_set_timing")

    def get_timing(self):
        """-> (dict stage -> mean ms per launch, number of forwards averaged).  Synchronises."""
        torch.cuda.synchronize(self.device)
        ms = (C.c_float * 16)()
        n = C.c_int()
        check(_lib.lib().vst_plan_get_timing(self._h, ms, C.byref(n)), "vst_plan_get_timing")
        return dict(zip(self.STAGE_NAMES, list(ms))), n.value

    def stage_flops(self):
        """Algorithmic FLOPs (2*MACs of the REFERENCE convolution) per tap-GEMM launch."""
        c1, c2, c3, d1, d2 = self._widths
        N, H, W = self.N, self.H, self.W
        f = [2 * N * H * W * c1 * self.in_ch * 81, 2 * N * (H // 2) * (W // 2) * c2 * c1 * 9,
             2 * N * (H // 4) * (W // 4) * c3 * c2 * 9]
        f += [2 * N * (H // 4) * (W // 4) * c3 * c3 * 9] * 10
        f += [2 * N * (H // 2) * (W // 2) * d1 * c3 * 9, 2 * N * H * W * d2 * d1 * 9, 2 * N * H * W * 3 * d2 * 81]
        return dict(zip(self.STAGE_NAMES, f))
