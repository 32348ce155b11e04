"""Functional operators over torch CUDA tensors, each a direct call into libvst_b200.so.

torch is used for device memory and streams only; every FLOP runs in this repo's kernels.
All functions take contiguous fp32 NCHW CUDA tensors (like the reference's tensors) and launch on
torch's current stream.  CPU tensors are rejected by the library (VST_EDEVICE): no fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check

PAD_ZERO, PAD_REFLECT = 0, 1
ACT_NONE, ACT_RELU, ACT_TANH, ACT_RECONET_OUT, ACT_RT_OUT = 0, 1, 2, 3, 4

_scratch = {}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else 0


def _f32(t: torch.Tensor, name: str = "tensor") -> torch.Tensor:
    if t.dtype != torch.float32:
        raise _lib.VstError(f"{name}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def reduce_scratch(device) -> torch.Tensor:
    """Zero-initialised scratch for the deterministic grid reductions (counter stays zeroed)."""
    key = (str(device), torch.cuda.current_stream().cuda_stream if torch.cuda.is_available() else 0)
    if key not in _scratch:
        _scratch[key] = torch.zeros(_lib.lib().vst_reduce_scratch_floats(), dtype=torch.float32, device=device)
    return _scratch[key]


def conv2d(x, w, bias=None, stride=1, pad=0, pad_mode=PAD_REFLECT, ups=1, act=ACT_NONE):
    """[pad] -> Conv2d -> bias -> act.  include/vst_b200.h: vst_conv2d_f32."""
    x, w = _f32(x, "x"), _f32(w, "w")
    N, Cin, H, W = x.shape
    Cout, Cin_w, k, _ = w.shape
    if Cin_w != Cin:
        raise _lib.VstError(f"conv2d: weight expects {Cin_w} input channels, got {Cin}")
    Ho = (H * ups + 2 * pad - k) // stride + 1
    Wo = (W * ups + 2 * pad - k) // stride + 1
    y = torch.empty((N, Cout, Ho, Wo), dtype=torch.float32, device=x.device)
    b = None if bias is None else _f32(bias, "bias")
    check(_lib.lib().vst_conv2d_f32(x.data_ptr(), w.data_ptr(), _ptr(b), y.data_ptr(), N, Cin, H, W, Cout, k, stride,
                                    pad, pad_mode, ups, act, _stream()), "vst_conv2d_f32")
    return y


def conv_transpose2d(x, w, bias=None):
    """ConvTranspose2d(k3, s2, p1, op1).  w: [Cin, Cout, 3, 3]."""
    x, w = _f32(x, "x"), _f32(w, "w")
    N, Cin, H, W = x.shape
    if w.shape[0] != Cin or tuple(w.shape[2:]) != (3, 3):
        raise _lib.VstError("conv_transpose2d: weight must be [Cin, Cout, 3, 3]")
    Cout = w.shape[1]
    y = torch.empty((N, Cout, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
    b = None if bias is None else _f32(bias, "bias")
    check(_lib.lib().vst_conv_transpose2d_f32(x.data_ptr(), w.data_ptr(), _ptr(b), y.data_ptr(), N, Cin, H, W, Cout,
                                              _stream()), "vst_conv_transpose2d_f32")
    return y


def instance_norm(x, gamma, beta, residual=None, act=ACT_NONE, eps=1e-5, return_stats=False):
    x, gamma, beta = _f32(x, "x"), _f32(gamma, "gamma"), _f32(beta, "beta")
    N, Cc, H, W = x.shape
    y = torch.empty_like(x)
    mean = rstd = None
    if return_stats:
        mean = torch.empty(N * Cc, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
    r = None if residual is None else _f32(residual, "residual")
    check(_lib.lib().vst_instance_norm_f32(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _ptr(r), y.data_ptr(),
                                           _ptr(mean), _ptr(rstd), N, Cc, H * W, eps, act, _stream()),
          "vst_instance_norm_f32")
    return (y, mean, rstd) if return_stats else y


def maxpool2(x):
    x = _f32(x, "x")
    N, Cc, H, W = x.shape
    y = torch.empty((N, Cc, H // 2, W // 2), dtype=torch.float32, device=x.device)
    check(_lib.lib().vst_maxpool2_f32(x.data_ptr(), y.data_ptr(), N * Cc, H, W, _stream()), "vst_maxpool2_f32")
    return y


def pack_bgr_u8(img, out=None):
    """[N,3,H,W] fp32 RGB -> [N,H,W,3] uint8 BGR with the reference's clamp(0,255) + astype(uint8) truncation."""
    img = _f32(img, "img")
    N, Cc, H, W = img.shape
    if Cc != 3:
        raise _lib.VstError("pack_bgr_u8 expects 3 channels")
    if out is None:
        out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=img.device)
    check(_lib.lib().vst_pack_bgr_u8(img.data_ptr(), out.data_ptr(), N, H, W, _stream()), "vst_pack_bgr_u8")
    return out


def vgg_normalize(batch, inplace_div: bool):
    """(batch/255 - mean)/std; with inplace_div the ARGUMENT is divided by 255 (RC semantics)."""
    if batch.dtype != torch.float32 or not batch.is_contiguous():
        if inplace_div:
            raise _lib.VstError("vgg_normalize(in place) needs a contiguous float32 tensor")
        batch = batch.float().contiguous()
    N, Cc, H, W = batch.shape
    if Cc != 3:
        raise _lib.VstError("vgg_normalize expects 3 channels")
    y = torch.empty_like(batch)
    check(_lib.lib().vst_vgg_normalize_f32(batch.data_ptr(), y.data_ptr(), N, H * W, int(inplace_div), _stream()),
          "vst_vgg_normalize_f32")
    return y


def warp(x, flo, return_corners=False):
    x, flo = _f32(x, "x"), _f32(flo, "flo")
    B, Cc, H, W = x.shape
    if tuple(flo.shape) != (B, 2, H, W):
        raise _lib.VstError(f"warp: flow must be [{B},2,{H},{W}], got {tuple(flo.shape)}")
    out = torch.empty_like(x)
    corners = torch.empty((B, H, W, 2), dtype=torch.int32, device=x.device) if return_corners else None
    check(_lib.lib().vst_warp_f32(x.data_ptr(), flo.data_ptr(), out.data_ptr(), _ptr(corners), B, Cc, H, W, _stream()),
          "vst_warp_f32")
    return (out, corners) if return_corners else out


def flow_warp_mask(flo01, flo10, threshold=2.0):
    """[2,H,W] flows -> [H,W] mask (reference signature) or [B,2,H,W] -> [B,H,W]."""
    unbatched = flo01.dim() == 3
    if unbatched:
        flo01, flo10 = flo01.unsqueeze(0), flo10.unsqueeze(0)
    flo01, flo10 = _f32(flo01, "flo01"), _f32(flo10, "flo10")
    B, two, H, W = flo01.shape
    if two != 2 or flo10.shape != flo01.shape:
        raise _lib.VstError("flow_warp_mask: flows must be [B,2,H,W] and equal in shape")
    mask = torch.empty((B, H, W), dtype=torch.float32, device=flo01.device)
    check(_lib.lib().vst_flow_warp_mask_f32(flo01.data_ptr(), flo10.data_ptr(), mask.data_ptr(), B, H, W,
                                            float(threshold), _stream()), "vst_flow_warp_mask_f32")
    return mask[0] if unbatched else mask


def resize_bilinear(x, size_hw, chan_scale=None):
    """F.interpolate(x, size=size_hw, mode="bilinear", align_corners=False) for [B,C,H,W] fp32, times an optional
    per-channel factor (the SceneFlow flow resize, RC/datasets.py:116-134)."""
    x = _f32(x, "x")
    B, Cc, Hs, Ws = x.shape
    Hd, Wd = int(size_hw[0]), int(size_hw[1])
    out = torch.empty((B, Cc, Hd, Wd), dtype=torch.float32, device=x.device)
    cs = None
    if chan_scale is not None:
        cs = torch.as_tensor(chan_scale, dtype=torch.float32, device=x.device).contiguous()
        if cs.numel() != Cc:
            raise _lib.VstError("resize_bilinear: chan_scale needs one factor per channel")
    check(_lib.lib().vst_resize_bilinear_f32(x.data_ptr(), out.data_ptr(), B * Cc, Hs, Ws, Hd, Wd,
                                             None if cs is None else cs.data_ptr(), Cc, _stream()), "vst_resize_bilinear_f32")
    return out


def motion_mask_(mask, motion):
    """mask *= (motion == 0), in place (RC/datasets.py:137-143: motion[motion != 0] = 1; mask * (1 - motion))."""
    if mask.dtype != torch.float32 or not mask.is_contiguous():
        raise _lib.VstError("motion_mask_: mask must be contiguous float32 (updated in place)")
    motion = _f32(motion, "motion")
    if mask.shape != motion.shape:
        raise _lib.VstError("motion_mask_: shapes differ")
    check(_lib.lib().vst_motion_mask_f32(mask.data_ptr(), motion.data_ptr(), mask.numel(), _stream()), "vst_motion_mask_f32")
    return mask


def gram(y, scale: float):
    y = _f32(y, "y")
    B, Cc, H, W = y.shape
    out = torch.empty((B, Cc, Cc), dtype=torch.float32, device=y.device)
    check(_lib.lib().vst_gram_f32(y.data_ptr(), out.data_ptr(), B, Cc, H * W, float(scale), _stream()), "vst_gram_f32")
    return out


def _out(out, n, device):
    if out is None:
        return torch.empty(n, dtype=torch.float32, device=device)
    if out.dtype != torch.float32 or out.numel() != n or not out.is_contiguous():
        raise _lib.VstError(f"out: expected {n} contiguous float32 values")
    return out


def feature_temporal_sums(f1, f2, flow, mask, out=None) -> torch.Tensor:
    """-> device tensor [sum mask_f*(f2-warp(f1))^2, C*sum(mask_f)]."""
    f1, f2, flow, mask = _f32(f1), _f32(f2), _f32(flow), _f32(mask)
    B, Cc, Hf, Wf = f1.shape
    H, W = flow.shape[2:]
    out = _out(out, 2, f1.device)
    check(_lib.lib().vst_feature_temporal_f32(f1.data_ptr(), f2.data_ptr(), flow.data_ptr(), mask.data_ptr(),
                                              out.data_ptr(), reduce_scratch(f1.device).data_ptr(), B, Cc, Hf, Wf, H, W,
                                              _stream()), "vst_feature_temporal_f32")
    return out


def output_temporal_sums(s1, s2, i1, i2, flow, mask, luminance=True, out=None) -> torch.Tensor:
    s1, s2, flow, mask = _f32(s1), _f32(s2), _f32(flow), _f32(mask)
    B, Cc, H, W = s1.shape
    if Cc != 3:
        raise _lib.VstError("output_temporal: 3-channel images expected")
    if luminance:
        i1, i2 = _f32(i1), _f32(i2)
    out = _out(out, 2, s1.device)
    check(_lib.lib().vst_output_temporal_f32(s1.data_ptr(), s2.data_ptr(), _ptr(i1) if luminance else None,
                                             _ptr(i2) if luminance else None, flow.data_ptr(), mask.data_ptr(),
                                             out.data_ptr(), reduce_scratch(s1.device).data_ptr(), B, H, W,
                                             int(luminance), _stream()), "vst_output_temporal_f32")
    return out


def sqdiff_sum(a, b, out=None) -> torch.Tensor:
    a, b = _f32(a), _f32(b)
    if a.shape != b.shape:
        raise _lib.VstError("sqdiff_sum: shape mismatch")
    out = _out(out, 1, a.device)
    check(_lib.lib().vst_sqdiff_sum_f32(a.data_ptr(), b.data_ptr(), out.data_ptr(), reduce_scratch(a.device).data_ptr(),
                                        a.numel(), _stream()), "vst_sqdiff_sum_f32")
    return out


def frame_diff_sqsum(x0, x1, y0, y1, lo=0.0, hi=255.0, out=None) -> torch.Tensor:
    """sum ((x1 - x0) - (clamp(y1) - clamp(y0)))^2 (the numerator of RC/utilities.py calculate_mse)."""
    x0, x1, y0, y1 = _f32(x0), _f32(x1), _f32(y0), _f32(y1)
    if not (x0.shape == x1.shape == y0.shape == y1.shape):
        raise _lib.VstError("frame_diff_sqsum: shape mismatch")
    out = _out(out, 1, x0.device)
    check(_lib.lib().vst_frame_diff_sqsum_f32(x0.data_ptr(), x1.data_ptr(), y0.data_ptr(), y1.data_ptr(), float(lo), float(hi),
                                              out.data_ptr(), reduce_scratch(x0.device).data_ptr(), x0.numel(), _stream()),
          "vst_frame_diff_sqsum_f32")
    return out


def tv_sum(x, mode: int, out=None) -> torch.Tensor:
    x = _f32(x)
    B, Cc, H, W = x.shape
    out = _out(out, 1, x.device)
    check(_lib.lib().vst_tv_f32(x.data_ptr(), out.data_ptr(), reduce_scratch(x.device).data_ptr(), B * Cc, H, W, mode,
                                _stream()), "vst_tv_f32")
    return out


def tc_conv3x3(x, w, pad_mode=PAD_REFLECT):
    """3x3 stride-1 convolution through the tcgen05 tap-GEMM path (bf16 operands, fp32 accumulate)."""
    x, w = _f32(x), _f32(w)
    N, Cin, H, W = x.shape
    Cout = w.shape[0]
    L = _lib.lib()
    ws_bytes = L.vst_tc_conv_workspace_bytes(N, Cin, H, W, Cout, 3)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    y = torch.empty((N, Cout, H, W), dtype=torch.float32, device=x.device)
    check(L.vst_tc_conv3x3_f32io(x.data_ptr(), w.data_ptr(), y.data_ptr(), N, Cin, H, W, Cout, pad_mode, ws.data_ptr(),
                                 ws_bytes, _stream()), "vst_tc_conv3x3_f32io")
    return y


# ---------------------------------------------------------------------------------------------
# Adjoints: one function per backward entry point of include/vst_b200.h.
# ---------------------------------------------------------------------------------------------
def weight_flip_transpose(w):
    w = _f32(w, "w")
    Cout, Cin, k, _ = w.shape
    wt = torch.empty((Cin, Cout, k, k), dtype=torch.float32, device=w.device)
    check(_lib.lib().vst_weight_flip_transpose_f32(w.data_ptr(), wt.data_ptr(), Cout, Cin, k, _stream()),
          "vst_weight_flip_transpose_f32")
    return wt


def fold_pad(dxp, Hs, Ws, ups, pad, pad_mode):
    dxp = _f32(dxp, "dxp")
    N, Cc, Hp, Wp = dxp.shape
    dx = torch.empty((N, Cc, Hs, Ws), dtype=torch.float32, device=dxp.device)
    check(_lib.lib().vst_fold_pad_f32(dxp.data_ptr(), dx.data_ptr(), N * Cc, Hs, Ws, ups, pad, pad_mode, Hp, Wp, _stream()),
          "vst_fold_pad_f32")
    return dx


def conv2d_dgrad(dy, w, in_hw, stride=1, pad=0, pad_mode=PAD_REFLECT, ups=1):
    """Data gradient of `conv2d(x, w, stride, pad, pad_mode, ups)`; in_hw = (H, W) of x."""
    dy, w = _f32(dy, "dy"), _f32(w, "w")
    N, Cout, Ho, Wo = dy.shape
    _, Cin, k, _ = w.shape
    Hs, Ws = in_hw
    Hp, Wp = Hs * ups + 2 * pad, Ws * ups + 2 * pad
    if stride == 1:
        if pad_mode == PAD_ZERO and ups == 1:   # dx = corr(dy, flip(w)^T) with zero pad k-1-pad (VGG body)
            return conv2d(dy, weight_flip_transpose(w), None, 1, k - 1 - pad, PAD_ZERO)
        dxp = conv2d(dy, weight_flip_transpose(w), None, 1, k - 1, PAD_ZERO)
    else:
        dxp = torch.empty((N, Cin, Hp, Wp), dtype=torch.float32, device=dy.device)
        check(_lib.lib().vst_conv_transpose_gather_f32(dy.data_ptr(), w.data_ptr(), None, dxp.data_ptr(), N, Cout, Ho, Wo,
                                                       Cin, Hp, Wp, k, stride, 0, _stream()),
              "vst_conv_transpose_gather_f32")
    if pad == 0 and ups == 1:
        return dxp
    return fold_pad(dxp, Hs, Ws, ups, pad, pad_mode)


def conv2d_wgrad(x, dy, k, stride=1, pad=0, pad_mode=PAD_REFLECT, ups=1):
    x, dy = _f32(x, "x"), _f32(dy, "dy")
    N, Cin, H, W = x.shape
    Cout = dy.shape[1]
    dw = torch.empty((Cout, Cin, k, k), dtype=torch.float32, device=x.device)
    check(_lib.lib().vst_conv2d_wgrad_f32(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, Cin, H, W, Cout, k, stride, pad,
                                          pad_mode, ups, _stream()), "vst_conv2d_wgrad_f32")
    return dw


def channel_sum(x):
    x = _f32(x, "x")
    N, Cc, H, W = x.shape
    out = torch.empty(Cc, dtype=torch.float32, device=x.device)
    check(_lib.lib().vst_channel_sum_f32(x.data_ptr(), out.data_ptr(), N, Cc, H * W, _stream()), "vst_channel_sum_f32")
    return out


def act_bwd(dy, y, act, out=None):
    dy, y = _f32(dy, "dy"), _f32(y, "y")
    dz = torch.empty_like(dy) if out is None else out
    check(_lib.lib().vst_act_bwd_f32(dy.data_ptr(), y.data_ptr(), dz.data_ptr(), dy.numel(), act, _stream()), "vst_act_bwd_f32")
    return dz


def instance_norm_bwd(x, dy, gamma, beta, mean, rstd, act=ACT_NONE):
    """-> (dx, dgamma, dbeta); `x` is the norm's input, dy the gradient w.r.t. act(norm(x))."""
    x, dy = _f32(x, "x"), _f32(dy, "dy")
    N, Cc, H, W = x.shape
    dx = torch.empty_like(x)
    dg = torch.empty(Cc, dtype=torch.float32, device=x.device)
    db = torch.empty_like(dg)
    check(_lib.lib().vst_instance_norm_bwd_f32(x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                               mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                               N, Cc, H * W, act, _stream()), "vst_instance_norm_bwd_f32")
    return dx, dg, db


def maxpool2_bwd(x, dy):
    x, dy = _f32(x, "x"), _f32(dy, "dy")
    N, Cc, H, W = x.shape
    dx = torch.empty_like(x)
    check(_lib.lib().vst_maxpool2_bwd_f32(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), N * Cc, H, W, _stream()),
          "vst_maxpool2_bwd_f32")
    return dx


def vgg_normalize_bwd(dy):
    dy = _f32(dy, "dy")
    N, Cc, H, W = dy.shape
    dx = torch.empty_like(dy)
    check(_lib.lib().vst_vgg_normalize_bwd_f32(dy.data_ptr(), dx.data_ptr(), N, H * W, _stream()), "vst_vgg_normalize_bwd_f32")
    return dx


def warp_bwd(dy, flo):
    dy, flo = _f32(dy, "dy"), _f32(flo, "flo")
    B, Cc, H, W = dy.shape
    dx = torch.empty_like(dy)
    check(_lib.lib().vst_warp_bwd_f32(dy.data_ptr(), flo.data_ptr(), dx.data_ptr(), B, Cc, H, W, _stream()), "vst_warp_bwd_f32")
    return dx


def feature_temporal_bwd(f1, f2, flow, mask, scale: float, scale_dev=None):
    """-> (df1, df2) for scale * scale_dev * sum mask_f (f2 - warp(f1))^2."""
    f1, f2, flow, mask = _f32(f1), _f32(f2), _f32(flow), _f32(mask)
    B, Cc, Hf, Wf = f1.shape
    H, W = flow.shape[2:]
    df1, df2 = torch.empty_like(f1), torch.empty_like(f2)
    check(_lib.lib().vst_feature_temporal_bwd_f32(f1.data_ptr(), f2.data_ptr(), flow.data_ptr(), mask.data_ptr(), float(scale),
                                                  _ptr(scale_dev), df1.data_ptr(), df2.data_ptr(), B, Cc, Hf, Wf, H, W,
                                                  _stream()), "vst_feature_temporal_bwd_f32")
    return df1, df2


def output_temporal_bwd(s1, s2, i1, i2, flow, mask, scale: float, scale_dev=None, luminance=True):
    s1, s2, flow, mask = _f32(s1), _f32(s2), _f32(flow), _f32(mask)
    B, Cc, H, W = s1.shape
    if luminance:
        i1, i2 = _f32(i1), _f32(i2)
    ds1, ds2 = torch.empty_like(s1), torch.empty_like(s2)
    check(_lib.lib().vst_output_temporal_bwd_f32(s1.data_ptr(), s2.data_ptr(), _ptr(i1) if luminance else None,
                                                 _ptr(i2) if luminance else None, flow.data_ptr(), mask.data_ptr(),
                                                 float(scale), _ptr(scale_dev), ds1.data_ptr(), ds2.data_ptr(), B, H, W,
                                                 int(luminance), _stream()), "vst_output_temporal_bwd_f32")
    return ds1, ds2


def sqdiff_bwd(a, b, scale: float, want_db=False):
    a, b = _f32(a), _f32(b)
    da = torch.empty_like(a)
    db = torch.empty_like(a) if want_db else None
    check(_lib.lib().vst_sqdiff_bwd_f32(a.data_ptr(), b.data_ptr(), float(scale), da.data_ptr(), _ptr(db), a.numel(), _stream()),
          "vst_sqdiff_bwd_f32")
    return (da, db) if want_db else da


def tv_bwd(x, scale: float, mode: int):
    x = _f32(x)
    B, Cc, H, W = x.shape
    dx = torch.empty_like(x)
    check(_lib.lib().vst_tv_bwd_f32(x.data_ptr(), float(scale), dx.data_ptr(), B * Cc, H, W, mode, _stream()), "vst_tv_bwd_f32")
    return dx


def gram_bwd(y, dG, scale: float):
    y, dG = _f32(y), _f32(dG)
    B, Cc, H, W = y.shape
    dy = torch.empty_like(y)
    check(_lib.lib().vst_gram_bwd_f32(y.data_ptr(), dG.data_ptr(), dy.data_ptr(), B, Cc, H * W, float(scale), _stream()),
          "vst_gram_bwd_f32")
    return dy


def axpy_(y, x, alpha: float = 1.0):
    """y += alpha * x (in place)."""
    if x.shape != y.shape or not y.is_contiguous():
        raise _lib.VstError("axpy_: shape mismatch or non-contiguous destination")
    x = _f32(x)
    check(_lib.lib().vst_axpy_f32(x.data_ptr(), y.data_ptr(), float(alpha), y.numel(), _stream()), "vst_axpy_f32")
    return y


def loss_terms(sums, entries, n_groups: int):
    """entries: list of (num_idx, den_idx|-1, coef, den_eps, group) -> (terms[n_groups+2], scales[len(entries)]);
    terms[:n_groups] per group, terms[n_groups] the total, terms[n_groups+1] = 1 if a denominator was exactly zero."""
    import ctypes as C

    n = len(entries)
    ia = lambda vals: (C.c_int * n)(*[int(v) for v in vals])
    fa = lambda vals: (C.c_float * n)(*[float(v) for v in vals])
    terms = torch.empty(n_groups + 2, dtype=torch.float32, device=sums.device)
    scales = torch.empty(n, dtype=torch.float32, device=sums.device)
    check(_lib.lib().vst_loss_terms_f32(sums.data_ptr(), ia(e[0] for e in entries), ia(e[1] for e in entries),
                                        fa(e[2] for e in entries), fa(e[3] for e in entries), ia(e[4] for e in entries),
                                        n, n_groups, terms.data_ptr(), scales.data_ptr(), _stream()), "vst_loss_terms_f32")
    return terms, scales


def adam_(p, g, m, v, step: int, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0, skip_flag=None):
    """In-place Adam over (flat) fp32 buffers; `skip_flag` (device float, nullable): nonzero -> the launch updates nothing."""
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.VstError("adam_: contiguous float32 buffers expected")
    check(_lib.lib().vst_adam_f32(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, b1, b2, eps, int(step),
                                  float(grad_scale), _ptr(skip_flag), _stream()), "vst_adam_f32")
