"""Training step of the frame path with hand-written adjoints.

The reference's train scripts (RC/train_single/train_starry-night.py:58-152, RT/train.py:97-143) build
the loss from ATen ops and call `loss.backward()` + `optim.Adam.step()`.  Here the same step is an
explicit forward tape and an explicit reverse sweep over this repo's CUDA kernels: torch supplies
device memory, streams and the NCCL process group only - there is no autograd graph anywhere.

Layout of one step (`PairTrainer.step`):
  1. stylizer forward on cat(img1, img2)            (InstanceNorm is per sample, so batching is exact)
  2. loss sums (one reduction kernel per term)      -> `sums` device buffer
  3. `vst_loss_terms_f32`                           -> loss terms + the 1/count backward scales (device)
  4. reverse sweep                                  -> flat fp32 gradient buffer
  5. bucketed all-reduce of the flat buffer (NCCL, overlapped with the sweep) and fused Adam.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib, ops
from .vggcfg import VGG_LAYOUTS

REFLECT, ZERO = ops.PAD_REFLECT, ops.PAD_ZERO


# =============================================================================================
# flat parameter / gradient storage
# =============================================================================================
class FlatParams:
    """All trainable tensors of a module re-pointed into ONE flat fp32 buffer (registration order),
    with a matching flat gradient buffer: one Adam launch and a few large all-reduce buckets per step
    instead of 62 small ones.  `state_dict()` of the module is unchanged (same keys, same values)."""

    def __init__(self, module: torch.nn.Module):
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        if not named:
            raise _lib.VstError("FlatParams: module has no trainable parameters")
        dev = named[0][1].device
        total = sum(p.numel() for _, p in named)
        self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.names: List[str] = []
        self.offsets: Dict[str, tuple] = {}
        off = 0
        for n, p in named:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            self.names.append(n)
            self.offsets[n] = (off, k, tuple(p.shape))
            off += k
        self.total = total

    def grad_view(self, name: str) -> torch.Tensor:
        off, k, shape = self.offsets[name]
        return self.grad[off:off + k].view(shape)

    def param_view(self, name: str) -> torch.Tensor:
        off, k, shape = self.offsets[name]
        return self.flat[off:off + k].view(shape)


class GradSink:
    """Receives parameter gradients as the reverse sweep produces them (last layer first) and fires an
    asynchronous all-reduce for each bucket of the flat buffer as soon as every tensor in it has been
    written, so the exchange hides under the rest of the sweep."""

    def __init__(self, flat: FlatParams, process_group=None, n_buckets: int = 4):
        self.flat, self.pg = flat, process_group
        self.world = 1
        self.defer = False      # True: no exchange during the sweep, `exchange_all()` afterwards (VST_DP_OVERLAP=0 under graph replay)
        self.streams = None     # (sweep stream, weight-gradient stream) while a two-stream sweep is running
        if process_group is not None:
            import torch.distributed as dist

            self.world = dist.get_world_size(process_group)
        # contiguous buckets of ~equal size, boundaries on tensor boundaries
        target = flat.total / max(1, n_buckets)
        self.bucket_of: Dict[str, int] = {}
        self.ranges: List[List[int]] = []
        start, b = 0, 0
        for n in flat.names:
            off, k, _ = flat.offsets[n]
            if off - start >= target and len(self.ranges) < n_buckets - 1:
                self.ranges.append([start, off])
                start, b = off, b + 1
            self.bucket_of[n] = b
        self.ranges.append([start, flat.total])
        self.members = [sum(1 for n in flat.names if self.bucket_of[n] == i) for i in range(len(self.ranges))]
        self.reset()

    def reset(self):
        self.pending = list(self.members)
        self.works = []
        self.written = set()

    def put(self, name: str, g: torch.Tensor):
        if name in self.written:
            raise _lib.VstError(f"gradient of {name} produced twice")
        self.written.add(name)
        self.flat.grad_view(name).copy_(g.view(self.flat.offsets[name][2]))
        self._ready(name)

    def _ready(self, name: str):
        b = self.bucket_of[name]
        self.pending[b] -= 1
        if self.pending[b] == 0 and self.world > 1 and not self.defer:
            self._exchange(b)

    def _exchange(self, b: int):
        import torch.distributed as dist

        a, e = self.ranges[b]
        if self.streams is not None:
            # a bucket holds gradients written on BOTH streams of the sweep (InstanceNorm adjoints on the sweep stream, weight
            # gradients on the side stream): the collective is issued from the side stream after it has caught up with the
            # sweep stream, so it is ordered after every writer.  Event waits only - the whole thing is graph-capturable, and
            # under CUDA-graph replay the all-reduce becomes a node that runs beside the rest of the sweep.
            main, side = self.streams
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                w = dist.all_reduce(self.flat.grad[a:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        else:
            w = dist.all_reduce(self.flat.grad[a:e], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        self.works.append(w)

    def exchange_all(self):
        """Deferred mode: all buckets, last-written first, after the captured sweep has been replayed."""
        if self.world > 1:
            for b in reversed(range(len(self.ranges))):
                self._exchange(b)
            for w in self.works:
                w.wait()
            self.works = []

    def mark(self, name: str):
        """The gradient of `name` was written in place into `flat.grad_view(name)` by a kernel."""
        if name in self.written:
            raise _lib.VstError(f"gradient of {name} produced twice")
        self.written.add(name)
        self._ready(name)

    def finish(self) -> float:
        """Wait for the exchanges; returns the factor Adam must scale the summed gradient by."""
        missing = [n for n in self.flat.names if n not in self.written]
        if missing:
            raise _lib.VstError(f"reverse sweep produced no gradient for {missing[:4]}...")
        for w in self.works:
            w.wait()
        return 1.0 / self.world


# =============================================================================================
# fp32 stylizer graphs: forward tape + reverse sweep, layer by layer through the C-ABI
# =============================================================================================
class _ConvIN:
    """[ups] -> reflect pad -> conv(+bias) -> InstanceNorm(affine) -> act (+ residual)."""

    def __init__(self, prefix_conv: str, prefix_norm: str, conv: torch.nn.Conv2d, norm, k, stride, ups, act):
        self.wn, self.bn = prefix_conv + ".weight", prefix_conv + ".bias"
        self.gn, self.ben = prefix_norm + ".weight", prefix_norm + ".bias"
        self.conv, self.norm, self.k, self.stride, self.ups, self.act = conv, norm, k, stride, ups, act

    def fwd(self, x, residual=None):
        raw = ops.conv2d(x, self.conv.weight, self.conv.bias, self.stride, self.k // 2, REFLECT, self.ups)
        y, mean, rstd = ops.instance_norm(raw, self.norm.weight, self.norm.bias, residual=residual, act=self.act,
                                          return_stats=True)
        self.ctx = (x, raw, mean, rstd)
        return y

    def bwd(self, dy, sink: GradSink, need_dx=True):
        x, raw, mean, rstd = self.ctx
        self.ctx = None
        draw, dg, db = ops.instance_norm_bwd(raw, dy, self.norm.weight, self.norm.bias, mean, rstd, self.act)
        sink.put(self.ben, db)
        sink.put(self.gn, dg)
        sink.put(self.bn, ops.channel_sum(draw))
        sink.put(self.wn, ops.conv2d_wgrad(x, draw, self.k, self.stride, self.k // 2, REFLECT, self.ups))
        if not need_dx:
            return None
        return ops.conv2d_dgrad(draw, self.conv.weight, x.shape[2:], self.stride, self.k // 2, REFLECT, self.ups)


class _DeconvIN:
    """ConvTranspose2d(k3,s2,p1,op1) -> IN -> act (RT/network.py:48-60)."""

    def __init__(self, prefix: str, m):
        self.wn, self.bn, self.gn, self.ben = prefix + ".deconv.weight", prefix + ".deconv.bias", prefix + ".norm.weight", \
            prefix + ".norm.bias"
        self.m = m

    def fwd(self, x):
        m = self.m
        raw = ops.conv_transpose2d(x, m.deconv.weight, m.deconv.bias)
        y, mean, rstd = ops.instance_norm(raw, m.norm.weight, m.norm.bias, act=m._act, return_stats=True)
        self.ctx = (x, raw, mean, rstd)
        return y

    def bwd(self, dy, sink: GradSink):
        m = self.m
        x, raw, mean, rstd = self.ctx
        self.ctx = None
        draw, dg, db = ops.instance_norm_bwd(raw, dy, m.norm.weight, m.norm.bias, mean, rstd, m._act)
        sink.put(self.ben, db)
        sink.put(self.gn, dg)
        sink.put(self.bn, ops.channel_sum(draw))
        # dw[ci][co] = sum x[ci] * draw_pad[co] at stride 2: the wgrad of a stride-2 zero-padded conv with roles swapped
        sink.put(self.wn, ops.conv2d_wgrad(draw, x, 3, 2, 1, ZERO, 1))
        # dx = stride-2 zero-padded conv of draw with the same weight tensor read as [Cout'=Cin][Cin'=Cout]
        return ops.conv2d(draw, m.deconv.weight, None, 2, 1, ZERO)


class ReCoNetGraphFp32:
    """RC/network.py:171-190 (and the SD1/SD2 variants) with its reverse sweep."""

    def __init__(self, model):
        self.model = model
        o = model._order
        self.head = []
        for name in o[:3]:
            m = getattr(model, name)
            self.head.append(_ConvIN(f"{name}.conv2d", f"{name}.instance", m.conv2d, m.instance, m.kernel_size, m.stride, 1,
                                     ops.ACT_RELU))
        self.res = []
        for name in o[3:8]:
            m = getattr(model, name)
            a = _ConvIN(f"{name}.conv1.conv2d", f"{name}.in1", m.conv1.conv2d, m.in1, 3, 1, 1, ops.ACT_RELU)
            b = _ConvIN(f"{name}.conv2.conv2d", f"{name}.in2", m.conv2.conv2d, m.in2, 3, 1, 1, ops.ACT_NONE)
            self.res.append((a, b))
        self.up = []
        for name in o[8:10]:
            m = getattr(model, name)
            self.up.append(_ConvIN(f"{name}.conv2d", f"{name}.instance", m.conv2d, m.instance, m.kernel_size, m.stride,
                                   m.upsample or 1, ops.ACT_RELU))
        self.out_name = o[10]

    def forward(self, x):
        for l in self.head:
            x = l.fwd(x)
        self.taps = {"conv3": x}                       # the tensors the SD forwards return first (RC/network.py:218,265)
        for a, b in self.res:
            x = b.fwd(a.fwd(x), residual=x)
        features = x
        for i, l in enumerate(self.up):
            x = l.fwd(x)
            if i == 0:
                self.taps["deconv1"] = x
        m = getattr(self.model, self.out_name)
        img = ops.conv2d(x, m.conv2d.weight, m.conv2d.bias, 1, m.kernel_size // 2, REFLECT, 1, ops.ACT_RECONET_OUT)
        self.ctx = (x, img)
        return features, img

    def tap(self, name: str) -> torch.Tensor:
        """"conv3" / "deconv1" activation of the last forward, fp32 NCHW (distillation logging)."""
        return self.taps[name]

    def backward(self, d_features: Optional[torch.Tensor], d_img: torch.Tensor, sink: GradSink):
        x, img = self.ctx
        self.ctx = None
        m = getattr(self.model, self.out_name)
        k = m.kernel_size
        dz = ops.act_bwd(d_img, img, ops.ACT_RECONET_OUT)
        sink.put(f"{self.out_name}.conv2d.bias", ops.channel_sum(dz))
        sink.put(f"{self.out_name}.conv2d.weight", ops.conv2d_wgrad(x, dz, k, 1, k // 2, REFLECT, 1))
        d = ops.conv2d_dgrad(dz, m.conv2d.weight, x.shape[2:], 1, k // 2, REFLECT, 1)
        for l in reversed(self.up):
            d = l.bwd(d, sink)
        if d_features is not None:
            ops.axpy_(d, d_features)
        for a, b in reversed(self.res):
            d_in = a.bwd(b.bwd(d, sink), sink)
            d = ops.axpy_(d_in, d)           # residual fan-out: d x = d out + path through the two convs
        for i, l in enumerate(reversed(self.head)):
            d = l.bwd(d, sink, need_dx=(i < len(self.head) - 1))


class RtnstvGraphFp32:
    """RT/network.py:78-91 with its reverse sweep."""

    def __init__(self, model):
        self.model = model
        mk = lambda name, m: _ConvIN(f"{name}.conv", f"{name}.norm", m.conv, m.norm, m.kernel_size, m.stride, 1, m._act)
        self.head = [mk(f"conv{i}", getattr(model, f"conv{i}")) for i in (1, 2, 3)]
        self.res = []
        for i in range(1, 6):
            r = getattr(model, f"res{i}")
            self.res.append((mk(f"res{i}.conv1", r.conv1), mk(f"res{i}.conv2", r.conv2)))
        self.up = [_DeconvIN("deconv1", model.deconv1), _DeconvIN("deconv2", model.deconv2)]
        c4 = model.conv4
        self.out = _ConvIN("conv4.conv", "conv4.norm", c4.conv, c4.norm, 3, 1, 1, ops.ACT_RT_OUT)

    def forward(self, x):
        for l in self.head:
            x = l.fwd(x)
        for a, b in self.res:
            x = b.fwd(a.fwd(x), residual=x)
        for l in self.up:
            x = l.fwd(x)
        return None, self.out.fwd(x)

    def backward(self, d_features, d_img, sink: GradSink):
        d = self.out.bwd(d_img, sink)
        for l in reversed(self.up):
            d = l.bwd(d, sink)
        for a, b in reversed(self.res):
            d_in = a.bwd(b.bwd(d, sink), sink)
            d = ops.axpy_(d_in, d)
        for i, l in enumerate(reversed(self.head)):
            d = l.bwd(d, sink, need_dx=(i < len(self.head) - 1))


# =============================================================================================
# frozen VGG: forward taps + data-gradient-only reverse sweep
# =============================================================================================
class VggGraphFp32:
    def __init__(self, vgg):
        self.vgg = vgg
        self.layout = VGG_LAYOUTS[vgg.kind]["slices"]

    def forward(self, x, n_slices: Optional[int] = None, save: bool = True) -> List[torch.Tensor]:
        taps, tape = [], []
        for si, sl in enumerate(self.layout[:n_slices]):
            seq = getattr(self.vgg, f"slice{si + 1}")
            for idx, op in sl:
                if op[0] == "conv":
                    m = getattr(seq, str(idx))
                    y = ops.conv2d(x, m.weight, m.bias, 1, 1, ZERO, 1, ops.ACT_RELU)
                    tape.append(("conv", m.weight, y, tuple(x.shape[2:])))
                    x = y
                elif op[0] == "pool":
                    y = ops.maxpool2(x)
                    tape.append(("pool", x))
                    x = y
            taps.append(x)
            tape.append(("tap", si))
        if save:
            self.tape = tape
        return taps

    def backward(self, tap_grads: Sequence[Optional[torch.Tensor]]) -> torch.Tensor:
        """tap_grads[i] = dL/d tap_i (or None); returns dL/d input of the VGG body."""
        d = None
        for rec in reversed(self.tape):
            if rec[0] == "tap":
                g = tap_grads[rec[1]] if rec[1] < len(tap_grads) else None
                if g is not None:
                    d = g if d is None else ops.axpy_(d, g)
            elif d is None:
                continue
            elif rec[0] == "conv":
                _, w, y, in_hw = rec
                d = ops.conv2d_dgrad(ops.act_bwd(d, y, ops.ACT_RELU), w, in_hw, 1, 1, ZERO, 1)
            else:
                d = ops.maxpool2_bwd(rec[1], d)
        self.tape = None
        return d


class PerceptualFp32:
    """Content + style terms on the VGG taps and their gradient w.r.t. the (normalised) styled frames.
    RC: :126-138 (content relu3_3, Gram / C*H*W);  RT: RT/train.py:36-53 (content relu4_2, Gram / H*W)."""

    def __init__(self, vgg, content_tap: int, gram_div_c: bool, style_grams: List[torch.Tensor]):
        self.graph = VggGraphFp32(vgg)
        self.content_tap, self.gram_div_c = content_tap, gram_div_c
        self.style_grams = style_grams            # [1,C,C] each

    def gram_scale(self, f):
        _, c, h, w = f.shape
        return 1.0 / (c * h * w) if self.gram_div_c else 1.0 / (h * w)

    def forward(self, styled_in, content_in, sums: torch.Tensor, i_content: int, i_style0: int):
        """Writes sum (sf_c - cf_c)^2 to sums[i_content] and sum (G_k - Gs_k)^2 to sums[i_style0 + k]."""
        cf = self.graph.forward(content_in, n_slices=self.content_tap + 1, save=False)[self.content_tap]
        sf = self.graph.forward(styled_in)
        ops.sqdiff_sum(sf[self.content_tap], cf, out=sums[i_content:i_content + 1])
        grams = []
        for k, f in enumerate(sf):
            g = ops.gram(f, self.gram_scale(f))
            gs = self.style_grams[k].expand(g.shape[0], -1, -1).contiguous()
            ops.sqdiff_sum(g, gs, out=sums[i_style0 + k:i_style0 + k + 1])
            grams.append((g, gs))
        self.ctx = (sf, cf, grams)

    def backward(self, content_scale: float, style_scales: Sequence[float]) -> torch.Tensor:
        """content_scale = dL/d(sum sq diff) of the content term; style_scales[k] likewise per tap."""
        sf, cf, grams = self.ctx
        self.ctx = None
        tap_grads = []
        for k, f in enumerate(sf):
            g, gs = grams[k]
            d = ops.gram_bwd(f, ops.sqdiff_bwd(g, gs, style_scales[k]), self.gram_scale(f))
            if k == self.content_tap:
                ops.axpy_(d, ops.sqdiff_bwd(f, cf, content_scale))
            tap_grads.append(d)
        return self.graph.backward(tap_grads)


# =============================================================================================
# the step
# =============================================================================================
class LossTerms:
    """Device-resident loss terms of one step; reading them synchronises (like `.item()` in the
    reference's logging, RC/...starry-night.py:155-166)."""

    def __init__(self, names: Sequence[str], terms: torch.Tensor, sums: torch.Tensor, count_idx: Sequence[int],
                 strict_count: bool, logged: Optional[Dict[str, torch.Tensor]] = None):
        self.names, self.terms, self.sums, self.count_idx, self.strict = list(names), terms, sums, count_idx, strict_count
        self.logged = logged or {}      # device scalars that are reported but NOT part of `loss` (the reference's SDL)

    @property
    def bad_flag(self) -> torch.Tensor:
        """Device float: 1 when a strict denominator of this step was zero (Adam skips the update on it)."""
        return self.terms[-1:]

    def to_dict(self) -> Dict[str, float]:
        t = self.terms.cpu().tolist()
        if self.strict:
            s = self.sums.cpu().tolist()
            if t[-1] != 0 or any(s[i] == 0 for i in self.count_idx):
                # the reference computes `1 / non_zero_count` with a Python int (RC/...starry-night.py:105,122); the step that
                # produced these terms left the weights and the Adam moments untouched (device-side flag), as the
                # reference's exception before backward() does
                raise ZeroDivisionError("division by zero: the occlusion mask of this batch is empty")
        d = dict(zip(self.names, t[:-2]))
        d["loss"] = t[-2]
        for k, v in self.logged.items():
            d[k] = float(v)
        return d


class PairTrainer:
    """One optimisation step on a batch of frame pairs - the body of the reference's training loops.

    family "reconet": RC/train_single/train_starry-night.py:58-152 (five terms, VGG16, normalised-domain OTL/TV)
    family "rtnstv":  RT/train.py:97-143 + spatial_loss :36-60 (content/style/sqrt-TV per frame + temporal)
    """

    def __init__(self, model, vgg, style_img255: torch.Tensor, family: str = "reconet", lr: float = 1e-3,
                 alpha=None, beta=None, gamma=None, lambda_f=None, lambda_o=None, process_group=None, n_buckets: int = 4, precision: str = "fp32",
                 teacher=None):
        if family not in ("reconet", "rtnstv"):
            raise ValueError("family must be 'reconet' or 'rtnstv'")
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.VstError("PairTrainer needs the model on a CUDA device (no CPU fallback)")
        self.family, self.model, self.vgg, self.lr = family, model, vgg, lr
        rc = family == "reconet"
        # constants at the top of the reference scripts (RC/...starry-night.py:24-28, RT/train.py:29-32)
        self.alpha = (1e5 if rc else 1e7) if alpha is None else alpha
        self.beta = (1e11 if rc else 5e7) if beta is None else beta
        self.gamma = (1e-2 if rc else 5e-1) if gamma is None else gamma
        self.lambda_f = 1e12 if lambda_f is None else lambda_f
        self.lambda_o = (1e7 if rc else 1e6) if lambda_o is None else lambda_o
        # Teacher / student scripts (RC/train_single/train_Flow_SD{1,2}.py:42-48,82-86,155-162): a frozen teacher's first output
        # against the student's `sd` tensor gives the symmetric-distillation term SDL = 0.01 * BETA * (MSE_1 + MSE_2), which the
        # reference LOGS BUT NEVER ADDS to the loss - reproduced as such: reported, no gradient.
        self.teacher = teacher
        if teacher is not None:
            if not rc:
                raise ValueError("teacher= is a ReCoNet-family option (train_Flow_SD1 / train_Flow_SD2)")
            for p_ in teacher.parameters():
                p_.requires_grad = False
            self._sd_tap = {"ReCoNetSD1": "deconv1", "ReCoNetSD2": "conv3"}.get(type(model).__name__, "deconv1")
        self.flat = FlatParams(model)
        self.sink = GradSink(self.flat, process_group, n_buckets)
        self.m = torch.zeros_like(self.flat.flat)
        self.v = torch.zeros_like(self.flat.flat)
        self.t = 0
        sin = ops.vgg_normalize(style_img255.to(dev).float().contiguous(), inplace_div=False)
        # style Gram matrices once (RC/...starry-night.py:55-56, RT/train.py:92-93)
        if precision == "bf16":
            from .tc_graph import PerceptualTC

            # the tensor-core stylizer graph is built for the first batch's shape (it owns per-shape buffers)
            self.net = None
            self.perc = PerceptualTC(vgg, content_tap=2 if rc else 3, gram_div_c=rc, style_grams=[])
            self.perc.style_grams = self.perc.style_grams_from(sin)
        else:
            self.net = ReCoNetGraphFp32(model) if rc else RtnstvGraphFp32(model)
            self.perc = PerceptualFp32(vgg, content_tap=2 if rc else 3, gram_div_c=rc, style_grams=[])
            feats = self.perc.graph.forward(sin, save=False)
            self.perc.style_grams = [ops.gram(f, self.perc.gram_scale(f)) for f in feats]
        self.frame_index = (getattr(model, "input_frame_num", 1) - 1) * 3

    # ---- the reference's loss, term by term -----------------------------------------------------
    def _forward_losses(self, img1, img2, flow, mask):
        rc = self.family == "reconet"
        B = img1.shape[0]
        x = torch.cat((img1, img2), 0).contiguous()
        if self.precision == "bf16" and (self.net is None or (self.net.N, self.net.H, self.net.W) != (x.shape[0], x.shape[2], x.shape[3])):
            from .tc_graph import ReCoNetTC, RtnstvTC

            self.net = (ReCoNetTC if rc else RtnstvTC)(self.model, x.shape[0], x.shape[2], x.shape[3])
        i0 = self.frame_index
        frames = x[:, i0:i0 + 3].contiguous()
        con_n = ops.vgg_normalize(frames, inplace_div=False)
        # Off-critical-path work of the bf16 step runs on a second stream (fork / join by stream waits, capturable):
        # the content-tap VGG pass under the stylizer forward, the temporal / TV reductions under the styled VGG pass.
        aux = self._aux_streams()
        cf = None
        if aux:
            main, side = aux
            side.wait_stream(main)
            with torch.cuda.stream(side):
                cf = self.perc.content_features(con_n)
        feat, img = self.net.forward(x)
        logged = {}
        if self.teacher is not None:
            f_t = self.teacher(x)[0]                                # `feature_t, ... = teacher(img)` on both frames at once
            f_s = self.net.tap(self._sd_tap)
            if f_t.shape != f_s.shape:                              # what nn.MSELoss raises on these two tensors in the reference
                raise RuntimeError(f"The size of tensor a ({f_t.shape[1]}) must match the size of tensor b ({f_s.shape[1]}) "
                                   "at non-singleton dimension 1")
            logged["SDL"] = ops.sqdiff_sum(f_t, f_s) * (0.01 * self.beta / (f_t.numel() // 2))
        sty_n = ops.vgg_normalize(img, inplace_div=False)          # what `styled_img` holds after :81-82 (Q2)
        sums = torch.zeros(16, dtype=torch.float32, device=x.device)
        H, W = img.shape[2:]
        if aux:
            main.wait_stream(side)
            cf.t.record_stream(main)                              # allocated on `side`, consumed (and freed) on `main`
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._temporal_and_tv_sums(rc, B, feat, img, sty_n, con_n, flow, mask, sums)
            self.perc.forward(sty_n, con_n, sums, 4, 5, cf=cf)
            main.wait_stream(side)
        else:
            self.perc.forward(sty_n, con_n, sums, 4, 5)
            self._temporal_and_tv_sums(rc, B, feat, img, sty_n, con_n, flow, mask, sums)
        c_t = self.perc.ctx[0][self.perc.content_tap]
        n_content = (c_t if torch.is_tensor(c_t) else c_t.t).numel() // 2   # MSELoss(mean) over one frame batch
        if rc:
            entries = [(0, 1, self.lambda_f, 0.0, 0), (2, 3, self.lambda_o, 0.0, 1), (4, -1, self.alpha / n_content, 0.0, 2)]
            names = ["FTL", "OTL", "CL", "SL", "RL"]
            style_group, reg_group, reg_coef = 3, 4, self.gamma
        else:
            entries = [(2, 3, self.lambda_o, 1e-8, 3), (4, -1, self.alpha / n_content, 0.0, 0)]
            names = ["CL", "SL", "RL", "TL"]
            style_group, reg_group = 1, 2
            reg_coef = self.gamma / (B * 3 * (H - 1) * (W - 1))     # .mean() of each frame batch (RT/train.py:58)
        self.i_style = len(entries)
        for k, (g, _) in enumerate(self.perc.ctx[2]):
            entries.append((5 + k, -1, self.beta / (g.numel() // 2), 0.0, style_group))
        entries.append((9, -1, reg_coef, 0.0, reg_group))
        terms, scales = ops.loss_terms(sums, entries, len(names))
        self.ctx = (B, feat, img, sty_n, con_n, flow, mask, entries, scales)
        return LossTerms(names, terms, sums, (1, 3) if rc else (), rc, logged)

    def _aux_streams(self):
        """(current stream, second stream) for the bf16 step unless `VST_AUX_STREAM=0`; None on the fp32 path."""
        if self.precision != "bf16" or os.environ.get("VST_AUX_STREAM", "1") == "0":
            return None
        if getattr(self, "_aux_stream", None) is None:
            self._aux_stream = torch.cuda.Stream()
        return torch.cuda.current_stream(), self._aux_stream

    def _temporal_and_tv_sums(self, rc, B, feat, img, sty_n, con_n, flow, mask, sums):
        """Feature- / output-temporal and TV reductions (RC :91-123,141-145; RT/train.py:55-58,125-131) into `sums`."""
        if rc:
            ops.feature_temporal_sums(feat[:B], feat[B:], flow, mask, out=sums[0:2])
            ops.output_temporal_sums(sty_n[:B], sty_n[B:], con_n[:B], con_n[B:], flow, mask, True, out=sums[2:4])
            ops.tv_sum(sty_n, 0, out=sums[9:10])
        else:
            ops.output_temporal_sums(img[:B], img[B:], None, None, flow, mask, False, out=sums[2:4])
            ops.tv_sum(img, 1, out=sums[9:10])

    def _loss_adjoints(self, rc, B, feat, img, sty_n, con_n, flow, mask, entries, scales):
        """Adjoints of the temporal / TV terms: (d TV, d styled frame 1, d styled frame 2, d features or None)."""
        if rc:
            tvb = ops.tv_bwd(sty_n, self.gamma, 0)
            ds1, ds2 = ops.output_temporal_bwd(sty_n[:B], sty_n[B:], con_n[:B], con_n[B:], flow, mask, 1.0, scales[1:2], True)
            df1, df2 = ops.feature_temporal_bwd(feat[:B], feat[B:], flow, mask, 1.0, scales[0:1])
            return tvb, ds1, ds2, torch.cat((df1, df2), 0)
        tvb = ops.tv_bwd(img, entries[-1][2], 1)
        ds1, ds2 = ops.output_temporal_bwd(img[:B], img[B:], None, None, flow, mask, 1.0, scales[0:1], False)
        return tvb, ds1, ds2, None

    def _backward(self):
        rc = self.family == "reconet"
        B, feat, img, sty_n, con_n, flow, mask, entries, scales = self.ctx
        self.ctx = None
        self.sink.reset()
        content_scale = entries[2 if rc else 1][2]
        style_scales = [e[2] for e in entries[self.i_style:self.i_style + 4]]
        aux = self._aux_streams()
        if aux:                                                     # small HBM-bound adjoints under the VGG data-gradient sweep
            main, side = aux
            side.wait_stream(main)
            with torch.cuda.stream(side):
                adj = self._loss_adjoints(rc, B, feat, img, sty_n, con_n, flow, mask, entries, scales)
        d_in = self.perc.backward(content_scale, style_scales)      # d L / d normalised styled frames  [2B,3,H,W]
        if aux:
            main.wait_stream(side)
            for t in adj:
                if t is not None:
                    t.record_stream(main)                           # allocated on `side`, consumed (and freed) on `main`
        else:
            adj = self._loss_adjoints(rc, B, feat, img, sty_n, con_n, flow, mask, entries, scales)
        tvb, ds1, ds2, d_feat = adj
        if rc:
            ops.axpy_(d_in, tvb)
            ops.axpy_(d_in[:B], ds1)
            ops.axpy_(d_in[B:], ds2)
            d_img = ops.vgg_normalize_bwd(d_in)
        else:
            d_img = ops.vgg_normalize_bwd(d_in)
            ops.axpy_(d_img, tvb)
            ops.axpy_(d_img[:B], ds1)
            ops.axpy_(d_img[B:], ds2)
        self.net.backward(d_feat, d_img, self.sink)

    def forward_backward(self, img1, img2, flow, mask) -> LossTerms:
        """Loss terms + gradients in `self.flat.grad` (averaged over ranks once `finish` ran); no update."""
        terms = self._forward_losses(img1.float().contiguous(), img2.float().contiguous(), flow.float().contiguous(),
                                     mask.float().contiguous())
        self._backward()
        self._gscale = self.sink.finish()
        return terms

    def enable_cuda_graph(self, enable: bool = True):
        """Replay forward + reverse sweep as ONE CUDA graph (the step is ~600 small launches; eager issue is
        host-bound).  Inputs are copied into static buffers; the gradient exchange and Adam run after the replay."""
        self._use_graph = enable
        self._graph = None
        return self

    def _graph_forward_backward(self, img1, img2, flow, mask) -> LossTerms:
        args = [t.float().contiguous() for t in (img1, img2, flow, mask)]
        if self._graph is None or any(a.shape != s.shape for a, s in zip(args, self._static_in)):
            self._static_in = [a.clone() for a in args]
            # the bucket all-reduces are captured INSIDE the graph (forked from the sweep as each bucket completes, joined
            # before the graph ends) unless VST_DP_OVERLAP=0, which restores the exchange after the replay
            self.sink.defer = os.environ.get("VST_DP_OVERLAP", "1") == "0"
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                   # warm-up outside the capture (lazy buffers, func attributes)
                self._forward_losses(*self._static_in)
                self._backward()
                self.sink.finish()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._static_terms = self._forward_losses(*self._static_in)
                self._backward()
                self.sink.finish()
        for s, a in zip(self._static_in, args):
            s.copy_(a, non_blocking=True)
        self._graph.replay()
        if self.sink.defer:
            self.sink.exchange_all()
        self._gscale = 1.0 / self.sink.world
        return self._static_terms

    def step(self, img1, img2, flow, mask) -> LossTerms:
        """forward + backward + Adam (RC/...starry-night.py:148-152)."""
        if getattr(self, "_use_graph", False):
            terms = self._graph_forward_backward(img1, img2, flow, mask)
        else:
            terms = self.forward_backward(img1, img2, flow, mask)
        self.t += 1
        ops.adam_(self.flat.flat, self.flat.grad, self.m, self.v, self.t, lr=self.lr, grad_scale=self._gscale,
                  skip_flag=terms.bad_flag if terms.strict else None)
        # parameters were written through raw pointers (no torch version bump): invalidate cached inference plans
        self.model._weights_generation = getattr(self.model, "_weights_generation", 0) + 1
        return terms

    def exchange_check(self, img1, img2, flow, mask) -> Dict[str, object]:
        """Data-parallel self-check (no update): one sweep with the exchange held back, the LOCAL flat gradients of all ranks
        all-gathered, then the bucketed exchange on the same buffer.  Returns every rank's gradient norm, the error of the
        exchanged average against the mean of the gathered per-rank gradients, and (under CUDA-graph replay) the captured,
        overlapped exchange against the same mean - that one carries the run-to-run floor of the split-K atomics in the
        weight-gradient GEMMs, which `graph_repeat_rel` (the captured step run twice) measures."""
        import torch.distributed as dist

        sink = self.sink
        if sink.world < 2:
            raise _lib.VstError("exchange_check needs a process group with world_size > 1")
        args = [t.float().contiguous() for t in (img1, img2, flow, mask)]
        saved = sink.defer
        sink.defer = True
        try:
            terms = self._forward_losses(*args)
            self._backward()
            sink.finish()
        finally:
            sink.defer = saved
        local = self.flat.grad.clone()
        gathered = [torch.empty_like(local) for _ in range(sink.world)]
        dist.all_gather(gathered, local, group=sink.pg)
        mean = torch.stack([g.double() for g in gathered]).mean(0)
        sink.exchange_all()
        red = self.flat.grad.double() / sink.world
        out = {"grad_norm_per_rank": [float(g.double().norm()) for g in gathered],
               "allreduce_vs_mean_rel": float((red - mean).norm() / mean.norm()),
               "loss_terms": terms.to_dict()}
        if getattr(self, "_use_graph", False):
            self._graph_forward_backward(*args)
            g1 = self.flat.grad.double() * self._gscale
            self._graph_forward_backward(*args)
            g2 = self.flat.grad.double() * self._gscale
            out["graph_overlapped_vs_mean_rel"] = float((g1 - mean).norm() / mean.norm())
            out["graph_repeat_rel"] = float((g2 - g1).norm() / g1.norm())
            out["overlapped"] = not sink.defer
        return out

    def grads(self) -> Dict[str, torch.Tensor]:
        s = getattr(self, "_gscale", 1.0)
        return {n: self.flat.grad_view(n) * s for n in self.flat.names}
