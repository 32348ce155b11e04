"""bf16 tensor-core graphs of the training step: forward tapes with saved channels-last activations and
explicit reverse sweeps built from `vst_b200.tc` primitives.

  VggTC          frozen VGG16/19 prefix: forward taps, data-gradient-only sweep   
  PerceptualTC   content + style (Gram) terms and their gradient                   (a19/a20, B6/B7)
  ReCoNetTC      the stylizer: forward with saved raw/act tensors, dgrad + wgrad   (a1-a7, B10-B15)

Interfaces mirror the fp32 graphs in train_core.py so `PairTrainer` can switch on `precision`.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops, tc
from ._lib import ActDesc, TapGemmDesc, check
from .tc import Act, BF16, ConvTC, REFLECT, REPLICATE, ZERO
from .vggcfg import VGG_LAYOUTS


# =============================================================================================
# VGG
# =============================================================================================
class VggTC:
    def __init__(self, vgg):
        self.vgg = vgg
        self.layout = VGG_LAYOUTS[vgg.kind]["slices"]
        dev = next(vgg.parameters()).device
        self.convs = {}
        first = True
        for si, sl in enumerate(self.layout):
            seq = getattr(vgg, f"slice{si + 1}")
            for idx, op in sl:
                if op[0] == "conv":
                    m = getattr(seq, str(idx))
                    c = ConvTC("vgg27" if first else "vgg", op[1], op[2], dev, need_dgrad=True, need_wgrad=False)
                    c.pack(m.weight.detach().contiguous())          # frozen: packed once
                    c.bias = m.bias.detach().float().contiguous()
                    c.first = first
                    first = False
                    self.convs[(si, idx)] = c

    def forward(self, x_nchw: torch.Tensor, n_slices: Optional[int] = None, save: bool = True) -> List[Act]:
        N, _, H, W = x_nchw.shape
        x = tc.prologue_x27(x_nchw.float())
        taps, tape = [], []
        for si, sl in enumerate(self.layout[:n_slices]):
            for idx, op in sl:
                if op[0] == "conv":
                    c = self.convs[(si, idx)]
                    y = Act(N, x.H, x.W, c.cout, device=x.t.device)
                    c.forward(x, y.t, (x.H, x.W), bias=c.bias, relu=True)
                    tape.append(["conv", c, y, (x.H, x.W)])
                    x = y
                elif op[0] == "pool":
                    y = tc.maxpool2(x)
                    tape.append(["pool"])
                    x = y
            taps.append(x)
            tape.append(["tap", si])
        if save:
            self.tape = tape
        return taps

    def backward(self, tap_grads: Sequence[Optional[torch.Tensor]]) -> torch.Tensor:
        """tap_grads[i]: bf16 gradient w.r.t. tap i ([N,H,W,C] flat) or None -> fp32 NCHW gradient w.r.t. the input."""
        g = None            # gradient flowing down (flat bf16), w.r.t. the tensor *after* the current position
        pooled = False      # g is at pooled resolution relative to the next conv output going down
        add = None
        out = None
        for rec in reversed(self.tape):
            if rec[0] == "tap":
                add = tap_grads[rec[1]] if rec[1] < len(tap_grads) else None
            elif rec[0] == "pool":
                pooled = g is not None
            else:
                _, c, y, in_hw = rec
                if g is None and add is None:
                    continue
                if g is None:
                    gm = tc.relu_pool_bwd(add, y, None, False)
                else:
                    gm = tc.relu_pool_bwd(g, y, add, pooled)
                add, pooled = None, False
                if c.first:
                    out = torch.empty((y.N, 3, in_hw[0], in_hw[1]), dtype=torch.float32, device=y.t.device)
                    c.dgrad(gm, in_hw, out_f32_nchw=out)
                else:
                    g = c.dgrad(gm, in_hw)
        self.tape = None
        return out


class PerceptualTC:
    """Same contract as train_core.PerceptualFp32, on the tensor-core VGG."""

    def __init__(self, vgg, content_tap: int, gram_div_c: bool, style_grams: List[torch.Tensor]):
        self.graph = VggTC(vgg)
        self.content_tap, self.gram_div_c = content_tap, gram_div_c
        self.style_grams = style_grams

    def gram_scale(self, f) -> float:
        if isinstance(f, Act):
            c, h, w = f.C, f.H, f.W
        else:
            _, c, h, w = f.shape
        return 1.0 / (c * h * w) if self.gram_div_c else 1.0 / (h * w)

    def style_grams_from(self, style_norm: torch.Tensor) -> List[torch.Tensor]:
        feats = self.graph.forward(style_norm, save=False)
        return [tc.gram(f, self.gram_scale(f)) for f in feats]

    def content_features(self, content_in):
        """The content tap of the un-styled frames (no gradient needed): independent of the stylizer, so the step may run
        it on a second stream under the stylizer forward and hand it to `forward(cf=...)`."""
        return self.graph.forward(content_in, n_slices=self.content_tap + 1, save=False)[self.content_tap]

    def forward(self, styled_in, content_in, sums: torch.Tensor, i_content: int, i_style0: int, cf=None):
        if cf is None:
            cf = self.content_features(content_in)
        sf = self.graph.forward(styled_in)
        tc.sqdiff_sum(sf[self.content_tap].t, cf.t, sums[i_content:i_content + 1])
        grams = []
        for k, f in enumerate(sf):
            g = tc.gram(f, self.gram_scale(f))
            gs = self.style_grams[k]
            gse = gs.expand(g.shape[0], -1, -1).contiguous()
            ops.sqdiff_sum(g, gse, out=sums[i_style0 + k:i_style0 + k + 1])
            grams.append((g, gs))
        self.ctx = (sf, cf, grams)

    def backward(self, content_scale: float, style_scales: Sequence[float]) -> torch.Tensor:
        sf, cf, grams = self.ctx
        self.ctx = None
        tap_grads = []
        for k, f in enumerate(sf):
            g, gs = grams[k]
            d = tc.gram_bwd(f, g, gs, 2.0 * style_scales[k] * self.gram_scale(f))
            if k == self.content_tap:
                tc.add_(d, tc.sqdiff_bwd(f.t, cf.t, content_scale))
            tap_grads.append(d)
        return self.graph.backward(tap_grads)


class _SideWgrad:
    """Weight gradients off the critical path.  In the reverse sweep only the data gradients and the InstanceNorm adjoints
    feed the next stage; a stage's weight gradient (a smem-bound pixel-contraction GEMM at about half of the tensor peak)
    has no consumer before Adam.  Unless `VST_WGRAD_STREAM=0` they are launched on a second stream - forked from the sweep
    by an event after the stage's IN adjoint, joined once at the end - so they fill the SMs under the HBM-bound IN / ReLU
    adjoint kernels of the following stages.  Fork/join through events is capturable, so the CUDA-graph replay keeps the
    two branches.  Operands are kept alive (Python references) until the join.  Single rank: the bucket marks are issued
    after the join.  Data-parallel: each weight gradient is marked as it is enqueued, so a bucket's all-reduce is forked (from
    the side stream, after an event wait on the sweep stream) the moment its last member is in flight - also inside the
    captured graph, where it becomes an NCCL node beside the rest of the sweep."""

    def __init__(self, net, sink):
        import os

        self.sink, self.on = sink, os.environ.get("VST_WGRAD_STREAM", "1") != "0"
        self.keep, self.late = [], []
        # data-parallel sweep with overlapped exchange: a weight gradient is marked the moment it is ENQUEUED on the side
        # stream, and the bucket's all-reduce is issued from that stream (GradSink._exchange orders it after both streams)
        self.mark_now = sink.world > 1 and not sink.defer
        if self.on:
            if getattr(net, "_wgrad_stream", None) is None:
                net._wgrad_stream = torch.cuda.Stream(net.dev)
            self.side = net._wgrad_stream
            self.main = torch.cuda.current_stream(net.dev)
            sink.streams = (self.main, self.side) if self.mark_now else None

    def run(self, fn, name: str, *operands):
        if not self.on:
            fn()
            self.sink.mark(name)
            return
        ev = torch.cuda.Event()
        ev.record(self.main)
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            fn()
        self.keep.extend(operands)
        if self.mark_now:
            self.sink.mark(name)
        else:
            self.late.append(name)

    def join(self):
        if self.on:
            self.main.wait_stream(self.side)
            for n in self.late:
                self.sink.mark(n)
            self.keep.clear()
            self.late.clear()
            self.sink.streams = None


def _sweep_stages(net, G, g_desc, fskip, sink, side: Optional[_SideWgrad] = None):
    """Reverse sweep over the 15 conv -> InstanceNorm (-> ReLU) (+ residual) stages shared by the ReCoNet and RTNSTV
    stylizers: per stage halo fold + IN backward (reduce, apply) -> weight gradient (pcgemm) -> data gradient (tapgemm).
    G / g_desc: the gradient over the padded input domain of the stage AFTER stage 14; fskip: optional extra gradient on
    the trunk output (ReCoNet's feature-temporal term)."""
    N, L, flat = net.N, net.layers, sink.flat
    skip = None
    side = side or _SideWgrad(net, sink)
    for i in range(14, -1, -1):
        l = L[i]
        conv = l.conv
        Ho, Wo = l.out_hw
        cout = conv.cout
        draw = Act(N, Ho, Wo, cout, 0, ZERO, 1 if conv.kind in ("up2", "tconv") else 0, net.dev)
        is_block_out = 4 <= i <= 12 and (i - 3) % 2 == 1
        sk, gsum = None, None
        if is_block_out:
            sk = fskip if i == 12 else skip
            gsum = torch.empty(N * Ho * Wo * cout, dtype=BF16, device=net.dev)
        elif i == 2:
            sk = skip                                                    # conv3 output feeds res1 and its skip path
        stc = net._stats_view(i, cout)
        tc.in_bwd(G, g_desc, net.raws[i], stc, l.gamma, l.beta, draw, l.relu, net.red, flat.grad_view(l.gn), flat.grad_view(l.ben),
                  skip=sk, gsum=gsum)
        sink.mark(l.gn)
        sink.mark(l.ben)
        if is_block_out:
            skip = gsum
        flat.grad_view(l.bn).zero_()                                     # bias in front of IN: gradient is exactly cancelled (Q6)
        sink.mark(l.bn)
        x_in = net.x_first if i == 0 else net.acts[i - 1]
        side.run(lambda: conv.wgrad(draw, x_in, l.out_hw, flat.grad_view(l.wn)), l.wn, draw)
        if i == 0:
            break
        G = conv.dgrad(draw, (x_in.H, x_in.W))
        g_desc = ActDesc(x_in.H, x_in.W, x_in.C, 0 if conv.kind == "tconv" else 1, x_in.kind, 0)
    side.join()


# =============================================================================================
# ReCoNet
# =============================================================================================
class _Layer:
    """conv -> InstanceNorm -> (ReLU) [+ residual]: buffers and parameters of one stage."""

    def __init__(self, name_conv, name_norm, conv: ConvTC, weight, gamma, beta, relu, out_hw):
        self.wn, self.bn, self.gn, self.ben = name_conv + ".weight", name_conv + ".bias", name_norm + ".weight", name_norm + ".bias"
        self.conv, self.weight, self.gamma, self.beta, self.relu, self.out_hw = conv, weight, gamma, beta, relu, out_hw


class ReCoNetTC:
    """RC/network.py:171-190 on the tensor-core path with everything the reverse sweep needs kept in HBM."""

    def __init__(self, model, N: int, H: int, W: int):
        if H % 4 or W % 4:
            raise _lib.VstError("ReCoNetTC: H and W must be multiples of 4 (SURVEY.md Q14)")
        self.model, self.N, self.H, self.W = model, N, H, W
        dev = next(model.parameters()).device
        self.dev = dev
        o = model._order
        c1, c2, c3, d1, d2 = model._widths
        cin = 3 * model.input_frame_num
        self.cin = cin
        H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
        mods = [getattr(model, n) for n in o]
        L: List[_Layer] = []
        mk = lambda nm, m_conv, m_norm, conv, relu, hw: L.append(_Layer(nm[0], nm[1], conv, m_conv.weight, m_norm.weight, m_norm.bias, relu, hw))
        mk((f"{o[0]}.conv2d", f"{o[0]}.instance"), mods[0].conv2d, mods[0].instance, ConvTC("row9", cin, c1, dev, need_dgrad=False), True, (H, W))
        mk((f"{o[1]}.conv2d", f"{o[1]}.instance"), mods[1].conv2d, mods[1].instance, ConvTC("s2", c1, c2, dev), True, (H2, W2))
        mk((f"{o[2]}.conv2d", f"{o[2]}.instance"), mods[2].conv2d, mods[2].instance, ConvTC("s2", c2, c3, dev), True, (H4, W4))
        for i in range(3, 8):
            r = mods[i]
            mk((f"{o[i]}.conv1.conv2d", f"{o[i]}.in1"), r.conv1.conv2d, r.in1, ConvTC("s1", c3, c3, dev), True, (H4, W4))
            mk((f"{o[i]}.conv2.conv2d", f"{o[i]}.in2"), r.conv2.conv2d, r.in2, ConvTC("s1", c3, c3, dev), False, (H4, W4))
        mk((f"{o[8]}.conv2d", f"{o[8]}.instance"), mods[8].conv2d, mods[8].instance, ConvTC("up2", c3, d1, dev), True, (H2, W2))
        mk((f"{o[9]}.conv2d", f"{o[9]}.instance"), mods[9].conv2d, mods[9].instance, ConvTC("up2", d1, d2, dev), True, (H, W))
        self.layers = L
        self.out_name = o[10]
        self.out_mod = mods[10]
        # ---- activations written by each stage, in the layout its consumer reads
        A = lambda h, w, c, pad, kind, par=0: Act(N, h, w, c, pad, kind, par, dev)
        self.acts = [A(H, W, c1, 1, REFLECT, 1), A(H2, W2, c2, 1, REFLECT, 1), A(H4, W4, c3, 1, REFLECT)]
        for i in range(5):
            self.acts.append(A(H4, W4, c3, 1, REFLECT))                                  # res.conv1 output
            self.acts.append(A(H4, W4, c3, 1, REFLECT if i < 4 else REPLICATE))          # block output
        self.acts += [A(H2, W2, d1, 1, REPLICATE), A(H, W, d2, 4, REFLECT)]
        chans = [c1, c2, c3] + [c3] * 10 + [d1, d2]
        self.raws = [torch.empty(N * l.out_hw[0] * l.out_hw[1] * c, dtype=BF16, device=dev) for l, c in zip(L, chans)]
        self.stats = torch.zeros((15, N, 256, 2), dtype=torch.float64, device=dev)
        self.red = torch.zeros(N * 256 * 2, dtype=torch.float32, device=dev)
        # deconv3 (k9, Cout 3): row convolution on the tensor cores, forward and both adjoints
        self.out_conv = tc.RowConvOutTC(d2, 3, mods[10].kernel_size, dev)
        self.d2 = d2
        self._packer = None
        self._wd_pool = tc.pool_wgrad_accumulators([l.conv for l in L] + [self.out_conv])

    # ---- forward ---------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor):
        """x fp32 NCHW [N, 3n, H, W] -> (features fp32 NCHW, img fp32 NCHW)."""
        N, H, W = self.N, self.H, self.W
        L = self.layers
        self.stats.zero_()
        if self._packer is None:
            self._packer = tc.MergedPack([(l.conv, l.weight) for l in L] + [(self.out_conv, self.out_mod.conv2d.weight)])
        self._packer.run()
        self.x_first = tc.prologue_x9(x.contiguous(), L[0].conv.KR)
        cur = self.x_first
        for i, l in enumerate(L):
            stc = self._stats_view(i, l.conv.cout)
            l.conv.forward(cur, self.raws[i], l.out_hw, stats=stc)
            res = None
            if 3 <= i <= 12 and (i - 3) % 2 == 1:       # second conv of a residual block: + block input
                res = self.acts[i - 2]
            tc.in_apply(self.raws[i], stc, l.gamma, l.beta, self.acts[i], l.relu, residual=res)
            cur = self.acts[i]
        self.feat_act = self.acts[12]
        features = self.feat_act.to_nchw()
        # deconv3
        m = self.out_mod
        img = torch.empty((N, 3, H, W), dtype=torch.float32, device=self.dev)
        self.out_conv.forward(self.acts[14], img, m.conv2d.bias, ops.ACT_RECONET_OUT)
        self.img = img
        return features, img

    def tap(self, name: str) -> torch.Tensor:
        """"conv3" / "deconv1" activation of the last forward as fp32 NCHW (distillation logging, train_core.PairTrainer)."""
        return self.acts[{"conv3": 2, "deconv1": 13}[name]].to_nchw()

    def _stats_view(self, i: int, cout: int) -> torch.Tensor:
        """[N][cout][2] contiguous block inside the stats slab of stage i."""
        flat = self.stats[i].reshape(-1)
        return flat[: self.N * cout * 2]

    # ---- reverse sweep ------------------------------------------------------------------------------
    def backward(self, d_features: Optional[torch.Tensor], d_img: torch.Tensor, sink):
        """d_features fp32 NCHW (or None), d_img fp32 NCHW -> parameter gradients into `sink`."""
        N, H, W = self.N, self.H, self.W
        L = self.layers
        flat = sink.flat
        m = self.out_mod
        k = m.kernel_size
        self._wd_pool.zero_()
        # ---- deconv3 (ConvTanh): tanh adjoint in fp32, then the row-convolution adjoints over E
        dz = ops.act_bwd(d_img, self.img, ops.ACT_RECONET_OUT)
        sink.put(f"{self.out_name}.conv2d.bias", ops.channel_sum(dz))
        E = self.out_conv.expand(dz)
        side = _SideWgrad(self, sink)
        wname = f"{self.out_name}.conv2d.weight"
        side.run(lambda: self.out_conv.wgrad(E, self.acts[14], flat.grad_view(wname)), wname, E)
        G = self.out_conv.dgrad(E)
        g_desc = ActDesc(H, W, self.d2, k // 2, REFLECT, 0)
        if d_features is not None:
            c3 = self.layers[12].conv.cout
            fskip = Act(N, H // 4, W // 4, c3, device=self.dev).from_nchw(d_features).t
        else:
            fskip = None
        _sweep_stages(self, G, g_desc, fskip, sink, side)


# =============================================================================================
# RTNSTV
# =============================================================================================
class RtnstvTC:
    """RT/network.py:63-91 on the tensor-core path: every conv (3x3, stride 1 / 2, ConvTranspose2d k3 s2) is a tap-GEMM,
    every InstanceNorm an `apply` pass, with the saved tensors of the reverse sweep kept in HBM.  The 16/32/48-channel
    layers are HBM/issue-bound rather than tensor-bound, but they share the ReCoNet kernels and sweep.
    The last stage (conv4 -> IN -> tanh map on 3 channels) keeps its InstanceNorm on the fp32 kernels."""

    def __init__(self, model, N: int, H: int, W: int):
        if H % 4 or W % 4:
            raise _lib.VstError("RtnstvTC: H and W must be multiples of 4")
        self.model, self.N, self.H, self.W = model, N, H, W
        dev = next(model.parameters()).device
        self.dev = dev
        H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
        L: List[_Layer] = []

        def mk(name, m, conv, hw, norm_attr="norm", conv_attr="conv"):
            cm, nm = getattr(m, conv_attr), getattr(m, norm_attr)
            L.append(_Layer(f"{name}.{conv_attr}", f"{name}.{norm_attr}", conv, cm.weight, nm.weight, nm.bias, m._act == ops.ACT_RELU, hw))

        mk("conv1", model.conv1, ConvTC("s1", 3, 16, dev, cin_pad=16, need_dgrad=False), (H, W))
        mk("conv2", model.conv2, ConvTC("s2", 16, 32, dev), (H2, W2))
        mk("conv3", model.conv3, ConvTC("s2", 32, 48, dev), (H4, W4))
        for i in range(1, 6):
            r = getattr(model, f"res{i}")
            mk(f"res{i}.conv1", r.conv1, ConvTC("s1", 48, 48, dev), (H4, W4))
            mk(f"res{i}.conv2", r.conv2, ConvTC("s1", 48, 48, dev), (H4, W4))
        mk("deconv1", model.deconv1, ConvTC("tconv", 48, 32, dev), (H2, W2), conv_attr="deconv")
        mk("deconv2", model.deconv2, ConvTC("tconv", 32, 16, dev), (H, W), conv_attr="deconv")
        self.layers = L
        A = lambda h, w, c, pad, kind, par=0: Act(N, h, w, c, pad, kind, par, dev)
        self.acts = [A(H, W, 16, 1, REFLECT, 1), A(H2, W2, 32, 1, REFLECT, 1), A(H4, W4, 48, 1, REFLECT)]
        for i in range(5):
            self.acts.append(A(H4, W4, 48, 1, REFLECT))
            self.acts.append(A(H4, W4, 48, 1, REFLECT) if i < 4 else A(H4, W4, 48, 0, ZERO))     # -> deconv1 reads it unpadded
        self.acts += [A(H2, W2, 32, 0, ZERO), A(H, W, 16, 1, REFLECT)]
        chans = [16, 32, 48] + [48] * 10 + [32, 16]
        self.raws = [torch.empty(N * l.out_hw[0] * l.out_hw[1] * c, dtype=BF16, device=dev) for l, c in zip(L, chans)]
        self.stats = torch.zeros((15, N, 256, 2), dtype=torch.float64, device=dev)
        self.red = torch.zeros(N * 256 * 2, dtype=torch.float32, device=dev)
        self.x_act = A(H, W, 16, 1, REFLECT)
        self.out_conv = ConvTC("s1", 16, 3, dev)
        self.raw_out = Act(N, H, W, 16, device=dev)
        self._packer = None
        self._wd_pool = tc.pool_wgrad_accumulators([l.conv for l in L] + [self.out_conv])

    def _stats_view(self, i: int, cout: int) -> torch.Tensor:
        return self.stats[i].reshape(-1)[: self.N * cout * 2]

    def forward(self, x: torch.Tensor):
        """x fp32 NCHW [N,3,H,W] in 0..255 -> (None, img fp32 NCHW)."""
        L = self.layers
        self.stats.zero_()
        if self._packer is None:
            self._packer = tc.MergedPack([(l.conv, l.weight) for l in L] + [(self.out_conv, self.model.conv4.conv.weight)])
            self._packer.run()
        elif getattr(self, "repack", True):            # training: the weights change every step; a frozen inference plan
            self._packer.run()                         # (infer.RtnstvStylizer) packs once and again on refresh_weights()
        self.x_first = self.x_act.from_nchw(x)
        cur = self.x_first
        for i, l in enumerate(L):
            stc = self._stats_view(i, l.conv.cout)
            l.conv.forward(cur, self.raws[i], l.out_hw, stats=stc)
            res = self.acts[i - 2] if (3 <= i <= 12 and (i - 3) % 2 == 1) else None
            tc.in_apply(self.raws[i], stc, l.gamma, l.beta, self.acts[i], l.relu, residual=res)
            cur = self.acts[i]
        c4 = self.model.conv4
        self.out_conv.forward(self.acts[14], self.raw_out.t, (self.H, self.W))
        self.raw3 = self.raw_out.to_nchw(channels=3)                    # bias in front of IN is cancelled by it
        img, self.o_mean, self.o_rstd = ops.instance_norm(self.raw3, c4.norm.weight, c4.norm.bias, act=ops.ACT_RT_OUT, return_stats=True)
        return None, img

    def backward(self, d_features, d_img: torch.Tensor, sink):
        N, H, W = self.N, self.H, self.W
        flat = sink.flat
        c4 = self.model.conv4
        self._wd_pool.zero_()
        draw3, dg, db = ops.instance_norm_bwd(self.raw3, d_img, c4.norm.weight, c4.norm.bias, self.o_mean, self.o_rstd, ops.ACT_RT_OUT)
        sink.put("conv4.norm.weight", dg)
        sink.put("conv4.norm.bias", db)
        flat.grad_view("conv4.conv.bias").zero_()
        sink.mark("conv4.conv.bias")
        draw = Act(N, H, W, 16, device=self.dev).from_nchw(draw3)
        self.out_conv.wgrad(draw, self.acts[14], (H, W), flat.grad_view("conv4.conv.weight"))
        sink.mark("conv4.conv.weight")
        G = self.out_conv.dgrad(draw, (H, W))
        _sweep_stages(self, G, ActDesc(H, W, 16, 1, REFLECT, 0), None, sink)
