"""Drop-in for RC/network.py: same class names, constructor arguments, forward signatures,
return tuples and state_dict keys, computed by this repo's CUDA kernels.

The torch.nn modules below are parameter containers only (they give the reference's key names and
its default initialisation order, so `torch.manual_seed(s); ReCoNet()` produces the reference's
weights); no torch.nn forward is ever called.  Two execution paths:
  precision "fp32" - reference-semantics CUDA-core kernels, layer by layer (vst_b200.ops);
  precision "bf16" - the tcgen05/TMA tensor-core plan (vst_b200.engine), whole network per call;
  precision "fp16" - the same plan with fp16 operands / storage and an fp32 residual stream (same tensor rate, 8x finer
                     rounding): the mode that holds 2e-2 on the reference's shipped checkpoints.
"""
from __future__ import annotations

from collections import namedtuple

import torch
import torch.nn as nn

from .. import ops
from ..vggcfg import VGG_LAYOUTS


class ConvLayer(nn.Module):
    """Reflection-padded convolution (RC/network.py:63-75)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, bias=True):
        super().__init__()
        self.kernel_size, self.stride = kernel_size, stride
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, bias=bias)

    def _conv(self, x, ups=1, act=ops.ACT_NONE):
        return ops.conv2d(x, self.conv2d.weight, self.conv2d.bias, self.stride, self.kernel_size // 2, ops.PAD_REFLECT,
                          ups, act)

    def forward(self, x):
        return self._conv(x)


class ConvTanh(ConvLayer):
    """tanh(conv/255)*150 + 255/2 (RC/network.py:78-85)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__(in_channels, out_channels, kernel_size, stride)

    def forward(self, x):
        return self._conv(x, act=ops.ACT_RECONET_OUT)


class ConvInstRelu(ConvLayer):
    """conv -> InstanceNorm(affine) -> ReLU (RC/network.py:88-98)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__(in_channels, out_channels, kernel_size, stride)
        self.instance = nn.InstanceNorm2d(out_channels, affine=True)

    def forward(self, x):
        return ops.instance_norm(self._conv(x), self.instance.weight, self.instance.bias, act=ops.ACT_RELU)


class UpsampleConvLayer(nn.Module):
    """nearest x`upsample` then reflection-padded conv (RC/network.py:101-120); the upsampled
    tensor is never materialised - the conv kernel reads src = dst // 2."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, upsample=None):
        super().__init__()
        self.upsample, self.kernel_size, self.stride = upsample, kernel_size, stride
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride)

    def _conv(self, x):
        return ops.conv2d(x, self.conv2d.weight, self.conv2d.bias, self.stride, self.kernel_size // 2, ops.PAD_REFLECT,
                          self.upsample or 1)

    def forward(self, x):
        return self._conv(x)


class UpsampleConvInstRelu(UpsampleConvLayer):
    def __init__(self, in_channels, out_channels, kernel_size, stride, upsample=None):
        super().__init__(in_channels, out_channels, kernel_size, stride, upsample)
        self.instance = nn.InstanceNorm2d(out_channels, affine=True)

    def forward(self, x):
        return ops.instance_norm(self._conv(x), self.instance.weight, self.instance.bias, act=ops.ACT_RELU)


class ResidualBlock(nn.Module):
    """x + IN2(conv2(ReLU(IN1(conv1(x))))), no ReLU after the add (RC/network.py:136-150)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1):
        super().__init__()
        self.conv1 = ConvLayer(in_channels, out_channels, kernel_size, stride)
        self.in1 = nn.InstanceNorm2d(out_channels, affine=True)
        self.conv2 = ConvLayer(out_channels, out_channels, kernel_size, stride)
        self.in2 = nn.InstanceNorm2d(out_channels, affine=True)

    def forward(self, x):
        y = ops.instance_norm(self.conv1(x), self.in1.weight, self.in1.bias, act=ops.ACT_RELU)
        return ops.instance_norm(self.conv2(y), self.in2.weight, self.in2.bias, residual=x)


class SelectiveLoadModule(nn.Module):
    """Only load layers present under the same name (RC/network.py:46-60)."""

    def forward(self, x):
        return x

    def load_state_dict(self, state_dict):
        own = self.state_dict()
        for name, param in state_dict.items():
            if name in own:
                own[name].copy_(param)


class _ReCoNetBase(nn.Module):
    """Shared machinery: layer order, precision switch, tensor-core plan cache."""

    _returns_conv3 = False
    _MAX_PLANS = 4   # cached (shape, slot) plans per model, oldest dropped first
    _order = ()      # module attribute names in forward order (11 entries)
    _widths = None   # (c1, c2, c3, d1, d2)

    def __init__(self, input_frame_num=1):
        super().__init__()
        self.input_frame_num = input_frame_num
        self.precision = "fp32"
        self._plans = {}

    # plans hold ctypes handles and device arenas: they are per-process caches, never part of a copy / pickle of the model
    def __getstate__(self):
        d = self.__dict__.copy()
        d["_plans"] = {}
        return d

    def __deepcopy__(self, memo):
        import copy

        plans, self._plans = self._plans, {}
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            self._plans = plans
        return new

    # -- tensor-core path ----------------------------------------------------------------
    def set_precision(self, precision: str):
        if precision not in ("fp32", "bf16", "fp16"):
            raise ValueError("precision must be 'fp32', 'bf16' or 'fp16'")
        self.precision = precision
        return self

    def _weights_version(self):
        # `_weights_generation` is bumped by PairTrainer.step: its Adam kernel writes the parameters through raw pointers,
        # which moves neither `_version` nor `data_ptr`
        return (getattr(self, "_weights_generation", 0),) + tuple(p._version for p in self.parameters()) + \
            tuple(p.data_ptr() for p in self.parameters())

    def plan(self, N, H, W, slot=0):
        """The tensor-core plan for this input shape (rebuilt when any parameter changed).  `slot` distinguishes
        independent plans of the same shape (each owns its arena), e.g. one per concurrent CUDA stream."""
        from ..engine import ReCoNetPlan

        fp16 = self.precision == "fp16"
        key = (N, H, W, str(next(self.parameters()).device), slot, fp16)
        ver = self._weights_version()
        hit = self._plans.get(key)
        if hit is None or hit[0] != ver:
            c1, c2, c3, d1, d2 = self._widths
            tensors = [self.state_dict()[k] for k in self.plan_state_keys()]
            hit = (ver, ReCoNetPlan(tensors, 3 * self.input_frame_num, c1, c2, c3, d1, d2, N, H, W,
                                    next(self.parameters()).device, fp16=fp16))
            self._plans[key] = hit
            while len(self._plans) > self._MAX_PLANS:          # bounded: each plan owns an arena of up to several GB
                self._plans.pop(next(iter(self._plans)))
        return hit[1]

    def plan_state_keys(self):
        """state_dict keys in the order vst_plan_create expects (the ReCoNet registration order)."""
        keys = []
        for name in self._order:
            m = getattr(self, name)
            if isinstance(m, ResidualBlock):
                keys += [f"{name}.conv1.conv2d.weight", f"{name}.conv1.conv2d.bias", f"{name}.in1.weight", f"{name}.in1.bias",
                         f"{name}.conv2.conv2d.weight", f"{name}.conv2.conv2d.bias", f"{name}.in2.weight", f"{name}.in2.bias"]
            elif hasattr(m, "instance"):
                keys += [f"{name}.conv2d.weight", f"{name}.conv2d.bias", f"{name}.instance.weight", f"{name}.instance.bias"]
            else:
                keys += [f"{name}.conv2d.weight", f"{name}.conv2d.bias"]
        return keys

    def _run(self, x):
        """-> dict of the tensors the reference's forwards return: conv3, features, deconv1, img."""
        if self.precision in ("bf16", "fp16"):
            N, _, H, W = x.shape
            p = self.plan(N, H, W)
            img, feat = p.forward(x, want_img=True, want_features=True)
            # deconv1's buffer (stage 13) is still intact after a full forward; conv3's (stage 2) is
            # recycled by the residual trunk, so the variants that return it re-run the first 3 stages.
            conv3 = p.forward_upto(x, 2) if self._returns_conv3 else None
            return {"conv3": conv3, "features": feat, "deconv1": p.activation(13), "img": img}
        o = self._order
        x = getattr(self, o[0])(x)
        x = getattr(self, o[1])(x)
        x = conv3 = getattr(self, o[2])(x)
        for i in range(3, 8):
            x = getattr(self, o[i])(x)
        features = x
        x = deconv1 = getattr(self, o[8])(x)
        x = getattr(self, o[9])(x)
        img = getattr(self, o[10])(x)
        return {"conv3": conv3, "features": features, "deconv1": deconv1, "img": img}


class ReCoNet(_ReCoNetBase):
    """RC/network.py:153-190.  forward -> (sd1, features, img)."""

    _order = ("conv1", "conv2", "conv3", "res1", "res2", "res3", "res4", "res5", "deconv1", "deconv2", "deconv3")
    _widths = (48, 96, 192, 96, 48)

    def __init__(self, input_frame_num=1):
        super().__init__(input_frame_num)
        self.conv1 = ConvInstRelu(3 * input_frame_num, 48, kernel_size=9, stride=1)
        self.conv2 = ConvInstRelu(48, 96, kernel_size=3, stride=2)
        self.conv3 = ConvInstRelu(96, 192, kernel_size=3, stride=2)
        for i in range(1, 6):
            setattr(self, f"res{i}", ResidualBlock(192, 192))
        self.deconv1 = UpsampleConvInstRelu(192, 96, kernel_size=3, stride=1, upsample=2)
        self.deconv2 = UpsampleConvInstRelu(96, 48, kernel_size=3, stride=1, upsample=2)
        self.deconv3 = ConvTanh(48, 3, kernel_size=9, stride=1)

    def forward(self, x):
        r = self._run(x)
        return (r["deconv1"], r["features"], r["img"])


class ReCoNetSD1(_ReCoNetBase):
    """RC/network.py:193-237.  forward -> (sd2, sd, features, img)."""

    _order = ("conv1", "conv2", "conv3_sd", "res1_sd", "res2_sd", "res3_sd", "res4_sd", "res5_sd", "deconv1_sd",
              "deconv2", "deconv3")
    _widths = (32, 64, 64, 64, 32)
    _returns_conv3 = True

    def __init__(self, input_frame_num=1):
        super().__init__(input_frame_num)
        self.conv1 = ConvInstRelu(3 * input_frame_num, 32, kernel_size=9, stride=1)
        self.conv2 = ConvInstRelu(32, 64, kernel_size=3, stride=2)
        self.conv3_sd = ConvInstRelu(64, 64, kernel_size=3, stride=2)
        for i in range(1, 6):
            setattr(self, f"res{i}_sd", ResidualBlock(64, 64))
        self.deconv1_sd = UpsampleConvInstRelu(64, 64, kernel_size=3, stride=1, upsample=2)
        self.deconv2 = UpsampleConvInstRelu(64, 32, kernel_size=3, stride=1, upsample=2)
        self.deconv3 = ConvTanh(32, 3, kernel_size=9, stride=1)

    def forward(self, x):
        r = self._run(x)
        return (r["conv3"], r["deconv1"], r["features"], r["img"])


class ReCoNetSD2(_ReCoNetBase):
    """RC/network.py:240-279.  forward -> (sd, features, img)."""

    _order = ("conv1_sd2", "conv2_sd2", "conv3_sd2", "res1_sd", "res2_sd", "res3_sd", "res4_sd", "res5_sd",
              "deconv1_sd2", "deconv2_sd2", "deconv3_sd2")
    _widths = (16, 32, 64, 32, 16)
    _returns_conv3 = True

    def __init__(self, input_frame_num=1):
        super().__init__(input_frame_num)
        self.conv1_sd2 = ConvInstRelu(3 * input_frame_num, 16, kernel_size=9, stride=1)
        self.conv2_sd2 = ConvInstRelu(16, 32, kernel_size=3, stride=2)
        self.conv3_sd2 = ConvInstRelu(32, 64, kernel_size=3, stride=2)
        for i in range(1, 6):
            setattr(self, f"res{i}_sd", ResidualBlock(64, 64))
        self.deconv1_sd2 = UpsampleConvInstRelu(64, 32, kernel_size=3, stride=1, upsample=2)
        self.deconv2_sd2 = UpsampleConvInstRelu(32, 16, kernel_size=3, stride=1, upsample=2)
        self.deconv3_sd2 = ConvTanh(16, 3, kernel_size=9, stride=1)

    def forward(self, x):
        r = self._run(x)
        return (r["conv3"], r["features"], r["img"])


VggOutputs = namedtuple("VggOutputs", ["relu1_2", "relu2_2", "relu3_3", "relu4_3"])


class _VggBody(nn.Module):
    """Frozen VGG `features` prefix with taps at the slice ends; keys `slice{k}.{idx}.{weight,bias}`."""

    def __init__(self, kind: str):
        super().__init__()
        self.kind = kind
        lay = VGG_LAYOUTS[kind]
        for si, sl in enumerate(lay["slices"]):
            seq = nn.Sequential()
            for idx, op in sl:
                if op[0] == "conv":
                    seq.add_module(str(idx), nn.Conv2d(op[1], op[2], 3, padding=1))
                elif op[0] == "relu":
                    seq.add_module(str(idx), nn.ReLU(inplace=True))
                else:
                    seq.add_module(str(idx), nn.MaxPool2d(2, 2))
            setattr(self, f"slice{si + 1}", seq)
        for p in self.parameters():
            p.requires_grad = False
        self.pretrained = False     # the reference constructs these with IMAGENET1K_V1 weights (a download); see load_torchvision

    def load_torchvision(self, src) -> "_VggBody":
        """Load torchvision VGG weights (`features.<idx>.weight|bias` keys of vgg16 / vgg19, a state_dict or a path to one) -
        what RC/network.py:12 / RT/vgg19.py:11 download in the constructor.  Marks the body as pretrained."""
        sd = torch.load(src, map_location="cpu", weights_only=True) if isinstance(src, (str, bytes)) or hasattr(src, "__fspath__") else src
        own = self.state_dict()
        for key in own:
            _, idx, kind = key.split(".")
            tv = f"features.{idx}.{kind}"
            if tv not in sd:
                raise KeyError(f"load_torchvision: {tv} missing from the given state_dict")
            own[key].copy_(sd[tv])
        self.pretrained = True
        return self

    def ensure_weights(self, who: str = "train()"):
        """Called by the train() entry points: use $VST_VGG_WEIGHTS (a torchvision vgg16 / vgg19 state_dict file) when set,
        otherwise warn LOUDLY that content / style losses are being taken against random features."""
        import os
        import warnings

        if self.pretrained:
            return self
        path = os.environ.get("VST_VGG_WEIGHTS")
        if path:
            return self.load_torchvision(path)
        warnings.warn(f"{who}: {type(self).__name__} is RANDOMLY INITIALISED - the reference loads torchvision IMAGENET1K_V1 weights "
                      "(no network here). Content / style losses are computed against random features; set VST_VGG_WEIGHTS to a "
                      "torchvision state_dict file or call .load_torchvision(...) for a meaningful stylisation.", RuntimeWarning, stacklevel=2)
        return self

    def taps(self, x):
        out = []
        for si, sl in enumerate(VGG_LAYOUTS[self.kind]["slices"]):
            seq = getattr(self, f"slice{si + 1}")
            for idx, op in sl:
                if op[0] == "conv":
                    m = getattr(seq, str(idx))
                    x = ops.conv2d(x, m.weight, m.bias, 1, 1, ops.PAD_ZERO, 1, ops.ACT_RELU)  # conv + bias + ReLU fused
                elif op[0] == "pool":
                    x = ops.maxpool2(x)
            out.append(x)
        return out


class Vgg16(_VggBody):
    """RC/network.py:9-40: VGG16 features[0:23], taps relu1_2/2_2/3_3/4_3 as a namedtuple.
    Weights are random-init here (no network for the ImageNet blob): `load_torchvision()` / $VST_VGG_WEIGHTS bring in the
    torchvision ones, and the train() entry points warn loudly when neither was used."""

    def __init__(self, device="cpu"):
        super().__init__("vgg16_rc")
        self.to(device)

    def forward(self, X):
        return VggOutputs(*self.taps(X))


if __name__ == "__main__":
    device = torch.device("cuda")
    model = ReCoNet().to(device)
    print(model(torch.randn(2, 3, 360, 640, device=device))[-1].shape)
