"""Drop-in for RT/network.py: Conv, Res, Deconv, StylizingNetwork with the reference's state_dict
keys (`conv{1..4}.conv.*`, `.norm.*`, `res{i}.conv{1,2}.*`, `deconv{1,2}.{deconv,norm}.*`)."""
from __future__ import annotations

from typing import Union

import torch
import torch.nn as nn

from .. import ops

_ACT = {None: ops.ACT_NONE, "relu": ops.ACT_RELU, "tanh": ops.ACT_TANH}


def _act_code(activation) -> int:
    if activation is None:
        return ops.ACT_NONE
    if isinstance(activation, nn.ReLU):
        return ops.ACT_RELU
    if isinstance(activation, nn.Tanh):
        return ops.ACT_TANH
    raise NotImplementedError(f"activation {activation!r} has no fused kernel")


class Conv(nn.Module):
    """reflect-pad -> conv -> InstanceNorm(affine) -> activation (RT/network.py:10-26)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int,
                 activation: Union[nn.Module, None] = None):
        super().__init__()
        self.kernel_size, self.stride = kernel_size, stride
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride)
        self.norm = nn.InstanceNorm2d(out_channels, affine=True)
        self.activation = activation
        self._act = _act_code(activation)

    def forward(self, x, residual=None):
        y = ops.conv2d(x, self.conv.weight, self.conv.bias, self.stride, self.kernel_size // 2, ops.PAD_REFLECT)
        return ops.instance_norm(y, self.norm.weight, self.norm.bias, residual=residual, act=self._act)


class Res(nn.Module):
    """conv-IN-ReLU, conv-IN, add (RT/network.py:29-45); the channel-pad branch (:40-42) only
    triggers when widths differ, which StylizingNetwork never does."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv1 = Conv(in_channels, out_channels, 3, 1, nn.ReLU())
        self.conv2 = Conv(out_channels, out_channels, 3, 1, None)

    def forward(self, x):
        y = self.conv1(x)
        if x.shape[1] != self.conv2.conv.out_channels:
            pad = self.conv2.conv.out_channels - x.shape[1]
            x = torch.cat([x, x.new_zeros(x.shape[0], pad, *x.shape[2:])], 1)
        return self.conv2(y, residual=x)


class Deconv(nn.Module):
    """ConvTranspose2d(k3,s2,p1,op1) -> IN -> activation (RT/network.py:48-60)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int,
                 activation: Union[nn.Module, None] = None):
        super().__init__()
        if (kernel_size, stride) != (3, 2):
            raise NotImplementedError("Deconv kernel is specialised for k=3, s=2 (the only use in the reference)")
        self.deconv = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride, padding=1, output_padding=1)
        self.norm = nn.InstanceNorm2d(out_channels, affine=True)
        self.activation = activation
        self._act = _act_code(activation)

    def forward(self, x):
        y = ops.conv_transpose2d(x, self.deconv.weight, self.deconv.bias)
        return ops.instance_norm(y, self.norm.weight, self.norm.bias, act=self._act)


class StylizingNetwork(nn.Module):
    """RT/network.py:63-91; output (tanh(IN(conv)) + 1)/2*255 in [0,255]."""

    def __init__(self):
        super().__init__()
        self.conv1 = Conv(3, 16, 3, 1, nn.ReLU())
        self.conv2 = Conv(16, 32, 3, 2, nn.ReLU())
        self.conv3 = Conv(32, 48, 3, 2, nn.ReLU())
        for i in range(1, 6):
            setattr(self, f"res{i}", Res(48, 48))
        self.deconv1 = Deconv(48, 32, 3, 2, nn.ReLU())
        self.deconv2 = Deconv(32, 16, 3, 2, nn.ReLU())
        self.conv4 = Conv(16, 3, 3, 1, nn.Tanh())

    precision = "fp32"

    def __getstate__(self):                                 # per-process device caches never travel with a copy / pickle
        d = self.__dict__.copy()
        d.pop("_tc_graphs", None)
        return d

    def __deepcopy__(self, memo):
        import copy

        cache = self.__dict__.pop("_tc_graphs", None)
        try:
            new = self.__class__.__new__(self.__class__)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            if cache is not None:
                self.__dict__["_tc_graphs"] = cache
        return new

    def set_precision(self, precision: str):
        """"fp32": reference-semantics CUDA-core kernels; "bf16": the tcgen05 tap-GEMM path (vst_b200.tc_graph.RtnstvTC)."""
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        return self

    def forward(self, x):
        if self.precision == "bf16":
            from ..tc_graph import RtnstvTC

            key = (x.shape[0], x.shape[2], x.shape[3], str(x.device))
            cache = self.__dict__.setdefault("_tc_graphs", {})
            if key not in cache:
                cache[key] = RtnstvTC(self, x.shape[0], x.shape[2], x.shape[3])
                while len(cache) > 4:                       # bounded: each graph owns its activation buffers
                    cache.pop(next(iter(cache)))
            return cache[key].forward(x.float().contiguous())[1]
        x = self.conv3(self.conv2(self.conv1(x)))
        for i in range(1, 6):
            x = getattr(self, f"res{i}")(x)
        x = self.deconv2(self.deconv1(x))
        # conv4 with the output map fused: (tanh(IN(conv)) + 1)/2*255
        c = self.conv4
        y = ops.conv2d(x, c.conv.weight, c.conv.bias, 1, 1, ops.PAD_REFLECT)
        return ops.instance_norm(y, c.norm.weight, c.norm.bias, act=ops.ACT_RT_OUT)
