"""CPU emulation of the bf16 tensor-core inference plan's ROUNDING POINTS (no kernels): every conv operand (activation and
weight) rounded to bf16, fp32 accumulation, the conv output stored as bf16 (`raw`), InstanceNorm statistics from the rounded
`raw`, the normalised activation stored as bf16.  Separates "the kernels are wrong" from "bf16 storage is this lossy on these
weights".   python tools/bf16_emulation.py SD1|SD2|ReCoNet_gain  [options: raw32 (keep raw in fp32), bias (keep conv bias)]"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vst_b200  # noqa: E402,F401
from oracle import ref_torch as O  # noqa: E402
from trained_fixtures import CROP360, VARIANT, centred_rel_l2, frame, state_dict  # noqa: E402


HALF = False   # "fp16" option: fp16 storage / operands instead of bf16 (range check printed)


def r16(t):
    if HALF:
        if float(t.abs().max()) > 65504:
            print("  fp16 OVERFLOW: |max| =", float(t.abs().max()))
        return t.half().float()
    return t.bfloat16().float()


def emulate(sd, x, variant, raw32=False, keep_bias=False, hp_blocks=(), hp_stream=False, trace=None):
    """hp_blocks: residual blocks (1..5) whose second conv keeps `raw` in fp32 and whose add reads / writes an fp32 copy of the
    residual stream; hp_stream: the residual stream is fp32 everywhere (conv operands are still its bf16 rounding)."""
    def conv(x, name, stride, up=False, keep32=False):
        w, b = sd[f"{name}.weight"], sd[f"{name}.bias"]
        if up:
            x = O.nearest_up2(x)
        y = O.reflect_conv(r16(x), r16(w), b if keep_bias else None, stride)
        return y if (raw32 or keep32) else r16(y)

    def inorm(y, g, b, relu, res=None, out32=False):
        mean = y.mean(dim=(2, 3), keepdim=True)
        var = (y.square().mean(dim=(2, 3), keepdim=True) - mean * mean).clamp_min(0)
        a = g.view(1, -1, 1, 1) * torch.rsqrt(var + 1e-5)
        o = y * a + (b.view(1, -1, 1, 1) - mean * a)
        if relu:
            o = F.relu(o)
        if res is not None:
            o = o + res
        return o if out32 else r16(o)

    x = r16(x)
    blk = 0
    for name, kind, stride in O._RECONET_LAYERS[variant]:
        if kind in ("cir", "up"):
            y = conv(x, f"{name}.conv2d", stride, kind == "up")
            x = inorm(y, sd[f"{name}.instance.weight"], sd[f"{name}.instance.bias"], True)
        elif kind == "res":
            blk += 1
            hp = blk in hp_blocks
            y = inorm(conv(x, f"{name}.conv1.conv2d", 1), sd[f"{name}.in1.weight"], sd[f"{name}.in1.bias"], True)
            # the stream entering an hp block (or an hp stream) is fp32: the producer wrote an fp32 copy next to the bf16 operand
            x = inorm(conv(y, f"{name}.conv2.conv2d", 1, keep32=hp), sd[f"{name}.in2.weight"], sd[f"{name}.in2.bias"], False, x,
                      out32=hp_stream or (blk + 1) in hp_blocks)
        else:
            y = O.reflect_conv(r16(x), r16(sd[f"{name}.conv2d.weight"]), sd[f"{name}.conv2d.bias"], 1)
            x = torch.tanh(y / 255) * 150 + 255 / 2
        if trace is not None:
            trace[name] = x
    return x


if __name__ == "__main__":
    case = sys.argv[1] if len(sys.argv) > 1 else "SD1"
    opts = set(sys.argv[2:])
    HALF = "fp16" in opts
    hp_blocks = tuple(int(o[2:]) for o in opts if o.startswith("hp") and o[2:].isdigit())
    sd, x = state_dict(case), frame(360)
    t_ref, t_emu = {}, {}
    with torch.no_grad():
        ref = O.reconet_forward(sd, x, VARIANT[case], trace=t_ref)[-1]
        emu = emulate(sd, x, VARIANT[case], raw32="raw32" in opts, keep_bias="bias" in opts, hp_blocks=hp_blocks,
                      hp_stream="hpstream" in opts, trace=t_emu)
    for k in t_ref:
        a, b = t_emu[k], t_ref[k]
        print(f"{k:12s} rel-L2 {O.rel_l2(a, b):.3e}   ref mean {float(b.mean()):8.3f} std {float(b.std()):8.3f}")
    print("frame centred rel-L2", centred_rel_l2(emu, ref))
