"""Deterministic synthetic frames / flows / weights.

Every BASELINE config is synthetic (the reference ships no data,
no golden vectors and its trained blobs are missing).  All generators here are
built on numpy's PCG64 `random()` stream, whose values are stable across
platforms and numpy versions, so a fixture made in the authoring container can
be regenerated bit-for-bit on the GPU box without shipping the tensors.
"""
from __future__ import annotations

import hashlib
from typing import Dict, Iterable, Tuple

import numpy as np
import torch


def _rng(tag: str, seed: int) -> np.random.Generator:
    h = hashlib.sha256(f"{tag}:{seed}".encode()).digest()
    return np.random.Generator(np.random.PCG64(int.from_bytes(h[:8], "little")))


def uniform(shape: Iterable[int], tag: str, seed: int = 0, lo: float = 0.0, hi: float = 1.0) -> torch.Tensor:
    """fp32 tensor of U[lo, hi) values, reproducible from (tag, seed)."""
    u = _rng(tag, seed).random(tuple(shape), dtype=np.float32)
    return torch.from_numpy(u * np.float32(hi - lo) + np.float32(lo))


def frames(n: int, h: int, w: int, tag: str = "frames", seed: int = 1234, c: int = 3) -> torch.Tensor:
    """[n, c, h, w] fp32 in 0..255 (`rand * 255`)."""
    return uniform((n, c, h, w), tag, seed, 0.0, 255.0)


def smooth_frames(n: int, h: int, w: int, tag: str = "smooth", seed: int = 1234, cell: int = 8) -> torch.Tensor:
    """Low-passed frames (noise at 1/cell resolution, bilinearly upsampled) - closer to video."""
    lo = uniform((n, 3, (h + cell - 1) // cell + 1, (w + cell - 1) // cell + 1), tag, seed, 0.0, 255.0)
    return torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=True).contiguous()


def flow(n: int, h: int, w: int, tag: str = "flow", seed: int = 7, mag: float = 4.0) -> torch.Tensor:
    """[n, 2, h, w] fp32 pixel flow; ch 0 = dx along W, ch 1 = dy along H (RC/datasets.py:100-146)."""
    # sum of 4 uniforms, centred: roughly gaussian with std = mag
    r = _rng(tag, seed).random((4, n, 2, h, w), dtype=np.float32)
    g = (r.sum(0) - np.float32(2.0)) * np.float32(mag / 0.57735)
    return torch.from_numpy(g.astype(np.float32))


def smooth_flow(n: int, h: int, w: int, tag: str = "sflow", seed: int = 7, mag: float = 3.0, cell: int = 16) -> torch.Tensor:
    """Spatially smooth flow (noise at 1/cell resolution, bilinearly upsampled)."""
    lo = flow(n, (h + cell - 1) // cell + 1, (w + cell - 1) // cell + 1, tag, seed, mag)
    return torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=True).contiguous()


def fb_flows(h: int, w: int, tag: str = "fb", seed: int = 7) -> Tuple[torch.Tensor, torch.Tensor]:
    """Forward / backward flow pair [2,h,w]: smooth forward flow, backward = -forward + noise,
    so the consistency mask (`flow_warp_mask`) comes out mostly valid with occluded speckle."""
    f01 = smooth_flow(1, h, w, tag + ":01", seed, 3.0)[0]
    noise = flow(1, h, w, tag + ":n", seed, 0.7)[0]
    return f01.contiguous(), (-f01 + noise).contiguous()


def mask(n: int, h: int, w: int, tag: str = "mask", seed: int = 7, keep: float = 0.85) -> torch.Tensor:
    """[n, h, w] fp32 {0,1} occlusion mask, 1 = valid."""
    return (uniform((n, h, w), tag, seed) < keep).float()


def fill_state_dict_(sd: Dict[str, torch.Tensor], tag: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Overwrite every tensor of a state_dict with reproducible values (in place).

    conv weights ~ U(-b, b) with b = 1/sqrt(fan_in) (the scale torch's default
    kaiming_uniform(a=sqrt(5)) produces), norm weights ~ U(0.5, 1.5), biases ~ U(-0.1, 0.1).
    """
    for name, t in sd.items():
        if not torch.is_floating_point(t):
            continue
        if t.dim() == 4:
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            b = 1.0 / float(np.sqrt(fan_in))
            v = uniform(t.shape, f"{tag}:{name}", seed, -b, b)
        elif name.endswith("weight"):
            v = uniform(t.shape, f"{tag}:{name}", seed, 0.5, 1.5)
        else:
            v = uniform(t.shape, f"{tag}:{name}", seed, -0.1, 0.1)
        t.copy_(v)
    return sd


def vgg_state_dict(kind: str, tag: str = "vgg", seed: int = 1) -> Dict[str, torch.Tensor]:
    """Random-init VGG16/19 `features[0:23]` weights under the reference's slice key names.

    kind: "vgg16_rc" (RC/network.py:17-24), "vgg19_rt" (RT/vgg19.py:19-32),
          "vgg19_aa" (AA/vgg19.py:19-37, sweep only, features[0:30]).
    He-scaled so activations stay O(1) through the 10+ layers.
    """
    from .vggcfg import VGG_LAYOUTS

    lay = VGG_LAYOUTS[kind]
    sd: Dict[str, torch.Tensor] = {}
    for si, sl in enumerate(lay["slices"]):
        for idx, op in sl:
            if op[0] != "conv":
                continue
            cin, cout = op[1], op[2]
            b = float(np.sqrt(6.0 / (cin * 9)))
            sd[f"slice{si + 1}.{idx}.weight"] = uniform((cout, cin, 3, 3), f"{tag}:{kind}:{idx}:w", seed, -b, b)
            sd[f"slice{si + 1}.{idx}.bias"] = uniform((cout,), f"{tag}:{kind}:{idx}:b", seed, -0.05, 0.05)
    return sd


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]
