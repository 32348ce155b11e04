"""CPU ORACLE - test infrastructure, NOT product code.

A plain torch-CPU fp32 restatement of the reference's frame path, written from
the reference's behaviour (file:line cited per function; prefixes as in
SURVEY.md: RC/ = ReCoNet dir, RT/ = RTNSTV dir, TORCH/ = ATen headers).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product package never does.

Pinning: the reference ships no golden vectors or tests, so this
restatement is pinned against outputs of the reference itself, generated in the
authoring container by `oracle/make_golden.py` (which imports /root/reference)
and committed under `tests/golden/`; `tests/test_oracle_golden.py` replays them.

Everything is functional: networks take a `state_dict` with the reference's key
names, so the same weights can be fed to the reference, the oracle and the CUDA path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# ----------------------------------------------------------------------------- building blocks


def reflect_conv(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], stride: int) -> torch.Tensor:
    """ReflectionPad2d(k//2) then Conv2d(k, stride)  (RC/network.py:63-75, RT/network.py:10-20)."""
    p = w.shape[-1] // 2
    return F.conv2d(F.pad(x, (p, p, p, p), mode="reflect"), w, b, stride=stride)


def instance_norm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """InstanceNorm2d(affine=True, track_running_stats=False): biased variance over H*W (RC/network.py:91)."""
    mean = x.mean(dim=(2, 3), keepdim=True)
    var = (x - mean).square().mean(dim=(2, 3), keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * g.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def nearest_up2(x: torch.Tensor) -> torch.Tensor:
    """F.interpolate(scale_factor=2) default mode nearest: src = dst // 2 (RC/network.py:117)."""
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


# ----------------------------------------------------------------------------- ReCoNet family

# (module name, kind, stride) in forward order; kinds: cir = conv+IN+ReLU, res, up = upsample conv+IN+ReLU, tanh
_RECONET_LAYERS = {
    "ReCoNet": [("conv1", "cir", 1), ("conv2", "cir", 2), ("conv3", "cir", 2)]
    + [(f"res{i}", "res", 1) for i in range(1, 6)]
    + [("deconv1", "up", 1), ("deconv2", "up", 1), ("deconv3", "tanh", 1)],
    "ReCoNetSD1": [("conv1", "cir", 1), ("conv2", "cir", 2), ("conv3_sd", "cir", 2)]
    + [(f"res{i}_sd", "res", 1) for i in range(1, 6)]
    + [("deconv1_sd", "up", 1), ("deconv2", "up", 1), ("deconv3", "tanh", 1)],
    "ReCoNetSD2": [("conv1_sd2", "cir", 1), ("conv2_sd2", "cir", 2), ("conv3_sd2", "cir", 2)]
    + [(f"res{i}_sd", "res", 1) for i in range(1, 6)]
    + [("deconv1_sd2", "up", 1), ("deconv2_sd2", "up", 1), ("deconv3_sd2", "tanh", 1)],
}


def reconet_forward(sd: SD, x: torch.Tensor, variant: str = "ReCoNet", trace: Optional[dict] = None):
    """ReCoNet / SD1 / SD2 forward (RC/network.py:171-190, :215-237, :262-279).

    Returns the same tuple as the reference: ReCoNet (sd1, features, img);
    SD1 (sd2, sd, features, img); SD2 (sd, features, img).
    `trace`, if given, receives every layer output by module name.
    """
    layers = _RECONET_LAYERS[variant]
    outs = {}
    for name, kind, stride in layers:
        if kind == "cir":
            y = reflect_conv(x, sd[f"{name}.conv2d.weight"], sd[f"{name}.conv2d.bias"], stride)
            x = F.relu(instance_norm(y, sd[f"{name}.instance.weight"], sd[f"{name}.instance.bias"]))
        elif kind == "res":
            # x + IN2(conv2(ReLU(IN1(conv1(x))))) - no ReLU after the add (RC/network.py:145-150)
            y = reflect_conv(x, sd[f"{name}.conv1.conv2d.weight"], sd[f"{name}.conv1.conv2d.bias"], 1)
            y = F.relu(instance_norm(y, sd[f"{name}.in1.weight"], sd[f"{name}.in1.bias"]))
            y = reflect_conv(y, sd[f"{name}.conv2.conv2d.weight"], sd[f"{name}.conv2.conv2d.bias"], 1)
            x = x + instance_norm(y, sd[f"{name}.in2.weight"], sd[f"{name}.in2.bias"])
        elif kind == "up":
            y = reflect_conv(nearest_up2(x), sd[f"{name}.conv2d.weight"], sd[f"{name}.conv2d.bias"], 1)
            x = F.relu(instance_norm(y, sd[f"{name}.instance.weight"], sd[f"{name}.instance.bias"]))
        elif kind == "tanh":
            # tanh(conv/255)*150 + 255/2  (RC/network.py:83-85)
            y = reflect_conv(x, sd[f"{name}.conv2d.weight"], sd[f"{name}.conv2d.bias"], 1)
            x = torch.tanh(y / 255) * 150 + 255 / 2
        outs[name] = x
        if trace is not None:
            trace[name] = x
    names = [n for n, _, _ in layers]
    conv3, res5, dec1, img = outs[names[2]], outs[names[7]], outs[names[8]], outs[names[10]]
    if variant == "ReCoNet":
        return dec1, res5, img
    if variant == "ReCoNetSD1":
        return conv3, dec1, res5, img
    return conv3, res5, img


def infer_frame_u8(sd: SD, x: torch.Tensor, variant: str = "ReCoNet") -> torch.Tensor:
    """What `Inference.__iter__` yields for one frame (RC/utilities.py:216-224):
    clamp(0,255), HWC, RGB->BGR, astype(uint8) truncation.  x: [1,3n,H,W] -> uint8 [H,W,3]."""
    img = reconet_forward(sd, x, variant)[-1].clamp(0, 255)
    return img[0].permute(1, 2, 0).flip(-1).to(torch.uint8).contiguous()


# ----------------------------------------------------------------------------- RTNSTV stylizer


def _rt_conv(sd: SD, name: str, x: torch.Tensor, stride: int, act: Optional[str]) -> torch.Tensor:
    """RT `Conv`: reflect-pad -> conv -> IN(affine) -> activation (RT/network.py:10-26)."""
    y = reflect_conv(x, sd[f"{name}.conv.weight"], sd[f"{name}.conv.bias"], stride)
    y = instance_norm(y, sd[f"{name}.norm.weight"], sd[f"{name}.norm.bias"])
    return F.relu(y) if act == "relu" else torch.tanh(y) if act == "tanh" else y


def _rt_deconv(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """RT `Deconv`: ConvTranspose2d(k3,s2,p1,op1) -> IN -> ReLU (RT/network.py:48-60)."""
    y = F.conv_transpose2d(x, sd[f"{name}.deconv.weight"], sd[f"{name}.deconv.bias"], stride=2, padding=1, output_padding=1)
    return F.relu(instance_norm(y, sd[f"{name}.norm.weight"], sd[f"{name}.norm.bias"]))


def rtnstv_forward(sd: SD, x: torch.Tensor, trace: Optional[dict] = None) -> torch.Tensor:
    """StylizingNetwork.forward (RT/network.py:78-91): output (tanh(IN(conv)) + 1)/2*255."""
    x = _rt_conv(sd, "conv1", x, 1, "relu")
    x = _rt_conv(sd, "conv2", x, 2, "relu")
    x = _rt_conv(sd, "conv3", x, 2, "relu")
    for i in range(1, 6):
        y = _rt_conv(sd, f"res{i}.conv1", x, 1, "relu")
        y = _rt_conv(sd, f"res{i}.conv2", y, 1, None)
        x = y + x
        if trace is not None:
            trace[f"res{i}"] = x
    x = _rt_deconv(sd, "deconv1", x)
    x = _rt_deconv(sd, "deconv2", x)
    x = _rt_conv(sd, "conv4", x, 1, "tanh")
    return (x + 1) / 2 * 255


# ----------------------------------------------------------------------------- VGG taps

_VGG16 = (2, 2, 3, 3)  # convs per block up to block 4
_VGG19 = (2, 2, 4, 4, 4)
_WIDTH = (64, 128, 256, 512, 512)


def _vgg_ops(depths: Sequence[int]) -> List[Tuple[str, int, int]]:
    """torchvision `features` op list: ('conv', cin, cout) / ('relu',) / ('pool',)."""
    ops, cin = [], 3
    for blk, n in enumerate(depths):
        for _ in range(n):
            ops += [("conv", cin, _WIDTH[blk]), ("relu", 0, 0)]
            cin = _WIDTH[blk]
        ops.append(("pool", 0, 0))
    return ops


_VGG_KINDS = {
    "vgg16_rc": (_VGG16 + (3,), [4, 9, 16, 23]),       # RC/network.py:17-24
    "vgg19_rt": (_VGG19, [4, 9, 14, 23]),              # RT/vgg19.py:19-32
    "vgg19_aa": (_VGG19, [2, 7, 12, 21, 30]),          # AA/vgg19.py:19-37
}


def vgg_taps(sd: SD, x: torch.Tensor, kind: str) -> List[torch.Tensor]:
    """Frozen VGG body with taps at the slice ends (RC/network.py:29-40, RT/vgg19.py:38-55).

    x is already ImageNet-normalised for "vgg16_rc"/"vgg19_aa"; for "vgg19_rt" the
    reference normalises inside forward (RT/vgg19.py:39) - callers use `vgg19_rt_forward`.
    conv = 3x3, zero pad 1, bias, ReLU; max-pool 2x2 stride 2 floor.
    """
    depths, ends = _VGG_KINDS[kind]
    ops = _vgg_ops(depths)
    taps, start = [], 0
    for si, end in enumerate(ends):
        for idx in range(start, end):
            op = ops[idx]
            if op[0] == "conv":
                x = F.conv2d(x, sd[f"slice{si + 1}.{idx}.weight"], sd[f"slice{si + 1}.{idx}.bias"], padding=1)
            elif op[0] == "relu":
                x = F.relu(x)
            else:
                x = F.max_pool2d(x, 2, 2)
        taps.append(x)
        start = end
    return taps


def vgg19_rt_forward(sd: SD, x255: torch.Tensor) -> Dict[str, torch.Tensor]:
    t = vgg_taps(sd, vgg_normalize_rt(x255), "vgg19_rt")
    return dict(zip(["relu1_2", "relu2_2", "relu3_2", "relu4_2"], t))


def vgg19_aa_forward(sd: SD, x255: torch.Tensor) -> Dict[str, torch.Tensor]:
    """AA/vgg19.py:43-63: normalise (AA/utilities.py:79-85, identical to RT's) then the five slices relu1_1 ... relu5_1."""
    t = vgg_taps(sd, vgg_normalize_rt(x255), "vgg19_aa")
    return dict(zip(["relu1_1", "relu2_1", "relu3_1", "relu4_1", "relu5_1"], t))


# ----------------------------------------------------------------------------- helpers (utilities.py)

_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


def vgg_normalize_rc(batch: torch.Tensor) -> torch.Tensor:
    """RC/utilities.py:101-106 - divides ITS ARGUMENT by 255 in place, returns (batch-mean)/std."""
    mean = batch.new_tensor(_MEAN).view(-1, 1, 1)
    std = batch.new_tensor(_STD).view(-1, 1, 1)
    batch.div_(255.0)
    return (batch - mean) / std


def vgg_normalize_rt(batch: torch.Tensor) -> torch.Tensor:
    """RT/utilities.py:163-169 - out of place."""
    batch = batch.float()
    mean = batch.new_tensor(_MEAN).view(-1, 1, 1)
    std = batch.new_tensor(_STD).view(-1, 1, 1)
    return (batch / 255.0 - mean) / std


def gram_matrix(y: torch.Tensor, family: str = "rc") -> torch.Tensor:
    """F F^T / (C*H*W) for RC (RC/utilities.py:93-98); / (H*W) for RT (RT/utilities.py:155-160)."""
    b, c, h, w = y.shape
    f = y.reshape(b, c, h * w)
    g = torch.bmm(f, f.transpose(1, 2))
    return g / (c * h * w) if family == "rc" else g / (h * w)


def warp_coords(flo: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Source sampling coordinates (ix, iy) [B,H,W] fp32 for `warp` - every step rounded to fp32
    in the reference's order: v = grid + flo; n = 2*v/max(W-1,1) - 1 (RC/utilities.py:50-54);
    ix = ((n + 1) * W - 1) / 2 (TORCH/GridSampler.h:27-36, align_corners=False)."""
    B, _, H, W = flo.shape
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, W).expand(B, H, W)
    yy = torch.arange(H, dtype=torch.float32).view(1, H, 1).expand(B, H, W)
    nx = 2.0 * (xx + flo[:, 0]) / max(W - 1, 1) - 1.0
    ny = 2.0 * (yy + flo[:, 1]) / max(H - 1, 1) - 1.0
    ix = ((nx + 1) * W - 1) / 2
    iy = ((ny + 1) * H - 1) / 2
    return ix, iy


def warp_corners(flo: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """int64 north-west corner indices floor(ix), floor(iy) - the bit-exact index target."""
    ix, iy = warp_coords(flo)
    return torch.floor(ix).long(), torch.floor(iy).long()


def warp(x: torch.Tensor, flo: torch.Tensor) -> torch.Tensor:
    """Backward warp by flow = grid_sample(bilinear, zeros, align_corners=False) on the
    reference's normalised grid (RC/utilities.py:39-57 == RT/utilities.py:59-77).
    Explicit gather; blend order nw, ne, sw, se (TORCH/GridSampler.h:164-172, 205-244)."""
    B, C, H, W = x.shape
    ix, iy = warp_coords(flo)
    x0f, y0f = torch.floor(ix), torch.floor(iy)
    x0, y0 = x0f.long(), y0f.long()
    x1, y1 = x0 + 1, y0 + 1
    wx1, wy1 = ix - x0f, iy - y0f          # weight of east / south
    wx0, wy0 = (x0f + 1) - ix, (y0f + 1) - iy
    flat = x.reshape(B, C, H * W)

    def tap(xi, yi, wgt):
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).view(B, 1, H * W).expand(B, C, H * W)
        v = torch.gather(flat, 2, idx).view(B, C, H, W)
        return v * (wgt * ok).unsqueeze(1)

    return tap(x0, y0, wx0 * wy0) + tap(x1, y0, wx1 * wy0) + tap(x0, y1, wx0 * wy1) + tap(x1, y1, wx1 * wy1)


def flow_warp_mask(flo01: torch.Tensor, flo10: torch.Tensor, threshold: float = 2) -> torch.Tensor:
    """Forward-backward consistency mask (RC/utilities.py:60-90; RT adds `threshold`, RT/utilities.py:80-110).
    flows [2,H,W] -> float {0,1} [H,W]: warp (grid+flo01) by flo10, mask = |dx|+|dy| < threshold."""
    _, H, W = flo01.shape
    xx = torch.arange(W, dtype=torch.float32).view(1, W).expand(H, W)
    yy = torch.arange(H, dtype=torch.float32).view(H, 1).expand(H, W)
    grid = torch.stack((xx, yy), 0)
    w = warp((grid + flo01).unsqueeze(0), flo10.unsqueeze(0))[0]
    err = (w - grid).abs().sum(0)
    return (err < threshold).float()


# ----------------------------------------------------------------------------- losses


def feature_flow_and_mask(flow: torch.Tensor, mask: torch.Tensor, hf: int, wf: int):
    """Flow / mask at feature resolution (RC/train_single/train_starry-night.py:91-100)."""
    H, W = flow.shape[2:]
    ff = F.interpolate(flow, size=(hf, wf), mode="bilinear")
    ff = torch.stack((ff[:, 0] * (float(wf) / W), ff[:, 1] * (float(hf) / H)), 1)
    fm = (F.interpolate(mask.unsqueeze(1), size=(hf, wf), mode="bilinear").squeeze(1) > 0).float()
    return ff, fm


def reconet_losses(sd: SD, vgg_sd: SD, style_gm: List[torch.Tensor], img1, img2, flow, mask,
                   alpha=1e5, beta=1e11, gamma=1e-2, lambda_f=1e12, lambda_o=1e7,
                   input_frame_num: int = 1, variant: str = "ReCoNet") -> Dict[str, torch.Tensor]:
    """The five ReCoNet loss terms exactly as the training loop composes them
    (RC/train_single/train_starry-night.py:76-148).  Differentiable w.r.t. `sd` tensors."""
    idx0 = (input_frame_num - 1) * 3
    *_, fmap1, sty1 = reconet_forward(sd, img1, variant)
    *_, fmap2, sty2 = reconet_forward(sd, img2, variant)
    # :81-84 - in-place /255 on the net outputs: from here on sty{1,2} are the /255 tensors
    sty1 = sty1 / 255.0
    sty2 = sty2 / 255.0
    mean = sty1.new_tensor(_MEAN).view(-1, 1, 1)
    std = sty1.new_tensor(_STD).view(-1, 1, 1)
    sty1 = (sty1 - mean) / std
    sty2 = (sty2 - mean) / std
    im1 = (img1[:, idx0:idx0 + 3] / 255.0 - mean) / std
    im2 = (img2[:, idx0:idx0 + 3] / 255.0 - mean) / std
    sf1, sf2 = vgg_taps(vgg_sd, sty1, "vgg16_rc"), vgg_taps(vgg_sd, sty2, "vgg16_rc")
    cf1, cf2 = vgg_taps(vgg_sd, im1, "vgg16_rc"), vgg_taps(vgg_sd, im2, "vgg16_rc")

    # feature-temporal (:91-106)
    ff, fm = feature_flow_and_mask(flow, mask, fmap1.shape[2], fmap1.shape[3])
    fme = fm.unsqueeze(1).expand(-1, fmap1.shape[1], -1, -1)
    ftl = torch.sum(fme * (fmap2 - warp(fmap1, ff)).square())
    ftl = ftl * (1 / int(torch.count_nonzero(fme)))
    ftl = ftl * lambda_f

    # output-temporal (:109-123) on the normalised tensors
    o = sty2 - warp(sty1, flow)
    i = im2 - warp(im1, flow)
    lum = (0.2126 * i[:, 0] + 0.7152 * i[:, 1] + 0.0722 * i[:, 2]).unsqueeze(1).expand(-1, 3, -1, -1)
    me = mask.unsqueeze(1).expand(-1, 3, -1, -1)
    otl = torch.sum(me * (o - lum).square())
    otl = otl * (1 / int(torch.count_nonzero(me)))
    otl = otl * lambda_o

    # content (:126-129): relu3_3
    cl = (F.mse_loss(sf1[2], cf1[2]) + F.mse_loss(sf2[2], cf2[2])) * alpha

    # style (:132-138)
    sl = 0
    for k, gs in enumerate(style_gm):
        g1, g2 = gram_matrix(sf1[k], "rc"), gram_matrix(sf2[k], "rc")
        sl = sl + F.mse_loss(g1, gs.expand(g1.shape[0], -1, -1)) + F.mse_loss(g2, gs.expand(g1.shape[0], -1, -1))
    sl = sl * beta

    # TV (:141-145) on the normalised styled frames, top-left (H-1)x(W-1) window
    def tv(s):
        return (s[:, :, :-1, 1:] - s[:, :, :-1, :-1]).square() + (s[:, :, 1:, :-1] - s[:, :, :-1, :-1]).square()

    rl = gamma * torch.sum(tv(sty1) + tv(sty2))
    return {"FTL": ftl, "OTL": otl, "CL": cl, "SL": sl, "RL": rl, "loss": ftl + otl + cl + sl + rl}


def sd_loss(teacher_sd: SD, student_sd: SD, img1, img2, teacher_variant: str = "ReCoNetSD1", student_variant: str = "ReCoNetSD2",
            beta: float = 1e11) -> torch.Tensor:
    """The symmetric-distillation term of RC/train_single/train_Flow_SD{1,2}.py:155-159 - computed and LOGGED by the reference,
    never added to `loss` (:162): 0.01 * BETA * (MSE(t(img1)[0], s_sd(img1)) + MSE(t(img2)[0], s_sd(img2))),
    with s_sd = output 1 of a ReCoNetSD1 student (`sd`, :85) or output 0 of a ReCoNetSD2 student (`sd`)."""
    si = 1 if student_variant == "ReCoNetSD1" else 0
    out = 0
    for img in (img1, img2):
        out = out + F.mse_loss(reconet_forward(teacher_sd, img, teacher_variant)[0], reconet_forward(student_sd, img, student_variant)[si])
    return out * (0.01 * beta)


def style_grams(vgg_sd: SD, style255: torch.Tensor, family: str = "rc") -> List[torch.Tensor]:
    """Style-image Gram matrices, computed once (RC/...starry-night.py:55-56, RT/train.py:92-93)."""
    if family == "rc":
        feats = vgg_taps(vgg_sd, vgg_normalize_rt(style255), "vgg16_rc")
    else:
        feats = list(vgg19_rt_forward(vgg_sd, style255).values())
    return [gram_matrix(f, family) for f in feats]


def rtnstv_losses(sd: SD, vgg_sd: SD, style_gm: List[torch.Tensor], img1, img2, flow, mask,
                  alpha=1e7, beta=5e7, gamma=5e-1, lam=1e6) -> Dict[str, torch.Tensor]:
    """RTNSTV step losses (RT/train.py:36-60, 113-139)."""
    s1, s2 = rtnstv_forward(sd, img1), rtnstv_forward(sd, img2)

    def spatial(content, styled):
        cf = vgg19_rt_forward(vgg_sd, content)["relu4_2"]
        sf = vgg19_rt_forward(vgg_sd, styled)
        cl = F.mse_loss(cf, sf["relu4_2"]) * alpha
        sl = 0
        for gs, f in zip(style_gm, sf.values()):
            g = gram_matrix(f, "rt")
            sl = sl + F.mse_loss(g, gs.expand(g.shape[0], -1, -1))
        sl = sl * beta
        r1 = (styled[:, :, :-1, 1:] - styled[:, :, :-1, :-1]).square()
        r2 = (styled[:, :, 1:, :-1] - styled[:, :, :-1, :-1]).square()
        rl = torch.sqrt((r1 + r2).clamp(min=1e-8)).mean() * gamma
        return cl, sl, rl

    c1, st1, r1 = spatial(img1, s1)
    c2, st2, r2 = spatial(img2, s2)
    me = mask.unsqueeze(1).expand(-1, 3, -1, -1)
    cnt = me.sum() + 1e-8
    tl = (me * (s2 - warp(s1, flow)).square()).sum() / cnt * lam
    cl, sl, rl = c1 + c2, st1 + st2, r1 + r2
    return {"CL": cl, "SL": sl, "RL": rl, "TL": tl, "loss": cl + sl + rl + tl}


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int,
              lr=1e-3, b1=0.9, b2=0.999, eps=1e-8) -> None:
    """torch.optim.Adam defaults (RC/...starry-night.py:44), single-tensor form, in place."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b|| / ||b|| - the tolerance metric of BASELINE.json."""
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def scene_flow_sample(flow_into_future, flow_into_past, motion, resolution_wh):
    """RC/datasets.py:114-143 for one sample: flows [2,H0,W0] at native size, motion [H,W] (already resized) ->
    (flow_into_past [2,H,W], mask [H,W]).  Keeps the reference's channel/ratio pairing (:131-134)."""
    W, H = resolution_wh
    shp = flow_into_past.shape
    fp = F.interpolate(flow_into_past.unsqueeze(0), size=(H, W), mode="bilinear", align_corners=False).squeeze(0)
    ff = F.interpolate(flow_into_future.unsqueeze(0), size=(H, W), mode="bilinear", align_corners=False).squeeze(0)
    ff, fp = ff.clone(), fp.clone()
    ff[0] *= ff.shape[1] / shp[1]
    ff[1] *= ff.shape[2] / shp[2]
    fp[0] *= fp.shape[1] / shp[1]
    fp[1] *= fp.shape[2] / shp[2]
    m = motion.clone()
    m[m != 0] = 1
    m = 1 - m
    mask = flow_warp_mask(ff, fp) * m
    return fp, mask
