"""GPU parity on frames with a real signal: the reference's shipped SD1 / SD2 checkpoints and a gained
full ReCoNet, at 640x360 (exact uint8 frame) and 1920x1080 (crops + block means), against outputs of the UNMODIFIED reference
(tests/golden/trained_*.npz, oracle/make_golden.py trained).  All errors are CENTRED (the reference frame's mean removed), so a
constant image scores 1.0.  Gates: fp32 <= 1e-4, bf16 tensor-core path <= 2e-2 on the frame (BASELINE.json), bytes within one
count on >= 99 % of the pixels.  Every measured value is appended to gpurun_out/parity_trained.jsonl for DESIGN.md's table."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

import vst_b200  # noqa: F401
from trained_fixtures import CASES, CROP360, CROP1080, U8CROP1080, centred_rel_l2, frame, model
from oracle import ref_torch as O

pytestmark = pytest.mark.gpu
LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_trained.jsonl")


def _log(**kw):
    os.makedirs(os.path.dirname(LOG), exist_ok=True)
    with open(LOG, "a") as f:
        f.write(json.dumps(kw) + "\n")
    print(kw)


def _measure(case, res, precision, golden):
    from vst_b200.infer import FrameStylizer

    g = golden(f"trained_{case}_{res}p")
    H, W = (360, 640) if res == 360 else (1080, 1920)
    crop = CROP360 if res == 360 else CROP1080
    x = frame(res)
    m = model(case).cuda().set_precision(precision)
    out0, *_, feat, img = m(x.cuda())
    img, feat, out0 = img.cpu(), feat.cpu(), out0.cpu()
    fk, fp = ("feat_pool10", 10) if res == 360 else ("feat_pool30", 30)
    r = dict(case=case, res=res, precision=precision,
             crop=centred_rel_l2(img[:, :, crop[0], crop[1]], g["img_crop"]),
             crop_plain=O.rel_l2(img[:, :, crop[0], crop[1]], g["img_crop"]),
             out0=O.rel_l2(F.avg_pool2d(out0, fp), g[f"out0_pool{fp}"]) if f"out0_pool{fp}" in g else None,
             pool8=centred_rel_l2(F.avg_pool2d(img, 8), g["img_pool8"]),
             feat=O.rel_l2(F.avg_pool2d(feat, fp), g[fk]),
             ref_std=float(g["img_std"]), ref_mean=float(g["img_mean"]))
    u8 = torch.from_numpy(FrameStylizer(m, H, W).stylize_u8(x)[0].copy())
    want = g["u8"] if res == 360 else g["u8_crop"]
    if res != 360:
        u8 = u8[U8CROP1080[0], U8CROP1080[1]]
    d = (u8.int() - want.int()).abs()
    r.update(u8_max=int(d.max()), u8_within1=float((d <= 1).float().mean()), u8_exact=float((d == 0).float().mean()),
             u8_mean_abs=float(d.float().mean()))
    _log(**r)
    return r


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("res", [360, 1080])
def test_trained_fp32(case, res, golden):
    """fp32 CUDA-core path.  The reference's OWN fp32 output sits 3.2e-5 (SD1) / 6.9e-5 (SD2) centred from the exact (fp64)
    result on these checkpoints; with a plain running fp32 sum in the direct convolution this path sat at 2.4e-4 on
    SD2, with the two-level compensated sum (csrc/fp32_ops.cu) it holds 1e-4 CENTRED on single pixels."""
    r = _measure(case, res, "fp32", golden)
    assert r["crop"] < 1e-4 and r["crop_plain"] < 1e-4 and r["pool8"] < 1e-4, r   # measured: <= 8.4e-5 centred (SD2, 360p)
    assert r["feat"] < 1.5e-4, r                                 # `features` has std 0.007 on these checkpoints (see below)
    assert r["u8_max"] <= 1 and r["u8_exact"] > 0.99, r           # truncation ties only


@pytest.mark.parametrize("res", [360, 1080])
def test_trained_bf16_gained_reconet(res, golden):
    """The benchmarked architecture with a frame std of 42 counts.  Plain rel-L2 (BASELINE.json's definition) <= 2e-2; the
    CENTRED error of single pixels is 3.5e-2 - exactly what bf16 storage of the 16 layers' operands gives (the CPU emulation
    of the plan's rounding points, tools/bf16_emulation.py ReCoNet_gain, gives 3.51e-2), block means hold 2e-2."""
    r = _measure("ReCoNet_gain", res, "bf16", golden)
    assert r["crop_plain"] < 2e-2 and r["pool8"] < 2e-2 and r["crop"] < 4.5e-2, r
    assert r["feat"] < 2e-2, r
    assert r["u8_max"] <= 12 and r["u8_mean_abs"] < 1.5, r


@pytest.mark.parametrize("case", ["SD1", "SD2"])
@pytest.mark.parametrize("res", [360, 1080])
def test_trained_bf16_shipped_checkpoints(case, res, golden):
    """bf16 storage CANNOT hold 2e-2 on the shipped checkpoints, and this test pins that honestly: training with
    LAMBDA_F = 1e12 drove `features` (the res5 output) to std 0.007 while the residual stream that feeds it has std 0.7-0.9,
    i.e. res5 cancels its input to 1 %; the 2^-9 relative rounding of a bf16 stream (1.4e-3 absolute) is then 20-60 % of the
    signal deconv1's InstanceNorm re-amplifies.  The CPU emulation of the plan's rounding points reproduces the GPU numbers
    (SD1 0.49 emulated / 0.50 measured), so the kernels compute what bf16 implies.  Upstream of the trunk (conv3, `out0`) the
    path is at bf16 accuracy.  The fp16 + fp32-residual-stream plan (precision="fp16") is the one that holds 2e-2 here."""
    r = _measure(case, res, "bf16", golden)
    assert r["out0"] < 1e-2, r
    assert r["crop"] < 0.9, r                                     # documented miss: regression guard only


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("res", [360, 1080])
def test_trained_fp16_plan(case, res, golden):
    """The 16-bit tensor-core plan that DOES hold BASELINE.json's 2e-2 on the shipped checkpoints, CENTRED: fp16 operands and
    activation storage (same tcgen05 rate as bf16, 11 instead of 8 significand bits; frames pre-scaled by 1/16 with conv1's
    eps rescaled so the result is unchanged and nothing leaves the fp16 range) + the residual stream and each block's second
    conv output in fp32, so the res5 cancellation happens in fp32.  CPU emulation of the same rounding points
    (tools/bf16_emulation.py <case> fp16 hp1 hp2 hp3 hp4 hp5): SD1 8.6e-3, SD2 1.7e-2, gained ReCoNet 4.4e-3."""
    r = _measure(case, res, "fp16", golden)
    assert r["crop"] < 2e-2 and r["pool8"] < 2e-2, r
    assert r["feat"] < 3e-2 and (r["out0"] is None or r["out0"] < 2e-3), r
    assert r["u8_max"] <= 12 and r["u8_mean_abs"] < 1.0, r
