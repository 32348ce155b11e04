"""Shared by the CPU and GPU tests on the high-signal pins (tests/golden/trained_*.npz, made by
`python oracle/make_golden.py trained` from the reference's own shipped checkpoints RC/models_old/SD{1,2}_*.pth and from a
full ReCoNet whose deconv3 kernel is scaled so the frame spans tens of counts).  The random-init fixtures give frames of
127.5 +/- 0.2 counts, on which a relative-L2 gate passes for a constant image; these have a frame std of 42-60 counts."""
import os

import numpy as np
import torch

from vst_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DECONV3_GAIN = 200.0          # oracle/make_golden.py:DECONV3_GAIN
CASES = ("SD1", "SD2", "ReCoNet_gain")
VARIANT = {"SD1": "ReCoNetSD1", "SD2": "ReCoNetSD2", "ReCoNet_gain": "ReCoNet"}
CROP360 = (slice(100, 228), slice(200, 392))
CROP1080 = (slice(400, 528), slice(800, 992))
U8CROP1080 = (slice(400, 656), slice(800, 1184))


def state_dict(case: str):
    """The reference's trained state_dict (from the committed npz) or the gained synthetic ReCoNet one."""
    if case == "ReCoNet_gain":
        from vst_b200.reconet.network import ReCoNet

        sd = synth.fill_state_dict_(ReCoNet(1).state_dict(), "gold:ReCoNet:1")
        sd["deconv3.conv2d.weight"].mul_(DECONV3_GAIN)
        return sd
    z = np.load(os.path.join(GOLDEN, f"trained_{case}_weights.npz"))
    return {k.replace("__", "."): torch.from_numpy(z[k]) for k in z.files}


def model(case: str):
    from vst_b200.reconet import network as N

    m = getattr(N, VARIANT[case])(1)
    m.load_state_dict(state_dict(case))
    return m


def frame(res: int):
    return synth.smooth_frames(1, 360, 640, "t:trained:x360") if res == 360 else synth.smooth_frames(1, 1080, 1920, "t:trained:x1080")


def centred_rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """Relative L2 with the reference tensor's own mean removed from both (so a constant image scores 1, not ~0)."""
    m = b.double().mean()
    return float(((a.double() - m) - (b.double() - m)).norm() / (b.double() - m).norm())
