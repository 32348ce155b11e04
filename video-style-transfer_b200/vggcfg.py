"""VGG body layouts and tap sets used by the reference.

Indices are torchvision `vggNN().features` indices, so the state_dict keys
`slice{k}.{idx}.{weight,bias}` match the reference's
(RC/network.py:17-24, RT/vgg19.py:19-32, AA/vgg19.py:19-37).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

_CFG16 = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"]
_CFG19 = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]


def _features(cfg) -> List[Tuple]:
    ops: List[Tuple] = []
    cin = 3
    for v in cfg:
        if v == "M":
            ops.append(("pool",))
        else:
            ops.append(("conv", cin, v))
            ops.append(("relu",))
            cin = v
    return ops


def _slices(cfg, bounds) -> List[List[Tuple[int, Tuple]]]:
    ops = _features(cfg)
    return [[(i, ops[i]) for i in range(a, b)] for a, b in bounds]


VGG_LAYOUTS: Dict[str, dict] = {
    # ReCoNet: VGG16 relu1_2 / relu2_2 / relu3_3 / relu4_3
    "vgg16_rc": {
        "slices": _slices(_CFG16, [(0, 4), (4, 9), (9, 16), (16, 23)]),
        "taps": ["relu1_2", "relu2_2", "relu3_3", "relu4_3"],
    },
    # RTNSTV: VGG19 relu1_2 / relu2_2 / relu3_2 / relu4_2
    "vgg19_rt": {
        "slices": _slices(_CFG19, [(0, 4), (4, 9), (9, 14), (14, 23)]),
        "taps": ["relu1_2", "relu2_2", "relu3_2", "relu4_2"],
    },
    # AdaAttN tap set (config-5 Gram sweep only): relu1_1 ... relu5_1
    "vgg19_aa": {
        "slices": _slices(_CFG19, [(0, 2), (2, 7), (7, 12), (12, 21), (21, 30)]),
        "taps": ["relu1_1", "relu2_1", "relu3_1", "relu4_1", "relu5_1"],
    },
}
