"""Drop-in for RT/vgg19.py: VGG19 features[0:23], normalises inside forward, returns a dict."""
from __future__ import annotations

from ..reconet.network import _VggBody
from .utilities import vgg_normalize


class VGG19(_VggBody):
    def __init__(self):
        super().__init__("vgg19_rt")

    def forward(self, x):
        t = self.taps(vgg_normalize(x))
        return {"relu1_2": t[0], "relu2_2": t[1], "relu3_2": t[2], "relu4_2": t[3]}
