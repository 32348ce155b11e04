"""Drop-in for AA/vgg19.py:8-63: VGG19 `features[0:30]` in five slices, normalises inside forward (AA/utilities.py:79-85 -
no in-place division, the RT semantics), returns {"relu1_1", ..., "relu5_1"}.  Same state_dict keys (`slice{k}.{idx}.*`)."""
from __future__ import annotations

from ..reconet.network import _VggBody
from ..rtnstv.utilities import vgg_normalize

TAPS = ("relu1_1", "relu2_1", "relu3_1", "relu4_1", "relu5_1")


class VGG19(_VggBody):
    """precision "fp32": reference-semantics CUDA-core kernels; "bf16": the tcgen05 tap-GEMM body (tc_graph.VggTC, weights
    packed once - the body is frozen, AA/vgg19.py:40-41)."""

    def __init__(self):
        super().__init__("vgg19_aa")
        self.precision = "fp32"
        self._tc = None

    def set_precision(self, precision: str):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        return self

    def forward(self, x, n_slices: int = 5):
        xn = vgg_normalize(x)
        if self.precision == "bf16":
            from ..tc_graph import VggTC

            if self._tc is None:
                self._tc = VggTC(self)
            taps = [t.to_nchw() for t in self._tc.forward(xn, n_slices=n_slices, save=False)]
        else:
            taps = self.taps(xn)[:n_slices]
        return dict(zip(TAPS, taps))
