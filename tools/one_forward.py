"""Minimal driver for ncu captures: N inference forwards of B 1080p frames on the bf16 plan (no timing, no host traffic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vst_b200  # noqa: E402,F401
from vst_b200 import synth  # noqa: E402
from vst_b200.infer import FrameStylizer  # noqa: E402
from vst_b200.reconet.network import ReCoNet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
model = ReCoNet(1).cuda().set_precision("bf16")
st = FrameStylizer(model, 1080, 1920, batch=B)
x = synth.frames(B, 1080, 1920, "bench:x").cuda()
for _ in range(n):
    st.run_device(x)
torch.cuda.synchronize()
print("done")
