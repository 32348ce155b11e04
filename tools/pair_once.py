"""One paired forward (vst_plan_forward_pair) at 1080p x 4 frames, for ncu captures of a tap-GEMM carrying an apply rider."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vst_b200  # noqa: E402,F401
from vst_b200 import synth  # noqa: E402
from vst_b200.infer import FrameStylizer  # noqa: E402
from vst_b200.reconet.network import ReCoNet  # noqa: E402

torch.manual_seed(0)
model = ReCoNet(1).cuda().set_precision("bf16")
st = FrameStylizer(model, 1080, 1920, batch=4, lanes=2)
x = synth.frames(4, 1080, 1920, "bench:x").cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    st.run_device(x)
torch.cuda.synchronize()
print("paired", st.paired)
