"""Per-kernel durations of one paired forward (vst_plan_forward_pair) from CUPTI (torch.profiler): which tap-GEMM launches the
apply riders stretch.    python tools/pair_timeline.py"""
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
for _ in range(4):
    st.run_device(x)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    st.run_device(x)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memset" not in e.name.lower()]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
print("paired", st.paired, "span us", max(e.time_range.end for e in ev) - t0)
for i, e in enumerate(ev):
    print(f"{i:3d} {e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:60]}")
