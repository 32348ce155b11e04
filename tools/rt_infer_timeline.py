"""Kernel list of one captured RTNSTV 640x360 x 4 inference replay (CUPTI): name, start, duration."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vst_b200  # noqa: E402,F401
from vst_b200 import synth  # noqa: E402
from vst_b200.infer import RtnstvStylizer  # noqa: E402
from vst_b200.rtnstv.network import StylizingNetwork  # noqa: E402

torch.manual_seed(0)
rt = RtnstvStylizer(StylizingNetwork().cuda().set_precision("bf16"), 360, 640, batch=4)
x = synth.frames(4, 360, 640, "bench:rt").cuda()
for _ in range(4):
    rt.run_device(x)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    rt.run_device(x)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
print("span us", max(e.time_range.end for e in ev) - t0, "sum", sum(e.time_range.end - e.time_range.start for e in ev))
for e in ev:
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:70]}")
