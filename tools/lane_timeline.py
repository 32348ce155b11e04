"""Kernel timeline of the inference forward (CUPTI through torch.profiler): start / duration / stream of every kernel of one
step, plus how much of the step had a tap-GEMM resident, an HBM-bound kernel resident, or both.
    python tools/lane_timeline.py [--lanes 2] [--frames-per-step 4] [--dump]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vst_b200  # noqa: E402,F401
from vst_b200 import synth  # noqa: E402
from vst_b200.infer import FrameStylizer  # noqa: E402
from vst_b200.reconet.network import ReCoNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lanes", type=int, default=2)
ap.add_argument("--frames-per-step", type=int, default=4)
ap.add_argument("--dump", action="store_true")
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--width", type=int, default=1920)
a = ap.parse_args()
torch.manual_seed(0)
model = ReCoNet(1).cuda().set_precision("bf16")
st = FrameStylizer(model, a.height, a.width, batch=a.frames_per_step, lanes=a.lanes)
x = synth.frames(a.frames_per_step, a.height, a.width, "bench:x").cuda()
for _ in range(4):
    st.run_device(x)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        st.run_device(x)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memset" not in e.name.lower()]
ev.sort(key=lambda e: e.time_range.start)
# keep the middle step: split by the prologue kernels
pro = [i for i, e in enumerate(ev) if "prologue" in e.name]
per_step = len(pro) // 3
lo, hi = pro[per_step], pro[2 * per_step]
ev = ev[lo:hi]
t0 = ev[0].time_range.start
span = max(e.time_range.end for e in ev) - t0
pts = []
for e in ev:
    kind = "gemm" if "tapgemm" in e.name else "hbm"
    pts.append((e.time_range.start - t0, 1, kind))
    pts.append((e.time_range.end - t0, -1, kind))
pts.sort()
cnt = {"gemm": 0, "hbm": 0}
acc = {"gemm_only": 0.0, "hbm_only": 0.0, "both": 0.0, "idle": 0.0}
last = 0.0
for t, d, k in pts:
    dt = t - last
    g, h = cnt["gemm"] > 0, cnt["hbm"] > 0
    acc["both" if g and h else "gemm_only" if g else "hbm_only" if h else "idle"] += dt
    last = t
    cnt[k] += d
print(f"lanes {a.lanes} frames/step {a.frames_per_step}: step span {span:.1f} us -> {a.frames_per_step / span * 1e6:.1f} frames/s; " +
      ", ".join(f"{k} {v:.1f} us ({100 * v / span:.0f} %)" for k, v in acc.items()))
byname = {}
for e in ev:
    n = e.name.split("(")[0].replace("void vst::", "")[:40]
    byname.setdefault(n, []).append(e.time_range.end - e.time_range.start)
for n, v in sorted(byname.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {n:42s} n {len(v):3d} sum {sum(v):8.1f} us mean {sum(v) / len(v):7.1f}")
if a.dump:
    for e in ev:
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name.split('(')[0][-44:]}")
