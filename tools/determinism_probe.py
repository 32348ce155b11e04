"""Run-to-run and batch-composition reproducibility of the bf16 inference plan (InstanceNorm statistics: deterministic per-CTA
partials + fp64 atomics).  Prints the fraction of differing bytes / the centred relative L2 of the fp32 frames."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vst_b200  # noqa: E402,F401
from vst_b200 import synth  # noqa: E402
from vst_b200.infer import FrameStylizer  # noqa: E402
from vst_b200.reconet.network import ReCoNet  # noqa: E402

for (H, W) in ((360, 640), (1080, 1920)):
    model = ReCoNet(1)
    model.load_state_dict(synth.fill_state_dict_(model.state_dict(), "gold:ReCoNet:1"))
    model = model.cuda().set_precision("bf16")
    x = synth.smooth_frames(4, H, W, "t:det:x")
    xd = x.cuda()
    a = model(xd)[-1].clone()
    b = model(xd)[-1].clone()
    solo = torch.cat([model(xd[i:i + 1])[-1] for i in range(4)])
    pair = torch.cat([model(xd[0:2])[-1], model(xd[2:4])[-1]])

    def rel(p, q):
        return float(((p - q).double().norm() / (q.double() - 127.5).norm()))

    print(f"{W}x{H}: run-to-run equal={bool(torch.equal(a, b))} rel={rel(a, b):.2e} | batch4 vs solo equal={bool(torch.equal(a, solo))} "
          f"rel={rel(a, solo):.2e} | batch4 vs pairs equal={bool(torch.equal(a, pair))} rel={rel(a, pair):.2e}")
    u4 = torch.from_numpy(FrameStylizer(model, H, W, batch=4).stylize_u8(x).copy())
    u2 = torch.from_numpy(FrameStylizer(model, H, W, batch=4, lanes=2).stylize_u8(x).copy())
    u1 = torch.cat([torch.from_numpy(FrameStylizer(model, H, W, batch=1).stylize_u8(x[i:i + 1]).copy()) for i in range(4)])
    print(f"   bytes: batch4 vs lanes2 differing {(u4 != u2).float().mean():.2e}, batch4 vs solo differing {(u4 != u1).float().mean():.2e}")
