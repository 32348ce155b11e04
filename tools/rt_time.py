"""Profile the RTNSTV hybrid bf16 training step (640x360, 4 pairs)."""
import sys, time, torch
sys.path.insert(0, ".")
import vst_b200  # noqa
from vst_b200 import synth
from vst_b200.rtnstv.network import StylizingNetwork
from vst_b200.rtnstv.vgg19 import VGG19
from vst_b200.train_core import PairTrainer
H, W, B = 360, 640, 4
model, vgg = StylizingNetwork().cuda(), VGG19().cuda()
vgg.load_state_dict(synth.vgg_state_dict("vgg19_rt"))
tr = PairTrainer(model, vgg, synth.smooth_frames(1, H, W, "style"), "rtnstv", precision=sys.argv[1] if len(sys.argv) > 1 else "bf16")
a = [synth.smooth_frames(B, H, W, "a").cuda(), synth.smooth_frames(B, H, W, "b").cuda(), synth.smooth_flow(B, H, W, "f").cuda(), synth.mask(B, H, W, "m").cuda()]
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); tr.step(*a); torch.cuda.synchronize(); print(f"step {i}: {(time.perf_counter()-t0)*1e3:.1f} ms")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as pr:
    tr.step(*a); torch.cuda.synchronize()
print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
