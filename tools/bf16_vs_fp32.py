"""bf16 tensor-core training step vs the fp32 step on the same inputs: loss terms and gradient rel-L2 per tensor."""
import sys
import torch
sys.path.insert(0, ".")
import vst_b200  # noqa
from oracle import ref_torch as O
from vst_b200 import synth
from vst_b200.reconet.network import ReCoNet, Vgg16
from vst_b200.train_core import PairTrainer
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 96)
B = 2
res = {}
for prec in ("fp32", "bf16"):
    model = ReCoNet(1)
    model.load_state_dict(synth.fill_state_dict_(model.state_dict(), "gold:ReCoNet:1"))
    vgg = Vgg16()
    vgg.load_state_dict(synth.vgg_state_dict("vgg16_rc"))
    tr = PairTrainer(model.cuda(), vgg.cuda(), synth.smooth_frames(1, H, W, "style"), "reconet", precision=prec)
    img1, img2 = synth.smooth_frames(B, H, W, "a").cuda(), synth.smooth_frames(B, H, W, "b").cuda()
    flow, mask = synth.smooth_flow(B, H, W, "f").cuda(), synth.mask(B, H, W, "m").cuda()
    terms = tr.forward_backward(img1, img2, flow, mask).to_dict()
    res[prec] = (terms, {k: v.clone().cpu() for k, v in tr.grads().items()})
    print(prec, terms)
t32, g32 = res["fp32"]
t16, g16 = res["bf16"]
for k in t32:
    print(f"term {k}: rel err {abs(t16[k] / t32[k] - 1):.2e}")
worst = 0
for k in g32:
    if k.endswith("conv2d.bias") and not k.startswith("deconv3"):
        continue
    e = O.rel_l2(g16[k], g32[k])
    worst = max(worst, e)
    print(f"grad {k}: {e:.2e}")
print("worst grad rel-L2", worst)
