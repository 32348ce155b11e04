"""Workload for compute-sanitizer (memcheck / racecheck / synccheck, one tool per run): a small ReCoNet inference forward on the
tensor-core plan (every tap-GEMM pipeline mode the environment selects: per-tap / dy-sharing / CTA pair / accumulator ring /
row ring, direct or staged epilogue), then one bf16 training step (tap-GEMM data gradients, pixel-contraction GEMM weight
gradients in plain and M-chunk form, Gram GEMMs) and one RTNSTV step.  Shapes are the smallest that still produce several
tiles per kernel and ragged edges.   compute-sanitizer --tool memcheck python tools/sanitize_run.py [infer|train|all]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vst_b200  # noqa: E402,F401
from vst_b200 import synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
H, W = 72, 136
if what in ("infer", "all"):
    from vst_b200.reconet.network import ReCoNet

    m = ReCoNet(1)
    m.load_state_dict(synth.fill_state_dict_(m.state_dict(), "gold:ReCoNet:1"))
    m = m.cuda().set_precision("bf16")
    y = m(synth.smooth_frames(2, H, W, "t:san").cuda())[-1]
    torch.cuda.synchronize()
    print("infer ok", float(y.mean()))
if what in ("train", "all"):
    from vst_b200.reconet.network import ReCoNet, Vgg16
    from vst_b200.rtnstv.network import StylizingNetwork
    from vst_b200.rtnstv.vgg19 import VGG19
    from vst_b200.train_core import PairTrainer

    h, w = 64, 96
    args = [t.cuda() for t in (synth.smooth_frames(1, h, w, "t:san:1"), synth.smooth_frames(1, h, w, "t:san:2"),
                               synth.smooth_flow(1, h, w, "t:san:f", mag=1.5), synth.mask(1, h, w, "t:san:m"))]
    style = synth.smooth_frames(1, h, w, "t:san:s")
    for fam, net, vgg, kind in (("reconet", ReCoNet(1), Vgg16(), "vgg16_rc"), ("rtnstv", StylizingNetwork(), VGG19(), "vgg19_rt")):
        vgg.load_state_dict(synth.vgg_state_dict(kind))
        tr = PairTrainer(net.cuda(), vgg.cuda(), style, fam, precision="bf16")
        t = tr.step(*args).to_dict()
        torch.cuda.synchronize()
        print(fam, "train ok", t["loss"])
