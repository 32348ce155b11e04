"""GPU bring-up diagnostics (not a test): per-stage rel-L2 of the bf16 plan against the oracle.
Writes gpurun_out/diag.txt."""
import os, sys, traceback
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vst_b200
from vst_b200 import ops, synth
from oracle import ref_torch as O

os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/diag.txt", "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s); out.write(s + "\n"); out.flush()

def main():
    P("device", torch.cuda.get_device_name(0))
    # 1. plain tc conv at a few shapes
    for cin, cout, hw in ((64, 64, (16, 32)), (192, 192, (24, 40)), (48, 96, (33, 17)), (96, 48, (19, 45)), (16, 32, (16, 16))):
        x = synth.uniform((1, cin, *hw), f"d:x:{cin}", lo=-1, hi=1); w = synth.uniform((cout, cin, 3, 3), f"d:w:{cin}", lo=-.1, hi=.1)
        try:
            got = ops.tc_conv3x3(x.cuda(), w.cuda(), ops.PAD_REFLECT); torch.cuda.synchronize()
            ref = O.reflect_conv(x.bfloat16().float(), w.bfloat16().float(), None, 1)
            P("tc_conv", cin, cout, hw, "rel_l2", O.rel_l2(got.cpu(), ref), "max|got|", got.abs().max().item(), "max|ref|", ref.abs().max().item())
        except Exception as e:
            P("tc_conv", cin, cout, hw, "FAILED", repr(e)); traceback.print_exc()
    # 2. plan stage by stage
    from vst_b200.reconet import network as N
    for variant, hw in (("ReCoNet", (32, 48)), ("ReCoNet", (72, 136)), ("ReCoNetSD2", (40, 64))):
        model = getattr(N, variant)(1)
        model.load_state_dict(synth.fill_state_dict_(model.state_dict(), f"gold:{variant}:1"))
        model = model.cuda().set_precision("bf16")
        x = synth.smooth_frames(2, *hw, "d:plan")
        tr = {}
        ref = O.reconet_forward({k: v.cpu() for k, v in model.state_dict().items()}, x, variant, trace=tr)
        try:
            p = model.plan(2, *hw)
            img, feat = p.forward(x.cuda(), want_img=True, want_features=True); torch.cuda.synchronize()
            names = list(tr)
            stage_of = {0: 0, 1: 1, 2: 2, 3: 4, 4: 6, 5: 8, 6: 10, 7: 12, 8: 13, 9: 14}
            for li, name in enumerate(names[:10]):
                a = p.forward_upto(x.cuda(), stage_of[li]).cpu()
                P(variant, hw, name, "rel_l2", O.rel_l2(a, tr[name]), "nan", bool(torch.isnan(a).any()))
            P(variant, hw, "img rel_l2", O.rel_l2(img.cpu(), ref[-1]), "features", O.rel_l2(feat.cpu(), tr[names[7]]))
        except Exception as e:
            P(variant, hw, "FAILED", repr(e)); traceback.print_exc()

main()
