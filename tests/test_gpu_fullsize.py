"""GPU parity at BASELINE.json's FULL sizes (the shapes bench.py times), through the same entry points:

* configs[3]  ReCoNet 1920x1080 frames: fp32 and bf16 paths against the REFERENCE's own output at that size (block means and
  a crop, tests/golden/fullsize_*.npz, made by oracle/make_golden.py), bf16 frames out of `FrameStylizer` (the benchmarked
  engine) against the CPU oracle run on the very same frame (the oracle needs ~2 s per 1080p frame), plus the property frame sharding rests on: a frame's bytes do not
  depend on which batch / rank it was stylised in (RC/network.py:171-190 is a pure function of one frame).
* configs[1]  one ReCoNet training step on 2 Sintel-shaped 1024x436 pairs: the five loss terms of the fp32 step (<= 1e-4) and of
  the bf16 tensor-core step (<= 1e-2, BASELINE.json) against the oracle's loss terms at that size; bf16 gradients against the
  fp32 step's.
* configs[4]  helper kernels at sweep sizes: `warp` at 2048^2 (corner indices bit-exact, values <= 1e-4, linearity in x),
  `flow_warp_mask` at 1024x436 bit-exact, Gram at relu1 size (64 x 1024^2) against the oracle, with the exact properties
  G = G^T and gram(2F) = 4 gram(F) (powers of two commute with every rounding on the way).
"""
import pytest
import torch

import vst_b200  # noqa: F401
from oracle import ref_torch as O
from vst_b200 import ops, synth

pytestmark = pytest.mark.gpu


def dev(t):
    return t.cuda().contiguous()


def _reconet(tag="gold:ReCoNet:1"):
    from vst_b200.reconet.network import ReCoNet

    model = ReCoNet(1)
    model.load_state_dict(synth.fill_state_dict_(model.state_dict(), tag))
    return model.cuda()


# ------------------------------------------------------------------ configs[3]: 1080p inference
def test_reconet_1080p_bf16_frames_vs_oracle_and_batch_independence(golden):
    import torch.nn.functional as F
    from vst_b200.infer import FrameStylizer

    H, W = 1080, 1920
    # the reference itself at this size (oracle/make_golden.py fullsize): block means of the frame / features + an exact crop
    g = golden("fullsize_reconet_1080p")
    x0 = dev(synth.smooth_frames(2, H, W, "t:full:x")[:1])
    _, f32, i32 = _reconet()(x0)                                  # fp32 CUDA-core path: <= 1e-4 (BASELINE.json)
    assert O.rel_l2(F.avg_pool2d(i32, 8).cpu() - 127.5, g["img_pool8"] - 127.5) < 1e-4
    assert O.rel_l2(F.avg_pool2d(f32, 10).cpu(), g["feat_pool10"]) < 1e-4
    assert O.rel_l2(i32[:, :, 500:532, 900:948].cpu() - 127.5, g["img_crop"] - 127.5) < 1e-4
    del f32, i32
    model = _reconet().set_precision("bf16")
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    x = synth.smooth_frames(2, H, W, "t:full:x")
    # the tensors the module surface returns (RC/network.py:190), frame 0 against the oracle
    with torch.no_grad():
        ref = O.reconet_forward(sd, x[:1])
    outs = model(dev(x[:1]))
    assert [tuple(o.shape) for o in outs] == [(1, 96, H // 2, W // 2), (1, 192, H // 4, W // 4), (1, 3, H, W)]
    e_img = O.rel_l2(outs[-1].cpu(), ref[-1])
    assert e_img < 2e-2, e_img                                   # BASELINE.json: stylised frames <= 2e-2
    assert O.rel_l2(outs[-1].cpu() - 127.5, ref[-1] - 127.5) < 0.1
    assert O.rel_l2(outs[1].cpu(), ref[1]) < 6e-2                 # features after 13 bf16 layers
    assert O.rel_l2(F.avg_pool2d(outs[-1], 8).cpu() - 127.5, g["img_pool8"] - 127.5) < 0.1      # bf16 vs the reference's frame
    assert O.rel_l2(F.avg_pool2d(outs[1], 10).cpu(), g["feat_pool10"]) < 6e-2
    # the benchmarked engine: uint8 BGR bytes of a 2-frame batch against the oracle's bytes of frame 0
    u8 = torch.from_numpy(FrameStylizer(model, H, W, batch=2).stylize_u8(x).copy())
    with torch.no_grad():
        want = O.infer_frame_u8(sd, x[:1])
    d = (u8[0].int() - want.int()).abs().float()
    assert d.mean() < 2.0, (d.mean(), d.max())
    # batch / shard independence, EXACT: the InstanceNorm statistics are deterministic per-CTA partials (fixed-order warp
    # butterfly, tiles of an image always grouped by the same residue class) that meet in fp64 atomics, so frame 1 stylised
    # alone has the very bytes - and the very fp32 values - it had inside the batch, run after run
    solo = torch.from_numpy(FrameStylizer(model, H, W, batch=1).stylize_u8(x[1:]).copy())
    assert torch.equal(solo[0], u8[1])
    both, alone = model(dev(x))[-1][1:], model(dev(x[1:]))[-1]
    assert torch.equal(alone, both)
    assert torch.equal(model(dev(x[1:]))[-1], alone)             # and twice the same call: bit-identical


# ------------------------------------------------------------------ configs[1]: the training step at 1024x436, batch 2
def test_reconet_train_step_1024x436_vs_oracle(golden):
    from vst_b200.reconet.network import Vgg16
    from vst_b200.train_core import PairTrainer

    H, W, B = 436, 1024, 2
    img1, img2 = synth.smooth_frames(B, H, W, "t:full:i1"), synth.smooth_frames(B, H, W, "t:full:i2")
    flow, mask = synth.smooth_flow(B, H, W, "t:full:flow"), synth.mask(B, H, W, "t:full:mask")
    style = synth.smooth_frames(1, H, W, "t:full:style")
    vgg_sd = synth.vgg_state_dict("vgg16_rc")
    m32, m16 = _reconet(), _reconet()
    sd = {k: v.detach().cpu().clone() for k, v in m32.state_dict().items()}
    with torch.no_grad():
        ref = O.reconet_losses(sd, vgg_sd, O.style_grams(vgg_sd, style, "rc"), img1, img2, flow, mask)
    args = (dev(img1), dev(img2), dev(flow), dev(mask))

    def trainer(model, precision):
        vgg = Vgg16()
        vgg.load_state_dict(vgg_sd)
        return PairTrainer(model, vgg.cuda(), style, "reconet", precision=precision)

    t32 = trainer(m32, "fp32")
    got32 = t32.forward_backward(*args).to_dict()
    gref = golden("fullsize_reconet_losses_1024x436")             # the reference's own loop body at this size
    for k in ("FTL", "OTL", "CL", "SL", "RL", "loss"):
        assert abs(got32[k] / float(ref[k]) - 1) < 1e-4, ("fp32", k, got32[k], float(ref[k]))
        assert abs(got32[k] / float(gref[k]) - 1) < 1e-4, ("fp32 vs reference", k, got32[k], float(gref[k]))
    g32 = {k: v.detach().float().cpu().clone() for k, v in t32.grads().items()}
    del t32
    torch.cuda.empty_cache()

    t16 = trainer(m16, "bf16")
    got16 = t16.forward_backward(*args).to_dict()
    for k in ("FTL", "OTL", "CL", "SL", "RL", "loss"):
        assert abs(got16[k] / float(ref[k]) - 1) < 1e-2, ("bf16", k, got16[k], float(ref[k]))
        assert abs(got16[k] / float(gref[k]) - 1) < 1e-2, ("bf16 vs reference", k, got16[k], float(gref[k]))
    g16 = t16.grads()
    assert O.rel_l2(g16["deconv3.conv2d.weight"].float().cpu(), g32["deconv3.conv2d.weight"]) < 2e-2
    for name, g in g32.items():
        if name.endswith("conv2d.bias") and not name.startswith("deconv3"):
            continue                                             # bias in front of InstanceNorm: zero gradient
        r = float(g16[name].double().norm()) / float(g.double().norm())
        assert abs(r - 1) < 0.15, (name, r)


# ------------------------------------------------------------------ configs[4]: helpers at sweep sizes
def test_decoder_frames_1080p_two_lanes_equal_float_path():
    """BASELINE configs[3] shape through the e2e path of bench.py: uint8 BGR decoder frames in (vst_plan_forward_bgr8, two lanes)
    give the bytes of the float path fed with cvframe_to_tensor's tensor."""
    from vst_b200.infer import FrameStylizer
    from vst_b200.reconet.network import ReCoNet

    model = ReCoNet(1)
    model.load_state_dict(synth.fill_state_dict_(model.state_dict(), "gold:ReCoNet:1"))
    model = model.cuda().set_precision("bf16")
    H, W = 1080, 1920
    frames = torch.randint(0, 256, (4, H, W, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(9))
    st = FrameStylizer(model, H, W, batch=4, lanes=2)
    want = torch.from_numpy(st.stylize_u8(frames.flip(-1).permute(0, 3, 1, 2).float().contiguous()).copy())
    got = torch.from_numpy(st.stylize_frames(frames).copy())
    assert torch.equal(want, got)


def test_warp_2048_corners_values_linearity():
    S = 2048
    x = synth.frames(1, S, S, "t:full:warp:x")
    y = synth.frames(1, S, S, "t:full:warp:y")
    flo = synth.flow(1, S, S, "t:full:warp:flo", mag=4.0)
    out, corners = ops.warp(dev(x), dev(flo), return_corners=True)
    x0, y0 = O.warp_corners(flo)
    assert torch.equal(corners.cpu()[..., 0].long(), x0) and torch.equal(corners.cpu()[..., 1].long(), y0)
    assert O.rel_l2(out.cpu(), O.warp(x, flo)) < 1e-4
    lin = ops.warp(dev(2.0 * x + 3.0 * y), dev(flo))
    assert O.rel_l2(lin.cpu(), (2.0 * out + 3.0 * ops.warp(dev(y), dev(flo))).cpu()) < 2e-6


def test_flow_warp_mask_1024x436_bit_exact():
    f01, f10 = synth.fb_flows(436, 1024, "t:full:fb")
    got = ops.flow_warp_mask(dev(f01[None]), dev(f10[None]), 2.0).cpu()[0]
    ref = O.flow_warp_mask(f01, f10)
    assert got.shape == ref.shape == (436, 1024)
    assert set(got.unique().tolist()) <= {0.0, 1.0}
    assert (got != ref).sum().item() == 0


def test_gram_relu1_size_vs_oracle_and_exact_properties():
    from vst_b200 import tc
    from vst_b200.tc import Act

    C, S = 64, 1024
    y = synth.uniform((1, C, S, S), "t:full:gram", lo=-1, hi=2).bfloat16().float()
    ref = O.gram_matrix(y, "rc")
    sc = 1.0 / (C * S * S)
    assert O.rel_l2(ops.gram(dev(y), sc).cpu(), ref) < 1e-4       # fp32 CUDA-core Gram
    fa = Act(1, S, S, C, device="cuda").from_nchw(dev(y))
    G = tc.gram(fa, sc).clone()
    assert O.rel_l2(G.cpu(), ref) < 2e-5                          # tcgen05 Gram on bf16-exact operands, fp32 accumulation
    assert O.rel_l2(G.cpu(), G.transpose(1, 2).cpu()) < 1e-6
    G2 = tc.gram(Act(1, S, S, C, device="cuda").from_nchw(dev(2.0 * y)), sc)
    # scaling by a power of two is exact in bf16 and fp32; only the order of the split-K atomics differs between launches
    assert O.rel_l2(G2.cpu(), 4.0 * G.cpu()) < 1e-6


# ------------------------------------------------------------------ configs[2]: the RTNSTV step at 640x360, batch 4
def test_rtnstv_train_step_640x360_vs_reference_golden(golden):
    """fp32 (<= 1e-4) and bf16 tensor-core (<= 1e-2) loss terms of one RTNSTV step on four 640x360 pairs against the
    reference's own loop body run at that size (tests/golden/fullsize_rtnstv_losses_640x360.npz)."""
    from vst_b200.rtnstv.network import StylizingNetwork
    from vst_b200.rtnstv.vgg19 import VGG19
    from vst_b200.train_core import PairTrainer

    g = golden("fullsize_rtnstv_losses_640x360")
    H, W, B = 360, 640, 4
    args = (dev(synth.smooth_frames(B, H, W, "t:full:rt:i1")), dev(synth.smooth_frames(B, H, W, "t:full:rt:i2")),
            dev(synth.smooth_flow(B, H, W, "t:full:rt:flow")), dev(synth.mask(B, H, W, "t:full:rt:mask")))
    style = synth.smooth_frames(1, H, W, "t:full:rt:style")
    for precision, tol in (("fp32", 1e-4), ("bf16", 1e-2)):
        model = StylizingNetwork()
        model.load_state_dict(synth.fill_state_dict_(model.state_dict(), "gold:rtnstv"))
        vgg = VGG19()
        vgg.load_state_dict(synth.vgg_state_dict("vgg19_rt"))
        tr = PairTrainer(model.cuda(), vgg.cuda(), style, "rtnstv", precision=precision)
        terms = tr.forward_backward(*args).to_dict()
        for k in ("CL", "SL", "RL", "TL", "loss"):
            assert abs(terms[k] / float(g[k]) - 1) < tol, (precision, k, terms[k], float(g[k]))
        assert torch.isfinite(tr.flat.grad).all()
        del tr
        torch.cuda.empty_cache()
