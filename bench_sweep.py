#!/usr/bin/env python
"""Kernel sweep of BASELINE configs[4]: Gram matrices of the VGG19 relu1_1..relu4_1 tap set on the
tcgen05 pixel-contraction GEMM and the bilinear flow warp, S = 256..2048, batch 1..32, as achieved TFLOP/s and GB/s
against the measured peaks (MEASURED_PEAKS.json).  Inputs are synthetic and rotate through a pool larger than L2.

    python bench_sweep.py [--quick] > profiles/<round>_c5_sweep.jsonl
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import vst_b200  # noqa: E402,F401
from vst_b200 import ops, tc  # noqa: E402
from vst_b200.tc import Act  # noqa: E402

L2_BYTES = 126 << 20


def peaks():
    d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    return d["bf16_tflops"], d["hbm_gbs"]


def timed(fn, n_variants, iters=10, warm=3):
    for i in range(warm):
        fn(i % n_variants)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_variants)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def gram_case(B, Cc, hw):
    H, W = hw
    nbytes = B * Cc * H * W * 2
    nv = max(1, min(8, (2 * L2_BYTES) // nbytes + 1))
    pool = []
    for i in range(nv):
        a = Act(B, H, W, Cc, device="cuda")
        a.t.copy_(torch.randn(a.t.numel(), device="cuda", generator=torch.Generator("cuda").manual_seed(99 + i)).relu_().bfloat16())
        pool.append(a)
    out = torch.zeros((B, Cc, Cc), dtype=torch.float32, device="cuda")
    t = timed(lambda i: tc.gram(pool[i], 1.0 / (Cc * H * W), out), nv)
    return t, 2.0 * Cc * Cc * H * W * B, nbytes + 4 * B * Cc * Cc


def warp_case(B, S, Cc=3, smooth=False):
    """smooth=False: spatially WHITE flow (randn * 4 per pixel - every lane of a gather lands on its own cache line, the
    worst case); smooth=True: optical-flow-like (noise at 1/16 resolution, bilinearly upsampled - neighbouring pixels
    sample neighbouring texels, what real Sintel / SceneFlow fields look like)."""
    nbytes = 4 * (2 * Cc + 2) * B * S * S
    nv = max(1, min(8, (2 * L2_BYTES) // nbytes + 1))
    g = torch.Generator("cuda").manual_seed(99)
    xs = [torch.rand((B, Cc, S, S), device="cuda", generator=g) * 255 for _ in range(nv)]
    if smooth:
        fl = [torch.nn.functional.interpolate(torch.randn((B, 2, S // 16 + 1, S // 16 + 1), device="cuda", generator=g) * 4, size=(S, S),
                                              mode="bilinear", align_corners=True).contiguous() for _ in range(nv)]
    else:
        fl = [torch.randn((B, 2, S, S), device="cuda", generator=g) * 4 for _ in range(nv)]
    t = timed(lambda i: ops.warp(xs[i], fl[i]), nv)
    return t, nbytes


_AA = None


def vgg_forward_case(B, S):
    """The VGG19 pass that PRODUCES the sweep's taps (AA/vgg19.py slices 1-4, relu1_1 ... relu4_1) on the tensor-core body:
    9 tap-GEMMs (3->64 ... 256->512) + 3 max-pools per image batch.  FLOPs = 2*MACs of the nine reference convolutions."""
    global _AA
    from vst_b200 import synth
    from vst_b200.adaattn.vgg19 import VGG19

    if _AA is None:
        _AA = VGG19()
        _AA.load_state_dict(synth.vgg_state_dict("vgg19_aa"))
        _AA = _AA.cuda().set_precision("bf16")
    g = torch.Generator("cuda").manual_seed(7)
    nv = max(1, min(4, (2 * L2_BYTES) // (B * 3 * S * S * 4) + 1))
    xs = [torch.rand((B, 3, S, S), device="cuda", generator=g) * 255 for _ in range(nv)]
    t = timed(lambda i: _AA(xs[i], n_slices=4), nv, iters=5, warm=2)
    macs = 0
    for cin, cout, div in ((3, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 2), (128, 256, 4), (256, 256, 4), (256, 256, 4),
                           (256, 256, 4), (256, 512, 8)):
        macs += 9 * cin * cout * (S // div) * (S // div)
    return t, 2.0 * macs * B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    tf_peak, hbm_peak = peaks()
    sizes = [256, 1024] if args.quick else [256, 512, 1024, 2048]
    batches = [1, 8] if args.quick else [1, 2, 4, 8, 16, 32]
    for S in sizes:
        for B in batches:
            if B * S * S > 32 * 2048 * 2048:          # the largest cell (2048^2 x 32: a 17 GB relu1_1 operand) still fits
                continue
            row = {"kernel": "gram(pcgemm, tcgen05)", "S": S, "batch": B, "taps": {}}
            tot_t = tot_f = 0.0
            for name, Cc, div in (("relu1_1", 64, 1), ("relu2_1", 128, 2), ("relu3_1", 256, 4), ("relu4_1", 512, 8)):
                t, fl, by = gram_case(B, Cc, (S // div, S // div))
                tot_t, tot_f = tot_t + t, tot_f + fl
                bound = "hbm" if Cc <= 128 else "tensor"
                row["taps"][name] = {"us": round(t * 1e6, 1), "tflops": round(fl / t / 1e12, 1), "gbs": round(by / t / 1e9, 1),
                                     "bound": bound, "frac": round((by / t / 1e9 / hbm_peak) if bound == "hbm" else (fl / t / 1e12 / tf_peak), 3)}
            row["tflops_4taps"] = round(tot_f / tot_t / 1e12, 1)
            print(json.dumps(row), flush=True)
            if B * S * S <= 8 * 2048 * 2048:          # the VGG pass keeps every activation of the batch: bound its memory
                t, fl = vgg_forward_case(B, S)
                print(json.dumps({"kernel": "vgg19_aa forward to relu4_1 (tapgemm, tcgen05)", "S": S, "batch": B, "us": round(t * 1e6, 1),
                                  "gflop": round(fl / 1e9, 1), "tflops": round(fl / t / 1e12, 1), "bound": "tensor",
                                  "frac": round(fl / t / 1e12 / tf_peak, 3)}), flush=True)
            t, by = warp_case(B, S)
            print(json.dumps({"kernel": "warp_f32 (white-noise flow)", "S": S, "batch": B, "us": round(t * 1e6, 1), "gbs": round(by / t / 1e9, 1),
                              "bound": "hbm", "frac": round(by / t / 1e9 / hbm_peak, 3)}), flush=True)
            t, by = warp_case(B, S, smooth=True)
            print(json.dumps({"kernel": "warp_f32 (smooth flow)", "S": S, "batch": B, "us": round(t * 1e6, 1), "gbs": round(by / t / 1e9, 1),
                              "bound": "hbm", "frac": round(by / t / 1e9 / hbm_peak, 3)}), flush=True)


if __name__ == "__main__":
    main()
