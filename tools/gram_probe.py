"""Gram (pixel-contraction GEMM) bandwidth probe at the C5 sweep's largest cells; prints TB/s and the fraction of the HBM peak."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vst_b200  # noqa
import bench_sweep as S
hbm = json.load(open(os.path.join(S.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
for (B, Cc, s) in ((8, 64, 2048), (16, 64, 1024), (32, 64, 2048), (8, 128, 1024), (4, 256, 512), (4, 512, 256)):
    t, fl, by = S.gram_case(B, Cc, (s, s))
    print(f"gram B {B} C {Cc} {s}x{s}: {t*1e6:8.1f} us  {by/t/1e12:.2f} TB/s ({by/t/1e9/hbm:.2f} of HBM)  {fl/t/1e12:.0f} TFLOP/s")
