import json, os, sys
import torch
sys.path.insert(0, '.')
import vst_b200
import bench_sweep as S
for (B, Cc, s) in ((8, 256, 512), (8, 512, 256), (8, 256, 512), (8, 512, 256)):
    t, fl, by = S.gram_case(B, Cc, (s, s))
    print(f"gram B {B} C {Cc} {s}x{s}: {t*1e6:8.1f} us  {fl/t/1e12:.0f} TFLOP/s")
