"""warp_f32 bandwidth probe: white-noise and smooth flow at the sweep's large cells, tiled (shared-memory staged) vs direct kernel,
and bit-equality of the two (values + corner indices)."""
import json, os, subprocess, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vst_b200  # noqa
import bench_sweep as S
from vst_b200 import ops
hbm = json.load(open(os.path.join(S.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
for (B, s) in ((8, 1024), (4, 2048), (32, 512), (2, 436)):
    for smooth in (False, True):
        t, by = S.warp_case(B, s, smooth=smooth)
        print(f"warp B {B} {s}x{s} {'smooth' if smooth else 'white '}: {t*1e6:8.1f} us {by/t/1e12:.2f} TB/s ({by/t/1e9/hbm:.2f} of HBM)")
if os.environ.get("VST_WARP_TILED", "1") != "0":
    g = torch.Generator("cuda").manual_seed(3)
    x = torch.rand((2, 5, 333, 517), device="cuda", generator=g) * 255
    f = torch.randn((2, 2, 333, 517), device="cuda", generator=g) * 9
    o, c = ops.warp(x, f, return_corners=True)
    torch.save((o.cpu(), c.cpu()), "/tmp/warp_tiled.pt")
    r = subprocess.run([sys.executable, "-c", "import sys,torch;sys.path.insert(0,'.');import vst_b200;from vst_b200 import ops;"
                        "g=torch.Generator('cuda').manual_seed(3);x=torch.rand((2,5,333,517),device='cuda',generator=g)*255;"
                        "f=torch.randn((2,2,333,517),device='cuda',generator=g)*9;o,c=ops.warp(x,f,return_corners=True);"
                        "a,b=torch.load('/tmp/warp_tiled.pt');print('tiled == direct:', torch.equal(a,o.cpu()), torch.equal(b,c.cpu()))"],
                       env=dict(os.environ, VST_WARP_TILED="0"), capture_output=True, text=True)
    print(r.stdout.strip(), r.stderr[-300:])
