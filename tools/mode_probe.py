"""Bisecting helper for tests/test_gpu_modes.py: runs its per-mode script under a few environment variants and prints each one's
centred distance from the per-tap-box baseline.  usage: python tools/mode_probe.py "VST_EPI8=96" "VST_EPI8=96 VST_TG_GENERIC=1" ..."""
import importlib.util, os, sys, tempfile, pathlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("test_gpu_modes", os.path.join(ROOT, "tests", "test_gpu_modes.py"))
M = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(M)

tmp = pathlib.Path(tempfile.mkdtemp())
base = M._run(tmp, "base", M.BASE)
b0 = base["f32"] - 127.5
for i, spec in enumerate(sys.argv[1:]):
    env = dict(kv.split("=") for kv in spec.split())
    got = M._run(tmp, f"m{i}", env)
    d = np.abs(got["u8"].astype(np.int32) - base["u8"].astype(np.int32))
    rel = np.linalg.norm(got["f32"] - base["f32"]) / np.linalg.norm(b0)
    print(f"{spec:50s} rel {rel:.4g} u8 max {d.max()} frac {float((d > 0).mean()):.4f} guards {bool(got['guards_ok'])}")
