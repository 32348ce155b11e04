"""Loader for the UNMODIFIED reference modules - test / baseline infrastructure, NOT product code.

Two roots hold the same files under the same relative paths:
  /root/reference          the read-only reference checkout (authoring container only);
  <repo>/baseline/_ref     a git-ignored copy of the eight files of the frame path, made by `stage()` (called from
                           `__graft_entry__.build()` whenever /root/reference is present).  It travels to the GPU box with
                           the gpurun snapshot, so `bench.py --impl reference` and the cpu_baseline leg can time the
                           reference itself there.  Nothing from it is ever committed.
RC and RT both define `network` / `utilities`, so every file is imported under a private module name; the VGG constructors
are rebound to `weights=None` (RC/network.py:12 and RT/vgg19.py:11 would download ImageNet weights).
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import textwrap
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference"
REF_COPY = os.path.join(ROOT, "baseline", "_ref")
RC_DIR = "Real-time-Coherent-Video-Style-Transfer-Network-(ReCoNet)"
RT_DIR = "Real-Time-Neural-Style-Transfer-for-Videos-(RTNSTV)"
AA_DIR = "Revisit-Attention-Mechanism-in-Arbitrary-Neural-Style-Transfer-(AdaAttN)"   # only vgg19.py + utilities.py (row a10)
FILES = [f"{RC_DIR}/network.py", f"{RC_DIR}/utilities.py", f"{RC_DIR}/train_single/train_starry-night.py",
         f"{RC_DIR}/train_single/train_Flow_SD2.py",
         f"{RT_DIR}/network.py", f"{RT_DIR}/vgg19.py", f"{RT_DIR}/utilities.py", f"{RT_DIR}/train.py"]


def stage(verbose: bool = True) -> bool:
    """Copy the frame path's eight reference files, byte for byte, to baseline/_ref (no-op without /root/reference)."""
    if not os.path.isdir(REF_SRC):
        return os.path.isdir(REF_COPY)
    for rel in FILES:
        dst = os.path.join(REF_COPY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, rel), dst)
    if verbose:
        print(f"staged {len(FILES)} unmodified reference files under baseline/_ref")
    return True


def find_root(prefer_copy: bool = False):
    """-> (root, kind) with kind "reference" when the reference itself is importable, else (None, "port")."""
    order = (REF_COPY, REF_SRC) if prefer_copy else (REF_SRC, REF_COPY)
    for r in order:
        if all(os.path.exists(os.path.join(r, f)) for f in FILES):
            return r, "reference"
    return None, "port"


def _load(name: str, path: str, alias: dict | None = None):
    """Import a reference file under a private module name; `alias` temporarily maps the bare names the file imports
    (e.g. `utilities`) to the right already-loaded module."""
    saved = {}
    for k, v in (alias or {}).items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


class Reference:
    """The reference's modules: rc_util, rc_net, rt_util, rt_net, rt_vgg (+ rt_train on demand)."""

    def __init__(self, root: str):
        import torchvision

        self.root = root
        self.rc_dir, self.rt_dir = os.path.join(root, RC_DIR), os.path.join(root, RT_DIR)
        self.rc_util = _load("ref_rc_utilities", os.path.join(self.rc_dir, "utilities.py"))
        self.rc_net = _load("ref_rc_network", os.path.join(self.rc_dir, "network.py"))
        self.rt_util = _load("ref_rt_utilities", os.path.join(self.rt_dir, "utilities.py"))
        self.rt_net = _load("ref_rt_network", os.path.join(self.rt_dir, "network.py"))
        self.rt_vgg = _load("ref_rt_vgg19", os.path.join(self.rt_dir, "vgg19.py"), alias={"utilities": self.rt_util})
        # weight download is impossible offline: rebind the names the modules looked up
        self.rc_net.vgg16 = lambda weights=None: torchvision.models.vgg16(weights=None)
        self.rt_vgg.vgg19 = lambda weights=None: torchvision.models.vgg19(weights=None)
        self._rt_train = None
        self._aa_vgg = None

    @property
    def aa_vgg(self):
        """AA/vgg19.py (the relu1_1..relu5_1 tap set); authoring container only - it is not staged."""
        if self._aa_vgg is None:
            import torchvision

            aa = os.path.join(self.root, AA_DIR)
            util = _load("ref_aa_utilities", os.path.join(aa, "utilities.py"))
            self._aa_vgg = _load("ref_aa_vgg19", os.path.join(aa, "vgg19.py"), alias={"utilities": util})
            self._aa_vgg.vgg19 = lambda weights=None: torchvision.models.vgg19(weights=None)
        return self._aa_vgg

    def modules(self):
        return self.rc_util, self.rc_net, self.rt_util, self.rt_net, self.rt_vgg

    @property
    def rt_train(self):
        """RT/train.py imported with matplotlib / datasets stubbed (only `spatial_loss` and the constants are used)."""
        if self._rt_train is None:
            stubs = {m: types.ModuleType(m) for m in ("matplotlib", "matplotlib.pyplot", "datasets")}
            stubs["matplotlib"].use = lambda *a, **k: None
            stubs["matplotlib"].pyplot = stubs["matplotlib.pyplot"]
            stubs["datasets"].Videvo = stubs["datasets"].FlyingThings3D_Monkaa = object
            self._rt_train = _load("ref_rt_train", os.path.join(self.rt_dir, "train.py"),
                                   alias={**stubs, "vgg19": self.rt_vgg, "network": self.rt_net, "utilities": self.rt_util})
        return self._rt_train

    def rc_loop_body(self) -> str:
        """The reference's own ReCoNet training-loop body (RC/train_single/train_starry-night.py, "# Forward pass" up to
        "# Backward pass"), read from the file at run time."""
        return loop_body(os.path.join(self.rc_dir, "train_single", "train_starry-night.py"), "# Forward pass", "# Backward pass")

    def rc_sd2_loop_body(self) -> str:
        """The teacher / student loop body of RC/train_single/train_Flow_SD2.py ("# Forward pass" .. "# Backward pass"):
        teacher ReCoNetSD1, student ReCoNetSD2, `sd_loss` computed and logged, not added."""
        return loop_body(os.path.join(self.rc_dir, "train_single", "train_Flow_SD2.py"), "# Forward pass", "# Backward pass")

    def rt_loop_body(self) -> str:
        body = loop_body(os.path.join(self.rt_dir, "train.py"), "# Forward pass", "# Backward pass")
        return "\n".join(l for l in body.splitlines() if not l.strip().startswith("loss_"))  # drop list logging


def loop_body(path: str, start_marker: str, end_marker: str) -> str:
    src = open(path).read().splitlines()
    a = next(i for i, l in enumerate(src) if start_marker in l)
    b = next(i for i, l in enumerate(src) if end_marker in l)
    return textwrap.dedent("\n".join(src[a:b]))


def infer_frame_u8(model, x):
    """The body of the reference's `Inference.__iter__` for one input tensor (RC/utilities.py:213-224), op for op: forward
    under no_grad, clamp, squeeze / cpu / permute / numpy, cv2 RGB->BGR, astype(uint8)."""
    import cv2
    import torch

    with torch.no_grad():
        *_, output_tensor = model(x)
        output_tensor = output_tensor.clamp(0, 255)
    output_image = output_tensor.squeeze(0).cpu().permute(1, 2, 0).numpy()
    output_image = cv2.cvtColor(output_image, cv2.COLOR_RGB2BGR)
    return output_image.astype("uint8")
