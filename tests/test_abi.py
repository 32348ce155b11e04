"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every symbol the header
declares, the Python prototypes cover them all, and host tensors are refused (no CPU fallback)."""
import os
import re

import pytest
import torch

import vst_b200  # noqa: F401
from vst_b200 import _lib, build, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "vst_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(vst_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vst_b200.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)


def test_abi_version_and_error_string(lib):
    assert lib.vst_abi_version() == 1
    assert isinstance(lib.vst_last_error(), bytes)
    assert lib.vst_reduce_scratch_floats() > 1000


def test_arena_size_scales(lib):
    import ctypes as C

    d1 = _lib.NetDesc(0, 3, 48, 96, 192, 96, 48, 1, 360, 640)
    d2 = _lib.NetDesc(0, 3, 48, 96, 192, 96, 48, 1, 1080, 1920)
    a1, a2 = lib.vst_plan_arena_bytes(C.byref(d1)), lib.vst_plan_arena_bytes(C.byref(d2))
    assert 50e6 < a1 < 1e9 and 8 < a2 / a1 < 10


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU error path")
def test_host_tensors_are_refused(lib):
    x = torch.zeros(1, 3, 8, 8)
    flo = torch.zeros(1, 2, 8, 8)
    with pytest.raises(_lib.VstError):
        ops.warp(x, flo)
    with pytest.raises(_lib.VstError):
        ops.conv2d(x, torch.zeros(4, 3, 3, 3), None, 1, 1)


def test_state_dict_keys_match_reference_contract():
    from vst_b200.reconet.network import ReCoNet, ReCoNetSD1, ReCoNetSD2, Vgg16
    from vst_b200.rtnstv.network import StylizingNetwork
    from vst_b200.rtnstv.vgg19 import VGG19

    sd = ReCoNet().state_dict()
    assert len(sd) == 62 and sum(v.numel() for v in sd.values()) == 3_763_011
    assert list(sd)[:4] == ["conv1.conv2d.weight", "conv1.conv2d.bias", "conv1.instance.weight", "conv1.instance.bias"]
    assert "res3.in2.bias" in sd and "deconv3.conv2d.bias" in sd
    assert sum(v.numel() for v in ReCoNetSD1().state_dict().values()) == 497_475
    assert sum(v.numel() for v in ReCoNetSD2().state_dict().values()) == 424_899
    assert "res5_sd.conv2.conv2d.weight" in ReCoNetSD2().state_dict()
    rt = StylizingNetwork().state_dict()
    assert sum(v.numel() for v in rt.values()) == 246_969
    assert rt["deconv1.deconv.weight"].shape == (48, 32, 3, 3)
    assert list(Vgg16().state_dict())[-1] == "slice4.21.bias"
    assert list(VGG19().state_dict())[-1] == "slice4.21.bias" and "slice3.12.weight" in VGG19().state_dict()
    assert ReCoNet().plan_state_keys() == list(sd)


def test_default_init_matches_torch_seed():
    """`torch.manual_seed(s); ReCoNet()` must reproduce the reference's default initialisation:
    same parameter creation order as RC/network.py:157-169."""
    from vst_b200.reconet.network import ReCoNet

    torch.manual_seed(0)
    a = ReCoNet().state_dict()
    torch.manual_seed(0)
    ref = torch.nn.Conv2d(3, 48, 9)  # first module the reference constructs
    assert torch.equal(a["conv1.conv2d.weight"], ref.weight) and torch.equal(a["conv1.conv2d.bias"], ref.bias)


def test_read_sintel_flow_roundtrip_and_errors(tmp_path):
    """`.flo` reader (RT/utilities.py:113-152): tag / size validation and channel order."""
    import struct

    import numpy as np

    from vst_b200.rtnstv.utilities import read_sintel_flow

    flow = np.arange(5 * 7 * 2, dtype=np.float32).reshape(5, 7, 2) / 3
    p = tmp_path / "a.flo"
    p.write_bytes(struct.pack("<f", 202021.25) + struct.pack("<ii", 7, 5) + flow.tobytes())
    got = read_sintel_flow(str(p))
    assert got.shape == (5, 7, 2) and np.array_equal(got, flow)
    (tmp_path / "bad.flo").write_bytes(struct.pack("<f", 1.0) + struct.pack("<ii", 7, 5) + flow.tobytes())
    with pytest.raises(ValueError, match="wrong tag"):
        read_sintel_flow(str(tmp_path / "bad.flo"))
    (tmp_path / "short.flo").write_bytes(struct.pack("<f", 202021.25) + struct.pack("<ii", 7, 5) + flow.tobytes()[:-4])
    with pytest.raises(ValueError, match="too short"):
        read_sintel_flow(str(tmp_path / "short.flo"))
    (tmp_path / "long.flo").write_bytes(struct.pack("<f", 202021.25) + struct.pack("<ii", 7, 5) + flow.tobytes() + b"x")
    with pytest.raises(ValueError, match="too long"):
        read_sintel_flow(str(tmp_path / "long.flo"))
