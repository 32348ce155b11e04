"""GPU parity of the tensor-core training primitives (csrc/tc_conv.cu tap-GEMM in its forward / data-gradient
roles, csrc/tc_pcgemm.cu pixel-contraction GEMM for weight gradients and Gram matrices, csrc/tc_train.cu
element-wise adjoints) against torch autograd of the CPU oracle on bf16-rounded operands.
bf16 outputs carry one rounding (<= 4e-3 rel-L2); fp32 outputs (weight gradients, Gram) must be ~1e-6."""
import pytest
import torch

import vst_b200  # noqa: F401
from tools import tc_diag as D
from vst_b200.tc import REFLECT, REPLICATE, ZERO

pytestmark = pytest.mark.gpu
BF16_OUT, F32_OUT = 4e-3, 2e-5


@pytest.mark.parametrize("case", [("s1", 192, 192, (24, 40)), ("s1", 64, 64, (16, 32)), ("s2", 48, 96, (32, 48)),
                                  ("s2", 96, 192, (20, 28)), ("up2", 192, 96, (10, 14)), ("up2", 96, 48, (12, 20)),
                                  ("vgg", 3, 64, (16, 24)), ("vgg", 64, 128, (20, 36)), ("vgg", 256, 512, (8, 12)),
                                  ("row9", 3, 48, (24, 40)), ("row9", 6, 48, (20, 24)), ("s1", 192, 192, (109, 256))])
def test_conv_forward_dgrad_wgrad(case):
    r = D.conv_case(*case)
    assert r["guards_ok"], r                 # no write outside the output / statistics / weight-gradient tensors
    assert r["fwd"] < BF16_OUT, r
    if "dgrad" in r:
        assert r["dgrad"] < BF16_OUT, r
    if "wgrad" in r:
        assert r["wgrad"] < F32_OUT, r
    if "stats" in r:
        assert r["stats"] < 5e-3, r


@pytest.mark.parametrize("c,hw", [(64, (16, 32)), (128, (9, 11)), (256, (27, 64)), (512, (6, 10))])
def test_gram_and_adjoint(c, hw):
    r = D.gram_case(c, hw)
    assert r["gram"] < F32_OUT and r["gram_bwd"] < BF16_OUT, r


@pytest.mark.parametrize("kind,pad,relu,skip,c", [(REFLECT, 1, True, False, 48), (REFLECT, 1, False, True, 192),
                                                  (REPLICATE, 1, True, True, 96), (ZERO, 0, True, False, 48),
                                                  (REFLECT, 4, True, False, 48)])
def test_instance_norm_backward_with_fold(kind, pad, relu, skip, c):
    r = D.in_bwd_case(kind, pad, relu, skip, Cc=c)
    # the halo of G is folded onto its interior IN PLACE (one extra bf16 rounding on border pixels, a quarter of this
    # tiny tensor); without a halo the parameter gradients are fp32-exact
    ptol = F32_OUT if pad == 0 else BF16_OUT
    assert r["draw"] < BF16_OUT and r["dgamma"] < ptol and r["dbeta"] < ptol, r


def test_pool_and_relu_adjoints():
    r = D.pool_case()
    assert r["pool"] == 0 and r["relu_bwd"] == 0 and r["pool_relu_bwd"] < BF16_OUT, r


@pytest.mark.parametrize("C,Cin,hw,pad", [(192, 192, (9, 70), 0), (40, 37, (5, 33), 0), (64, 64, (6, 31), 1), (16, 3, (7, 20), 0),
                                          (256, 256, (3, 65), 2)])
def test_nchw_act_converters_roundtrip(C, Cin, hw, pad):
    """fp32 NCHW <-> bf16 channels-last: the tiled transposes (C >= 32) and the element-wise kernels agree with torch."""
    from vst_b200 import tc, synth

    H, W = hw
    x = synth.uniform((2, Cin, H, W), f"t:conv:nchw:{C}", lo=-2, hi=2).cuda()
    a = tc.Act(2, H, W, C, pad=pad, kind=tc.REFLECT if pad else tc.ZERO).from_nchw(x)
    want = x.bfloat16().float()
    inner = a.nhwc()[:, pad:pad + H, pad:pad + W, :].float().permute(0, 3, 1, 2)
    assert torch.equal(inner[:, :Cin], want)
    assert torch.count_nonzero(inner[:, Cin:]) == 0
    if pad:   # mirrored halo written by the converter
        assert torch.equal(a.nhwc()[:, 0, pad:pad + W, :Cin].float(), want[:, :, pad, :].permute(0, 2, 1))
    back = a.to_nchw()
    assert back.shape == (2, C, H, W)
    assert torch.equal(back[:, :Cin], want)


@pytest.mark.parametrize("shape", [(2, 24, 333), (1, 37, 45), (3, 16, 256), (1, 9, 31)])
def test_prologue_x9_bit_exact(shape):
    """conv1's operand X9[n][yp][px][3*kx + c] = bf16(x[n][c][reflect(yp - 4)][reflect(px + kx - 4)]), zero above k = 27
    (the 9 horizontal taps of RC/network.py:20-23's ReflectionPad2d(4) + 9x9 convolution), for widths that end inside a warp:
    the staged store (swizzled shared-memory tile, 512-byte warp stores) must write exactly the rows that exist."""
    from vst_b200 import tc

    N, H, W = shape
    g = torch.Generator("cuda").manual_seed(H * W)
    x = torch.rand((N, 3, H, W), device="cuda", generator=g) * 255 - 100
    a = tc.prologue_x9(x, 32)
    guard = a.t.clone()
    got = a.nhwc().float().cpu()                                            # [N, H+8, W, 32]
    xp = torch.nn.functional.pad(x.cpu(), (4, 4, 4, 4), mode="reflect")     # [N, 3, H+8, W+8]
    want = torch.zeros(N, H + 8, W, 32)
    for kx in range(9):
        want[..., 3 * kx:3 * kx + 3] = xp[:, :, :, kx:kx + W].permute(0, 2, 3, 1)
    want = want.to(torch.bfloat16).float()
    assert torch.equal(got, want)
    assert torch.equal(guard, a.t)
