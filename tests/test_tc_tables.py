"""CPU checks of the host logic of the tensor-core training path: the tap tables, weight-packing tables and
weight-gradient unpacking tables of `vst_b200.tc.ConvTC` are executed by a plain torch emulation of the
documented tap-GEMM / pixel-contraction-GEMM semantics (include/vst_b200.h) and compared with autograd of
the reference convolutions.  No kernel runs here - this pins the index bookkeeping the GPU kernels consume."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import vst_b200  # noqa: F401
from oracle import ref_torch as O
from vst_b200 import synth, tc
from vst_b200.tc import Act, ConvTC, REFLECT, REPLICATE, ZERO


def act_from_nchw(x, pad, kind, parity, Cp=None):
    """CPU stand-in for vst_tc_nchw_to_act."""
    N, Cc, H, W = x.shape
    Cp = Cp or tc.round_up(Cc, 8)
    a = Act(N, H, W, Cp, pad, kind, parity, "cpu")
    xp = x
    if pad:
        xp = F.pad(x, (pad,) * 4, mode={REFLECT: "reflect", REPLICATE: "replicate", ZERO: "constant"}[kind])
    xp = F.pad(xp, (0, 0, 0, 0, 0, Cp - Cc)).permute(0, 2, 3, 1).contiguous()      # N, Hp, Wp, C
    if parity:
        planes = [xp[:, py::2, px::2] for py in (0, 1) for px in (0, 1)]
        xp = torch.stack(planes, 0).contiguous()
    a.t = xp.reshape(-1).clone()
    return a


def dense(act_or_dims, flat):
    Cc, X, Y, N, P = act_or_dims
    return flat.float().view(P, N, Y, X, Cc)


def read_box(t5, pl, y0, x0, gh, gw):
    """t5[pl][:, y0:y0+gh, x0:x0+gw, :] with zero fill outside -> [N, gh, gw, C]"""
    P, N, Y, X, Cc = t5.shape
    out = torch.zeros((N, gh, gw, Cc))
    ys, ye, xs, xe = max(y0, 0), min(y0 + gh, Y), max(x0, 0), min(x0 + gw, X)
    if ys < ye and xs < xe:
        out[:, ys - y0:ye - y0, xs - x0:xe - x0] = t5[pl][:, ys:ye, xs:xe]
    return out


def emu_tapgemm(d, a_flat, b_flat):
    a = dense((d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P), a_flat)
    b = b_flat.float().view(d.b_rows, d.b_K)
    om = max(d.out_mul, 1)
    out = torch.zeros((d.a_N, d.Hout, d.Wout, d.Cout))
    kt = d.kb_per_tap * d.BK
    for ph in range(d.n_phase):
        acc = torch.zeros((d.a_N, d.grid_h, d.grid_w, d.n_ntile * d.N_mma))
        for t in range(d.n_taps):
            i = ph * d.n_taps + t
            box = read_box(a, d.tap_pl[i], d.tap_dy[i], d.tap_dx[i], d.grid_h, d.grid_w)        # N, gh, gw, C
            cc = min(d.a_C, kt)
            w = b[ph * d.n_ntile * d.N_mma:(ph + 1) * d.n_ntile * d.N_mma, t * kt:t * kt + cc]   # rows, C
            acc += torch.einsum("nyxc,rc->nyxr", box[..., :cc], w)
        oy, ox = (d.ph_oy[ph], d.ph_ox[ph]) if om > 1 else (0, 0)
        out[:, oy::om, ox::om][:, :d.grid_h, :d.grid_w] = acc[..., :d.Cout]
    return out


def emu_pcgemm(d, a_flat, b_flat):
    a = dense((d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P), a_flat)
    b = dense((d.b_C, d.b_X, d.b_Y, d.b_N, d.b_P), b_flat)
    D = torch.zeros((d.n_taps, d.M, d.N))
    for t in range(d.n_taps):
        ab = read_box(a, d.a_pl[t], d.a_dy[t], d.a_dx[t], d.grid_h, d.grid_w)[..., :d.M]
        bb = read_box(b, d.b_pl[t], d.b_dy[t], d.b_dx[t], d.grid_h, d.grid_w)[..., :d.N]
        D[t] = torch.einsum("nyxm,nyxk->mk", ab, bb) * d.scale
    return D


def emu_gather(src, tab):
    s = torch.cat((src.reshape(-1).float(), torch.zeros(1)))
    return s[tab.long().clamp(min=-1)].sum(1) if True else None


def ref_conv(kind, x, w):
    if kind == "s1":
        return F.conv2d(F.pad(x, (1,) * 4, mode="reflect"), w)
    if kind == "s2":
        return F.conv2d(F.pad(x, (1,) * 4, mode="reflect"), w, stride=2)
    if kind == "up2":
        return F.conv2d(F.pad(O.nearest_up2(x), (1,) * 4, mode="reflect"), w)
    if kind in ("vgg", "vgg27"):
        return F.conv2d(x, w, padding=1)
    if kind == "tconv":
        return F.conv_transpose2d(x, w, None, stride=2, padding=1, output_padding=1)
    return F.conv2d(F.pad(x, (4,) * 4, mode="reflect"), w)


def x27_from_nchw(x):
    N, Cc, H, W = x.shape
    xp = F.pad(x, (1,) * 4)
    cols = [xp[:, :, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)]      # each N, 3, H, W
    t = torch.stack(cols, 1).permute(0, 3, 4, 1, 2).reshape(N, H, W, 27)
    a = Act(N, H, W, 32, device="cpu")
    a.t = F.pad(t, (0, 5)).reshape(-1).clone()
    return a


def x9_from_nchw(x, KR):
    N, Cc, H, W = x.shape
    xp = F.pad(x, (4,) * 4, mode="reflect")                       # N, C, H+8, W+8
    cols = [xp[:, :, :, kx:kx + W] for kx in range(9)]            # each N, C, H+8, W
    t = torch.stack(cols, 1).permute(0, 3, 4, 1, 2).reshape(N, H + 8, W, 9 * Cc)
    t = F.pad(t, (0, KR - 9 * Cc))
    a = Act(N, H + 8, W, KR, device="cpu")
    a.t = t.reshape(-1).clone()
    return a


@pytest.mark.parametrize("kind,cin,cout,hw", [("s1", 16, 24, (6, 10)), ("s1", 72, 40, (5, 7)), ("s2", 16, 32, (8, 12)),
                                              ("s2", 40, 16, (6, 8)), ("up2", 24, 16, (4, 6)), ("up2", 72, 8, (3, 5)),
                                              ("vgg", 3, 16, (6, 9)), ("vgg", 24, 40, (5, 8)), ("vgg27", 3, 16, (6, 9)), ("tconv", 16, 24, (5, 7)), ("tconv", 48, 32, (4, 6)),
                                              ("row9", 3, 16, (10, 12)),
                                              ("row9", 6, 8, (9, 11))])
def test_conv_tables_against_autograd(kind, cin, cout, hw):
    N = 2
    tag = f"{kind}:{cin}:{cout}"
    x = synth.uniform((N, cin, *hw), "tt:x:" + tag, lo=-1, hi=1).requires_grad_(True)
    k = 9 if kind == "row9" else 3
    wshape = (cin, cout, k, k) if kind == "tconv" else (cout, cin, k, k)
    w = synth.uniform(wshape, "tt:w:" + tag, lo=-0.3, hi=0.3).requires_grad_(True)
    y = ref_conv(kind, x, w)
    dy = synth.uniform(tuple(y.shape), "tt:dy:" + tag, lo=-1, hi=1)
    y.backward(dy)
    Ho, Wo = y.shape[2:]
    c = ConvTC(kind, cin, cout, "cpu", need_dgrad=kind != "row9", need_wgrad=not kind.startswith("vgg"))
    wp = emu_gather(w.detach(), c.f_tab)
    xd = x.detach()
    if kind == "row9":
        xa = x9_from_nchw(xd, c.KR)
    elif kind == "vgg27":
        xa = x27_from_nchw(xd)
    else:
        pad, knd, par = {"s1": (1, REFLECT, 0), "s2": (1, REFLECT, 1), "up2": (1, REPLICATE, 0), "vgg": (0, ZERO, 0),
                         "tconv": (0, ZERO, 0)}[kind]
        xa = act_from_nchw(xd, pad, knd, par)
    cp = tc.round_up(cout, 8)
    raw = torch.zeros(N * Ho * Wo * cp)
    d = c.fwd_desc(xa, raw, (Ho, Wo))
    got = emu_tapgemm(d, xa.t, wp)[..., :cout].permute(0, 3, 1, 2)
    assert O.rel_l2(got, y.detach()) < 1e-5

    da = act_from_nchw(dy, 0, ZERO, 1 if kind in ("up2", "tconv") else 0)
    if kind != "row9":
        wd = emu_gather(w.detach(), c.d_tab)
        p = 0 if kind.startswith("vgg") or kind == "tconv" else 1
        dd = c.dgrad_desc(da, hw, torch.zeros(1))
        G = emu_tapgemm(dd, da.t, wd)[..., :cin].permute(0, 3, 1, 2)          # N, cin, Hp, Wp
        assert tuple(G.shape[2:]) == (hw[0] + 2 * p, hw[1] + 2 * p)
        if p:
            xs = torch.zeros((N, cin, *hw), requires_grad=True)
            F.pad(xs, (1,) * 4, mode="replicate" if kind == "up2" else "reflect").backward(G)
            G = xs.grad
        assert O.rel_l2(G, x.grad) < 1e-5
    if not kind.startswith("vgg"):
        wdsc = c.wgrad_desc(da, xa, (Ho, Wo))
        D = emu_pcgemm(wdsc, da.t, xa.t)
        dw = emu_gather(D, c.w_tab).view(wshape)
        assert O.rel_l2(dw, w.grad) < 1e-5


def test_gradsink_buckets_cover_every_parameter():
    from vst_b200.reconet.network import ReCoNet
    from vst_b200.train_core import FlatParams, GradSink

    m = ReCoNet(1)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    flat = FlatParams(m)
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k])                       # re-pointing keeps keys and values
    sink = GradSink(flat, None, 4)
    assert sum(sink.members) == len(flat.names) == 62
    assert sink.ranges[0][0] == 0 and sink.ranges[-1][1] == flat.total
    for (a, b), (c, _) in zip(sink.ranges, sink.ranges[1:]):
        assert b == c
    for n in reversed(flat.names):
        sink.put(n, torch.ones(flat.offsets[n][2]))
    assert sink.finish() == 1.0 and float(flat.grad.sum()) == flat.total


def test_rowconv_out_tables_against_autograd():
    """ConvTanh's conv (48 -> 3, k = 9) as a row convolution: forward GEMM + kx-shift-sum, and both adjoints over the
    expanded gradient E (vst_tc_rowconv_expand), emulated on the CPU."""
    N, cin, cout, k, H, W = 2, 16, 3, 9, 10, 13
    x = synth.uniform((N, cin, H, W), "tt:rc:x", lo=-1, hi=1).requires_grad_(True)
    w = synth.uniform((cout, cin, k, k), "tt:rc:w", lo=-0.3, hi=0.3).requires_grad_(True)
    y = F.conv2d(F.pad(x, (4,) * 4, mode="reflect"), w)
    dz = synth.uniform(tuple(y.shape), "tt:rc:dz", lo=-1, hi=1)
    y.backward(dz)
    c = tc.RowConvOutTC(cin, cout, k, "cpu")
    xa = act_from_nchw(x.detach(), 4, REFLECT, 0)
    # forward: D[x'][(kx,co)] then out[x][co] = sum_kx D[x + kx][(kx,co)]
    d = c.fwd_desc(xa, torch.zeros(1), None, 0)
    d.grid_w, d.Wout, d.Cout = W + 8, W + 8, 32           # emulate the raw GEMM over all padded columns
    Dm = emu_tapgemm(d, xa.t, emu_gather(w.detach(), c.f_tab))        # N, H, W+8, 32
    out = sum(Dm[:, :, kx:kx + W, kx * cout:(kx + 1) * cout] for kx in range(k)).permute(0, 3, 1, 2)
    assert O.rel_l2(out, y.detach()) < 1e-5
    # E
    E = Act(N, H, W + k - 1, 32, device="cpu")
    e = torch.zeros((N, H, W + k - 1, 32))
    for kx in range(k):
        e[:, :, kx:kx + W, kx * cout:(kx + 1) * cout] = dz.permute(0, 2, 3, 1)
    E.t = e.reshape(-1)
    G = emu_tapgemm(c.dgrad_desc(E, torch.zeros(1)), E.t, emu_gather(w.detach(), c.d_tab)).permute(0, 3, 1, 2)
    xs = torch.zeros((N, cin, H, W), requires_grad=True)
    F.pad(xs, (4,) * 4, mode="reflect").backward(G)
    assert O.rel_l2(xs.grad, x.grad) < 1e-5
    Dw = emu_pcgemm(c.wgrad_desc(E, xa), E.t, xa.t)
    assert O.rel_l2(emu_gather(Dw, c.w_tab).view(cout, cin, k, k), w.grad) < 1e-5


# ---- pipeline planner (host-only ABI query: runs without a GPU) -------------------------------------------------
def _fwd_desc(kind, cin, cout, hw):
    c = ConvTC(kind, cin, cout, "cpu", need_dgrad=False, need_wgrad=False)
    H, W = hw
    if kind == "row9":
        xa, out_hw = Act(1, H + 8, W, c.KR, 0, ZERO, 0, "cpu"), (H, W)
    elif kind == "s2":
        xa, out_hw = Act(1, H, W, tc.round_up(cin, 8), 1, REFLECT, 1, "cpu"), (H // 2, W // 2)
    elif kind == "up2":
        xa, out_hw = Act(1, H, W, tc.round_up(cin, 8), 1, REPLICATE, 0, "cpu"), (2 * H, 2 * W)
    elif kind == "vgg":
        xa, out_hw = Act(1, H, W, tc.round_up(cin, 8), 0, ZERO, 0, "cpu"), (H, W)
    else:
        xa, out_hw = Act(1, H, W, tc.round_up(cin, 8), 1, REFLECT, 0, "cpu"), (H, W)
    return c.fwd_desc(xa, torch.zeros(8), out_hw)


@pytest.mark.parametrize("kind,cin,cout,hw,want_dyshare", [("s1", 192, 192, (270, 480), False),   # weight tiles too large
                                                           ("s1", 48, 48, (90, 160), True), ("s2", 48, 96, (540, 960), True),
                                                           ("up2", 96, 48, (270, 480), True), ("up2", 192, 96, (135, 240), True),
                                                           ("row9", 3, 48, (1080, 1920), True), ("vgg", 64, 64, (436, 1024), True),
                                                           ("vgg", 512, 512, (54, 128), False)])
def test_planner_columns_partition_the_taps(kind, cin, cout, hw, want_dyshare):
    """dy-sharing: every tap of every phase must appear in exactly one column, at the row offset and weight-tap index the
    kernel derives from the column table; the box must hold TH + longest column - 1 rows; tiles must cover the grid."""
    d = _fwd_desc(kind, cin, cout, hw)
    info = tc.plan_info(d)
    assert info.stream == 0
    assert bool(info.dyshare) == want_dyshare, (kind, cin, cout, info.dyshare)
    assert info.TW * info.TH == 128 * info.MT and info.TW >= 8
    assert info.tiles_x * info.TW >= d.grid_w and info.tiles_y * info.TH >= d.grid_h
    if not info.dyshare:
        assert info.box_rows == info.TH
        # wide single-phase layers that keep per-tap boxes run as CTA pairs (M = 256 MMAs, half the weight tile per CTA)
        assert info.cta2 == int(d.N_mma >= 128 and d.n_ntile == 1 and info.MT == 1)
        return
    assert info.cta2 == 0
    assert info.box_rows == info.TH + info.dy_max - 1 and info.dy_max >= 2
    for ph in range(d.n_phase):
        seen = {}
        for c in range(info.n_cols):
            i = ph * info.n_cols + c
            for j in range(info.col_n[i]):
                t = info.col_t0[i] + j * info.col_ts[i]
                assert 0 <= t < d.n_taps and t not in seen
                seen[t] = (info.col_dx[i], info.col_dy0[i] + j, info.col_pl[i])
        assert len(seen) == d.n_taps
        for t, (dx, dy, pl) in seen.items():
            k = ph * d.n_taps + t
            assert (d.tap_dx[k], d.tap_dy[k], d.tap_pl[k]) == (dx, dy, pl)


def test_planner_row_convolution_streams_through_tmem():
    """ConvTanh (48 -> 3, k = 9): N = 32 fits 16 accumulator slots -> accumulator-ring streaming, one 128-pixel row per tile."""
    rc = tc.RowConvOutTC(48, 3, 9, "cpu")
    x = Act(1, 64, 200, 48, 4, REFLECT, 0, "cpu")
    info = tc.plan_info(rc.fwd_desc(x, torch.zeros(8), None, 0))
    assert (info.stream, info.dyshare, info.TW, info.TH, info.MT) == (2, 0, 128, 1, 1)
    assert info.tiles_x == -(-200 // 120) and info.tiles_y == 64


def test_side_wgrad_falls_back_to_inline_order_without_a_second_stream(monkeypatch):
    """tc_graph._SideWgrad host logic (no GPU needed for the in-line mode): with the switch off the weight gradient
    runs in line and is marked at once, single- or multi-rank."""
    from vst_b200.tc_graph import _SideWgrad

    class Sink:
        def __init__(self, world, defer):
            self.world, self.defer, self.marked = world, defer, []

        def mark(self, name):
            self.marked.append(name)

    ran = []
    monkeypatch.setenv("VST_WGRAD_STREAM", "0")
    s = Sink(1, False)
    w = _SideWgrad(object(), s)
    assert not w.on
    w.run(lambda: ran.append("a"), "conv.weight", object())
    assert ran == ["a"] and s.marked == ["conv.weight"] and w.keep == []
    w.join()
    assert s.marked == ["conv.weight"]
    s = Sink(2, False)                      # data-parallel with the switch off: still in line, marked at once
    w = _SideWgrad(object(), s)
    assert not w.on and w.mark_now          # (with a second stream the mark happens as the gradient is ENQUEUED, so the
    w.run(lambda: ran.append("b"), "res.weight")   # bucket's all-reduce forks mid-sweep - GPU-side: bench.py dp_check)
    assert ran == ["a", "b"] and s.marked == ["res.weight"]
    s = Sink(2, True)                       # exchange deferred to after the sweep: marks wait for the join
    assert not _SideWgrad(object(), s).mark_now
