"""Tensor-core primitives of the bf16 training step (include/vst_b200.h, "Tensor-core primitives").

Everything here works on channels-last bf16 buffers (`Act`) and launches this repo's kernels:
  * `ConvTC`          - one convolution layer in its three roles: forward (tap-GEMM), data gradient (tap-GEMM
                         over the output gradient with re-packed weights) and weight gradient (pixel-contraction
                         GEMM), for the five layer geometries of the networks:
                         "s1" 3x3 reflect, "s2" 3x3 stride 2 (parity planes), "up2" nearest-x2 + 3x3 (4 phases on
                         the low-res tensor), "row9" conv1 (k = 9 over the X9 row-window operand), "vgg" 3x3 zero pad;
  * `gram` / `gram_bwd` - F F^T and its adjoint on the same kernels;
  * thin wrappers over the InstanceNorm / ReLU / max-pool forward and adjoint kernels.
Weight (re)packing and weight-gradient unpacking are table-driven gathers: the index tables are built once per
layer with numpy (pure index bookkeeping), the arithmetic runs in `vst_gather_sum_f32`.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import ActDesc, PcGemmDesc, TapGemmDesc, TapGemmPlanInfo, check

REFLECT, REPLICATE, ZERO = 0, 1, 2
EPI_BF16, EPI_F32_NCHW, EPI_ROWCONV = 0, 1, 2
BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def choose_bk(c: int) -> Tuple[int, int]:
    """(BK, k-blocks per tap) for an operand with c channels (csrc/engine.cu choose_bk)."""
    if c % 64 == 0:
        return 64, c // 64
    if c % 32 == 0:
        return 32, c // 32
    if c <= 64:
        return 64, 1            # TMA zero-fills channels c..63
    return 16, (c + 15) // 16


class Act:
    """Channels-last bf16 activation buffer [N][H+2p][W+2p][C] (or 4 parity planes of it)."""

    def __init__(self, N, H, W, Cc, pad=0, kind=ZERO, parity=0, device="cuda", zero=False):
        if Cc % 8:
            raise _lib.VstError("Act: channels must be a multiple of 8")
        if parity and ((H + 2 * pad) % 2 or (W + 2 * pad) % 2):
            raise _lib.VstError("Act: parity planes need even padded extents")
        self.N, self.H, self.W, self.C, self.pad, self.kind, self.parity = N, H, W, Cc, pad, kind, parity
        n = N * (H + 2 * pad) * (W + 2 * pad) * Cc
        self.t = (torch.zeros if zero else torch.empty)(n, dtype=BF16, device=device)

    @property
    def desc(self) -> ActDesc:
        return ActDesc(self.H, self.W, self.C, self.pad, self.kind, self.parity)

    def dims(self):
        """(C, X, Y, N, P) of the dense tensor the TMA maps see."""
        Hp, Wp = self.H + 2 * self.pad, self.W + 2 * self.pad
        if self.parity:
            return self.C, Wp // 2, Hp // 2, self.N, 4
        return self.C, Wp, Hp, self.N, 1

    def ptr(self) -> int:
        return self.t.data_ptr()

    def nhwc(self) -> torch.Tensor:
        """View [N, H+2p, W+2p, C] (plain layouts only)."""
        assert not self.parity
        return self.t.view(self.N, self.H + 2 * self.pad, self.W + 2 * self.pad, self.C)

    def from_nchw(self, x: torch.Tensor) -> "Act":
        x = x.float().contiguous()
        check(_lib.lib().vst_tc_nchw_to_act(x.data_ptr(), x.shape[1], self.ptr(), self.desc, self.N, _stream()), "vst_tc_nchw_to_act")
        return self

    def to_nchw(self, channels: Optional[int] = None) -> torch.Tensor:
        """fp32 NCHW copy of the interior; `channels` (<= 8) keeps only the first channels of a padded operand."""
        if channels is not None and channels < self.C:
            out = torch.empty((self.N, channels, self.H, self.W), dtype=torch.float32, device=self.t.device)
            check(_lib.lib().vst_tc_act_to_nchw_first(self.ptr(), self.desc, self.N, channels, out.data_ptr(), _stream()),
                  "vst_tc_act_to_nchw_first")
            return out
        out = torch.empty((self.N, self.C, self.H, self.W), dtype=torch.float32, device=self.t.device)
        check(_lib.lib().vst_tc_act_to_nchw(self.ptr(), self.desc, self.N, out.data_ptr(), _stream()), "vst_tc_act_to_nchw")
        return out


def _i8(vals, n=96):
    arr = (C.c_byte * n)()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr


def plan_info(d: TapGemmDesc) -> TapGemmPlanInfo:
    """Pipeline mode / tiling `vst_tc_tapgemm` would choose for `d` (host-only query, runs without a GPU)."""
    info = TapGemmPlanInfo()
    check(_lib.lib().vst_tc_tapgemm_plan(C.byref(d), C.byref(info)), "vst_tc_tapgemm_plan")
    return info


def gather_sum(src: torch.Tensor, table: torch.Tensor, dst: torch.Tensor):
    """dst[i] = sum_j src.flat[table[i, j]] (table < 0 skipped); dst fp32 or bf16."""
    n, terms = table.shape
    if dst.numel() != n or src.dtype != torch.float32 or not src.is_contiguous():
        raise _lib.VstError("gather_sum: shape / dtype mismatch")
    check(_lib.lib().vst_gather_sum_f32(src.data_ptr(), table.data_ptr(), terms, dst.data_ptr(), n, int(dst.dtype == BF16), _stream()),
          "vst_gather_sum_f32")
    return dst


class MergedPack:
    """One `gather_sum` launch that packs the weights of ALL layers of a network (forward and data-gradient operand of
    each) instead of two launches per layer.  Needs every weight to be a contiguous fp32 view of one storage (the flat
    parameter buffer of `train_core.FlatParams`); the per-layer packed buffers are re-pointed into one pooled bf16
    buffer.  Falls back to the per-layer `pack` when the weights live in separate storages."""

    def __init__(self, pairs):
        self.pairs = list(pairs)                       # [(conv object with f_tab/f_w[/d_tab/d_w], weight tensor)]
        self.ok = False
        self._sig = None
        self._build()

    def _signature(self):
        return tuple((w.data_ptr(), w.untyped_storage().data_ptr()) for _, w in self.pairs)

    def _build(self):
        ws = [w for _, w in self.pairs]
        self._sig = self._signature()
        st = ws[0].untyped_storage().data_ptr()
        self.ok = all(w.untyped_storage().data_ptr() == st and w.is_contiguous() and w.dtype == torch.float32 for w in ws)
        if not self.ok:
            return
        dev = ws[0].device
        jobs = []
        for conv, w in self.pairs:
            jobs.append((conv, "f_w", conv.f_tab, w))
            if getattr(conv, "d_tab", None) is not None:
                jobs.append((conv, "d_w", conv.d_tab, w))
        terms = max(j[2].shape[1] for j in jobs)
        tabs, off, slots = [], 0, []
        for conv, attr, tab, w in jobs:
            n = tab.shape[0]
            n_al = (n + 63) // 64 * 64                 # 128-byte aligned slices (TMA base addresses)
            t = torch.full((n_al, terms), -1, dtype=torch.int32, device=dev)
            t[:n, : tab.shape[1]] = torch.where(tab >= 0, tab + w.storage_offset(), tab)
            tabs.append(t)
            slots.append((conv, attr, off, n))
            off += n_al
        self.table = torch.cat(tabs, 0).contiguous()
        self.pool = torch.zeros(off, dtype=BF16, device=dev)
        self.src = torch.empty(0, dtype=torch.float32, device=dev).set_(ws[0].untyped_storage())
        for conv, attr, o, n in slots:
            old = getattr(conv, attr)
            setattr(conv, attr, self.pool[o:o + n].view(old.shape))

    def run(self):
        if self._signature() != self._sig:             # parameters were re-pointed (e.g. FlatParams built later)
            self._build()
        if not self.ok:
            for conv, w in self.pairs:
                conv.pack(w.detach())
            return
        gather_sum(self.src, self.table, self.pool)


def pool_wgrad_accumulators(convs) -> torch.Tensor:
    """Re-points the fp32 weight-gradient accumulators (`w_D`, filled by atomics) of `convs` into ONE buffer, so a sweep
    clears them with a single fill instead of one per layer.  Returns the pool; the caller zeroes it before the sweep."""
    convs = [c for c in convs if getattr(c, "w_D", None) is not None]
    sizes = [(c.w_D.numel() + 31) // 32 * 32 for c in convs]
    pool = torch.zeros(sum(sizes), dtype=torch.float32, device=convs[0].w_D.device)
    off = 0
    for c, n in zip(convs, sizes):
        c.w_D = pool[off:off + c.w_D.numel()].view(c.w_D.shape)
        c._wd_pooled = True
        off += n
    return pool


# up2 phase structure: S(p, d) = the original kernel rows that phase p / low-res tap d sums (csrc/engine.cu)
_S = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}


class ConvTC:
    """One convolution (fp32 master weight [Cout, Cin, k, k]) on the tensor-core kernels."""

    def __init__(self, kind: str, cin: int, cout: int, device, cin_pad: Optional[int] = None, need_dgrad=True, need_wgrad=True):
        self.kind, self.cin, self.cout, self.dev = kind, cin, cout, device
        self.k = 9 if kind == "row9" else 3
        self.cin_p = cin_pad or round_up(cin, 8)            # channels of the input Act
        self.n_mma = min(256, round_up(cout, 16))
        self.n_ntile = (cout + self.n_mma - 1) // self.n_mma
        self.rows = self.n_mma * self.n_ntile
        self._build_fwd_table()
        if need_dgrad:
            self._build_dgrad_table()
        if need_wgrad:
            self._build_wgrad_table()

    # ---- index bookkeeping ---------------------------------------------------------------------
    def _widx(self, co, ci, ky, kx):
        if self.kind == "tconv":       # ConvTranspose2d weight layout [Cin, Cout, k, k] (RT/network.py:51)
            return ((ci * self.cout + co) * self.k + ky) * self.k + kx
        return ((co * self.cin + ci) * self.k + ky) * self.k + kx

    def _to_dev(self, tab: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(tab.astype(np.int32))).to(self.dev)

    def _build_fwd_table(self):
        kind, cin, cout, k = self.kind, self.cin, self.cout, self.k
        if kind == "row9":
            self.KR = 32 if 9 * cin <= 32 else round_up(9 * cin, 64)
            self.f_BK, self.f_kbpt = (32, 1) if self.KR == 32 else (64, self.KR // 64)
            K = 9 * self.KR
            tab = -np.ones((self.rows, K, 1), np.int64)
            co, ky, kx, c = np.meshgrid(np.arange(cout), np.arange(9), np.arange(9), np.arange(cin), indexing="ij")
            tab[co, ky * self.KR + kx * cin + c, 0] = self._widx(co, c, ky, kx)
            self.f_ntaps, self.f_nphase = 9, 1
        elif kind == "vgg27":
            # first VGG conv (3 -> 64): im2col operand X27 [N][H][W][32], ONE tap of K = 32 (vst_tc_prologue_x27)
            self.f_BK, self.f_kbpt = 32, 1
            tab = -np.ones((self.rows, 32, 1), np.int64)
            co, c, ky, kx = np.meshgrid(np.arange(cout), np.arange(cin), np.arange(3), np.arange(3), indexing="ij")
            tab[co, (ky * 3 + kx) * cin + c, 0] = self._widx(co, c, ky, kx)
            self.f_ntaps, self.f_nphase = 1, 1
        elif kind == "tconv":
            # ConvTranspose2d(k3, s2, p1, op1) as 4 output-parity phases over the input grid: even outputs use kernel index 1
            # at offset 0; odd outputs use kernel index 0 at offset +1 and kernel index 2 at offset 0 (zero fill past the edge)
            BK, kbpt = choose_bk(self.cin_p)
            self.f_BK, self.f_kbpt = BK, kbpt
            tab = -np.ones((4 * self.rows, 4 * kbpt * BK, 1), np.int64)
            co, ci = np.meshgrid(np.arange(cout), np.arange(cin), indexing="ij")
            self.f_taps = []
            for ph in range(4):
                py, px = ph >> 1, ph & 1
                ty = [(0, 1)] if py == 0 else [(1, 0), (0, 2)]          # (offset, kernel index)
                tx = [(0, 1)] if px == 0 else [(1, 0), (0, 2)]
                taps = [(a, b) for a in ty for b in tx]
                for t in range(4):
                    if t < len(taps):
                        (dy, ky), (dx, kx) = taps[t]
                        tab[ph * self.rows + co, t * kbpt * BK + ci, 0] = self._widx(co, ci, ky, kx)
                        self.f_taps.append((dy, dx, 0))
                    else:
                        self.f_taps.append((0, 0, 0))
            self.f_ntaps, self.f_nphase = 4, 4
        else:
            BK, kbpt = choose_bk(self.cin_p)
            self.f_BK, self.f_kbpt = BK, kbpt
            if kind == "up2":
                K = 4 * kbpt * BK
                tab = -np.ones((4 * self.rows, K, 4), np.int64)
                for ph in range(4):
                    py, px = ph >> 1, ph & 1
                    for d in range(4):
                        dy, dx = d >> 1, d & 1
                        j = 0
                        for ky in _S[(py, dy)]:
                            for kx in _S[(px, dx)]:
                                co, ci = np.meshgrid(np.arange(cout), np.arange(cin), indexing="ij")
                                tab[ph * self.rows + co, d * kbpt * BK + ci, j] = self._widx(co, ci, ky, kx)
                                j += 1
                self.f_ntaps, self.f_nphase = 4, 4
            else:
                K = 9 * kbpt * BK
                tab = -np.ones((self.rows, K, 1), np.int64)
                co, ci, t = np.meshgrid(np.arange(cout), np.arange(cin), np.arange(9), indexing="ij")
                tab[co, t * kbpt * BK + ci, 0] = self._widx(co, ci, t // 3, t % 3)
                self.f_ntaps, self.f_nphase = 9, 1
        self.f_K = tab.shape[1]
        self.f_tab = self._to_dev(tab.reshape(-1, tab.shape[2]))
        self.f_w = torch.empty(tab.shape[0] * tab.shape[1], dtype=BF16, device=self.dev)

    def _build_dgrad_table(self):
        """Packed weights of the data gradient: rows = cin (the GEMM's N), K = taps x cout blocks."""
        kind, cin, cout = self.kind, self.cin, self.cout
        cout_p = round_up(cout, 8)
        BK, kbpt = choose_bk(cout_p)
        self.d_BK, self.d_kbpt = BK, kbpt
        self.d_nmma = min(256, round_up(cin, 16))
        self.d_ntile = (cin + self.d_nmma - 1) // self.d_nmma
        rows = self.d_nmma * self.d_ntile
        ci, co = np.meshgrid(np.arange(cin), np.arange(cout), indexing="ij")
        if kind in ("s1", "vgg", "vgg27"):
            nt, nph = 9, 1
            tab = -np.ones((rows, nt * kbpt * BK, 1), np.int64)
            for t in range(9):
                tab[ci, t * kbpt * BK + co, 0] = self._widx(co, ci, t // 3, t % 3)
            off = 0 if kind == "s1" else 1
            self.d_taps = [(off - t // 3, off - t % 3, 0) for t in range(9)]   # (dy, dx, plane)
        elif kind == "s2":
            nt, nph = 4, 4
            tab = -np.ones((4 * rows, nt * kbpt * BK, 1), np.int64)
            self.d_taps = []
            for ph in range(4):
                py, px = ph >> 1, ph & 1
                kys = (0, 2) if py == 0 else (1,)
                kxs = (0, 2) if px == 0 else (1,)
                taps = [(ky, kx) for ky in kys for kx in kxs]
                for t in range(4):
                    if t < len(taps):
                        ky, kx = taps[t]
                        tab[ph * rows + ci, t * kbpt * BK + co, 0] = self._widx(co, ci, ky, kx)
                        self.d_taps.append((-(ky >> 1), -(kx >> 1), 0))
                    else:
                        self.d_taps.append((0, 0, 0))            # zero-weight filler tap
        elif kind == "tconv":
            # dx[i] = sum_k dout[2i - 1 + k] w[k]: kernel index 1 reads the even plane at i, 0 / 2 the odd plane at i-1 / i
            nt, nph = 9, 1
            tab = -np.ones((rows, nt * kbpt * BK, 1), np.int64)
            self.d_taps = []
            for t in range(9):
                ky, kx = t // 3, t % 3
                tab[ci, t * kbpt * BK + co, 0] = self._widx(co, ci, ky, kx)
                self.d_taps.append((-1 if ky == 0 else 0, -1 if kx == 0 else 0, (ky != 1) * 2 + (kx != 1)))
        elif kind == "up2":
            nt, nph = 16, 1
            tab = -np.ones((rows, nt * kbpt * BK, 4), np.int64)
            self.d_taps = []
            for ph in range(4):
                py, px = ph >> 1, ph & 1
                for d in range(4):
                    dy, dx = d >> 1, d & 1
                    j = 0
                    for ky in _S[(py, dy)]:
                        for kx in _S[(px, dx)]:
                            tab[ci, (ph * 4 + d) * kbpt * BK + co, j] = self._widx(co, ci, ky, kx)
                            j += 1
                    self.d_taps.append((-(py + dy), -(px + dx), ph))
        else:
            raise _lib.VstError(f"no tensor-core data gradient for kind {kind}")
        self.d_ntaps, self.d_nphase, self.d_rows, self.d_K = nt, nph, tab.shape[0], tab.shape[1]
        self.d_tab = self._to_dev(tab.reshape(-1, tab.shape[2]))
        self.d_w = torch.empty(tab.shape[0] * tab.shape[1], dtype=BF16, device=self.dev)

    def _build_wgrad_table(self):
        """D[t][co][n] (pixel-contraction output) -> dw[co][ci][ky][kx]."""
        kind, cin, cout, k = self.kind, self.cin, self.cout, self.k
        co, ci, ky, kx = np.meshgrid(np.arange(cout), np.arange(cin), np.arange(k), np.arange(k), indexing="ij")
        if kind in ("s1", "s2"):
            self.w_N = cin
            self.w_taps = []
            for t in range(9):
                ty, tx = t // 3, t % 3
                if kind == "s1":
                    self.w_taps.append(((0, 0, 0), (ty, tx, 0)))                      # (A tap, B tap) = (dy, dx, plane)
                else:
                    self.w_taps.append(((0, 0, 0), (ty >> 1, tx >> 1, (ty & 1) * 2 + (tx & 1))))
            tab = (((ky * 3 + kx) * cout + co) * cin + ci)[..., None]
        elif kind == "up2":
            self.w_N = cin
            self.w_taps = []
            for ph in range(4):
                py, px = ph >> 1, ph & 1
                for d in range(4):
                    self.w_taps.append(((0, 0, ph), (py + (d >> 1), px + (d & 1), 0)))
            tab = -np.ones((cout, cin, 3, 3, 4), np.int64)
            cnt = np.zeros((3, 3), np.int64)
            for ph in range(4):
                py, px = ph >> 1, ph & 1
                for d in range(4):
                    for a in _S[(py, d >> 1)]:
                        for b in _S[(px, d & 1)]:
                            tab[:, :, a, b, cnt[a, b]] = (((ph * 4 + d) * cout + co[:, :, 0, 0]) * cin + ci[:, :, 0, 0])
                            cnt[a, b] += 1
            assert (cnt == 4).all()
        elif kind == "tconv":
            self.w_N = cin
            self.w_taps = []
            for t in range(9):
                ty, tx = t // 3, t % 3
                self.w_taps.append(((-1 if ty == 0 else 0, -1 if tx == 0 else 0, (ty != 1) * 2 + (tx != 1)), (0, 0, 0)))
            tab = (((ky * 3 + kx) * cout + co) * cin + ci)[..., None]          # dw_t[ci][co][ky][kx] <- D[t][co][ci]
            tab = np.transpose(tab, (1, 0, 2, 3, 4))
        elif kind == "row9":
            self.w_N = 9 * cin                                    # columns (kx, c) of the X9 operand
            self.w_taps = [((0, 0, 0), (t, 0, 0)) for t in range(9)]
            tab = ((ky * cout + co) * self.w_N + kx * cin + ci)[..., None]
        else:
            raise _lib.VstError(f"no tensor-core weight gradient for kind {kind}")
        self.w_tab = self._to_dev(tab.reshape(-1, tab.shape[-1]))
        self.w_D = torch.empty(len(self.w_taps) * cout * self.w_N, dtype=torch.float32, device=self.dev)

    # ---- per-step weight packing --------------------------------------------------------------------
    def pack(self, w: torch.Tensor, dgrad=True):
        gather_sum(w, self.f_tab, self.f_w)
        if dgrad and hasattr(self, "d_tab"):
            gather_sum(w, self.d_tab, self.d_w)

    # ---- forward ----------------------------------------------------------------------------------------
    def forward(self, x: Act, out_raw: torch.Tensor, out_hw, stats: Optional[torch.Tensor] = None, bias=None, relu=False):
        """x: the layer's operand Act (layout per kind); out_raw: bf16 [N, Ho, Wo, cout_p] (cout_p = out_raw channel stride)."""
        d = self.fwd_desc(x, out_raw, out_hw, stats, bias, relu)
        check(_lib.lib().vst_tc_tapgemm(C.byref(d), _stream()), f"vst_tc_tapgemm(fwd {self.kind})")
        return out_raw

    def fwd_desc(self, x: Act, out_raw: torch.Tensor, out_hw, stats=None, bias=None, relu=False) -> TapGemmDesc:
        Ho, Wo = out_hw
        d = TapGemmDesc()
        a_C, a_X, a_Y, a_N, a_P = x.dims()
        d.a, d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P = x.ptr(), a_C, a_X, a_Y, a_N, a_P
        d.b, d.b_K, d.b_rows = self.f_w.data_ptr(), self.f_K, self.f_w.numel() // self.f_K
        d.BK, d.kb_per_tap, d.n_taps, d.n_phase = self.f_BK, self.f_kbpt, self.f_ntaps, self.f_nphase
        d.n_ntile, d.N_mma = self.n_ntile, self.n_mma
        kind = self.kind
        if kind == "tconv":
            d.grid_h, d.grid_w, d.out_mul = Ho // 2, Wo // 2, 2
            taps = self.f_taps
            d.ph_oy, d.ph_ox = _i8([0, 0, 1, 1], 4), _i8([0, 1, 0, 1], 4)
        elif kind == "up2":
            d.grid_h, d.grid_w, d.out_mul = Ho // 2, Wo // 2, 2
            taps, oy, ox = [], [], []
            for ph in range(4):
                py, px = ph >> 1, ph & 1
                oy.append(py), ox.append(px)
                taps += [(py + (t >> 1), px + (t & 1), 0) for t in range(4)]
            d.ph_oy, d.ph_ox = _i8(oy, 4), _i8(ox, 4)
        else:
            d.grid_h, d.grid_w, d.out_mul = Ho, Wo, 1
            if kind == "s1":
                taps = [(t // 3, t % 3, 0) for t in range(9)]
            elif kind == "vgg":
                taps = [(t // 3 - 1, t % 3 - 1, 0) for t in range(9)]
            elif kind == "vgg27":
                taps = [(0, 0, 0)]
            elif kind == "s2":
                taps = [((t // 3) >> 1, (t % 3) >> 1, ((t // 3) & 1) * 2 + ((t % 3) & 1)) for t in range(9)]
            else:  # row9
                taps = [(t, 0, 0) for t in range(9)]
        d.tap_dy, d.tap_dx, d.tap_pl = _i8(t[0] for t in taps), _i8(t[1] for t in taps), _i8(t[2] for t in taps)
        cstride = out_raw.numel() // (x.N * Ho * Wo)
        d.Hout, d.Wout, d.Cout, d.out_cstride = Ho, Wo, min(cstride, self.rows), cstride
        d.epi_mode, d.relu = EPI_BF16, int(relu)
        d.out = out_raw.data_ptr()
        d.bias = None if bias is None else bias.data_ptr()
        d.stats = None if stats is None else stats.data_ptr()
        return d

    # ---- data gradient ------------------------------------------------------------------------------------
    def dgrad(self, draw: Act, in_hw, out: Optional[torch.Tensor] = None, out_f32_nchw: Optional[torch.Tensor] = None):
        """draw: gradient w.r.t. the conv output (Act, pad 0; parity planes for "up2").  Returns the gradient over the
        layer's PADDED input domain [N][Hs+2p][Ws+2p][cin_p] bf16 (p = 1; "vgg": p = 0, i.e. the input itself)."""
        Hs, Ws = in_hw
        p = 0 if self.kind in ("vgg", "vgg27", "tconv") else 1
        Hp, Wp = Hs + 2 * p, Ws + 2 * p
        if out is None and out_f32_nchw is None:
            out = torch.empty(draw.N * Hp * Wp * self.cin_p, dtype=BF16, device=draw.t.device)
        d = self.dgrad_desc(draw, in_hw, out, out_f32_nchw)
        check(_lib.lib().vst_tc_tapgemm(C.byref(d), _stream()), f"vst_tc_tapgemm(dgrad {self.kind})")
        return out if out_f32_nchw is None else out_f32_nchw

    def dgrad_desc(self, draw: Act, in_hw, out, out_f32_nchw=None) -> TapGemmDesc:
        Hs, Ws = in_hw
        p = 0 if self.kind in ("vgg", "vgg27", "tconv") else 1
        Hp, Wp = Hs + 2 * p, Ws + 2 * p
        d = TapGemmDesc()
        a_C, a_X, a_Y, a_N, a_P = draw.dims()
        d.a, d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P = draw.ptr(), a_C, a_X, a_Y, a_N, a_P
        d.b, d.b_K, d.b_rows = self.d_w.data_ptr(), self.d_K, self.d_rows
        d.BK, d.kb_per_tap, d.n_taps, d.n_phase = self.d_BK, self.d_kbpt, self.d_ntaps, self.d_nphase
        d.n_ntile, d.N_mma = self.d_ntile, self.d_nmma
        if self.kind == "s2":
            d.grid_h, d.grid_w, d.out_mul = Hp // 2, Wp // 2, 2
            d.ph_oy, d.ph_ox = _i8([0, 0, 1, 1], 4), _i8([0, 1, 0, 1], 4)
        else:
            d.grid_h, d.grid_w, d.out_mul = Hp, Wp, 1
        d.tap_dy, d.tap_dx, d.tap_pl = _i8(t[0] for t in self.d_taps), _i8(t[1] for t in self.d_taps), _i8(t[2] for t in self.d_taps)
        if out_f32_nchw is not None:
            d.epi_mode, d.out, d.Cout, d.out_cstride = EPI_F32_NCHW, out_f32_nchw.data_ptr(), self.cin, self.cin
        else:
            d.epi_mode, d.out, d.Cout, d.out_cstride = EPI_BF16, out.data_ptr(), self.cin_p, self.cin_p
        d.Hout, d.Wout = Hp, Wp
        return d

    # ---- weight gradient ------------------------------------------------------------------------------------
    def wgrad(self, draw: Act, x: Act, out_hw, dw_out: torch.Tensor):
        """dw_out (fp32 [cout, cin, k, k], e.g. a view of the flat gradient buffer) <- sum over pixels and images."""
        d = self.wgrad_desc(draw, x, out_hw)
        if not getattr(self, "_wd_pooled", False):
            self.w_D.zero_()
        check(_lib.lib().vst_tc_pcgemm(C.byref(d), _stream()), f"vst_tc_pcgemm(wgrad {self.kind})")
        if not dw_out.is_contiguous() or dw_out.dtype != torch.float32:
            raise _lib.VstError("wgrad: dw_out must be contiguous float32")
        return gather_sum(self.w_D, self.w_tab, dw_out)

    def wgrad_desc(self, draw: Act, x: Act, out_hw) -> PcGemmDesc:
        Ho, Wo = out_hw
        d = PcGemmDesc()
        d.a, (d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P) = draw.ptr(), draw.dims()
        d.b, (d.b_C, d.b_X, d.b_Y, d.b_N, d.b_P) = x.ptr(), x.dims()
        d.n_img = draw.N
        d.grid_h, d.grid_w = (Ho // 2, Wo // 2) if self.kind in ("up2", "tconv") else (Ho, Wo)
        d.n_taps, d.M, d.N, d.per_image, d.k_splits, d.scale = len(self.w_taps), self.cout, self.w_N, 0, 0, 1.0
        d.out = self.w_D.data_ptr()
        d.a_dy, d.a_dx, d.a_pl = _i8(t[0][0] for t in self.w_taps), _i8(t[0][1] for t in self.w_taps), _i8(t[0][2] for t in self.w_taps)
        d.b_dy, d.b_dx, d.b_pl = _i8(t[1][0] for t in self.w_taps), _i8(t[1][1] for t in self.w_taps), _i8(t[1][2] for t in self.w_taps)
        return d


class RowConvOutTC:
    """k x k reflect-padded convolution with few output channels (ConvTanh: 48 -> 3, k = 9, RC/network.py:78-85) as a
    ROW convolution: GEMM columns are (kx, co), K runs over (ky, c); the epilogue sums the kx-shifted columns
    (csrc/tc_conv.cu TG_EPI_ROWCONV).  Its adjoints use the expanded gradient E[y][x'][(kx,co)] = dz[y][x'-kx][co]:
    weight gradient = 9-tap pixel contraction of E with the padded input, data gradient = 9-tap GEMM over E."""

    KE = 32

    def __init__(self, cin: int, cout: int, k: int, device):
        if k * cout > self.KE or cin % 8:
            raise _lib.VstError("RowConvOutTC: k*cout must be <= 32 and cin a multiple of 8")
        self.cin, self.cout, self.k, self.dev = cin, cout, k, device
        BK, kbpt = choose_bk(cin)
        self.f_BK, self.f_kbpt = BK, kbpt
        kt = kbpt * BK
        co, c, ky, kx = np.meshgrid(np.arange(cout), np.arange(cin), np.arange(k), np.arange(k), indexing="ij")
        widx = ((co * cin + c) * k + ky) * k + kx
        tab = -np.ones((self.KE, k * kt), np.int64)
        tab[kx * cout + co, ky * kt + c] = widx
        self.f_K = k * kt
        self.f_tab = torch.from_numpy(tab.reshape(-1, 1).astype(np.int32)).to(device)
        self.f_w = torch.empty(self.KE * self.f_K, dtype=BF16, device=device)
        # data gradient: rows = input channel c (N of the GEMM), K = ky * KE + (kx, co)
        self.d_nmma = round_up(cin, 16)
        # taps are stored in REVERSED ky order (tap t' <-> ky = k-1-t', row offset t' - (k-1)) so that the row offsets
        # increase with the tap index: a pure row stencil, eligible for the row-streaming kernel mode
        tab = -np.ones((self.d_nmma, k * self.KE), np.int64)
        tab[c, (k - 1 - ky) * self.KE + kx * cout + co] = widx
        self.d_K = k * self.KE
        self.d_tab = torch.from_numpy(tab.reshape(-1, 1).astype(np.int32)).to(device)
        self.d_w = torch.empty(self.d_nmma * self.d_K, dtype=BF16, device=device)
        # weight gradient: D[ky][(kx,co)][c] -> dw[co][c][ky][kx]
        self.w_M = k * cout
        self.w_tab = torch.from_numpy((((ky * self.w_M + kx * cout + co) * cin + c).reshape(-1, 1)).astype(np.int32)).to(device)
        self.w_D = torch.empty(k * self.w_M * cin, dtype=torch.float32, device=device)

    def pack(self, w: torch.Tensor, dgrad=True):
        gather_sum(w, self.f_tab, self.f_w)
        if dgrad:
            gather_sum(w, self.d_tab, self.d_w)

    def fwd_desc(self, x: Act, img_out: torch.Tensor, bias, act: int) -> TapGemmDesc:
        k = self.k
        H, W = x.H, x.W
        d = TapGemmDesc()
        d.a, (d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P) = x.ptr(), x.dims()
        d.b, d.b_K, d.b_rows = self.f_w.data_ptr(), self.f_K, self.KE
        d.BK, d.kb_per_tap, d.n_taps, d.n_phase, d.n_ntile, d.N_mma = self.f_BK, self.f_kbpt, k, 1, 1, self.KE
        d.grid_h, d.grid_w, d.out_mul = H, W, 1
        d.Hout, d.Wout, d.Cout, d.out_cstride = H, W, self.cout, self.cout
        d.epi_mode, d.act, d.rc_k, d.rc_co = EPI_ROWCONV, act, k, self.cout
        d.tile_step_x, d.TW, d.TH, d.MT = 128 - k + 1, 128, 2, 2       # 2 output rows x 120 pixels per CTA tile
        d.out = img_out.data_ptr()
        d.bias = None if bias is None else bias.data_ptr()
        d.tap_dy, d.tap_dx, d.tap_pl = _i8(range(k)), _i8([0] * k), _i8([0] * k)
        return d

    def forward(self, x: Act, img_out: torch.Tensor, bias, act: int):
        """x: Act padded by k//2 (reflect); img_out: fp32 NCHW [N, cout, H, W] = act(conv + bias)."""
        d = self.fwd_desc(x, img_out, bias, act)
        check(_lib.lib().vst_tc_tapgemm(C.byref(d), _stream()), "vst_tc_tapgemm(rowconv)")
        return img_out

    def expand(self, dz: torch.Tensor) -> Act:
        """dz fp32 NCHW [N, cout, H, W] (gradient w.r.t. conv + bias) -> E as an Act [N][H][W+k-1][32]."""
        N, Co, H, W = dz.shape
        E = Act(N, H, W + self.k - 1, self.KE, device=dz.device)
        check(_lib.lib().vst_tc_rowconv_expand(dz.data_ptr(), E.ptr(), N, Co, H, W, self.k, self.KE, _stream()), "vst_tc_rowconv_expand")
        return E

    def dgrad_desc(self, E: Act, out: torch.Tensor) -> TapGemmDesc:
        k, p = self.k, self.k // 2
        H, Wp = E.H, E.W
        d = TapGemmDesc()
        d.a, (d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P) = E.ptr(), E.dims()
        d.b, d.b_K, d.b_rows = self.d_w.data_ptr(), self.d_K, self.d_nmma
        d.BK, d.kb_per_tap, d.n_taps, d.n_phase, d.n_ntile, d.N_mma = 32, 1, k, 1, 1, self.d_nmma
        d.grid_h, d.grid_w, d.out_mul = H + 2 * p, Wp, 1
        d.Hout, d.Wout, d.Cout, d.out_cstride = H + 2 * p, Wp, self.cin, self.cin
        d.epi_mode, d.out = EPI_BF16, out.data_ptr()
        d.tap_dy, d.tap_dx, d.tap_pl = _i8(t - (k - 1) for t in range(k)), _i8([0] * k), _i8([0] * k)
        return d

    def dgrad(self, E: Act, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """-> gradient over the padded input domain [N][H+2p][W+2p][cin] bf16 (fold with reflect, pad k//2)."""
        p = self.k // 2
        if out is None:
            out = torch.empty(E.N * (E.H + 2 * p) * E.W * self.cin, dtype=BF16, device=E.t.device)
        d = self.dgrad_desc(E, out)
        check(_lib.lib().vst_tc_tapgemm(C.byref(d), _stream()), "vst_tc_tapgemm(rowconv dgrad)")
        return out

    def wgrad_desc(self, E: Act, x: Act) -> PcGemmDesc:
        k = self.k
        d = PcGemmDesc()
        d.a, (d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P) = E.ptr(), E.dims()
        d.b, (d.b_C, d.b_X, d.b_Y, d.b_N, d.b_P) = x.ptr(), x.dims()
        d.n_img, d.grid_h, d.grid_w = E.N, E.H, E.W
        d.n_taps, d.M, d.N, d.per_image, d.k_splits, d.scale = k, self.w_M, self.cin, 0, 0, 1.0
        d.out = self.w_D.data_ptr()
        d.a_dy, d.a_dx, d.a_pl = _i8([0] * k), _i8([0] * k), _i8([0] * k)
        d.b_dy, d.b_dx, d.b_pl = _i8(range(k)), _i8([0] * k), _i8([0] * k)
        return d

    def wgrad(self, E: Act, x: Act, dw_out: torch.Tensor):
        d = self.wgrad_desc(E, x)
        if not getattr(self, "_wd_pooled", False):
            self.w_D.zero_()
        check(_lib.lib().vst_tc_pcgemm(C.byref(d), _stream()), "vst_tc_pcgemm(rowconv wgrad)")
        return gather_sum(self.w_D, self.w_tab, dw_out)


# =============================================================================================
# Gram matrices
# =============================================================================================
def gram(f: Act, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[N, C, C] fp32 = scale * F F^T per image (RC/utilities.py:93-98); f: plain unpadded Act."""
    if out is None:
        out = torch.zeros((f.N, f.C, f.C), dtype=torch.float32, device=f.t.device)
    else:
        out.zero_()
    d = PcGemmDesc()
    d.a, (d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P) = f.ptr(), f.dims()
    d.b, (d.b_C, d.b_X, d.b_Y, d.b_N, d.b_P) = f.ptr(), f.dims()
    d.n_img, d.grid_h, d.grid_w = f.N, f.H, f.W
    d.n_taps, d.M, d.N, d.per_image, d.k_splits, d.scale = 1, f.C, f.C, 1, 0, float(scale)
    d.out = out.data_ptr()
    check(_lib.lib().vst_tc_pcgemm(C.byref(d), _stream()), "vst_tc_pcgemm(gram)")
    return out


def gram_bwd(f: Act, G: torch.Tensor, Gs: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dF [N][H][W][C] bf16 for L = c * sum (G - Gs)^2 with G = s * F F^T: `scale` = 2 * c * s."""
    N, Cc = f.N, f.C
    if Cc % 64:
        raise _lib.VstError("gram_bwd: channels must be a multiple of 64")
    S = torch.empty((N, Cc, Cc), dtype=BF16, device=f.t.device)
    check(_lib.lib().vst_tc_gram_grad_weights(G.data_ptr(), Gs.data_ptr(), Gs.shape[0], float(scale), S.data_ptr(), N, Cc, _stream()),
          "vst_tc_gram_grad_weights")
    if out is None:
        out = torch.empty(N * f.H * f.W * Cc, dtype=BF16, device=f.t.device)
    d = TapGemmDesc()
    d.a, (d.a_C, d.a_X, d.a_Y, d.a_N, d.a_P) = f.ptr(), f.dims()
    d.b, d.b_K, d.b_rows, d.b_img_rows = S.data_ptr(), Cc, N * Cc, Cc
    d.BK, d.kb_per_tap, d.n_taps, d.n_phase = 64, Cc // 64, 1, 1
    d.N_mma = min(256, Cc)
    d.n_ntile = Cc // d.N_mma
    d.grid_h, d.grid_w, d.out_mul = f.H, f.W, 1
    d.Hout, d.Wout, d.Cout, d.out_cstride = f.H, f.W, Cc, Cc
    d.epi_mode, d.out = EPI_BF16, out.data_ptr()
    d.tap_dy, d.tap_dx, d.tap_pl = _i8([0]), _i8([0]), _i8([0])
    check(_lib.lib().vst_tc_tapgemm(C.byref(d), _stream()), "vst_tc_tapgemm(gram_bwd)")
    return out


# =============================================================================================
# element-wise companions
# =============================================================================================
def in_apply(raw: torch.Tensor, stats, gamma, beta, dst: Act, relu: bool, residual: Optional[Act] = None, eps=1e-5):
    rd = residual.desc if residual is not None else dst.desc
    check(_lib.lib().vst_tc_in_apply(raw.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                     None if residual is None else residual.ptr(), rd, dst.ptr(), dst.desc, dst.N, eps, int(relu),
                                     _stream()), "vst_tc_in_apply")
    return dst


def in_bwd(G: torch.Tensor, g_desc: ActDesc, raw: torch.Tensor, stats, gamma, beta, draw: Act, relu: bool, red: torch.Tensor,
           dgamma: torch.Tensor, dbeta: torch.Tensor, skip: Optional[torch.Tensor] = None, gsum: Optional[torch.Tensor] = None,
           eps=1e-5):
    """InstanceNorm(+ReLU) backward: G (bf16, padded domain described by g_desc) -> draw; dgamma / dbeta written."""
    L = _lib.lib()
    N = draw.N
    sk = None if skip is None else skip.data_ptr()
    check(L.vst_tc_in_bwd_reduce(G.data_ptr(), g_desc, sk, raw.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                 red.data_ptr(), N, eps, int(relu), _stream()), "vst_tc_in_bwd_reduce")
    check(L.vst_tc_in_bwd_apply(G.data_ptr(), g_desc, sk, raw.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                red.data_ptr(), draw.ptr(), draw.desc, None if gsum is None else gsum.data_ptr(), N, eps,
                                int(relu), _stream()), "vst_tc_in_bwd_apply")
    check(L.vst_tc_in_param_grads(red.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), N, draw.C, _stream()), "vst_tc_in_param_grads")
    return draw


def maxpool2(x: Act) -> Act:
    y = Act(x.N, x.H // 2, x.W // 2, x.C, device=x.t.device)
    check(_lib.lib().vst_tc_maxpool2(x.ptr(), y.ptr(), x.N, x.H, x.W, x.C, _stream()), "vst_tc_maxpool2")
    return y


def relu_pool_bwd(g: torch.Tensor, y: Act, add: Optional[torch.Tensor], pooled: bool, out: Optional[Act] = None) -> Act:
    gm = out or Act(y.N, y.H, y.W, y.C, device=y.t.device)
    check(_lib.lib().vst_tc_relu_pool_bwd(g.data_ptr(), y.ptr(), None if add is None else add.data_ptr(), gm.ptr(), y.N, y.H, y.W, y.C,
                                          int(pooled), _stream()), "vst_tc_relu_pool_bwd")
    return gm


_bf16_scratch = {}


def _scratch(device):
    key = (str(device), _stream())
    if key not in _bf16_scratch:
        _bf16_scratch[key] = torch.zeros(1024 + 8, dtype=torch.float32, device=device)
    return _bf16_scratch[key]


def sqdiff_sum(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor):
    check(_lib.lib().vst_tc_sqdiff_sum_bf16(a.data_ptr(), b.data_ptr(), out.data_ptr(), _scratch(a.device).data_ptr(), a.numel(),
                                            _stream()), "vst_tc_sqdiff_sum_bf16")
    return out


def sqdiff_bwd(a: torch.Tensor, b: torch.Tensor, scale: float) -> torch.Tensor:
    da = torch.empty_like(a)
    check(_lib.lib().vst_tc_sqdiff_bwd_bf16(a.data_ptr(), b.data_ptr(), float(scale), da.data_ptr(), a.numel(), _stream()),
          "vst_tc_sqdiff_bwd_bf16")
    return da


def add_(y: torch.Tensor, x: torch.Tensor):
    check(_lib.lib().vst_tc_add_bf16(x.data_ptr(), y.data_ptr(), y.numel(), _stream()), "vst_tc_add_bf16")
    return y


def prologue_x27(x: torch.Tensor) -> Act:
    """fp32 NCHW [N,3,H,W] -> the first VGG convolution's im2col operand [N][H][W][32]."""
    N, Cin, H, W = x.shape
    if Cin != 3:
        raise _lib.VstError("prologue_x27: 3-channel input expected")
    a = Act(N, H, W, 32, device=x.device)
    check(_lib.lib().vst_tc_prologue_x27(x.contiguous().data_ptr(), a.ptr(), N, H, W, _stream()), "vst_tc_prologue_x27")
    return a


def prologue_x9(x: torch.Tensor, KR: int) -> Act:
    """fp32 NCHW frames -> the conv1 operand X9 [N][H+8][W][KR] (an Act with H+8 rows, pad 0)."""
    N, Cin, H, W = x.shape
    a = Act(N, H + 8, W, KR, device=x.device)
    check(_lib.lib().vst_tc_prologue_x9(x.data_ptr(), a.ptr(), N, Cin, H, W, KR, _stream()), "vst_tc_prologue_x9")
    return a
