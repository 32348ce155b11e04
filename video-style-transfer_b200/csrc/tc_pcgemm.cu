// Pixel-contraction GEMM on tcgen05 tensor cores (sm_100a): the contraction index is the PIXEL.
//
//   D_t[m][n] (+)= scale * sum_{img, y, x}  A[img][plA_t][y + dyA_t][x + dxA_t][m] * B[img][plB_t][y + dyB_t][x + dxB_t][n]
//
// Two users:
//   * convolution weight gradients: A = dL/d(conv output) (NHWC bf16), B = the layer's padded input
//     activation, one tap t per filter position -> dW_t[co][ci];
//   * Gram matrices F F^T: A = B = the feature map, one tap, one output per image.
//
// Both operands are channels-last, so a TMA box of P pixels x 64 channels lands in shared memory as P rows
// of 128 bytes - exactly the canonical *MN-major* SWIZZLE_128B operand layout of tcgen05.mma (rows are the K
// index, 8-row swizzle atoms are SBO = 1024 bytes apart, 64-channel chunks LBO apart).  No transposed
// copies of the activations are ever made: the instruction descriptor's a_major/b_major bits do the work.
// Channel counts that are not multiples of 64 use 32- or 16-channel chunks (SWIZZLE_64B / 32B).
//
// M-chunk mode (weight gradients whose taps all read the SAME A tile and whose M is not a multiple of 128, i.e. the
// 192 -> 192 trunk): the operands swap roles inside the kernel - the shifted input activation becomes the M side, the
// output gradient the (single, shared) N operand - and the M space is the concatenation of (tap, 64-channel chunk)
// pieces: 9 taps x 3 chunks = 27 chunks = 13.5 tiles of 128 rows instead of 9 taps x (128 + 64 half-empty) rows.  Each
// 64-channel chunk of an M tile is its own TMA box (own tap offset); a CTA owns up to 512 / N_mma consecutive M tiles,
// which share the one N-operand box per K step.  The output keeps the [tap][M][N] layout (the epilogue transposes).
//
// CTA = one (image?, M tile of 128, N tile, tap, K split); K = a strided subset of the 64-pixel tiles.
// Warp roles: 0 A producer, 1 MMA issuer (+TMEM alloc), 2..5 epilogue, 6 B producer.  Partial sums of the
// K splits are combined with fp32 atomics into the caller-zeroed output.
#include <stdlib.h>
#include "tc_conv.cuh"
#include "tc_ptx.cuh"

namespace vst {

constexpr int PC_THREADS = 224;
constexpr int PC_PK = 64;          // pixels (K) per pipeline stage
constexpr int PC_MAX_STAGES = 12;   // round 2: was 4 - a narrow Gram (8 KB of new data per stage) then had 32 KB in flight per SM and
                                     // sat at 3.4 TB/s, the latency bound of that depth (C5 sweep relu1_1 at 0.5 of HBM)

struct PcGemmParams {
  CUtensorMap tmA, tmB;  // 5-D (c, X, Y, chunk, img*plane) bf16; box (cw, TW, TH, chunks, 1)
  int cwA, cwB;          // channels per chunk (64 / 32 / 16)
  int a_chunks, b_chunks;  // chunks per box: a_chunks*cwA == 128 (M_mma), b_chunks*cwB == N_mma
  int N_mma;
  int m_tiles, n_tiles, n_taps, n_out_img;  // n_out_img = n_img when per_image else 1
  int n_img, tiles_x, tiles_y, TW, TH;      // TW*TH == PC_PK
  int k_splits, per_image;
  int n_groups, tpc, stages;               // tap groups (taps sharing one A tile), max taps per group, pipeline stages
  short grp_first[TG_MAX_TAPS], grp_cnt[TG_MAX_TAPS];
  int M, N;              // real extents of one output matrix
  float scale;
  float* out;            // [n_out_img][n_taps][M][N] fp32, accumulated
  int tapA[TG_MAX_TAPS], tapB[TG_MAX_TAPS];  // (dx & 0xff) | (dy & 0xff) << 8 | plane << 16
  int dbg;
  // M-chunk mode: kernel A = caller's B (tap-shifted, one TMA box per chunk), kernel B = caller's A (tap tapB[0])
  int xfast;                     // K tiles walk x fastest (single-tap contractions)
  int alias_b;                   // Gram with C <= 128: B is the A tile itself (one TMA box per K step, no second producer)
  int mchunk, rows_total, cpt;   // rows_total = n_taps * N (caller), cpt = chunks per tap = N / cwA
};

// MN-major smem descriptor: LBO = bytes between channel chunks, SBO = bytes between 8-pixel groups.
__device__ __forceinline__ uint64_t smem_desc_mn_hi(int cw, uint32_t lbo_bytes, int swap) {
  const int row_bytes = cw * 2;
  const uint64_t layout = row_bytes == 128 ? 2 : row_bytes == 64 ? 4 : 6;
  uint64_t lbo = lbo_bytes >> 4, sbo = (uint64_t)(8 * row_bytes) >> 4;
  if (swap) { const uint64_t t = lbo; lbo = sbo; sbo = t; }
  return (lbo << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

__global__ void __launch_bounds__(PC_THREADS, 1) pcgemm_kernel(const __grid_constant__ PcGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = 128 * PC_PK * 2;                 // 16 KB
  const int b_bytes = p.N_mma * PC_PK * 2;             // <= 32 KB per tap
  const int b_al = (b_bytes + 1023) & ~1023;
  const int stage_bytes = p.alias_b ? a_bytes : p.mchunk ? p.tpc * a_bytes + b_al : a_bytes + p.tpc * b_al;
  const int S = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint64_t* empty = full + PC_MAX_STAGES;
  uint64_t* accfull = empty + PC_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- which output block and which K tiles
  int bid = blockIdx.x;
  const int split = bid % p.k_splits; bid /= p.k_splits;
  const int grp = bid % p.n_groups; bid /= p.n_groups;
  // plain: taps tap .. tap+cnt-1 share the A tile.  M-chunk: M tiles tap .. tap+cnt-1 (of the chunked M space) share the B tile
  const int tap = p.mchunk ? grp * p.tpc : p.grp_first[grp];
  const int cnt = p.mchunk ? min(p.tpc, p.m_tiles - tap) : p.grp_cnt[grp];
  const int nt = p.mchunk ? 0 : bid % p.n_tiles; if (!p.mchunk) bid /= p.n_tiles;
  const int mt = p.mchunk ? 0 : bid % p.m_tiles; if (!p.mchunk) bid /= p.m_tiles;
  const int oimg = p.mchunk ? 0 : bid;   // 0 unless per_image
  const int tiles_img = p.tiles_x * p.tiles_y;
  const int k_total = (p.per_image ? 1 : p.n_img) * tiles_img;
  const int per = (k_total + p.k_splits - 1) / p.k_splits;             // contiguous K range per split
  const int k_begin = split * per;
  const int my_k = max(0, min(per, k_total - k_begin));

  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.tpc * p.N_mma) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], p.alias_b ? 1 : 2);
      mbar_init(&empty[s], 1);
    }
    mbar_init(accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_s = smem_u32(smem), full_s = smem_u32(full), empty_s = smem_u32(empty);
  pdl_wait();   // programmatic dependent launch: the prologue above overlaps the previous kernel's tail (common.cuh)
  pdl_trigger();

  if (warp == 0 || warp == 6) {
    // ================================ producers: A (warp 0) / B (warp 6) ================
    if (p.mchunk) {
      if (elect_one() && my_k > 0) {
        const bool isA = warp == 0;
        const int chunk_bytes = PC_PK * p.cwA * 2;
        // A side: up to tpc * a_chunks chunk boxes per K step, each with its own tap offset; B side: one box
        int cdx[16], cdy[16], cpl[16], ccc[16];
        uint32_t coff[16];
        int nload = 0;
        if (isA) {
          for (int j = 0; j < cnt; ++j)
            for (int c = 0; c < p.a_chunks; ++c) {
              const int g = (tap + j) * p.a_chunks + c;          // global chunk index in the (tap, channel chunk) M space
              if (g >= p.n_taps * p.cpt) continue;                 // tail of the last tile: rows never stored
              const int tp = p.tapA[g / p.cpt];
              cdx[nload] = (int)(signed char)(tp & 0xff); cdy[nload] = (int)(signed char)((tp >> 8) & 0xff); cpl[nload] = tp >> 16;
              ccc[nload] = g % p.cpt;
              coff[nload] = (uint32_t)(j * a_bytes + c * chunk_bytes);
              ++nload;
            }
        } else {
          const int tp = p.tapB[0];
          cdx[0] = (int)(signed char)(tp & 0xff); cdy[0] = (int)(signed char)((tp >> 8) & 0xff); cpl[0] = tp >> 16;
          ccc[0] = 0;
          coff[0] = (uint32_t)(p.tpc * a_bytes);
          nload = 1;
        }
        const CUtensorMap* tm = isA ? &p.tmA : &p.tmB;
        const uint32_t bytes = isA ? (uint32_t)(nload * chunk_bytes) : (uint32_t)b_bytes;
        int s = 0;
        uint32_t ph = 0;
        for (int i = 0, kt = k_begin; i < my_k; ++i, ++kt) {
          const int ty = kt % p.tiles_y, tx = (kt / p.tiles_y) % p.tiles_x;
          const int n = kt / tiles_img;
          mbar_wait_a(empty_s + s * 8, ph ^ 1);
          const uint32_t bar = full_s + s * 8;
          mbar_expect_tx_a(bar, bytes);
          const uint32_t base = smem_s + s * stage_bytes;
          for (int j = 0; j < nload; ++j)
            tma_load_5d_a(base + coff[j], tm, bar, 0, tx * p.TW + cdx[j], ty * p.TH + cdy[j], ccc[j], cpl[j] * p.n_img + n);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    } else if (elect_one() && my_k > 0 && !(p.alias_b && warp == 6)) {
      const bool isA = warp == 0;
      const CUtensorMap* tm = isA ? &p.tmA : &p.tmB;
      const int chunk0 = isA ? mt * p.a_chunks : nt * p.b_chunks;
      const int nload = isA ? 1 : cnt;
      int dx[TG_MAX_TAPS > 16 ? 16 : TG_MAX_TAPS], dy[16], pl[16];
      for (int j = 0; j < nload; ++j) {
        const int tp = isA ? p.tapA[tap] : p.tapB[tap + j];
        dx[j] = (int)(signed char)(tp & 0xff); dy[j] = (int)(signed char)((tp >> 8) & 0xff); pl[j] = tp >> 16;
      }
      const uint32_t bytes = isA ? (uint32_t)a_bytes : (uint32_t)(cnt * b_bytes);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0, kt = k_begin; i < my_k; ++i, ++kt) {
        // K tiles walk DOWN a 64-pixel column strip (y fastest): taps that differ by a row offset re-read the rows of the
        // previous steps while they are still in L2 (x-fastest put a whole image row of tiles between the two uses)
        // (a single-tap contraction - a Gram - has no such reuse: it walks x fastest, i.e. straight through memory)
        const int ty = p.xfast ? (kt / p.tiles_x) % p.tiles_y : kt % p.tiles_y, tx = p.xfast ? kt % p.tiles_x : (kt / p.tiles_y) % p.tiles_x;
        const int n = p.per_image ? oimg : kt / tiles_img;
        mbar_wait_a(empty_s + s * 8, ph ^ 1);
        const uint32_t bar = full_s + s * 8;
        mbar_expect_tx_a(bar, bytes);
        uint32_t dst = smem_s + s * stage_bytes + (isA ? 0u : (uint32_t)a_bytes);
        for (int j = 0; j < nload; ++j) {
          tma_load_5d_a(dst, tm, bar, 0, tx * p.TW + dx[j], ty * p.TH + dy[j], chunk0, pl[j] * p.n_img + n);
          dst += b_al;
        }
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =========================================
    if (elect_one() && my_k > 0) {
      // kind::f16, D = f32, A = B = bf16, both operands MN-major (bits 15, 16)
      const uint32_t idesc = make_idesc(128, p.N_mma) | (1u << 15) | (1u << 16);
      const int swap = (p.dbg >> 4) & 1;
      const uint64_t hiA = smem_desc_mn_hi(p.cwA, (uint32_t)(PC_PK * p.cwA * 2), swap);
      const uint64_t hiB = smem_desc_mn_hi(p.cwB, (uint32_t)(PC_PK * p.cwB * 2), swap);
      const uint32_t kstepA = (uint32_t)(16 * p.cwA * 2) >> 4, kstepB = (uint32_t)(16 * p.cwB * 2) >> 4;  // 16 pixel rows
      const uint32_t bal_d = (uint32_t)b_al >> 4;
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_k; ++i) {
        mbar_wait_a(full_s + s * 8, ph);
        tc_fence_after();
        const uint32_t sa = smem_s + s * stage_bytes;
        uint64_t da = hiA | (uint64_t)((sa & 0x3FFFFu) >> 4);
        uint64_t db = hiB | (uint64_t)(((sa + (p.alias_b ? 0 : p.mchunk ? p.tpc * a_bytes : a_bytes)) & 0x3FFFFu) >> 4);
        uint32_t dt = tmem_base;
        for (int j = 0; j < cnt; ++j) {
          uint64_t ak = da, bk = db;
          umma_bf16(dt, ak, bk, idesc, i != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 1; k < PC_PK / 16; ++k) {
            ak += kstepA; bk += kstepB;
            umma_bf16_acc(dt, ak, bk, idesc);
          }
          if (p.mchunk) da += (uint32_t)a_bytes >> 4;     // next M tile, same N operand
          else db += bal_d;                                // next tap, same A tile
          dt += p.N_mma;
        }
        umma_commit_a(empty_s + s * 8);
        if (++s == S) { s = 0; ph ^= 1; }
      }
      umma_commit(accfull);
    }
  } else if (my_k > 0) {
    // ================================ epilogue (warps 2..5) ==============================
    const int lg = warp & 3, row = lg * 32 + lane;
    mbar_wait(accfull, 0);
    tc_fence_after();
    const int m = mt * 128 + row;
    const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16);
    for (int j = 0; j < cnt && p.mchunk; ++j) {
      // row = (tap', ci) of the chunked M space, column = co: out[tap'][co][ci] (the caller's [tap][M][N] layout);
      // lanes hold consecutive ci, so each atomic instruction of a warp covers 128 contiguous bytes
      const int grow = (tap + j) * 128 + row;
      const bool ok = grow < p.rows_total;
      const int tp = ok ? grow / p.N : 0, ci = ok ? grow - tp * p.N : 0;
      float* obase = p.out + (size_t)tp * p.M * p.N + ci;
      for (int c0 = 0; c0 < p.N_mma; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + j * p.N_mma + c0, r);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (c0 + q < p.M) atomicAdd(obase + (size_t)(c0 + q) * p.N, __uint_as_float(r[q]) * p.scale);
        }
      }
    }
    for (int j = 0; j < cnt && !p.mchunk; ++j) {
      float* obase = p.out + (((size_t)oimg * p.n_taps + tap + j) * p.M + m) * p.N;
      for (int c0 = 0; c0 < p.N_mma; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + j * p.N_mma + c0, r);
        tmem_ld_wait();
        if (m < p.M) {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int n = nt * p.N_mma + c0 + q;
            if (n < p.N) atomicAdd(obase + n, __uint_as_float(r[q]) * p.scale);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// 5-D map (c, X, Y, chunk, img*plane) over a dense channels-last tensor [NP][Y][X][C]
static int make_tmap_pc(CUtensorMap* out, const void* base, int C, int X, int Y, int NP, int cw, int TW, int TH, int chunks) {
  // reuse the activation encoder: dims (cw, X, Y, C/cw, NP), strides (C, X*C, cw, Y*X*C) elements
  return make_tmap_act_generic(out, base, cw, X, Y, C / cw, NP, (size_t)C, (size_t)X * C, (size_t)cw, (size_t)Y * X * C, cw, TW, TH,
                               chunks);
}

static int chunk_width(int C) { return C % 64 == 0 ? 64 : C % 32 == 0 ? 32 : 16; }

}  // namespace vst

using namespace vst;

extern "C" {

int vst_tc_pcgemm(const vst_pcgemm_desc* d, void* stream) {
  VST_CHECK_ARG(d, "pcgemm: NULL descriptor");
  VST_CHECK_ARG(d->n_taps >= 1 && d->n_taps <= TG_MAX_TAPS, "pcgemm: 1..%d taps", TG_MAX_TAPS);
  VST_CHECK_ARG(d->a_C % 16 == 0 && d->b_C % 16 == 0, "pcgemm: channel counts must be multiples of 16");
  VST_CHECK_ARG(d->M >= 1 && d->M <= d->a_C && d->N >= 1 && d->N <= d->b_C, "pcgemm: M/N exceed the operand channels");
  VST_CHECK_ARG(d->grid_h >= 1 && d->grid_w >= 1 && d->n_img >= 1, "pcgemm: empty pixel grid");
  VST_DEVPTR(d->a); VST_DEVPTR(d->b); VST_DEVPTR(d->out);
  PcGemmParams p;
  memset(&p, 0, sizeof(p));
  p.cwA = chunk_width(d->a_C);
  p.cwB = chunk_width(d->b_C);
  p.a_chunks = 128 / p.cwA;
  const int n_pad = cdiv(d->N, 16) * 16;
  p.N_mma = n_pad > 256 ? 256 : n_pad;
  if (p.N_mma % p.cwB) p.N_mma = cdiv(p.N_mma, p.cwB) * p.cwB;
  VST_CHECK_ARG(p.N_mma <= 256, "pcgemm: N tile %d", p.N_mma);
  p.b_chunks = p.N_mma / p.cwB;
  p.m_tiles = cdiv(d->M, 128);
  p.n_tiles = cdiv(d->N, p.N_mma);
  p.n_taps = d->n_taps;
  p.n_img = d->n_img;
  p.per_image = d->per_image ? 1 : 0;
  p.n_out_img = p.per_image ? d->n_img : 1;
  // pixel tile: 64 wide when the grid is wide, else 16 x 4 / 8 x 8
  if (d->grid_w >= 48) { p.TW = 64; p.TH = 1; }
  else if (d->grid_w >= 12) { p.TW = 16; p.TH = 4; }
  else { p.TW = 8; p.TH = 8; }
  p.tiles_x = cdiv(d->grid_w, p.TW);
  p.tiles_y = cdiv(d->grid_h, p.TH);
  p.M = d->M; p.N = d->N; p.scale = d->scale; p.out = d->out;
  for (int t = 0; t < d->n_taps; ++t) {
    p.tapA[t] = (d->a_dx[t] & 0xff) | ((d->a_dy[t] & 0xff) << 8) | ((int)d->a_pl[t] << 16);
    p.tapB[t] = (d->b_dx[t] & 0xff) | ((d->b_dy[t] & 0xff) << 8) | ((int)d->b_pl[t] << 16);
    VST_CHECK_ARG(d->a_pl[t] >= 0 && d->a_pl[t] < d->a_P && d->b_pl[t] >= 0 && d->b_pl[t] < d->b_P, "pcgemm: tap plane out of range");
  }
  { const char* e = getenv("VST_PC_DBG"); p.dbg = e ? atoi(e) : 0; }
  // tap groups: consecutive taps that read the same A tile share one CTA (A is loaded once per K step and each tap
  // accumulates into its own TMEM columns); bounded by the 512 TMEM columns, 16 taps and the smem budget
  const int a_bytes = 128 * PC_PK * 2, b_al = (p.N_mma * PC_PK * 2 + 1023) & ~1023;
  int tpc_max = 512 / p.N_mma;
  if (tpc_max > 16) tpc_max = 16;
  while (tpc_max > 1 && 2 * (a_bytes + tpc_max * b_al) > 200 * 1024) --tpc_max;
  { const char* e = getenv("VST_PC_TPC"); if (e && atoi(e) > 0 && atoi(e) < tpc_max) tpc_max = atoi(e); }
  p.n_groups = 0; p.tpc = 1;
  for (int t = 0; t < d->n_taps;) {
    int c = 1;
    while (t + c < d->n_taps && c < tpc_max && p.tapA[t + c] == p.tapA[t]) ++c;
    // balance: do not leave a short tail group (e.g. 9 taps at 2 per CTA -> 2,2,2,2,1); prefer near-equal sizes
    p.grp_first[p.n_groups] = (short)t; p.grp_cnt[p.n_groups] = (short)c;
    if (c > p.tpc) p.tpc = c;
    ++p.n_groups;
    t += c;
  }
  p.stages = (200 * 1024) / (a_bytes + p.tpc * b_al);
  if (p.stages > PC_MAX_STAGES) p.stages = PC_MAX_STAGES;
  int k_total = (p.per_image ? 1 : p.n_img) * p.tiles_x * p.tiles_y;
  int blocks = p.n_out_img * p.m_tiles * p.n_tiles * p.n_groups;
  // ---- M-chunk mode (see the header): all taps share one A tile, 64-channel chunks on the shifted side, and M leaves a
  // half-empty 128-row tile (the 192 -> 192 trunk: 14 M tiles of useful rows instead of 18, and one shared N-operand
  // box per K step instead of one per tap -> a third less shared-memory traffic).  VST_PC_MCHUNK=0 disables it.
  static const bool mchunk_on = [] { const char* e = getenv("VST_PC_MCHUNK"); return !e || atoi(e) != 0; }();
  bool same_a = true;
  for (int t = 1; t < d->n_taps; ++t) same_a = same_a && p.tapA[t] == p.tapA[0];
  const void *baseA = d->a, *baseB = d->b;
  int aC = d->a_C, aX = d->a_X, aY = d->a_Y, aNP = d->a_N * d->a_P, bC = d->b_C, bX = d->b_X, bY = d->b_Y, bNP = d->b_N * d->b_P;
  if (mchunk_on && same_a && d->n_taps > 1 && !p.per_image && p.cwB == 64 && d->N % 64 == 0 && d->M % 128 != 0 && d->M % 16 == 0 &&
      d->M <= 256 && p.n_tiles == 1) {
    p.mchunk = 1;
    const int cw_out = p.cwA;                       // chunk width of the caller's A (the shared N operand from here on)
    p.cwA = 64; p.a_chunks = 2;
    p.cwB = cw_out;
    p.N_mma = cdiv(d->M, 16) * 16;
    if (p.N_mma % p.cwB) p.N_mma = cdiv(p.N_mma, p.cwB) * p.cwB;
    p.b_chunks = p.N_mma / p.cwB;
    p.cpt = d->N / 64;
    p.rows_total = d->n_taps * d->N;
    p.m_tiles = cdiv(p.rows_total, 128);
    const int b_al2 = (p.N_mma * PC_PK * 2 + 1023) & ~1023;
    p.tpc = 512 / p.N_mma;
    if (p.tpc > 4) p.tpc = 4;                       // <= 8 chunk boxes per K step from the one producing thread
    while (p.tpc > 1 && 2 * (p.tpc * a_bytes + b_al2) > 200 * 1024) --p.tpc;
    p.n_groups = cdiv(p.m_tiles, p.tpc);
    p.stages = (200 * 1024) / (p.tpc * a_bytes + b_al2);
    if (p.stages > PC_MAX_STAGES) p.stages = PC_MAX_STAGES;
    blocks = p.n_groups;
    const int tA0 = p.tapA[0];
    for (int t = 0; t < d->n_taps; ++t) p.tapA[t] = p.tapB[t];
    p.tapB[0] = tA0;
    baseA = d->b; aC = d->b_C; aX = d->b_X; aY = d->b_Y; aNP = d->b_N * d->b_P;
    baseB = d->a; bC = d->a_C; bX = d->a_X; bY = d->a_Y; bNP = d->a_N * d->a_P;
  }
  // Gram of a narrow tensor (A == B, C <= 128: relu1_1 / relu2_1): the B operand is a prefix of the A tile, so the kernel loads
  // ONE box per K step and points both descriptors at it - the second copy of F, a third of the shared-memory fill for C = 64,
  // and one of the two producers disappear.  VST_PC_ALIAS=0 disables it.
  static const bool alias_on = [] { const char* e = getenv("VST_PC_ALIAS"); return !e || atoi(e) != 0; }();
  { static const int xf = [] { const char* e = getenv("VST_PC_XFAST"); return e ? atoi(e) : 0; }(); p.xfast = (xf && d->n_taps == 1 && !p.mchunk) ? 1 : 0; }
  p.alias_b = 0;
  if (alias_on && !p.mchunk && d->a == d->b && d->n_taps == 1 && p.tapA[0] == p.tapB[0] && p.cwA == p.cwB && p.m_tiles == 1 &&
      p.n_tiles == 1 && p.b_chunks <= p.a_chunks && d->a_C == d->b_C && d->a_X == d->b_X && d->a_Y == d->b_Y && d->a_P == d->b_P) {
    p.alias_b = 1;
    p.stages = (200 * 1024) / a_bytes;
    if (p.stages > PC_MAX_STAGES) p.stages = PC_MAX_STAGES;
  }
  // Narrow Gram (one box per K step): a CTA is ONE serial producer -> MMA -> commit chain, and that chain - not TMA, not HBM -
  // bounded the C = 64 Gram at ~0.5 of HBM whatever the pipeline depth.  Two CTAs per SM (half the ring each) run two chains.
  // Measured at 2048^2 x 8, C = 64: 1 CTA per SM 2.60 TB/s, 2: 4.72, 3: 6.29 TB/s = 0.97 of the measured HBM peak.
  static const int gram_cps = [] { const char* e = getenv("VST_PC_GRAM_CPS"); return e ? atoi(e) : 0; }();   // 0 = by width
  int cps = 1;
  if (p.alias_b && gram_cps != 1 && p.cwA == 64 && d->a_C <= 128) {
    cps = gram_cps > 1 ? gram_cps : (d->a_C <= 64 ? 3 : 2);
    const int cap = (200 * 1024) / (cps * (a_bytes + 1280));
    if (p.stages > cap) p.stages = cap;
  }
  // the same for the general form (weight gradients, wide Grams), as an experiment switch: VST_PC_CPS=2 halves the ring of
  // every CTA and doubles the K splits (twice the epilogue atomics)
  static const int gen_cps = [] { const char* e = getenv("VST_PC_CPS"); return e ? atoi(e) : 1; }();
  if (cps == 1 && gen_cps > 1) {
    cps = gen_cps;
    const int per_stage = p.mchunk ? p.tpc * a_bytes + ((p.N_mma * PC_PK * 2 + 1023) & ~1023) : p.alias_b ? a_bytes : a_bytes + p.tpc * b_al;
    const int cap = (int)((200 * 1024) / cps - 1280) / per_stage;
    if (cap >= 2 && p.stages > cap) p.stages = cap;
    else if (cap < 2) cps = 1;
  }
  // one CTA per SM (smem-bound): fill whole waves of 148
  int ks = d->k_splits > 0 ? d->k_splits : (blocks >= cps * kNumSMs ? 1 : cps * kNumSMs / blocks);
  if (ks > k_total) ks = k_total;
  if (ks < 1) ks = 1;
  p.k_splits = ks;
  int r = make_tmap_pc(&p.tmA, baseA, aC, aX, aY, aNP, p.cwA, p.TW, p.TH, p.mchunk ? 1 : p.a_chunks);
  if (r != VST_OK) return r;
  r = make_tmap_pc(&p.tmB, baseB, bC, bX, bY, bNP, p.cwB, p.TW, p.TH, p.b_chunks);
  if (r != VST_OK) return r;
  VST_CHECK_ARG(d->a_N == d->n_img && d->b_N == d->n_img, "pcgemm: operand image counts differ from n_img");
  const size_t smem = (size_t)p.stages * (p.alias_b ? a_bytes : p.mchunk ? p.tpc * a_bytes + ((p.N_mma * PC_PK * 2 + 1023) & ~1023) : a_bytes + p.tpc * b_al) + 1024 + 256;
  static bool attr_done = false;
  if (!attr_done) {
    VST_CUDA(cudaFuncSetAttribute(pcgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  vst::launch(pcgemm_kernel, blocks * ks, PC_THREADS, smem, (cudaStream_t)stream, p);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

}  // extern "C"
