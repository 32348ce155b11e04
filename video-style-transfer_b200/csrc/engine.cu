// bf16 tensor-core inference engine for the ReCoNet family (RC/network.py:153-279).
//
// Data layout in HBM (all bf16, channels-last so a filter tap is one TMA box):
//   X9   [N][H+8][W][KR]            conv1 operand: row-window im2col over kx (KR = 9*Cin padded),
//                                    reflect-padded in y so the 9 ky taps are plain row offsets
//   raw  [N][Ho][Wo][C]             conv output before InstanceNorm (one scratch buffer, reused)
//   act  [N][H+2p][W+2p][C]         IN + ReLU (+ residual) applied, written ALREADY PADDED for its
//                                    consumer: reflect (3x3 / 9x9 convs), replicate (the x2-upsample
//                                    convs run as 4 output phases on the low-res tensor), or split
//                                    into 4 row/col-parity planes (stride-2 consumers)
// One forward = prologue + 16 x (tapgemm [+ stats] + apply) + output epilogue, on one stream.
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "tc_conv.cuh"
#include "tc_layout.cuh"

namespace vst {

// 16-bit storage format of a plan: bf16 (default) or fp16 (the "fp16" plan, VST_PLAN_FP16): same bytes, same kernels.
__device__ __forceinline__ uint16_t f2h16(float v, int half) {
  if (half) return __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float h162f(uint16_t u, int half) {
  return half ? __half2float(__ushort_as_half(u)) : __bfloat162float(__ushort_as_bfloat16(u));
}
template <bool HALF>
__device__ __forceinline__ float2 cvt2(uint32_t u) {
  if (HALF) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
}
template <bool HALF>
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  if (HALF) {
    __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Store of a warp's 32 consecutive 64-byte X9 rows (KR = 32).  Direct form: every thread writes its own row, i.e. a store
// instruction is 32 sixteen-byte pieces 64 bytes apart (each 32-byte sector half filled).  Staged form: the rows pass through
// a swizzled 2 KB shared-memory tile of the warp (chunk j of row l at l*64 + (j ^ ((l>>1)&3))*16 - conflict-free both ways)
// and leave as four 512-byte contiguous warp stores.  `valid` = this lane's pixel exists; px0 = the warp's first pixel.
__device__ __forceinline__ void x9_store_row32(__nv_bfloat16* __restrict__ rowbase /* row of pixel px0 */, const uint4* row, bool valid,
                                               int px0, int W, int staged, uint4* tile /* [128] of this warp */) {
  const int l = threadIdx.x & 31;
  if (!staged) {
    if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(rowbase) + 4 * l;
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = row[j];
    }
    return;
  }
  const int sw = (l >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) tile[4 * l + (j ^ sw)] = row[j];
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(rowbase);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int g = j * 32 + l, r = g >> 2, q = g & 3;
    if (px0 + r < W) dst[g] = tile[4 * r + (q ^ ((r >> 1) & 3))];
  }
}

// ---- prologue: fp32 NCHW frame -> X9 ------------------------------------------------------
// one thread per (padded row, pixel): gathers the 9*Cin window once (neighbouring threads share
// it through L1) and writes the KR-element row with 16-byte stores.  No integer divisions.
template <int KR, int CIN>   // CIN > 0: compile-time channel count (fully unrolled, registers only)
__global__ void __launch_bounds__(256) prologue_x9_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ x9,
                                                          int N, int Cin_rt, int H, int W, int half = 0, float scale = 1.f,
                                                          int staged = 0) {
  vst::pdl_grid_sync();
  __shared__ uint4 tile[KR == 32 ? 8 * 128 : 1];
  const int Cin = CIN > 0 ? CIN : Cin_rt;
  const int pxr = blockIdx.x * blockDim.x + threadIdx.x;
  const int yp = blockIdx.y, n = blockIdx.z;
  const bool valid = pxr < W;
  if (!valid && !(KR == 32 && staged)) return;
  const int px = valid ? pxr : W - 1;
  const int sy = reflect_idx(yp - 4, H);
  const float* xrow = x + ((size_t)n * Cin * H + sy) * W;
  int sx[9];
#pragma unroll
  for (int kx = 0; kx < 9; ++kx) sx[kx] = reflect_idx(px + kx - 4, W);
  __nv_bfloat16* dst = x9 + (((size_t)n * (H + 8) + yp) * W + px) * KR;
  __align__(16) uint16_t row[KR];
  int k = 0;
#pragma unroll
  for (int kx = 0; kx < 9; ++kx)
#pragma unroll
    for (int c = 0; c < Cin; ++c)
      if (k < KR) row[k++] = f2h16(__ldg(xrow + (size_t)c * H * W + sx[kx]) * scale, half);
  for (; k < KR; ++k) row[k] = 0;
  if (KR == 32) {
    const int px0 = pxr - (threadIdx.x & 31);
    x9_store_row32(x9 + (((size_t)n * (H + 8) + yp) * W + px0) * KR, reinterpret_cast<const uint4*>(row), valid, px0, W, staged,
                   tile + (threadIdx.x >> 5) * 128);
    return;
  }
#pragma unroll
  for (int j = 0; j < KR / 8; ++j) reinterpret_cast<uint4*>(dst)[j] = reinterpret_cast<const uint4*>(row)[j];
}

// Same X9 operand straight from decoder frames: uint8 [N][H][W][3] in BGR order (what cv2.VideoCapture.read returns and what
// RC/utilities.py:119-123 `cvframe_to_tensor` turns into a float RGB tensor ON THE HOST).  u8 -> float is exact, so the
// operand - and every frame - is bit-identical to the fp32 entry fed with cvframe_to_tensor's output; the upload is 4x smaller.
template <int KR>
__global__ void __launch_bounds__(256) prologue_x9_bgr8_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ x9,
                                                               int N, int H, int W, int half, float scale, int staged) {
  vst::pdl_grid_sync();
  __shared__ uint4 tile[KR == 32 ? 8 * 128 : 1];
  const int pxr = blockIdx.x * blockDim.x + threadIdx.x;
  const int yp = blockIdx.y, n = blockIdx.z;
  const bool valid = pxr < W;
  if (!valid && !(KR == 32 && staged)) return;
  const int px = valid ? pxr : W - 1;
  const int sy = reflect_idx(yp - 4, H);
  const uint8_t* xrow = x + ((size_t)n * H + sy) * W * 3;
  __nv_bfloat16* dst = x9 + (((size_t)n * (H + 8) + yp) * W + px) * KR;
  __align__(16) uint16_t row[KR];
  int k = 0;
#pragma unroll
  for (int kx = 0; kx < 9; ++kx) {
    const uint8_t* pxl = xrow + (size_t)reflect_idx(px + kx - 4, W) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (k < KR) row[k++] = f2h16((float)__ldg(pxl + (2 - c)) * scale, half);   // RGB channel c = BGR byte 2 - c
  }
  for (; k < KR; ++k) row[k] = 0;
  if (KR == 32) {
    const int px0 = pxr - (threadIdx.x & 31);
    x9_store_row32(x9 + (((size_t)n * (H + 8) + yp) * W + px0) * KR, reinterpret_cast<const uint4*>(row), valid, px0, W, staged,
                   tile + (threadIdx.x >> 5) * 128);
    return;
  }
#pragma unroll
  for (int j = 0; j < KR / 8; ++j) reinterpret_cast<uint4*>(dst)[j] = reinterpret_cast<const uint4*>(row)[j];
}
// Third form (default): the block first stages its 256 + 8 source pixels in shared memory, already converted to 16 bits in RGB
// order (each source element is read, mirrored, scaled and rounded ONCE, with coalesced loads, instead of nine times through L1).
// Pixel t's X9 row is then 27 consecutive 16-bit elements of that segment starting at element 3 * t; the warp rebuilds its 32
// rows directly in output order (below).  Same rounding of the same products -> same bits.
template <bool BGR8>
__global__ void __launch_bounds__(256) prologue_x9_seg_kernel(const void* __restrict__ xv, __nv_bfloat16* __restrict__ x9,
                                                              int N, int H, int W, int half, float scale) {
  vst::pdl_grid_sync();
  constexpr int SEGP = 256 + 8;                         // pixels of the segment
  __shared__ __align__(16) uint32_t seg32[(SEGP * 3 + 8) / 2];
  uint16_t* seg = reinterpret_cast<uint16_t*>(seg32);
  const int bx0 = blockIdx.x * 256;                     // first pixel of the block
  const int yp = blockIdx.y, n = blockIdx.z;
  const int sy = reflect_idx(yp - 4, H);
  for (int i = threadIdx.x; i < SEGP * 3 + 8; i += 256) {
    uint16_t v = 0;
    if (i < SEGP * 3) {
      int p, c;                                         // segment pixel, RGB channel
      if (BGR8) { p = i / 3; c = 2 - (i - 3 * p); }     // consecutive threads read consecutive bytes of the frame row
      else { c = i / SEGP; p = i - c * SEGP; }          // consecutive threads read consecutive floats of one channel row
      const int xs = bx0 + p - 4;
      if (xs < W + 4) {
        const int sx = reflect_idx(xs, W);
        float f;
        if (BGR8) f = (float)__ldg(reinterpret_cast<const uint8_t*>(xv) + (((size_t)n * H + sy) * W + sx) * 3 + (2 - c));
        else f = __ldg(reinterpret_cast<const float*>(xv) + (((size_t)n * 3 + c) * H + sy) * W + sx);
        v = f2h16(f * scale, half);
      }
      seg[p * 3 + c] = v;
    } else {
      seg[i] = 0;
    }
  }
  __syncthreads();
  // Output chunk g of the warp (16 bytes = elements 8q .. 8q+7 of row r = g >> 2, q = g & 3) is 8 consecutive segment elements
  // from 3 * (pixel) + 8q: lane l takes g = j*32 + l, so a warp store is 512 contiguous bytes, and its five 32-bit shared loads
  // hit words floor(1.5 r') + 4q + i (r' = 0..7) - at most 23 distinct words, all in distinct banks: conflict-free.
  const int l = threadIdx.x & 31, tw0 = threadIdx.x & ~31;
  const int px0 = bx0 + tw0, q = l & 3;
  uint4* dst = reinterpret_cast<uint4*>(x9 + (((size_t)n * (H + 8) + yp) * W + px0) * 32);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = j * 8 + (l >> 2);
    const int e0 = 3 * (tw0 + r) + 8 * q;
    const uint32_t* ws = seg32 + (e0 >> 1);
    uint32_t w[5], o[4];
#pragma unroll
    for (int i = 0; i < 5; ++i) w[i] = ws[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = (e0 & 1) ? __funnelshift_r(w[i], w[i + 1], 16) : w[i];
    if (q == 3) { o[1] &= 0xffffu; o[2] = 0; o[3] = 0; }   // elements 27..31 of the row are zero padding
    if (px0 + r < W) dst[j * 32 + l] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
static int x9_staged() {   // VST_X9_STAGED (A/B switch): 0 every thread stores its own 64-byte row directly, 1 staged store only,
                           // 2 (default) staged source segment + staged store (prologue_x9_seg_kernel)
  static const int on = [] { const char* e = getenv("VST_X9_STAGED"); const int v = e ? atoi(e) : 2; return v < 0 || v > 2 ? 2 : v; }();
  return on;
}

// ---- InstanceNorm statistics over a raw NHWC bf16 tensor -----------------------------------
// grid (chunks, N); each thread owns one 8-channel group and strides over pixels.
__global__ void __launch_bounds__(256) stats_kernel(const __nv_bfloat16* __restrict__ raw, double* __restrict__ stats,
                                                    int HW, int C, int pix_per_block) {
  vst::pdl_grid_sync();
  extern __shared__ float sh[];  // [2][C]
  const int n = blockIdx.y, groups = C / 8;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int g = threadIdx.x % groups, lane_pix = threadIdx.x / groups, pix_step = blockDim.x / groups;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (lane_pix < pix_step) {
    const uint4* src = reinterpret_cast<const uint4*>(raw + (size_t)n * HW * C);
    for (int p = p0 + lane_pix; p < p1; p += pix_step) {
      const uint4 q = src[(size_t)p * groups + g];
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h[j]);
        s1[2 * j] += f.x;
        s2[2 * j] = fmaf(f.x, f.x, s2[2 * j]);
        s1[2 * j + 1] += f.y;
        s2[2 * j + 1] = fmaf(f.y, f.y, s2[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[g * 8 + j], s1[j]);
      atomicAdd(&sh[C + g * 8 + j], s2[j]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&stats[((size_t)n * C + c) * 2], (double)sh[c]);
    atomicAdd(&stats[((size_t)n * C + c) * 2 + 1], (double)sh[C + c]);
  }
}

// ---- apply: y = act(IN(raw)) (+ residual), written into the consumer's padded layout ------
// grid (row bands, N).  Thread -> (8-channel group g, pixel lane): the inner loop walks a padded row
// with no divisions, two independent 16-byte loads in flight per thread.
__device__ __forceinline__ uint4 apply_one(const uint4 q, const uint4 rq, bool has_res, const float (&ca)[8], const float (&cb)[8],
                                           int relu) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
  const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rq);
  uint4 o;
  __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    float a = fmaf(f.x, ca[2 * j], cb[2 * j]);
    float b = fmaf(f.y, ca[2 * j + 1], cb[2 * j + 1]);
    if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
    if (has_res) {
      const float2 r = __bfloat1622float2(rh[j]);
      a += r.x; b += r.y;
    }
    oh[j] = __floats2bfloat162_rn(a, b);
  }
  return o;
}

template <int PX, int MINB, bool RES>
__global__ void __launch_bounds__(256, MINB) apply_kernel(const __nv_bfloat16* __restrict__ raw, const double* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const __nv_bfloat16* __restrict__ residual, ActLayout RL,
                                                       __nv_bfloat16* __restrict__ dst, ActLayout DL, int N, float eps,
                                                       int relu, int rows_per_block) {
  vst::pdl_grid_sync();
  extern __shared__ float sh[];  // a[C], b[C]
  const int n = blockIdx.y, C = DL.C, H = DL.H, W = DL.W;
  const float inv_cnt = 1.f / (float)(H * W);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s1 = stats[((size_t)n * C + c) * 2], s2 = stats[((size_t)n * C + c) * 2 + 1];
    const double mean_d = s1 * (double)inv_cnt;
    const float mean = (float)mean_d;
    const float var = fmaxf((float)(s2 * (double)inv_cnt - mean_d * mean_d), 0.f);
    const float a = gamma[c] * rsqrtf(var + eps);
    sh[c] = a;
    sh[C + c] = beta[c] - mean * a;
  }
  __syncthreads();
  const int groups = C >> 3, Hp = H + 2 * DL.pad, Wp = W + 2 * DL.pad;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups, step = blockDim.x / groups;
  if (pl >= step) return;
  // this thread's channel group never changes: scale / shift live in registers (ncu: with 16 shared-memory loads per
  // 16-byte vector the L1TEX pipe, not DRAM, was the saturated unit - 92 % busy at 34 % of DRAM throughput)
  float ca[8], cb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { ca[j] = sh[g * 8 + j]; cb[j] = sh[C + g * 8 + j]; }
  const int y_begin = blockIdx.x * rows_per_block, y_end = min(Hp, y_begin + rows_per_block);
  for (int yp = y_begin; yp < y_end; ++yp) {
    bool oky;
    const int sy = map_pad(yp - DL.pad, H, DL.kind, oky);
    const __nv_bfloat16* rrow = raw + ((size_t)n * H + sy) * W * C + g * 8;
    int xp = pl;
    // PX pixels per iteration: all loads issued before the first use
    for (; xp + (PX - 1) * step < Wp; xp += PX * step) {
      uint4 q[PX], r[PX];
      bool v[PX];
      int sx[PX];
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        bool okx;
        sx[u] = map_pad(xp + u * step - DL.pad, W, DL.kind, okx);
        v[u] = oky && okx;
        q[u] = make_uint4(0, 0, 0, 0);
        r[u] = q[u];
        if (v[u]) q[u] = __ldg(reinterpret_cast<const uint4*>(rrow + (size_t)sx[u] * C));
      }
      if (RES) {
#pragma unroll
        for (int u = 0; u < PX; ++u)
          if (v[u]) r[u] = __ldg(reinterpret_cast<const uint4*>(residual + act_offset(RL, N, n, sy + RL.pad, sx[u] + RL.pad) + g * 8));
      }
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        const uint4 o = v[u] ? apply_one(q[u], r[u], RES, ca, cb, relu) : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(dst + act_offset(DL, N, n, yp, xp + u * step) + g * 8) = o;
      }
    }
    // remainder: one pixel per iteration
    for (; xp < Wp; xp += step) {
      bool ok0;
      const int sx0 = map_pad(xp - DL.pad, W, DL.kind, ok0);
      uint4 o0 = make_uint4(0, 0, 0, 0);
      if (oky && ok0) {
        const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(rrow + (size_t)sx0 * C));
        uint4 r0 = make_uint4(0, 0, 0, 0);
        if (RES) r0 = __ldg(reinterpret_cast<const uint4*>(residual + act_offset(RL, N, n, sy + RL.pad, sx0 + RL.pad) + g * 8));
        o0 = apply_one(q0, r0, RES, ca, cb, relu);
      }
      *reinterpret_cast<uint4*>(dst + act_offset(DL, N, n, yp, xp) + g * 8) = o0;
    }
  }
}

// High-occupancy variant: scale / shift stay in shared memory as conflict-free float4 rows ([a.lo | a.hi | b.lo | b.hi][group])
// and are re-read per vector with volatile 128-bit loads, so the kernel fits 32-40 registers and 6-8 blocks per SM.
__device__ __forceinline__ float4 lds128(const float4* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
  return v;
}
template <bool HALF>
__device__ __forceinline__ uint2 apply_half(const uint2 q, const uint2 rq, bool has_res, const float4 a, const float4 b, int relu) {
  uint2 o;
  const float2 f0 = cvt2<HALF>(q.x), f1 = cvt2<HALF>(q.y);
  float v0 = fmaf(f0.x, a.x, b.x), v1 = fmaf(f0.y, a.y, b.y), v2 = fmaf(f1.x, a.z, b.z), v3 = fmaf(f1.y, a.w, b.w);
  if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
  if (has_res) {
    const float2 r0 = cvt2<HALF>(rq.x), r1 = cvt2<HALF>(rq.y);
    v0 += r0.x; v1 += r0.y; v2 += r1.x; v3 += r1.y;
  }
  o.x = pk2<HALF>(v0, v1);
  o.y = pk2<HALF>(v2, v3);
  return o;
}

template <int PX, int MINB, bool RES, bool HALF = false>
__global__ void __launch_bounds__(256, MINB) apply_lds_kernel(const __nv_bfloat16* __restrict__ raw, const double* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const __nv_bfloat16* __restrict__ residual, ActLayout RL,
                                                           __nv_bfloat16* __restrict__ dst, ActLayout DL, int N, float eps,
                                                           int relu, int rows_per_block) {
  vst::pdl_grid_sync();
  extern __shared__ float4 sh4[];  // [4][groups]
  const int n = blockIdx.y, C = DL.C, H = DL.H, W = DL.W;
  const int groups = C >> 3;
  const float inv_cnt = 1.f / (float)(H * W);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s1 = stats[((size_t)n * C + c) * 2], s2 = stats[((size_t)n * C + c) * 2 + 1];
    const double mean_d = s1 * (double)inv_cnt;
    const float mean = (float)mean_d;
    const float var = fmaxf((float)(s2 * (double)inv_cnt - mean_d * mean_d), 0.f);
    const float a = gamma[c] * rsqrtf(var + eps);
    float* shf = reinterpret_cast<float*>(sh4);
    const int slot = (((c >> 2) & 1) * groups + (c >> 3)) * 4 + (c & 3);
    shf[slot] = a;
    shf[slot + 8 * groups] = beta[c] - mean * a;
  }
  __syncthreads();
  const int Hp = H + 2 * DL.pad, Wp = W + 2 * DL.pad;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups, step = blockDim.x / groups;
  if (pl >= step) return;
  const float4* cst = sh4 + g;
  const int y_begin = blockIdx.x * rows_per_block, y_end = min(Hp, y_begin + rows_per_block);
  for (int yp = y_begin; yp < y_end; ++yp) {
    bool oky;
    const int sy = map_pad(yp - DL.pad, H, DL.kind, oky);
    const __nv_bfloat16* rrow = raw + ((size_t)n * H + sy) * W * C + g * 8;
    // pixel x of a row sits at base[x & parity] + (x >> parity) * C in either layout
    const __nv_bfloat16* res0 = nullptr;
    const __nv_bfloat16* res1 = nullptr;
    if (RES) {
      res0 = residual + act_offset(RL, N, n, sy + RL.pad, 0) + g * 8;
      res1 = residual + act_offset(RL, N, n, sy + RL.pad, RL.parity) + g * 8;
    }
    __nv_bfloat16* d0 = dst + act_offset(DL, N, n, yp, 0) + g * 8;
    __nv_bfloat16* d1 = dst + act_offset(DL, N, n, yp, DL.parity) + g * 8;
    for (int xp = pl; xp < Wp; xp += PX * step) {
      uint4 q[PX], r[PX];
      bool v[PX];
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        bool okx;
        const int x = xp + u * step;
        const int sx = map_pad(x - DL.pad, W, DL.kind, okx);
        v[u] = oky && okx && x < Wp;
        q[u] = make_uint4(0, 0, 0, 0);
        r[u] = q[u];
        if (v[u]) {
          q[u] = __ldg(reinterpret_cast<const uint4*>(rrow + (size_t)sx * C));
          if (RES) {
            const int rx = sx + RL.pad;
            r[u] = __ldg(reinterpret_cast<const uint4*>(((rx & RL.parity) ? res1 : res0) + (size_t)(rx >> RL.parity) * RL.C));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        const int x = xp + u * step;
        if (x >= Wp) break;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (v[u]) {
          const uint2 lo = apply_half<HALF>(make_uint2(q[u].x, q[u].y), make_uint2(r[u].x, r[u].y), RES, lds128(cst), lds128(cst + 2 * groups), relu);
          const uint2 hi = apply_half<HALF>(make_uint2(q[u].z, q[u].w), make_uint2(r[u].z, r[u].w), RES, lds128(cst + groups), lds128(cst + 3 * groups), relu);
          o = make_uint4(lo.x, lo.y, hi.x, hi.y);
        }
        *reinterpret_cast<uint4*>(((x & DL.parity) ? d1 : d0) + (size_t)(x >> DL.parity) * DL.C) = o;
      }
    }
  }
}

// one launcher for the plan and the stand-alone entry.  VST_APPLY_VARIANT (0..2) / VST_APPLY_BPS pick the unroll,
// residency and rows-per-block for tuning runs.
static void launch_apply(const __nv_bfloat16* raw, const double* stats, const float* gamma, const float* beta,
                         const __nv_bfloat16* res_buf, const ActLayout& RL, __nv_bfloat16* dst, const ActLayout& DL, int N,
                         float eps, int relu, cudaStream_t st, int half = 0) {
  static const int variant = [] { const char* e = getenv("VST_APPLY_VARIANT"); return e ? atoi(e) : 0; }();
  // Co-residency with the persistent tap-GEMM CTAs of another stream: a kernel can only join an SM whose shared-memory /
  // L1 split already matches its own preference, and the tap-GEMMs run at the maximum-shared split.  These streaming
  // kernels have no use for L1, so they ask for the same split (VST_APPLY_CARVEOUT=1; the default leaves the driver's choice, see below).
  static const bool carve_done = [] {
    const char* e = getenv("VST_APPLY_CARVEOUT");
    if (!e || atoi(e) == 0) return true;   // default off: measured -6 % on the apply kernels (tiny L1 = few loads in flight), and
                                           // co-residency with the tap-GEMM CTAs happens either way (tools/lane_timeline.py: 68 % overlap)
    const int mx = cudaSharedmemCarveoutMaxShared;
    cudaFuncSetAttribute(apply_lds_kernel<1, 8, true>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(apply_lds_kernel<1, 8, false>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(apply_lds_kernel<2, 6, true>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(apply_lds_kernel<2, 6, false>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(apply_kernel<2, 4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(apply_kernel<2, 4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    cudaFuncSetAttribute(prologue_x9_kernel<32, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, mx);
    return true;
  }();
  (void)carve_done;
  static const int bps_env = [] { const char* e = getenv("VST_APPLY_BPS"); return e ? atoi(e) : 0; }();
  const int Hp = DL.H + 2 * DL.pad;
  const int bps = bps_env > 0 ? bps_env : 16;
  int rpb = cdiv(Hp * N, kNumSMs * bps);
  if (rpb < 1) rpb = 1;
  dim3 grid(cdiv(Hp, rpb), N);
  const size_t smem = 2 * DL.C * sizeof(float);
#define VST_APPLY_GO(PX, MINB)                                                                                        \
  do {                                                                                                                 \
    if (res_buf) vst::launch(apply_kernel<PX, MINB, true>, grid, 256, smem, st, raw, stats, gamma, beta, res_buf, RL, dst, DL, N, eps, relu, rpb); \
    else vst::launch(apply_kernel<PX, MINB, false>, grid, 256, smem, st, raw, stats, gamma, beta, res_buf, RL, dst, DL, N, eps, relu, rpb);        \
  } while (0)
#define VST_APPLY_LDS(PX, MINB)                                                                                       \
  do {                                                                                                                 \
    if (res_buf) vst::launch(apply_lds_kernel<PX, MINB, true>, grid, 256, smem, st, raw, stats, gamma, beta, res_buf, RL, dst, DL, N, eps, relu, rpb); \
    else vst::launch(apply_lds_kernel<PX, MINB, false>, grid, 256, smem, st, raw, stats, gamma, beta, res_buf, RL, dst, DL, N, eps, relu, rpb);        \
  } while (0)
  if (half) {   // fp16 plan: the full-occupancy kernel with fp16 conversions
    if (res_buf) vst::launch(apply_lds_kernel<1, 8, true, true>, grid, 256, smem, st, raw, stats, gamma, beta, res_buf, RL, dst, DL, N, eps, relu, rpb);
    else vst::launch(apply_lds_kernel<1, 8, false, true>, grid, 256, smem, st, raw, stats, gamma, beta, res_buf, RL, dst, DL, N, eps, relu, rpb);
    return;
  }
  switch (variant) {
    case 1: VST_APPLY_GO(2, 4); break;   // constants in registers, 50 % occupancy
    case 2: VST_APPLY_LDS(2, 6); break;
    default: VST_APPLY_LDS(1, 8); break;  // measured best: full occupancy beats per-thread unrolling (485 vs 459 fps)
  }
#undef VST_APPLY_GO
#undef VST_APPLY_LDS
}

// ---- "fp16" plan, residual trunk: InstanceNorm (+ReLU) (+ fp32 residual) with an fp32 copy of the result -----------------
// The trained ReCoNet checkpoints drive `features` (the res5 output) to ~1 % of the residual stream that feeds it, so the
// stream must not be rounded to 16 bits between blocks: it lives in fp32 NHWC (`out32`, unpadded), the blocks' second conv
// writes its raw output in fp32 (`RAW32`), and only the operand copy the next convolution reads (dst, padded, fp16) is rounded.
// Thread -> (pixel, 8-channel group); halo pixels of dst recompute the mirrored interior pixel (a few percent of the work).
template <bool RAW32, bool RES32>
__global__ void __launch_bounds__(256) apply_hp_kernel(const void* __restrict__ raw_v, const double* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ res32, __nv_bfloat16* __restrict__ dst, ActLayout DL,
                                                       float* __restrict__ out32, int N, float eps, int relu, int rows_per_block) {
  vst::pdl_grid_sync();
  extern __shared__ float sh[];  // a[C], b[C]
  const int n = blockIdx.y, C = DL.C, H = DL.H, W = DL.W;
  const float inv_cnt = 1.f / (float)(H * W);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s1 = stats[((size_t)n * C + c) * 2], s2 = stats[((size_t)n * C + c) * 2 + 1];
    const double mean_d = s1 * (double)inv_cnt;
    const float var = fmaxf((float)(s2 * (double)inv_cnt - mean_d * mean_d), 0.f);
    const float a = gamma[c] * rsqrtf(var + eps);
    sh[c] = a;
    sh[C + c] = beta[c] - (float)mean_d * a;
  }
  __syncthreads();
  const int groups = C >> 3, Hp = H + 2 * DL.pad, Wp = W + 2 * DL.pad;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups, step = blockDim.x / groups;
  if (pl >= step) return;
  const int y_begin = blockIdx.x * rows_per_block, y_end = min(Hp, y_begin + rows_per_block);
  for (int yp = y_begin; yp < y_end; ++yp) {
    bool oky;
    const int sy = map_pad(yp - DL.pad, H, DL.kind, oky);
    for (int xp = pl; xp < Wp; xp += step) {
      bool okx;
      const int sx = map_pad(xp - DL.pad, W, DL.kind, okx);
      uint4 o = make_uint4(0, 0, 0, 0);
      if (oky && okx) {
        const size_t src = (((size_t)n * H + sy) * W + sx) * C + g * 8;
        float v[8];
        if (RAW32) {
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(raw_v) + src));
          const float4 r1 = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(raw_v) + src) + 1);
          v[0] = r0.x; v[1] = r0.y; v[2] = r0.z; v[3] = r0.w; v[4] = r1.x; v[5] = r1.y; v[6] = r1.z; v[7] = r1.w;
        } else {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(raw_v) + src));
          const float2 f0 = cvt2<true>(q.x), f1 = cvt2<true>(q.y), f2 = cvt2<true>(q.z), f3 = cvt2<true>(q.w);
          v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = fmaf(v[j], sh[g * 8 + j], sh[C + g * 8 + j]);
          if (relu) v[j] = fmaxf(v[j], 0.f);
        }
        if (RES32) {
          const float4 r0 = __ldg(reinterpret_cast<const float4*>(res32 + src)), r1 = __ldg(reinterpret_cast<const float4*>(res32 + src) + 1);
          v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
        const bool interior = (yp - DL.pad == sy) && (xp - DL.pad == sx);
        if (out32 && interior) {
          reinterpret_cast<float4*>(out32 + src)[0] = make_float4(v[0], v[1], v[2], v[3]);
          reinterpret_cast<float4*>(out32 + src)[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        o = make_uint4(pk2<true>(v[0], v[1]), pk2<true>(v[2], v[3]), pk2<true>(v[4], v[5]), pk2<true>(v[6], v[7]));
      }
      *reinterpret_cast<uint4*>(dst + act_offset(DL, N, n, yp, xp) + g * 8) = o;
    }
  }
}

// interior of a padded activation -> fp32 NCHW (features output, debug hook)
__global__ void __launch_bounds__(256) act_to_nchw_kernel(const __nv_bfloat16* __restrict__ act, ActLayout L, int N,
                                                          float* __restrict__ out, int half = 0) {
  vst::pdl_grid_sync();
  const size_t total = (size_t)N * L.C * L.H * L.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i, hw = (unsigned)(L.W * L.H);   // launchers guarantee total < 2^32
    const int x = iu % L.W, y = (iu / L.W) % L.H, c = (iu / hw) % L.C, n = iu / (hw * L.C);
    out[i] = h162f(reinterpret_cast<const uint16_t*>(act)[act_offset(L, N, n, y + L.pad, x + L.pad) + c], half);
  }
}

// Tiled transposes for the wide tensors of the training step (192-channel features and their gradient): one block moves
// 32 pixels of a row x all channels through shared memory, so both sides are coalesced (16-byte NHWC vectors, 128-byte
// NCHW rows).  The element-per-thread kernels below read 2 useful bytes per 32-byte sector on one side (~1 TB/s).
constexpr int TR_PX = 32, TR_PITCH = 33;
__global__ void __launch_bounds__(256) act_to_nchw_tiled_kernel(const __nv_bfloat16* __restrict__ act, ActLayout L, int N,
                                                                float* __restrict__ out) {
  vst::pdl_grid_sync();
  extern __shared__ float trs[];   // [C][TR_PITCH]
  const int C = L.C, groups = C >> 3;
  const int x0 = blockIdx.x * TR_PX, y = blockIdx.y, n = blockIdx.z;
  const int npx = min(TR_PX, L.W - x0);
  for (int v = threadIdx.x; v < npx * groups; v += blockDim.x) {
    const int px = v / groups, g = v - px * groups;
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(act + act_offset(L, N, n, y + L.pad, x0 + px + L.pad) + g * 8));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      trs[(g * 8 + 2 * j) * TR_PITCH + px] = f.x;
      trs[(g * 8 + 2 * j + 1) * TR_PITCH + px] = f.y;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < npx)
    for (int c = warp; c < C; c += 8) out[(((size_t)n * C + c) * L.H + y) * L.W + x0 + lane] = trs[c * TR_PITCH + lane];
}
// pad == 0, plain layout only (the gradient tensors the sweep builds)
__global__ void __launch_bounds__(256) nchw_to_act_tiled_kernel(const float* __restrict__ x, int Cin, __nv_bfloat16* __restrict__ dst,
                                                                ActLayout L, int N) {
  vst::pdl_grid_sync();
  extern __shared__ float trs[];   // [C][TR_PITCH]
  const int C = L.C, groups = C >> 3;
  const int x0 = blockIdx.x * TR_PX, y = blockIdx.y, n = blockIdx.z;
  const int npx = min(TR_PX, L.W - x0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < C; c += 8)
    trs[c * TR_PITCH + lane] = (c < Cin && lane < npx) ? __ldg(x + (((size_t)n * Cin + c) * L.H + y) * L.W + x0 + lane) : 0.f;
  __syncthreads();
  for (int v = threadIdx.x; v < npx * groups; v += blockDim.x) {
    const int px = v / groups, g = v - px * groups;
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      h[j] = __floats2bfloat162_rn(trs[(g * 8 + 2 * j) * TR_PITCH + px], trs[(g * 8 + 2 * j + 1) * TR_PITCH + px]);
    *reinterpret_cast<uint4*>(dst + act_offset(L, N, n, y, x0 + px) + g * 8) = q;
  }
}

// generic fp32 NCHW -> padded NHWC bf16 (stand-alone conv entry + tests)
__global__ void __launch_bounds__(256) nchw_to_act_kernel(const float* __restrict__ x, int Cin, __nv_bfloat16* __restrict__ dst,
                                                          ActLayout L, int N) {
  vst::pdl_grid_sync();
  const int Hp = L.H + 2 * L.pad, Wp = L.W + 2 * L.pad;
  const size_t total = (size_t)N * Hp * Wp * L.C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;
    const int c = iu % L.C, xp = (iu / L.C) % Wp, yp = (iu / ((unsigned)L.C * Wp)) % Hp, n = iu / ((unsigned)L.C * Wp * Hp);
    bool oky, okx;
    const int sy = map_pad(yp - L.pad, L.H, L.kind, oky), sx = map_pad(xp - L.pad, L.W, L.kind, okx);
    float v = 0.f;
    if (c < Cin && oky && okx) v = x[(((size_t)n * Cin + c) * L.H + sy) * L.W + sx];
    dst[act_offset(L, N, n, yp, xp) + c] = __float2bfloat16_rn(v);
  }
}

// few input channels (the 3-channel frames of a stylising net, padded to a 16-channel operand with a reflect halo): one thread
// per padded pixel, the planes read as coalesced rows, the pixel written as whole 16-byte vectors - the element-wise kernel
// above spends four integer divisions and a 2-byte store per element (143 us for eight 640x360 frames; this one: ~20)
__global__ void __launch_bounds__(256) nchw_to_act_narrow_kernel(const float* __restrict__ x, int Cin, __nv_bfloat16* __restrict__ dst,
                                                                 ActLayout L, int N) {
  vst::pdl_grid_sync();
  const int Hp = L.H + 2 * L.pad, Wp = L.W + 2 * L.pad;
  const int xp = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y, n = blockIdx.z;
  if (xp >= Wp) return;
  bool oky, okx;
  const int sy = map_pad(yp - L.pad, L.H, L.kind, oky), sx = map_pad(xp - L.pad, L.W, L.kind, okx);
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = 0.f;
  if (oky && okx) {
    const float* p = x + ((size_t)n * Cin * L.H + sy) * L.W + sx;
    const size_t plane = (size_t)L.H * L.W;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < Cin) v[c] = __ldg(p + c * plane);
  }
  uint4 q;
  q.x = pk2<false>(v[0], v[1]); q.y = pk2<false>(v[2], v[3]); q.z = pk2<false>(v[4], v[5]); q.w = pk2<false>(v[6], v[7]);
  uint4* d = reinterpret_cast<uint4*>(dst + act_offset(L, N, n, yp, xp));
  d[0] = q;
  for (int k = 1; k < (L.C >> 3); ++k) d[k] = make_uint4(0, 0, 0, 0);
  (void)Hp;
}

// first channels of a padded operand -> fp32 NCHW: one thread per interior pixel, one 16-byte load, coalesced plane stores
__global__ void __launch_bounds__(256) act_to_nchw_first_kernel(const __nv_bfloat16* __restrict__ act, ActLayout L, int N, int c_count,
                                                                float* __restrict__ out) {
  vst::pdl_grid_sync();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  if (x >= L.W) return;
  const uint4 q = *reinterpret_cast<const uint4*>(act + act_offset(L, N, n, y + L.pad, x + L.pad));
  const float2 f0 = cvt2<false>(q.x), f1 = cvt2<false>(q.y), f2 = cvt2<false>(q.z), f3 = cvt2<false>(q.w);
  const float v[8] = {f0.x, f0.y, f1.x, f1.y, f2.x, f2.y, f3.x, f3.y};
  const size_t plane = (size_t)L.H * L.W;
  float* o = out + (size_t)n * c_count * plane + (size_t)y * L.W + x;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c < c_count) o[c * plane] = v[c];
}

// ---- weight packing (device; runs once at plan creation) -----------------------------------
// B[row][k]: row = cout (padded with zero rows), k = (tap*kbpt + kb)*BK + cl with cin = kb*BK + cl.
__global__ void pack_w_taps_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ B, int Cout, int Cin, int ksz,
                                   int rows, int kbpt, int BK, int half = 0) {
  vst::pdl_grid_sync();
  const int K = ksz * ksz * kbpt * BK;
  const size_t total = (size_t)rows * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % K, co = i / K;
    const int cl = k % BK, kb = (k / BK) % kbpt, t = k / (BK * kbpt);
    const int ci = kb * BK + cl;
    float v = 0.f;
    if (co < Cout && ci < Cin) v = w[((size_t)co * Cin + ci) * ksz * ksz + t];
    reinterpret_cast<uint16_t*>(B)[i] = f2h16(v, half);
  }
}
// conv1 (k=9): B[co][ky*KR + kx*Cin + c]
__global__ void pack_w_conv1_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ B, int Cout, int Cin, int rows,
                                    int KR, int half = 0) {
  vst::pdl_grid_sync();
  const int K = 9 * KR;
  const size_t total = (size_t)rows * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % K, co = i / K;
    const int ky = k / KR, r = k % KR;
    float v = 0.f;
    if (co < Cout && r < 9 * Cin) {
      const int kx = r / Cin, c = r % Cin;
      v = w[(((size_t)co * Cin + c) * 9 + ky) * 9 + kx];
    }
    reinterpret_cast<uint16_t*>(B)[i] = f2h16(v, half);
  }
}
// nearest-x2 upsample + 3x3 conv == 4 output phases of a 2x2 conv on the low-res tensor with
// pre-summed weights: phase (py,px), tap (dy,dx): sum over ky in S(py,dy), kx in S(px,dx),
// S(0,0)={0} S(0,1)={1,2} S(1,0)={0,1} S(1,1)={2}.   B[(ph*rows + co)][((dy*2+dx)*kbpt + kb)*BK + cl]
__global__ void pack_w_upphase_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ B, int Cout, int Cin,
                                      int rows, int kbpt, int BK, int half = 0) {
  vst::pdl_grid_sync();
  const int K = 4 * kbpt * BK;
  const size_t total = (size_t)4 * rows * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % K, co = (i / K) % rows, ph = i / ((size_t)K * rows);
    const int cl = k % BK, kb = (k / BK) % kbpt, t = k / (BK * kbpt);
    const int ci = kb * BK + cl, py = ph >> 1, px = ph & 1, dy = t >> 1, dx = t & 1;
    float v = 0.f;
    if (co < Cout && ci < Cin) {
      const int ky0 = (py == 0) ? (dy == 0 ? 0 : 1) : (dy == 0 ? 0 : 2), ky1 = (py == 0) ? (dy == 0 ? 0 : 2) : (dy == 0 ? 1 : 2);
      const int kx0 = (px == 0) ? (dx == 0 ? 0 : 1) : (dx == 0 ? 0 : 2), kx1 = (px == 0) ? (dx == 0 ? 0 : 2) : (dx == 0 ? 1 : 2);
      for (int ky = ky0; ky <= ky1; ++ky)
        for (int kx = kx0; kx <= kx1; ++kx) v += w[(((size_t)co * Cin + ci) * 3 + ky) * 3 + kx];
    }
    reinterpret_cast<uint16_t*>(B)[i] = f2h16(v, half);
  }
}

// deconv3 as a row convolution: B[kx*3 + co][ky*kbpt*BK + c] = w[co][c][ky][kx]; rows 27..31 zero.
__global__ void pack_w_rowconv_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ B, int Cout, int Cin,
                                      int ksz, int rows, int kbpt, int BK, int half = 0) {
  vst::pdl_grid_sync();
  const int K = ksz * kbpt * BK;
  const size_t total = (size_t)rows * K;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = i % K, row = i / K;
    const int ky = k / (kbpt * BK), ci = k % (kbpt * BK);
    float v = 0.f;
    if (row < ksz * Cout && ci < Cin) {
      const int kx = row / Cout, co = row % Cout;
      v = w[(((size_t)co * Cin + ci) * ksz + ky) * ksz + kx];
    }
    reinterpret_cast<uint16_t*>(B)[i] = f2h16(v, half);
  }
}

// experiment switch (VST_RC_BK16=1): 16-channel k-blocks for deconv3 cover its 48 channels exactly (3 MMAs per tap instead of
// the 4 of a zero-filled 64-channel block) - measured SLOWER (1.41 vs 0.92 ms: three 32-byte-row TMA boxes per tap), so off
static inline bool rowconv_bk16() {
  static const bool on = [] { const char* e = getenv("VST_RC_BK16"); return e ? atoi(e) != 0 : false; }();
  return on;
}
static inline int ew_grid(size_t total) {
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// BK choice for a channel count: 64 (SWIZZLE_128B) when C is a multiple of 64 or small enough that
// one zero-filled block covers it; 32 (SWIZZLE_64B) for multiples of 32; 16 otherwise.
static void choose_bk(int C, int* BK, int* kbpt) {
  if (C % 64 == 0) { *BK = 64; *kbpt = C / 64; }
  else if (C % 32 == 0) { *BK = 32; *kbpt = C / 32; }
  else if (C <= 64) { *BK = 64; *kbpt = 1; }   // TMA zero-fills channels C..63
  else { *BK = 16; *kbpt = cdiv(C, 16); }
}

// ---- one convolution stage of the plan -----------------------------------------------------
struct ConvStage {
  TapGemmParams tg;
  int BK;
  // IN parameters / stats / destination
  int C;                  // output channels
  int Ho, Wo;             // output extent
  double* stats;          // [N][C][2]
  const float* gamma;
  const float* beta;
  ActLayout dst;          // layout apply writes
  __nv_bfloat16* dst_buf;
  const __nv_bfloat16* res_buf;  // residual (padded layout) or null
  ActLayout res;
  int relu;
  float eps;              // 1e-5, rescaled for a stage whose input was pre-scaled (conv1 of the fp16 plan)
  // "fp16" plan, residual trunk (apply_hp_kernel): raw32 = this stage's conv output in fp32 NHWC (else P->raw, 16-bit),
  // res32 = the fp32 residual stream to add, out32 = where the fp32 copy of the result goes; hp = use the hp kernel
  int skip_apply;         // the consumer normalises this stage's raw output itself (fused input normalisation)
  int hp;
  float* raw32;
  const float* res32;
  float* out32;
};

}  // namespace vst

using namespace vst;

struct vst_plan {
  vst_net_desc d;
  std::vector<ConvStage> stages;   // 15 IN stages
  TapGemmParams final_tg;          // deconv3 (tanh epilogue)
  int final_BK;
  // buffers
  __nv_bfloat16* x9;
  int KR;
  __nv_bfloat16* raw;
  double* stats_all;
  size_t stats_bytes;
  float* final_bias;
  ActLayout feat_layout;
  __nv_bfloat16* feat_buf;
  std::vector<std::pair<__nv_bfloat16*, ActLayout>> act_bufs;  // per stage, for the debug hook
  int fuse_stats;
  int launches;
  int half = 0;                   // VST_PLAN_FP16: fp16 storage / operands + fp32 residual stream
  float in_scale = 1.f;           // frames are multiplied by this in the prologue (fp16 range); conv1's eps compensates exactly
  // optional per-launch CUDA-event timing of the 16 tap-GEMM launches (bench.py roofline)
  static constexpr int kTimingRing = 32;
  int stop_after = -1;            // test hook: run only stages 0..stop_after
  bool timing = false;
  std::vector<cudaEvent_t> ev;   // [ring][16][2]
  long fwd_count = 0;
  ~vst_plan() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
  }
};

namespace {

struct Arena {
  uint8_t* base;
  size_t size, off;
  void* take(size_t bytes) {
    off = (off + 1023) & ~(size_t)1023;
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

struct ReCoNetLayout {
  // channel widths along the net
  int cin, c1, c2, c3, d1, d2;
  int N, H, W;
};

// Total arena bytes; when `a.base` is non-null the same walk hands out the pointers.
struct Buffers {
  __nv_bfloat16 *x9, *raw, *p1, *p2, *t[3], *rep, *u1, *u2;
  float *t32[3], *raw32;   // fp16 plan: fp32 residual stream (unpadded NHWC) and the fp32 raw output of a block's second conv
  double* stats;
  float* gb;        // gamma/beta for 15 stages, then final bias
  __nv_bfloat16* wpk[16];
  float* wstage;    // fp32 staging for one raw weight tensor
  size_t wpk_elems[16];
  int KR;
};

static int conv1_kr(int cin) {
  const int k = 9 * cin;
  return k <= 32 ? 32 : round_up(k, 64);
}

static void plan_buffers(const vst_net_desc& d, Arena& a, Buffers& b) {
  const int N = d.N, H = d.H, W = d.W;
  b.KR = conv1_kr(d.in_ch);
  b.x9 = (__nv_bfloat16*)a.take((size_t)N * (H + 8) * W * b.KR * 2);
  size_t raw_elems = (size_t)N * H * W * (d.c1 > d.d2 ? d.c1 : d.d2);
  raw_elems = std::max(raw_elems, (size_t)N * (H / 2) * (W / 2) * (size_t)std::max(d.c2, d.d1));
  raw_elems = std::max(raw_elems, (size_t)N * (H / 4) * (W / 4) * (size_t)d.c3);
  b.raw = (__nv_bfloat16*)a.take(raw_elems * 2);
  b.p1 = (__nv_bfloat16*)a.take((size_t)N * (H + 2) * (W + 2) * d.c1 * 2);
  b.p2 = (__nv_bfloat16*)a.take((size_t)N * (H / 2 + 2) * (W / 2 + 2) * d.c2 * 2);
  for (int i = 0; i < 3; ++i) b.t[i] = (__nv_bfloat16*)a.take((size_t)N * (H / 4 + 2) * (W / 4 + 2) * d.c3 * 2);
  b.rep = (__nv_bfloat16*)a.take((size_t)N * (H / 4 + 2) * (W / 4 + 2) * d.c3 * 2);
  b.u1 = (__nv_bfloat16*)a.take((size_t)N * (H / 2 + 2) * (W / 2 + 2) * d.d1 * 2);
  b.u2 = (__nv_bfloat16*)a.take((size_t)N * (H + 8) * (W + 8) * d.d2 * 2);
  for (int i = 0; i < 3; ++i) b.t32[i] = nullptr;
  b.raw32 = nullptr;
  if (d.flags & VST_PLAN_FP16) {
    const size_t e32 = (size_t)N * (H / 4) * (W / 4) * d.c3 * sizeof(float);
    for (int i = 0; i < 3; ++i) b.t32[i] = (float*)a.take(e32);
    b.raw32 = (float*)a.take(e32);
  }
  b.stats = (double*)a.take((size_t)15 * N * 256 * 2 * sizeof(double));
  b.gb = (float*)a.take((size_t)(15 * 2 * 256 + 16) * sizeof(float));
  // packed weights
  const int cins[16] = {d.in_ch, d.c1, d.c2, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.d1, d.d2};
  const int couts[16] = {d.c1, d.c2, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.c3, d.d1, d.d2, 3};
  size_t max_w = 0;
  for (int l = 0; l < 16; ++l) {
    const int rows = round_up(couts[l], 16);
    size_t elems;
    if (l == 0) elems = (size_t)rows * 9 * b.KR;
    else {
      int BK, kbpt;
      choose_bk(cins[l], &BK, &kbpt);
      const int taps = (l == 15) ? 9 /*row taps*/ : (l == 13 || l == 14) ? 16 /*4 phases x 4 taps*/ : 9;
      elems = (size_t)(l == 15 ? 32 : rows) * taps * kbpt * BK;
    }
    b.wpk_elems[l] = elems;
    b.wpk[l] = (__nv_bfloat16*)a.take(elems * 2);
    const int ks = (l == 0 || l == 15) ? 9 : 3;
    max_w = std::max(max_w, (size_t)couts[l] * cins[l] * ks * ks);
  }
  b.wstage = (float*)a.take(max_w * sizeof(float));
}

static void fill_taps_3x3_s1(TapGemmParams& p) {
  p.n_taps = 9;
  for (int t = 0; t < 9; ++t) { p.tap_dy[t] = t / 3; p.tap_dx[t] = t % 3; p.tap_pl[t] = 0; }
}
static void fill_taps_3x3_s2(TapGemmParams& p) {
  p.n_taps = 9;
  for (int t = 0; t < 9; ++t) {
    const int ky = t / 3, kx = t % 3;
    p.tap_dy[t] = ky >> 1; p.tap_dx[t] = kx >> 1; p.tap_pl[t] = (ky & 1) * 2 + (kx & 1);
  }
}
static void fill_taps_upphase(TapGemmParams& p) {
  p.n_taps = 4;
  p.n_phase = 4;
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    p.ph_oy[ph] = py; p.ph_ox[ph] = px;
    for (int t = 0; t < 4; ++t) {
      p.tap_dy[ph * 4 + t] = py + (t >> 1);
      p.tap_dx[ph * 4 + t] = px + (t & 1);
      p.tap_pl[ph * 4 + t] = 0;
    }
  }
}

static void tg_defaults(TapGemmParams& p, int N) {
  memset(&p, 0, sizeof(p));
  p.n_img = N;
  p.n_phase = 1;
  p.n_ntile = 1;
  p.out_mul = 1;
}

// A-operand tensor map over a padded activation buffer
static int tmap_for_act(CUtensorMap* m, const __nv_bfloat16* buf, const ActLayout& L, int N, int BK, int TW, int TH) {
  const int Hp = L.H + 2 * L.pad, Wp = L.W + 2 * L.pad;
  if (L.parity) {
    const size_t img = (size_t)(Hp / 2) * (Wp / 2) * L.C;
    return make_tmap_act(m, buf, L.C, Wp / 2, Hp / 2, N, 4, L.C, (size_t)(Wp / 2) * L.C, img, img * N, BK, TW, TH);
  }
  const size_t img = (size_t)Hp * Wp * L.C;
  return make_tmap_act(m, buf, L.C, Wp, Hp, N, 1, L.C, (size_t)Wp * L.C, img, img * N, BK, TW, TH);
}

}  // namespace

extern "C" {

size_t vst_plan_arena_bytes(const vst_net_desc* d) {
  if (!d) return 0;
  Arena a{nullptr, 0, 0};
  Buffers b;
  plan_buffers(*d, a, b);
  return a.off + 4096;
}

void vst_plan_destroy(vst_plan* p) { delete p; }
int vst_plan_launches(const vst_plan* p) { return p ? p->launches : 0; }

int vst_plan_create(const vst_net_desc* d, const float* const* weights_host, int n_tensors, void* arena,
                    size_t arena_bytes, void* stream, vst_plan** out) {
  VST_CHECK_ARG(d && weights_host && arena && out, "plan_create: NULL argument");
  VST_CHECK_ARG(d->net == VST_NET_RECONET, "plan_create: unknown net %d", d->net);
  VST_CHECK_ARG(d->N > 0 && d->H >= 16 && d->W >= 16 && d->H % 4 == 0 && d->W % 4 == 0,
                "plan_create: H and W must be multiples of 4 (SURVEY.md Q14)");
  VST_CHECK_ARG(d->c1 % 8 == 0 && d->c2 % 8 == 0 && d->c3 % 8 == 0 && d->d1 % 8 == 0 && d->d2 % 8 == 0,
                "plan_create: channel widths must be multiples of 8");
  VST_CHECK_ARG(d->c3 <= 256 && d->c1 <= 256 && d->c2 <= 256, "plan_create: widths > 256 unsupported");
  VST_CHECK_ARG(d->in_ch >= 3 && d->in_ch % 3 == 0 && d->in_ch <= 27, "plan_create: in_ch must be 3*frames");
  VST_CHECK_ARG(n_tensors == 62, "plan_create: expected the 62 state_dict tensors, got %d", n_tensors);
  VST_DEVPTR(arena);
  if (arena_bytes < vst_plan_arena_bytes(d)) {
    set_error("plan_create: arena %zu < %zu bytes", arena_bytes, vst_plan_arena_bytes(d));
    return VST_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  vst_plan* P = new vst_plan();
  P->d = *d;
  const char* fs = getenv("VST_FUSE_STATS");
  P->fuse_stats = fs ? atoi(fs) : 1;
  P->half = (d->flags & VST_PLAN_FP16) ? 1 : 0;
  if (P->half) {
    P->fuse_stats = 1;           // the fp32 raw tensors have no stand-alone statistics pass
    P->in_scale = 1.f / 16.f;    // frames 0..255 -> 0..16: conv1's raw output stays far inside the fp16 range (|w| * 243 * 16)
  }
  const int half = P->half;
  Arena a{(uint8_t*)arena, arena_bytes, 0};
  Buffers b;
  plan_buffers(*d, a, b);
  P->x9 = b.x9; P->KR = b.KR; P->raw = b.raw; P->stats_all = b.stats;
  P->stats_bytes = (size_t)15 * d->N * 256 * 2 * sizeof(double);
  const int N = d->N, H = d->H, W = d->W;

  // ---- the state_dict, in the reference's registration order (RC/network.py:157-169):
  // conv{1,2,3}: conv2d.weight, conv2d.bias, instance.weight, instance.bias            (3 x 4)
  // res{1..5}: conv1.conv2d.{w,b}, in1.{w,b}, conv2.conv2d.{w,b}, in2.{w,b}            (5 x 8)
  // deconv{1,2}: conv2d.{w,b}, instance.{w,b}                                          (2 x 4)
  // deconv3: conv2d.{w,b}                                                              (2)
  const float* conv_w[16]; const float* gam[15]; const float* bet[15]; const float* last_bias;
  {
    int t = 0, l = 0;
    for (int i = 0; i < 3; ++i) { conv_w[l] = weights_host[t]; gam[l] = weights_host[t + 2]; bet[l] = weights_host[t + 3]; t += 4; ++l; }
    for (int i = 0; i < 5; ++i) {
      conv_w[l] = weights_host[t]; gam[l] = weights_host[t + 2]; bet[l] = weights_host[t + 3]; ++l;
      conv_w[l] = weights_host[t + 4]; gam[l] = weights_host[t + 6]; bet[l] = weights_host[t + 7]; ++l;
      t += 8;
    }
    for (int i = 0; i < 2; ++i) { conv_w[l] = weights_host[t]; gam[l] = weights_host[t + 2]; bet[l] = weights_host[t + 3]; t += 4; ++l; }
    conv_w[15] = weights_host[t]; last_bias = weights_host[t + 1];
  }
  const int cins[16] = {d->in_ch, d->c1, d->c2, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->d1, d->d2};
  const int couts[16] = {d->c1, d->c2, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->c3, d->d1, d->d2, 3};

  // gamma / beta / final bias -> device.  (The conv biases in front of an InstanceNorm are
  // mathematically cancelled by it - SURVEY.md Q6 - and are not uploaded.)
  for (int l = 0; l < 15; ++l) {
    VST_CUDA(cudaMemcpyAsync(b.gb + (size_t)l * 512, gam[l], couts[l] * sizeof(float), cudaMemcpyHostToDevice, st));
    VST_CUDA(cudaMemcpyAsync(b.gb + (size_t)l * 512 + 256, bet[l], couts[l] * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  P->final_bias = b.gb + (size_t)15 * 512;
  VST_CUDA(cudaMemcpyAsync(P->final_bias, last_bias, 3 * sizeof(float), cudaMemcpyHostToDevice, st));

  // ---- pack weights on the device
  for (int l = 0; l < 16; ++l) {
    const int ks = (l == 0 || l == 15) ? 9 : 3;
    const size_t wn = (size_t)couts[l] * cins[l] * ks * ks;
    VST_CUDA(cudaMemcpyAsync(b.wstage, conv_w[l], wn * sizeof(float), cudaMemcpyHostToDevice, st));
    const int rows = round_up(couts[l], 16);
    if (l == 0) {
      vst::launch(pack_w_conv1_kernel, ew_grid(b.wpk_elems[l]), 256, 0, st, b.wstage, b.wpk[l], couts[l], cins[l], rows, b.KR, half);
    } else {
      int BK, kbpt;
      choose_bk(cins[l], &BK, &kbpt);
      if (l == 15 && rowconv_bk16() && cins[l] % 16 == 0 && cins[l] < 64) { BK = 16; kbpt = cins[l] / 16; }
      if (l == 15)
        vst::launch(pack_w_rowconv_kernel, ew_grid(b.wpk_elems[l]), 256, 0, st, b.wstage, b.wpk[l], 3, cins[l], 9, 32, kbpt, BK, half);
      else if (l == 13 || l == 14)
        vst::launch(pack_w_upphase_kernel, ew_grid(b.wpk_elems[l]), 256, 0, st, b.wstage, b.wpk[l], couts[l], cins[l], rows, kbpt, BK, half);
      else
        vst::launch(pack_w_taps_kernel, ew_grid(b.wpk_elems[l]), 256, 0, st, b.wstage, b.wpk[l], couts[l], cins[l], ks, rows, kbpt, BK, half);
    }
    VST_LAUNCH_CHECK();
    // the staging buffer is reused: the next H2D copy is stream-ordered after this kernel
  }

  // ---- activation layouts
  const ActLayout L_p1{H, W, d->c1, 1, PADK_REFLECT, 1};                 // conv1 out -> conv2 (s2)
  const ActLayout L_p2{H / 2, W / 2, d->c2, 1, PADK_REFLECT, 1};         // conv2 out -> conv3 (s2)
  const ActLayout L_t{H / 4, W / 4, d->c3, 1, PADK_REFLECT, 0};          // trunk
  const ActLayout L_rep{H / 4, W / 4, d->c3, 1, PADK_REPLICATE, 0};      // res5 out -> deconv1 phases
  const ActLayout L_u1{H / 2, W / 2, d->d1, 1, PADK_REPLICATE, 0};       // deconv1 out -> deconv2 phases
  const ActLayout L_u2{H, W, d->d2, 4, PADK_REFLECT, 0};                 // deconv2 out -> deconv3 (k9)

  auto add_stage = [&](int l, const CUtensorMap* /*unused*/, TapGemmParams tg, int BK, int C, int Ho, int Wo,
                       const ActLayout& dst, __nv_bfloat16* dst_buf, const __nv_bfloat16* res_buf, const ActLayout& res,
                       int relu) {
    ConvStage s;
    s.tg = tg; s.BK = BK; s.C = C; s.Ho = Ho; s.Wo = Wo;
    s.stats = b.stats + (size_t)l * N * 256 * 2;
    s.gamma = b.gb + (size_t)l * 512; s.beta = b.gb + (size_t)l * 512 + 256;
    s.dst = dst; s.dst_buf = dst_buf; s.res_buf = res_buf; s.res = res; s.relu = relu;
    s.eps = 1e-5f; s.hp = 0; s.raw32 = nullptr; s.res32 = nullptr; s.out32 = nullptr; s.skip_apply = 0;
    s.tg.stats = P->fuse_stats ? s.stats : nullptr;
    P->stages.push_back(s);
    P->act_bufs.push_back({dst_buf, dst});
  };

  // generic builder for a conv reading a padded activation buffer
  auto build = [&](int l, const __nv_bfloat16* in_buf, const ActLayout& in, int Ho, int Wo, int kind /*0 s1,1 s2,2 up*/,
                   TapGemmParams& tg, int& BK) -> int {
    tg_defaults(tg, N);
    tg.half = half;
    int kbpt;
    choose_bk(cins[l], &BK, &kbpt);
    tg.kb_per_tap = kbpt;
    const int gh = (kind == 2) ? Ho / 2 : Ho, gw = (kind == 2) ? Wo / 2 : Wo;  // tile grid extent
    tg.MT = choose_mt(round_up(couts[l], 16));
    choose_tile(gh, gw, tg.MT, &tg.TW, &tg.TH);
    tg.tiles_x = cdiv(gw, tg.TW); tg.tiles_y = cdiv(gh, tg.TH);
    tg.Ho = gh; tg.Wo = gw;
    tg.N_mma = round_up(couts[l], 16);
    tg.Cout = couts[l];
    tg.Hout = Ho; tg.Wout = Wo; tg.out_cstride = couts[l];
    tg.epi_mode = TG_EPI_BF16_NHWC;
    tg.out0 = b.raw;
    if (kind == 0) fill_taps_3x3_s1(tg);
    else if (kind == 1) fill_taps_3x3_s2(tg);
    else { fill_taps_upphase(tg); tg.out_mul = 2; }
    tapgemm_plan(tg, BK);
    int r = tmap_for_act(&tg.tmA, in_buf, in, N, BK, tg.TW, tapgemm_box_rows(tg));
    if (r != VST_OK) return r;
    const int K = tg.n_taps * kbpt * BK;
    return make_tmap_wgt(&tg.tmB, b.wpk[l], K, tg.N_mma * tg.n_phase, BK, tapgemm_b_box_rows(tg));
  };

  TapGemmParams tg;
  int BK, r;
  // conv1: 9 row taps over X9
  {
    tg_defaults(tg, N);
    tg.half = half;
    BK = b.KR <= 32 ? 32 : 64;
    tg.kb_per_tap = b.KR / BK;
    tg.MT = choose_mt(round_up(d->c1, 16));
    choose_tile(H, W, tg.MT, &tg.TW, &tg.TH);
    tg.tiles_x = cdiv(W, tg.TW); tg.tiles_y = cdiv(H, tg.TH);
    tg.Ho = H; tg.Wo = W; tg.N_mma = round_up(d->c1, 16); tg.Cout = d->c1;
    tg.Hout = H; tg.Wout = W; tg.out_cstride = d->c1; tg.epi_mode = TG_EPI_BF16_NHWC; tg.out0 = b.raw;
    tg.n_taps = 9;
    for (int t = 0; t < 9; ++t) { tg.tap_dy[t] = t; tg.tap_dx[t] = 0; tg.tap_pl[t] = 0; }
    tapgemm_plan(tg, BK);
    const size_t img = (size_t)(H + 8) * W * b.KR;
    r = make_tmap_act(&tg.tmA, b.x9, b.KR, W, H + 8, N, 1, b.KR, (size_t)W * b.KR, img, img * N, BK, tg.TW, tapgemm_box_rows(tg));
    if (r != VST_OK) { delete P; return r; }
    r = make_tmap_wgt(&tg.tmB, b.wpk[0], 9 * b.KR, tg.N_mma, BK, tapgemm_b_box_rows(tg));
    if (r != VST_OK) { delete P; return r; }
    add_stage(0, nullptr, tg, BK, d->c1, H, W, L_p1, b.p1, nullptr, L_p1, 1);
    P->stages.back().eps = 1e-5f * P->in_scale * P->in_scale;   // IN(conv(s x)) == IN(conv(x)) when eps scales by s^2
  }
  // conv2, conv3 (stride 2 on parity planes)
  r = build(1, b.p1, L_p1, H / 2, W / 2, 1, tg, BK); if (r != VST_OK) { delete P; return r; }
  add_stage(1, nullptr, tg, BK, d->c2, H / 2, W / 2, L_p2, b.p2, nullptr, L_p2, 1);
  r = build(2, b.p2, L_p2, H / 4, W / 4, 1, tg, BK); if (r != VST_OK) { delete P; return r; }
  add_stage(2, nullptr, tg, BK, d->c3, H / 4, W / 4, L_t, b.t[0], nullptr, L_t, 1);
  if (half) { ConvStage& s2 = P->stages.back(); s2.hp = 1; s2.out32 = b.t32[0]; }   // conv3 output opens the fp32 residual stream
  // residual blocks: x = t[cur]; mid = t[(cur+1)%3]; out = t[(cur+2)%3] (last block -> rep)
  int cur = 0;
  for (int blk = 0; blk < 5; ++blk) {
    const int mid = (cur + 1) % 3, nxt = (cur + 2) % 3;
    r = build(3 + 2 * blk, b.t[cur], L_t, H / 4, W / 4, 0, tg, BK); if (r != VST_OK) { delete P; return r; }
    add_stage(3 + 2 * blk, nullptr, tg, BK, d->c3, H / 4, W / 4, L_t, b.t[mid], nullptr, L_t, 1);
    r = build(4 + 2 * blk, b.t[mid], L_t, H / 4, W / 4, 0, tg, BK); if (r != VST_OK) { delete P; return r; }
    if (half) { tg.out_f32 = 1; tg.out0 = b.raw32; }            // fp32 raw output: its IN + residual add run in fp32
    if (blk < 4) add_stage(4 + 2 * blk, nullptr, tg, BK, d->c3, H / 4, W / 4, L_t, b.t[nxt], b.t[cur], L_t, 0);
    else add_stage(4 + 2 * blk, nullptr, tg, BK, d->c3, H / 4, W / 4, L_rep, b.rep, b.t[cur], L_t, 0);
    if (half) {
      ConvStage& sb = P->stages.back();
      sb.hp = 1; sb.raw32 = b.raw32; sb.res32 = b.t32[cur]; sb.out32 = b.t32[nxt];
    }
    cur = nxt;
  }
  P->feat_layout = L_rep; P->feat_buf = b.rep;
  // deconv1 / deconv2: 4-phase 2x2 convs on replicate-padded low-res tensors
  r = build(13, b.rep, L_rep, H / 2, W / 2, 2, tg, BK); if (r != VST_OK) { delete P; return r; }
  add_stage(13, nullptr, tg, BK, d->d1, H / 2, W / 2, L_u1, b.u1, nullptr, L_u1, 1);
  r = build(14, b.u1, L_u1, H, W, 2, tg, BK); if (r != VST_OK) { delete P; return r; }
  add_stage(14, nullptr, tg, BK, d->d2, H, W, L_u2, b.u2, nullptr, L_u2, 1);
  // deconv3 (k9, Cout=3): row convolution - 9 ky taps, N = 27 (kx,co) columns padded to 32,
  // 120 output pixels per 128-pixel tile, kx-shift-sum + tanh map + uint8 pack in the epilogue
  {
    TapGemmParams& f = P->final_tg;
    tg_defaults(f, N);
    f.half = half;
    int kbpt;
    choose_bk(d->d2, &P->final_BK, &kbpt);
    // row-streaming wants the smallest ring slot: 16-channel k-blocks cover 48 channels exactly (no zero-filled lanes)
    if (rowconv_bk16() && d->d2 % 16 == 0 && d->d2 < 64) { P->final_BK = 16; kbpt = d->d2 / 16; }
    f.kb_per_tap = kbpt;
    { const char* e = getenv("VST_RC_MT"); f.MT = e ? atoi(e) : 2; if (f.MT != 1 && f.MT != 2 && f.MT != 4) f.MT = 2; }
    f.TW = 128; f.TH = f.MT; f.tile_step_x = 120;   // MT output rows x 120 pixels per CTA tile
    f.tiles_x = cdiv(W, 120); f.tiles_y = cdiv(H, f.MT);
    f.Ho = H; f.Wo = W; f.N_mma = 32; f.Cout = 3; f.Hout = H; f.Wout = W; f.out_cstride = 3;
    f.epi_mode = TG_EPI_ROWCONV; f.rc_k = 9; f.rc_co = 3; f.act = VST_ACT_RECONET_OUT; f.bias = P->final_bias;
    f.n_taps = 9;
    for (int t = 0; t < 9; ++t) { f.tap_dy[t] = t; f.tap_dx[t] = 0; f.tap_pl[t] = 0; }
    tapgemm_plan(f, P->final_BK);
    // Fused input normalisation (VST_FUSE_DC3=0 disables): in the accumulator-ring mode every input row is loaded once, so
    // deconv3 can normalise deconv2's RAW output inside its own shared-memory ring and the full-resolution apply pass of stage
    // 14 - 0.8 GB read + 0.76 GB written per 4 frames at 1080p - disappears.  The operand is then the unpadded raw tensor:
    // taps are relative to the frame (dy = t - 4, dx = -4), rows mirror in the producer, edge columns in the transform warps.
    static const bool fuse_env = [] { const char* e = getenv("VST_FUSE_DC3"); return e ? atoi(e) != 0 : true; }();
    const bool fuse = fuse_env && f.stream == 2 && P->final_BK == 64 && kbpt == 1 && d->d2 <= 64 && P->fuse_stats;
    if (fuse) {
      for (int t = 0; t < 9; ++t) { f.tap_dy[t] = t - 4; f.tap_dx[t] = -4; }
      tapgemm_plan(f, P->final_BK);               // same mode, s_dy0 = -4
      ConvStage& s14 = P->stages.back();
      s14.skip_apply = 1;
      f.fuse_in = 1; f.in_relu = 1; f.in_H = H; f.in_W = W; f.in_C = d->d2;
      f.in_stats = s14.stats; f.in_gamma = s14.gamma; f.in_beta = s14.beta; f.in_eps = s14.eps;
      const size_t img = (size_t)H * W * d->d2;
      r = make_tmap_act(&f.tmA, b.raw, d->d2, W, H, N, 1, d->d2, (size_t)W * d->d2, img, img * N, P->final_BK, f.TW, tapgemm_box_rows(f));
      if (r != VST_OK) { delete P; return r; }
    } else {
      r = tmap_for_act(&f.tmA, b.u2, L_u2, N, P->final_BK, f.TW, tapgemm_box_rows(f)); if (r != VST_OK) { delete P; return r; }
    }
    r = make_tmap_wgt(&f.tmB, b.wpk[15], 9 * kbpt * P->final_BK, 32, P->final_BK, 32); if (r != VST_OK) { delete P; return r; }
  }
  VST_CUDA(cudaStreamSynchronize(st));
  P->launches = 2 /*memset+prologue*/ + 15 * (P->fuse_stats ? 2 : 3) + 1 - (P->final_tg.fuse_in ? 1 : 0);
  *out = P;
  return VST_OK;
}

static int plan_forward_impl(vst_plan* P, const float* x, const uint8_t* x_bgr8, float* img_out, uint8_t* u8_out, float* features_out,
                             void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const vst_net_desc& d = P->d;
  const int N = d.N;
  VST_CUDA(cudaMemsetAsync(P->stats_all, 0, P->stats_bytes, st));
  if (x_bgr8) {
    dim3 grid(cdiv(d.W, 256), d.H + 8, N);
    if (x9_staged() == 2) vst::launch(prologue_x9_seg_kernel<true>, grid, 256, 0, st, (const void*)x_bgr8, P->x9, N, d.H, d.W, P->half, P->in_scale);
    else vst::launch(prologue_x9_bgr8_kernel<32>, grid, 256, 0, st, x_bgr8, P->x9, N, d.H, d.W, P->half, P->in_scale, x9_staged());
    VST_LAUNCH_CHECK();
  } else {
    dim3 grid(cdiv(d.W, 256), d.H + 8, N);
    const int hf = P->half;
    const float sc = P->in_scale;
    const int sg = x9_staged();
    if (P->KR == 32 && d.in_ch == 3 && sg == 2) vst::launch(prologue_x9_seg_kernel<false>, grid, 256, 0, st, (const void*)x, P->x9, N, d.H, d.W, hf, sc);
    else if (P->KR == 32 && d.in_ch == 3) vst::launch(prologue_x9_kernel<32, 3>, grid, 256, 0, st, x, P->x9, N, d.in_ch, d.H, d.W, hf, sc, sg);
    else if (P->KR == 32) vst::launch(prologue_x9_kernel<32, 0>, grid, 256, 0, st, x, P->x9, N, d.in_ch, d.H, d.W, hf, sc, sg);
    else if (P->KR == 64) vst::launch(prologue_x9_kernel<64, 0>, grid, 256, 0, st, x, P->x9, N, d.in_ch, d.H, d.W, hf, sc, 0);
    else if (P->KR == 128) vst::launch(prologue_x9_kernel<128, 0>, grid, 256, 0, st, x, P->x9, N, d.in_ch, d.H, d.W, hf, sc, 0);
    else if (P->KR == 192) vst::launch(prologue_x9_kernel<192, 0>, grid, 256, 0, st, x, P->x9, N, d.in_ch, d.H, d.W, hf, sc, 0);
    else if (P->KR == 256) vst::launch(prologue_x9_kernel<256, 0>, grid, 256, 0, st, x, P->x9, N, d.in_ch, d.H, d.W, hf, sc, 0);
    else { set_error("plan_forward: KR=%d unsupported", P->KR); return VST_EUNSUPPORTED; }
    VST_LAUNCH_CHECK();
  }
  cudaEvent_t* evs = P->timing ? &P->ev[(size_t)(P->fwd_count % vst_plan::kTimingRing) * 32] : nullptr;
  for (size_t i = 0; i < P->stages.size(); ++i) {
    ConvStage& s = P->stages[i];
    if (evs) cudaEventRecord(evs[2 * i], st);
    int r = launch_tapgemm(s.tg, s.BK, st);
    if (evs) cudaEventRecord(evs[2 * i + 1], st);
    if (r != VST_OK) return r;
    const int HW = s.Ho * s.Wo;
    if (!P->fuse_stats) {
      const int groups = s.C / 8;
      const int threads = 256 / groups * groups;  // whole number of pixel lanes
      const int ppb = std::max(64, cdiv(HW, kNumSMs * 4));
      dim3 grid(cdiv(HW, ppb), N);
      vst::launch(stats_kernel, grid, threads, 2 * s.C * sizeof(float), st, P->raw, s.stats, HW, s.C, ppb);
      VST_LAUNCH_CHECK();
    }
    if (s.skip_apply) {
      if ((int)i == P->stop_after) return VST_OK;
      continue;
    }
    if (s.hp) {
      // fp32 residual stream of the fp16 plan (see apply_hp_kernel)
      const int Hp = s.dst.H + 2 * s.dst.pad;
      int rpb = cdiv(Hp * N, kNumSMs * 8);
      if (rpb < 1) rpb = 1;
      dim3 grid(cdiv(Hp, rpb), N);
      const int groups = s.C / 8, threads = 256 / groups * groups;
      const size_t smem = 2 * s.C * sizeof(float);
      if (s.raw32)
        vst::launch(apply_hp_kernel<true, true>, grid, threads, smem, st, s.raw32, s.stats, s.gamma, s.beta, s.res32, s.dst_buf, s.dst, s.out32, N,
                                                                 s.eps, s.relu, rpb);
      else
        vst::launch(apply_hp_kernel<false, false>, grid, threads, smem, st, P->raw, s.stats, s.gamma, s.beta, nullptr, s.dst_buf, s.dst, s.out32, N,
                                                                   s.eps, s.relu, rpb);
    } else {
      launch_apply(P->raw, s.stats, s.gamma, s.beta, s.res_buf, s.res, s.dst_buf, s.dst, N, s.eps, s.relu, st, P->half);
    }
    VST_LAUNCH_CHECK();
    if ((int)i == P->stop_after) return VST_OK;
  }
  if (features_out) {
    vst::launch(act_to_nchw_kernel, ew_grid((size_t)N * P->feat_layout.C * P->feat_layout.H * P->feat_layout.W), 256, 0, st, P->feat_buf, P->feat_layout, N, features_out, P->half);
    VST_LAUNCH_CHECK();
  }
  P->final_tg.out0 = img_out;
  P->final_tg.out_u8 = u8_out;
  if (evs) cudaEventRecord(evs[30], st);
  int r = launch_tapgemm(P->final_tg, P->final_BK, st);
  if (evs) cudaEventRecord(evs[31], st);
  P->fwd_count++;
  return r;
}

// Two half-batch plans in lock step on ONE stream: every tap-GEMM launch of one plan carries the pending InstanceNorm apply
// pass of the other as a rider (ApplyRider, tc_conv.cuh), so the 14 HBM-bound apply passes per plan run beside the MMAs
// instead of between them.  Order:  G_a(0) | G_b(0)+Ap_a(0) | G_a(1)+Ap_b(0) | G_b(1)+Ap_a(1) | ... | final_a | final_b
// - stream order alone gives every dependency (Ap(i) after G(i), G(i+1) after Ap(i), each plan has its own buffers).
static void set_rider(TapGemmParams& tg, const vst_plan* Q, const ConvStage& s) {
  ApplyRider& r = tg.rider;
  static const int dbg = [] { const char* e = getenv("VST_RIDER_DBG"); return e ? atoi(e) : 0; }();   // 1: riders launched but idle (timing experiments; wrong frames)
  r.on = dbg == 1 ? 2 : dbg == 2 ? 3 : 1;
  r.raw = Q->raw; r.stats = s.stats; r.gamma = s.gamma; r.beta = s.beta;
  r.residual = s.res_buf; r.RL = s.res; r.dst = s.dst_buf; r.DL = s.dst;
  r.N = Q->d.N; r.relu = s.relu; r.eps = s.eps;
}

static int launch_prologue(vst_plan* P, const float* x, cudaStream_t st) {
  const vst_net_desc& d = P->d;
  dim3 grid(cdiv(d.W, 256), d.H + 8, d.N);
  if (P->KR == 32 && d.in_ch == 3 && x9_staged() == 2)
    vst::launch(prologue_x9_seg_kernel<false>, grid, 256, 0, st, (const void*)x, P->x9, d.N, d.H, d.W, P->half, P->in_scale);
  else if (P->KR == 32 && d.in_ch == 3)
    vst::launch(prologue_x9_kernel<32, 3>, grid, 256, 0, st, x, P->x9, d.N, d.in_ch, d.H, d.W, P->half, P->in_scale, x9_staged());
  else return VST_EUNSUPPORTED;
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_plan_forward_pair(vst_plan* Pa, vst_plan* Pb, const float* xa, const float* xb, uint8_t* u8_a, uint8_t* u8_b,
                          float* img_a, float* img_b, void* stream) {
  VST_CHECK_ARG(Pa && Pb && xa && xb, "plan_forward_pair: NULL argument");
  VST_CHECK_ARG((u8_a || img_a) && (u8_b || img_b), "plan_forward_pair: no output requested");
  VST_CHECK_ARG(Pa != Pb, "plan_forward_pair: the two plans must be distinct (own buffers)");
  VST_DEVPTR(xa);
  VST_DEVPTR(xb);
  vst_plan* P[2] = {Pa, Pb};
  for (int k = 0; k < 2; ++k) {
    if (P[k]->half || !P[k]->fuse_stats || P[k]->timing || P[k]->stop_after >= 0 || P[k]->KR != 32 || P[k]->d.in_ch != 3 ||
        P[k]->stages.size() != Pa->stages.size() || P[k]->final_tg.rider.on) {
      set_error("plan_forward_pair: needs two bf16 plans of the same network (no timing / stop hooks)");
      return VST_EUNSUPPORTED;
    }
    for (const ConvStage& s : P[k]->stages)
      if (s.hp || s.dst.C % 8 != 0 || s.dst.C > 1024) { set_error("plan_forward_pair: unsupported stage"); return VST_EUNSUPPORTED; }
  }
  cudaStream_t st = (cudaStream_t)stream;
  VST_CUDA(cudaMemsetAsync(Pa->stats_all, 0, Pa->stats_bytes, st));
  VST_CUDA(cudaMemsetAsync(Pb->stats_all, 0, Pb->stats_bytes, st));
  int r = launch_prologue(Pa, xa, st); if (r != VST_OK) return r;
  r = launch_prologue(Pb, xb, st); if (r != VST_OK) return r;
  const size_t S = Pa->stages.size();
  for (size_t i = 0; i < S; ++i) {
    for (int k = 0; k < 2; ++k) {
      vst_plan* me = P[k];
      vst_plan* other = P[k ^ 1];
      ConvStage& s = me->stages[i];
      // pending apply of the other plan: plan a carries Ap_b(i-1), plan b carries Ap_a(i)
      const ConvStage* pend = nullptr;
      if (k == 0 && i > 0) pend = &other->stages[i - 1];
      if (k == 1) pend = &other->stages[i];
      if (pend && pend->skip_apply) pend = nullptr;
      if (pend && !s.tg.fuse_in) set_rider(s.tg, other, *pend);
      else if (pend) {   // a fused-input tap-GEMM has no spare warps: run the apply as its own launch first
        launch_apply(other->raw, pend->stats, pend->gamma, pend->beta, pend->res_buf, pend->res, pend->dst_buf, pend->dst, other->d.N,
                     pend->eps, pend->relu, st, 0);
        VST_LAUNCH_CHECK();
      }
      r = launch_tapgemm(s.tg, s.BK, st);
      s.tg.rider.on = 0;
      if (r != VST_OK) return r;
    }
  }
  // the last stage of plan b: its apply (unless the final layer normalises its own input) has no tap-GEMM left to ride on
  {
    const ConvStage& s = Pb->stages[S - 1];
    if (!s.skip_apply) {
      launch_apply(Pb->raw, s.stats, s.gamma, s.beta, s.res_buf, s.res, s.dst_buf, s.dst, Pb->d.N, s.eps, s.relu, st, 0);
      VST_LAUNCH_CHECK();
    }
  }
  float* imgs[2] = {img_a, img_b};
  uint8_t* u8s[2] = {u8_a, u8_b};
  for (int k = 0; k < 2; ++k) {
    P[k]->final_tg.out0 = imgs[k];
    P[k]->final_tg.out_u8 = u8s[k];
    r = launch_tapgemm(P[k]->final_tg, P[k]->final_BK, st);
    if (r != VST_OK) return r;
    P[k]->fwd_count++;
  }
  return VST_OK;
}

int vst_plan_forward(vst_plan* P, const float* x, float* img_out, uint8_t* u8_out, float* features_out, void* stream) {
  VST_CHECK_ARG(P && x, "plan_forward: NULL argument");
  VST_CHECK_ARG(img_out || u8_out, "plan_forward: no output requested");
  VST_DEVPTR(x);
  return plan_forward_impl(P, x, nullptr, img_out, u8_out, features_out, stream);
}

int vst_plan_forward_bgr8(vst_plan* P, const uint8_t* frames_bgr, float* img_out, uint8_t* u8_out, float* features_out, void* stream) {
  VST_CHECK_ARG(P && frames_bgr, "plan_forward_bgr8: NULL argument");
  VST_CHECK_ARG(img_out || u8_out, "plan_forward_bgr8: no output requested");
  VST_CHECK_ARG(P->d.in_ch == 3 && P->KR == 32, "plan_forward_bgr8: single-frame networks only (in_ch == 3)");
  VST_DEVPTR(frames_bgr);
  return plan_forward_impl(P, nullptr, frames_bgr, img_out, u8_out, features_out, stream);
}

int vst_plan_set_timing(vst_plan* P, int enable) {
  VST_CHECK_ARG(P, "set_timing: NULL plan");
  if (enable && P->ev.empty()) {
    P->ev.resize((size_t)vst_plan::kTimingRing * 32);
    for (auto& e : P->ev) VST_CUDA(cudaEventCreate(&e));
  }
  P->timing = enable != 0;
  P->fwd_count = 0;
  return VST_OK;
}

int vst_plan_get_timing(vst_plan* P, float* ms_out, int* n_forwards) {
  VST_CHECK_ARG(P && ms_out && n_forwards, "get_timing: NULL argument");
  VST_CHECK_ARG(!P->ev.empty(), "get_timing: timing was never enabled");
  const int n = (int)std::min<long>(P->fwd_count, vst_plan::kTimingRing);
  for (int i = 0; i < 16; ++i) ms_out[i] = 0.f;
  for (int f = 0; f < n; ++f)
    for (int i = 0; i < 16; ++i) {
      float ms = 0.f;
      VST_CUDA(cudaEventElapsedTime(&ms, P->ev[(size_t)f * 32 + 2 * i], P->ev[(size_t)f * 32 + 2 * i + 1]));
      ms_out[i] += ms / (float)n;
    }
  *n_forwards = n;
  return VST_OK;
}

int vst_plan_set_stop_after(vst_plan* P, int stage) {
  VST_CHECK_ARG(P, "set_stop_after: NULL plan");
  P->stop_after = stage;
  return VST_OK;
}

int vst_plan_debug_activation(vst_plan* P, int layer, float* out_nchw, size_t out_elems, void* stream) {
  VST_CHECK_ARG(P && out_nchw, "debug_activation: NULL argument");
  VST_CHECK_ARG(layer >= 0 && layer < (int)P->act_bufs.size(), "debug_activation: layer %d out of range", layer);
  VST_CHECK_ARG(!P->stages[layer].skip_apply, "debug_activation: stage %d is normalised inside its consumer (fused), it has no activation buffer", layer);
  const ActLayout& L = P->act_bufs[layer].second;
  const size_t need = (size_t)P->d.N * L.C * L.H * L.W;
  VST_CHECK_ARG(out_elems >= need, "debug_activation: need %zu elements", need);
  vst::launch(act_to_nchw_kernel, ew_grid(need), 256, 0, (cudaStream_t)stream, P->act_bufs[layer].first, L, P->d.N, out_nchw, P->half);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

// ---- stand-alone 3x3 stride-1 convolution through the tensor-core path ----------------------
size_t vst_tc_conv_workspace_bytes(int N, int Cin, int H, int W, int Cout, int k) {
  (void)k;
  const int Cp = round_up(Cin, 8);
  int BK, kbpt;
  choose_bk(Cp, &BK, &kbpt);
  const size_t act = (size_t)N * (H + 2) * (W + 2) * Cp * 2;
  const size_t wpk = (size_t)round_up(Cout, 16) * 9 * kbpt * BK * 2;
  return act + wpk + 4096;
}

int vst_tc_conv3x3_f32io(const float* x_nchw, const float* w, float* y_nchw, int N, int Cin, int H, int W, int Cout,
                         int pad_mode, void* workspace, size_t workspace_bytes, void* stream) {
  VST_CHECK_ARG(N > 0 && Cin > 0 && H >= 2 && W >= 2 && Cout > 0, "tc_conv3x3: bad shape");
  VST_DEVPTR(x_nchw); VST_DEVPTR(w); VST_DEVPTR(y_nchw); VST_DEVPTR(workspace);
  if (workspace_bytes < vst_tc_conv_workspace_bytes(N, Cin, H, W, Cout, 3)) {
    set_error("tc_conv3x3: workspace too small");
    return VST_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = round_up(Cin, 8);
  int BK, kbpt;
  choose_bk(Cp, &BK, &kbpt);
  const ActLayout L{H, W, Cp, 1, pad_mode == VST_PAD_REFLECT ? PADK_REFLECT : PADK_ZERO, 0};
  __nv_bfloat16* act = (__nv_bfloat16*)workspace;
  const size_t act_bytes = ((size_t)N * (H + 2) * (W + 2) * Cp * 2 + 1023) & ~(size_t)1023;
  __nv_bfloat16* wpk = (__nv_bfloat16*)((uint8_t*)workspace + act_bytes);
  vst::launch(nchw_to_act_kernel, ew_grid(act_elems(L, N)), 256, 0, st, x_nchw, Cin, act, L, N);
  VST_LAUNCH_CHECK();
  const int n_mma = Cout > 256 ? 256 : round_up(Cout, 16);
  const int n_ntile = cdiv(Cout, n_mma);
  const int rows = n_mma * n_ntile;
  vst::launch(pack_w_taps_kernel, ew_grid((size_t)rows * 9 * kbpt * BK), 256, 0, st, w, wpk, Cout, Cin, 3, rows, kbpt, BK, 0);
  VST_LAUNCH_CHECK();
  TapGemmParams tg;
  tg_defaults(tg, N);
  tg.kb_per_tap = kbpt;
  tg.MT = choose_mt(n_mma);
  choose_tile(H, W, tg.MT, &tg.TW, &tg.TH);
  tg.tiles_x = cdiv(W, tg.TW); tg.tiles_y = cdiv(H, tg.TH);
  tg.Ho = H; tg.Wo = W; tg.N_mma = n_mma; tg.n_ntile = n_ntile; tg.Cout = Cout;
  tg.Hout = H; tg.Wout = W; tg.out_cstride = Cout; tg.epi_mode = TG_EPI_F32_NCHW; tg.act = VST_ACT_NONE;
  tg.out0 = y_nchw;
  fill_taps_3x3_s1(tg);
  tapgemm_plan(tg, BK);
  int r = tmap_for_act(&tg.tmA, act, L, N, BK, tg.TW, tapgemm_box_rows(tg));
  if (r != VST_OK) return r;
  r = make_tmap_wgt(&tg.tmB, wpk, 9 * kbpt * BK, rows, BK, n_mma);
  if (r != VST_OK) return r;
  return launch_tapgemm(tg, BK, st);
}


// ---- generic entry points for the training primitives (vst_b200/tc.py) ------------------------
static inline ActLayout to_layout(const vst_act_desc& d) { return ActLayout{d.H, d.W, d.C, d.pad, d.kind, d.parity}; }

// descriptor -> kernel parameters + pipeline plan (everything but the tensor maps): shared by the launch and the
// host-only plan query
static int tapgemm_params_from_desc(const vst_tapgemm_desc* d, TapGemmParams& tg) {
  VST_CHECK_ARG(d, "tapgemm: NULL descriptor");
  VST_CHECK_ARG(d->n_taps >= 1 && d->n_phase >= 1 && d->n_phase <= 4 && d->n_phase * d->n_taps <= TG_MAX_TAPS, "tapgemm: tap table size");
  VST_CHECK_ARG(d->BK == 16 || d->BK == 32 || d->BK == 64, "tapgemm: BK must be 16/32/64");
  VST_CHECK_ARG(d->kb_per_tap >= 1 && d->b_K == d->n_taps * d->kb_per_tap * d->BK, "tapgemm: b_K != n_taps*kb_per_tap*BK");
  VST_CHECK_ARG(d->grid_h >= 1 && d->grid_w >= 1 && d->a_N >= 1 && d->a_P >= 1, "tapgemm: empty grid");
  VST_CHECK_ARG(d->a_C % 8 == 0, "tapgemm: operand channels must be a multiple of 8");
  VST_CHECK_ARG(d->N_mma % 16 == 0 && d->N_mma >= 16 && d->N_mma <= 256, "tapgemm: N_mma=%d invalid", d->N_mma);
  tg_defaults(tg, d->a_N);
  tg.kb_per_tap = d->kb_per_tap;
  tg.n_taps = d->n_taps; tg.n_phase = d->n_phase; tg.n_ntile = d->n_ntile > 0 ? d->n_ntile : 1;
  tg.N_mma = d->N_mma;
  tg.MT = d->MT > 0 ? d->MT : choose_mt(d->N_mma);
  if (d->TW > 0 && d->TH > 0) { tg.TW = d->TW; tg.TH = d->TH; }
  else choose_tile(d->grid_h, d->grid_w, tg.MT, &tg.TW, &tg.TH);
  tg.tile_step_x = d->tile_step_x;
  const int step_x = d->tile_step_x > 0 ? d->tile_step_x : tg.TW;
  tg.tiles_x = cdiv(d->grid_w, step_x); tg.tiles_y = cdiv(d->grid_h, tg.TH);
  tg.Ho = d->grid_h; tg.Wo = d->grid_w;
  tg.Cout = d->Cout; tg.out_mul = d->out_mul > 0 ? d->out_mul : 1;
  tg.Hout = d->Hout; tg.Wout = d->Wout; tg.out_cstride = d->out_cstride;
  tg.epi_mode = d->epi_mode; tg.act = d->act; tg.relu = d->relu;
  tg.rc_k = d->rc_k; tg.rc_co = d->rc_co;
  tg.out0 = d->out; tg.out_u8 = d->out_u8; tg.bias = d->bias; tg.stats = d->stats;
  tg.b_img_rows = d->b_img_rows;
  for (int i = 0; i < d->n_phase * d->n_taps; ++i) {
    tg.tap_dx[i] = d->tap_dx[i]; tg.tap_dy[i] = d->tap_dy[i]; tg.tap_pl[i] = d->tap_pl[i];
    VST_CHECK_ARG(d->tap_pl[i] >= 0 && d->tap_pl[i] < d->a_P, "tapgemm: tap plane out of range");
  }
  for (int i = 0; i < 4; ++i) { tg.ph_oy[i] = d->ph_oy[i]; tg.ph_ox[i] = d->ph_ox[i]; }
  tapgemm_plan(tg, d->BK);
  return VST_OK;
}

int vst_tc_tapgemm_plan(const vst_tapgemm_desc* d, vst_tapgemm_plan_info* info) {
  VST_CHECK_ARG(info, "tapgemm_plan: NULL info");
  TapGemmParams tg;
  int r = tapgemm_params_from_desc(d, tg);
  if (r != VST_OK) return r;
  memset(info, 0, sizeof(*info));
  info->stream = tg.stream; info->dyshare = tg.dyshare; info->n_cols = tg.n_cols; info->dy_max = tg.dy_max;
  info->box_rows = tapgemm_box_rows(tg); info->TW = tg.TW; info->TH = tg.TH; info->MT = tg.MT;
  info->tiles_x = tg.tiles_x; info->tiles_y = tg.tiles_y; info->cta2 = tg.cta2;
  if (tg.dyshare)
    for (int i = 0; i < tg.n_phase * tg.n_cols && i < 48; ++i) {
      info->col_dx[i] = tg.col_dx[i]; info->col_dy0[i] = tg.col_dy0[i]; info->col_pl[i] = tg.col_pl[i];
      info->col_n[i] = tg.col_n[i]; info->col_t0[i] = tg.col_t0[i]; info->col_ts[i] = tg.col_ts[i];
    }
  return VST_OK;
}

int vst_tc_tapgemm(const vst_tapgemm_desc* d, void* stream) {
  TapGemmParams tg;
  int r = tapgemm_params_from_desc(d, tg);
  if (r != VST_OK) return r;
  VST_DEVPTR(d->a); VST_DEVPTR(d->b);
  if (d->out) VST_DEVPTR(d->out);
  const size_t img = (size_t)d->a_Y * d->a_X * d->a_C;
  r = make_tmap_act(&tg.tmA, d->a, d->a_C, d->a_X, d->a_Y, d->a_N, d->a_P, d->a_C, (size_t)d->a_X * d->a_C, img,
                    img * d->a_N, d->BK, tg.TW, tapgemm_box_rows(tg));
  if (r != VST_OK) return r;
  r = make_tmap_wgt(&tg.tmB, d->b, d->b_K, d->b_rows, d->BK, tapgemm_b_box_rows(tg));
  if (r != VST_OK) return r;
  return launch_tapgemm(tg, d->BK, (cudaStream_t)stream);
}

int vst_tc_nchw_to_act(const float* x, int Cin, void* dst, vst_act_desc L, int N, void* stream) {
  VST_CHECK_ARG(N > 0 && Cin > 0 && Cin <= L.C && L.C % 8 == 0 && L.H > 0 && L.W > 0, "nchw_to_act: bad shape");
  VST_CHECK_ARG((size_t)N * (L.H + 2 * L.pad) * (L.W + 2 * L.pad) * L.C < ((size_t)1 << 32), "nchw_to_act: tensor too large for 32-bit indexing");
  VST_DEVPTR(x); VST_DEVPTR(dst);
  const ActLayout A = to_layout(L);
  if (A.pad == 0 && !A.parity && A.C >= 8 && A.C % 8 == 0 && A.C <= 256 && A.H <= 65535 && N <= 65535) {
    dim3 grid(cdiv(A.W, TR_PX), A.H, N);
    vst::launch(nchw_to_act_tiled_kernel, grid, 256, (size_t)A.C * TR_PITCH * sizeof(float), (cudaStream_t)stream, x, Cin, (__nv_bfloat16*)dst, A, N);
  } else if (Cin <= 8 && (A.H + 2 * A.pad) <= 65535 && N <= 65535) {
    const int Wp = A.W + 2 * A.pad;
    dim3 grid(cdiv(Wp, 256), A.H + 2 * A.pad, N);
    vst::launch(nchw_to_act_narrow_kernel, grid, 256, 0, (cudaStream_t)stream, x, Cin, (__nv_bfloat16*)dst, A, N);
  } else {
    vst::launch(nchw_to_act_kernel, ew_grid(act_elems(A, N)), 256, 0, (cudaStream_t)stream, x, Cin, (__nv_bfloat16*)dst, A, N);
  }
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_act_to_nchw_first(const void* act, vst_act_desc L, int N, int c_count, float* out, void* stream) {
  VST_CHECK_ARG(N > 0 && L.C >= 8 && L.C % 8 == 0 && L.H > 0 && L.W > 0 && c_count >= 1 && c_count <= 8, "act_to_nchw_first: bad shape");
  VST_CHECK_ARG(L.H <= 65535 && N <= 65535, "act_to_nchw_first: H and N must be <= 65535");
  VST_DEVPTR(act); VST_DEVPTR(out);
  const ActLayout A = to_layout(L);
  vst::launch(act_to_nchw_first_kernel, dim3(cdiv(A.W, 256), A.H, N), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)act, A, N, c_count, out);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_act_to_nchw(const void* act, vst_act_desc L, int N, float* out, void* stream) {
  VST_CHECK_ARG(N > 0 && L.C > 0 && L.H > 0 && L.W > 0 && (size_t)N * L.C * L.H * L.W < ((size_t)1 << 32), "act_to_nchw: bad shape");
  VST_DEVPTR(act); VST_DEVPTR(out);
  const ActLayout A = to_layout(L);
  if (!A.parity && A.C % 8 == 0 && A.C >= 8 && A.C <= 256 && A.H <= 65535 && N <= 65535) {
    dim3 grid(cdiv(A.W, TR_PX), A.H, N);
    vst::launch(act_to_nchw_tiled_kernel, grid, 256, (size_t)A.C * TR_PITCH * sizeof(float), (cudaStream_t)stream, (const __nv_bfloat16*)act, A, N, out);
  } else {
    vst::launch(act_to_nchw_kernel, ew_grid((size_t)N * A.C * A.H * A.W), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)act, A, N, out, 0);
  }
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_prologue_x9(const float* x, void* x9v, int N, int Cin, int H, int W, int KR, void* stream) {
  VST_CHECK_ARG(N > 0 && Cin > 0 && H > 4 && W > 4 && KR >= 9 * Cin, "prologue_x9: bad shape");
  VST_DEVPTR(x); VST_DEVPTR(x9v);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* x9 = (__nv_bfloat16*)x9v;
  dim3 grid(cdiv(W, 256), H + 8, N);
  if (KR == 32 && Cin == 3 && x9_staged() == 2) vst::launch(prologue_x9_seg_kernel<false>, grid, 256, 0, st, (const void*)x, x9, N, H, W, 0, 1.f);
  else if (KR == 32 && Cin == 3) vst::launch(prologue_x9_kernel<32, 3>, grid, 256, 0, st, x, x9, N, Cin, H, W, 0, 1.f, x9_staged());
  else if (KR == 32) vst::launch(prologue_x9_kernel<32, 0>, grid, 256, 0, st, x, x9, N, Cin, H, W, 0, 1.f, 0);
  else if (KR == 64) vst::launch(prologue_x9_kernel<64, 0>, grid, 256, 0, st, x, x9, N, Cin, H, W, 0, 1.f, 0);
  else if (KR == 128) vst::launch(prologue_x9_kernel<128, 0>, grid, 256, 0, st, x, x9, N, Cin, H, W, 0, 1.f, 0);
  else if (KR == 192) vst::launch(prologue_x9_kernel<192, 0>, grid, 256, 0, st, x, x9, N, Cin, H, W, 0, 1.f, 0);
  else if (KR == 256) vst::launch(prologue_x9_kernel<256, 0>, grid, 256, 0, st, x, x9, N, Cin, H, W, 0, 1.f, 0);
  else { set_error("prologue_x9: KR=%d unsupported", KR); return VST_EUNSUPPORTED; }
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_in_apply(const void* raw, const double* stats, const float* gamma, const float* beta, const void* residual,
                    vst_act_desc res_desc, void* dst, vst_act_desc dst_desc, int N, float eps, int relu, void* stream) {
  VST_CHECK_ARG(N > 0 && dst_desc.C % 8 == 0 && dst_desc.C <= 256 * 8, "in_apply: bad shape");
  VST_DEVPTR(raw); VST_DEVPTR(stats); VST_DEVPTR(gamma); VST_DEVPTR(beta); VST_DEVPTR(dst);
  const ActLayout D = to_layout(dst_desc), R = to_layout(res_desc);
  launch_apply((const __nv_bfloat16*)raw, stats, gamma, beta, (const __nv_bfloat16*)residual, R, (__nv_bfloat16*)dst, D, N, eps, relu,
               (cudaStream_t)stream);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

}  // extern "C"
