// Bandwidth-bound helpers of the loss path: flow backward-warp (bilinear grid_sample semantics),
// forward-backward occlusion mask, fused temporal losses, MSE / TV reductions, fp32 Gram.
// All index arithmetic follows the reference's fp32 operation order (see common.cuh).
#include <stdlib.h>
#include "common.cuh"

namespace vst {

struct Bilin {
  int x0, y0;
  float wnw, wne, wsw, wse;
};

// Corner indices + weights for one output pixel, as ATen's grid_sampler_2d computes them
// (TORCH/GridSampler.h:164-172): weights from the opposite corners.
__device__ __forceinline__ Bilin bilin_setup(int px, int py, float fx, float fy, int W, int H) {
  const float ix = warp_src_coord(px, fx, W), iy = warp_src_coord(py, fy, H);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float x1f = __fadd_rn(x0f, 1.f), y1f = __fadd_rn(y0f, 1.f);
  Bilin b;
  b.x0 = (int)x0f;
  b.y0 = (int)y0f;
  const float wx0 = __fsub_rn(x1f, ix), wx1 = __fsub_rn(ix, x0f);
  const float wy0 = __fsub_rn(y1f, iy), wy1 = __fsub_rn(iy, y0f);
  b.wnw = __fmul_rn(wx0, wy0);
  b.wne = __fmul_rn(wx1, wy0);
  b.wsw = __fmul_rn(wx0, wy1);
  b.wse = __fmul_rn(wx1, wy1);
  return b;
}

// zeros padding: an out-of-bounds corner contributes 0 (TORCH/GridSampler.h:205-207)
__device__ __forceinline__ float bilin_sample(const float* __restrict__ p, const Bilin& b, int W, int H) {
  const bool xl = b.x0 >= 0 && b.x0 < W, xr = b.x0 + 1 >= 0 && b.x0 + 1 < W;
  const bool yt = b.y0 >= 0 && b.y0 < H, yb = b.y0 + 1 >= 0 && b.y0 + 1 < H;
  const float* r0 = p + (size_t)b.y0 * W + b.x0;
  float acc = (xl && yt) ? __fmul_rn(r0[0], b.wnw) : 0.f;
  acc = __fadd_rn(acc, (xr && yt) ? __fmul_rn(r0[1], b.wne) : 0.f);
  acc = __fadd_rn(acc, (xl && yb) ? __fmul_rn(r0[W], b.wsw) : 0.f);
  acc = __fadd_rn(acc, (xr && yb) ? __fmul_rn(r0[W + 1], b.wse) : 0.f);
  return acc;
}

// grid (ceil(W/256), H, B): one thread per output pixel, no integer divisions (they, not memory, bounded the first
// version at 0.3 of HBM peak); the flow reads and the C output stores are coalesced rows, the 4C gathers hit L1/L2.
__global__ void __launch_bounds__(256) warp_f32_kernel(const float* __restrict__ x, const float* __restrict__ flo,
                                                       float* __restrict__ out, int32_t* __restrict__ corner,
                                                       int B, int C, int H, int W) {
  vst::pdl_grid_sync();
  const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y, b = blockIdx.z;
  if (px >= W) return;
  const size_t HW = (size_t)H * W, pix = (size_t)py * W + px;
  const float* f = flo + (size_t)b * 2 * HW + pix;
  const Bilin bl = bilin_setup(px, py, __ldg(f), __ldg(f + HW), W, H);
  if (corner) {
    const size_t i = (size_t)b * HW + pix;
    corner[2 * i] = bl.x0;
    corner[2 * i + 1] = bl.y0;
  }
  const float* xp = x + (size_t)b * C * HW;
  float* op = out + (size_t)b * C * HW + pix;
  if (C == 3) {   // the image warps of the temporal losses: all 12 gathers in flight together
    const float v0 = bilin_sample(xp, bl, W, H), v1 = bilin_sample(xp + HW, bl, W, H), v2 = bilin_sample(xp + 2 * HW, bl, W, H);
    op[0] = v0; op[HW] = v1; op[2 * HW] = v2;
    return;
  }
  for (int c = 0; c < C; ++c) op[c * HW] = bilin_sample(xp + c * HW, bl, W, H);
}

// ---- the same warp with half the instructions (round 2, last session) --------------------------------------------------
// ncu on warp_f32_kernel: 25 % of DRAM throughput, L1/TEX busy - and ~280 SASS instructions per output pixel for 32 bytes of
// algorithmic traffic: with smooth (optical-flow-like) fields the kernel is bound by instruction issue, not by memory.  What
// the instructions were: four IEEE divisions per pixel in the reference's coordinate formula, 64-bit index arithmetic, and
// four bounds-predicated loads per channel.  Here:
//  * x / 2 is x * 0.5 (exact, hence the same rounding); the division by the invariant d = max(S - 1, 1) is Markstein's
//    correctly-rounded quotient from the correctly-rounded reciprocal r = RN(1 / d) computed on the host (IEEE single
//    division): q0 = RN(a r), rem = fma(-q0, d, a) (exact), q = RN(q0 + rem r) - the same bits as __fdiv_rn(a, d) for every
//    d whose significand is not all ones (image extents never are);
//  * offsets are 32-bit (H * W < 2^31, else the launcher keeps the first kernel), the four corner offsets are computed once
//    from CLAMPED coordinates and shared by all channels, loads are unconditional and an out-of-bounds corner's product is
//    replaced by 0.f after the multiply - the reference's zeros padding, same operation order, same bits.
__device__ __forceinline__ float div_invariant(float a, float d, float r) {
  const float q0 = __fmul_rn(a, r);
  const float rem = __fmaf_rn(-q0, d, a);
  return __fmaf_rn(rem, r, q0);
}
__device__ __forceinline__ float warp_src_coord_inv(int p, float f, float S, float d, float r) {
  const float v = __fadd_rn((float)p, f);
  const float n = __fsub_rn(div_invariant(__fmul_rn(2.0f, v), d, r), 1.0f);
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(n, 1.0f), S), 1.0f), 0.5f);
}

constexpr int WL_ROWS = 8;   // 32 x 8 tiles a block walks down the image, the next tile's flow in flight under this tile's gathers
template <int CFIX>   // 3: the image warps (all 12 gathers in flight); 0: any channel count, two channels in flight
__global__ void __launch_bounds__(256, 6) warp_f32_lean_kernel(const float* __restrict__ x, const float* __restrict__ flo,
                                                            float* __restrict__ out, int32_t* __restrict__ corner,
                                                            int C, int H, int W, float dW, float rW, float dH, float rH) {
  vst::pdl_grid_sync();
  // a block is a 32 x 8 pixel tile (one warp per row): the source rows y0 / y0 + 1 of consecutive output rows overlap, so
  // 9 source rows serve 8 output rows out of this SM's L1 instead of 2 per row out of L2.  A pixel is a chain of two
  // dependent memory latencies (flow -> gathers), so the block walks WL_ROWS tiles down the image with the NEXT tile's two
  // flow values already in flight: one latency per pixel instead of two at the same occupancy.
  const int px = blockIdx.x * 32 + (threadIdx.x & 31), b = blockIdx.z;
  if (px >= W) return;
  int py_n = blockIdx.y * (8 * WL_ROWS) + (threadIdx.x >> 5);
  const uint32_t HW = (uint32_t)H * (uint32_t)W;
  const float* f = flo + (size_t)b * 2 * HW;
  const float* xb = x + (size_t)b * C * HW;
  float* ob = out + (size_t)b * C * HW;
  float fx_n = 0.f, fy_n = 0.f;
  if (py_n < H) {
    const uint32_t pn = (uint32_t)py_n * (uint32_t)W + (uint32_t)px;
    fx_n = __ldg(f + pn);
    fy_n = __ldg(f + HW + pn);
  }
#pragma unroll 1
  for (int k = 0; k < WL_ROWS; ++k) {
    const int py = py_n;
    if (py >= H) break;
    const float fx = fx_n, fy = fy_n;
    py_n += 8;
    if (k + 1 < WL_ROWS && py_n < H) {
      const uint32_t pn = (uint32_t)py_n * (uint32_t)W + (uint32_t)px;
      fx_n = __ldg(f + pn);
      fy_n = __ldg(f + HW + pn);
    }
    const uint32_t pix = (uint32_t)py * (uint32_t)W + (uint32_t)px;
    const float ix = warp_src_coord_inv(px, fx, (float)W, dW, rW);
    const float iy = warp_src_coord_inv(py, fy, (float)H, dH, rH);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float x1f = __fadd_rn(x0f, 1.f), y1f = __fadd_rn(y0f, 1.f);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float wx0 = __fsub_rn(x1f, ix), wx1 = __fsub_rn(ix, x0f);
    const float wy0 = __fsub_rn(y1f, iy), wy1 = __fsub_rn(iy, y0f);
    const float wnw = __fmul_rn(wx0, wy0), wne = __fmul_rn(wx1, wy0), wsw = __fmul_rn(wx0, wy1), wse = __fmul_rn(wx1, wy1);
    if (corner) {
      int2* cp = reinterpret_cast<int2*>(corner) + (size_t)b * HW + pix;
      *cp = make_int2(x0, y0);
    }
    const bool xl = (unsigned)x0 < (unsigned)W, xr = (unsigned)(x0 + 1) < (unsigned)W;
    const bool yt = (unsigned)y0 < (unsigned)H, yb = (unsigned)(y0 + 1) < (unsigned)H;
    const bool vnw = xl && yt, vne = xr && yt, vsw = xl && yb, vse = xr && yb;
    const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x0 + 1, 0), W - 1);
    const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y0 + 1, 0), H - 1);
    const uint32_t r0 = (uint32_t)cy0 * (uint32_t)W, r1 = (uint32_t)cy1 * (uint32_t)W;
    const uint32_t onw = r0 + cx0, one = r0 + cx1, osw = r1 + cx0, ose = r1 + cx1;
    const float* xp = xb;
    float* op = ob + pix;
    auto sample = [&](const float* __restrict__ p) {
      const float a = __ldg(p + onw), bq = __ldg(p + one), c = __ldg(p + osw), d = __ldg(p + ose);
      float acc = vnw ? __fmul_rn(a, wnw) : 0.f;
      acc = __fadd_rn(acc, vne ? __fmul_rn(bq, wne) : 0.f);
      acc = __fadd_rn(acc, vsw ? __fmul_rn(c, wsw) : 0.f);
      acc = __fadd_rn(acc, vse ? __fmul_rn(d, wse) : 0.f);
      return acc;
    };
    if constexpr (CFIX == 3) {   // the image warps of the temporal losses: all 12 gathers in flight together
      const float v0 = sample(xp), v1 = sample(xp + HW), v2 = sample(xp + 2 * (size_t)HW);
      op[0] = v0; op[HW] = v1; op[2 * (size_t)HW] = v2;
    } else {
#pragma unroll 2
      for (int c = 0; c < C; ++c, xp += HW, op += HW) *op = sample(xp);
    }
  }
}

// Shared-memory-staged variant (round 2; north_star (3) "shared-memory staging"; OPT-IN, see vst_warp_f32 for the measurement): a block owns a 64 x 32 output tile and
// stages, per channel, the (64 + 2*16 + 4) x (32 + 2*16 + 1) window of x around it with coalesced row loads (zeros outside
// the image, which IS the zeros-padding rule of grid_sample); the four corner reads of every pixel whose displacement stays
// within +-16 px then come from shared memory (a random gather costs ~3 bank wavefronts there against 32 L1 wavefronts for
// 32 lanes on 32 different lines - ncu on the direct kernel: L1/TEX 83 % busy at 25 % of DRAM throughput).  Pixels that
// reach further fall back to the global gather.  Same arithmetic, same operation order: values and corner indices are
// bit-identical to the direct kernel.
constexpr int WT_X = 64, WT_Y = 32, WT_HALO = 16, WT_RW = WT_X + 2 * WT_HALO + 4, WT_RH = WT_Y + 2 * WT_HALO + 1;
__global__ void __launch_bounds__(256) warp_f32_tiled_kernel(const float* __restrict__ x, const float* __restrict__ flo,
                                                             float* __restrict__ out, int32_t* __restrict__ corner,
                                                             int B, int C, int H, int W) {
  vst::pdl_grid_sync();
  __shared__ float reg[WT_RH * WT_RW];
  const int b = blockIdx.z, tx0 = blockIdx.x * WT_X, ty0 = blockIdx.y * WT_Y;
  const int lx = threadIdx.x & 63, ly = threadIdx.x >> 6;      // pixel column, first row (rows ly, ly+4, ...)
  const int px = tx0 + lx;
  const size_t HW = (size_t)H * W;
  const int rx0 = tx0 - WT_HALO, ry0 = ty0 - WT_HALO;           // frame coordinates of reg[0][0]
  Bilin bl[WT_Y / 4];
  bool inreg[WT_Y / 4];
#pragma unroll
  for (int k = 0; k < WT_Y / 4; ++k) {
    const int py = ty0 + ly + 4 * k;
    inreg[k] = false;
    if (px < W && py < H) {
      const float* f = flo + (size_t)b * 2 * HW + (size_t)py * W + px;
      bl[k] = bilin_setup(px, py, __ldg(f), __ldg(f + HW), W, H);
      if (corner) {
        const size_t i = (size_t)b * HW + (size_t)py * W + px;
        corner[2 * i] = bl[k].x0;
        corner[2 * i + 1] = bl[k].y0;
      }
      inreg[k] = bl[k].x0 >= rx0 && bl[k].x0 + 1 < rx0 + WT_RW && bl[k].y0 >= ry0 && bl[k].y0 + 1 < ry0 + WT_RH;
    }
  }
  for (int c = 0; c < C; ++c) {
    const float* xp = x + ((size_t)b * C + c) * HW;
    __syncthreads();                                            // the previous channel's gathers are done
    {   // warp w stages rows w, w+8, ...: lanes walk a row (coalesced 128-byte segments), no integer divisions
      const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 3
      for (int r = wid; r < WT_RH; r += 8) {
        const int gy = ry0 + r;
        const bool oky = gy >= 0 && gy < H;
        const float* src = xp + (size_t)(oky ? gy : 0) * W;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = lane + 32 * j, gx = rx0 + q;
          v[j] = (oky && q < WT_RW && gx >= 0 && gx < W) ? __ldg(src + gx) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (lane + 32 * j < WT_RW) reg[r * WT_RW + lane + 32 * j] = v[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < WT_Y / 4; ++k) {
      const int py = ty0 + ly + 4 * k;
      if (px >= W || py >= H) continue;
      float v;
      if (inreg[k]) {
        const float* r0 = reg + (bl[k].y0 - ry0) * WT_RW + (bl[k].x0 - rx0);
        v = __fmul_rn(r0[0], bl[k].wnw);
        v = __fadd_rn(v, __fmul_rn(r0[1], bl[k].wne));
        v = __fadd_rn(v, __fmul_rn(r0[WT_RW], bl[k].wsw));
        v = __fadd_rn(v, __fmul_rn(r0[WT_RW + 1], bl[k].wse));
      } else {
        v = bilin_sample(xp, bl[k], W, H);
      }
      out[((size_t)b * C + c) * HW + (size_t)py * W + px] = v;
    }
  }
}

// mask = |warp(grid + f01, f10) - grid|_1 < thr.  The warped field is formed in fp32 first
// (`flo01 = grid + flo01`, RC/utilities.py:72) and then blended, as the reference does.
__global__ void __launch_bounds__(256) flow_warp_mask_kernel(const float* __restrict__ f01, const float* __restrict__ f10,
                                                             float* __restrict__ mask, int B, int H, int W, float thr) {
  vst::pdl_grid_sync();
  const size_t HW = (size_t)H * W, total = (size_t)B * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;   // B*H*W < 2^32 (checked by the launcher)
    const int px = iu % W, py = (iu / W) % H, b = iu / (unsigned)HW;
    const float* fb = f10 + (size_t)b * 2 * HW + (size_t)py * W + px;
    const Bilin bl = bilin_setup(px, py, fb[0], fb[HW], W, H);
    const float* fu = f01 + (size_t)b * 2 * HW;
    const float* fv = fu + HW;
    float wx = 0.f, wy = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cx = bl.x0 + (k & 1), cy = bl.y0 + (k >> 1);
      const float wgt = k == 0 ? bl.wnw : k == 1 ? bl.wne : k == 2 ? bl.wsw : bl.wse;
      float vx = 0.f, vy = 0.f;
      if (cx >= 0 && cx < W && cy >= 0 && cy < H) {
        const size_t o = (size_t)cy * W + cx;
        vx = __fmul_rn(__fadd_rn((float)cx, fu[o]), wgt);
        vy = __fmul_rn(__fadd_rn((float)cy, fv[o]), wgt);
      }
      wx = __fadd_rn(wx, vx);
      wy = __fadd_rn(wy, vy);
    }
    const float err = __fadd_rn(fabsf(__fsub_rn(wx, (float)px)), fabsf(__fsub_rn(wy, (float)py)));
    mask[i] = err < thr ? 1.f : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// Deterministic grid reduction: every block stores its partial sums, the last block to finish
// adds them in block order and writes the result.  scratch = [gridDim*NV partials | counter].
// The counter is left at zero, so a zero-initialised scratch can be reused across launches
// on one stream.
// ------------------------------------------------------------------------------------------
constexpr int kRedMaxBlocks = kNumSMs * 8;
constexpr int kRedMaxVals = 4;

template <int NV>
__device__ void grid_reduce_finish(float (&v)[NV], float* __restrict__ out, float* __restrict__ scratch) {
  __shared__ float red[32];
  __shared__ bool is_last;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = block_sum(v[k], red);
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + (size_t)kRedMaxBlocks * kRedMaxVals);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) scratch[(size_t)blockIdx.x * NV + k] = v[k];
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.0;
  // fixed order: thread t sums blocks t, t+T, ...; then a fixed-shape tree
  __shared__ double dred[NV][256];
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] += (double)__ldcg(&scratch[(size_t)b * NV + k]);
#pragma unroll
  for (int k = 0; k < NV; ++k) dred[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s)
#pragma unroll
      for (int k = 0; k < NV; ++k) dred[k][threadIdx.x] += dred[k][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = (float)dred[k][0];
    *counter = 0u;
  }
}

// ATen upsample_bilinear2d (align_corners=False) source index: max(scale*(dst+0.5)-0.5, 0)
// (TORCH/UpSample.h:289-315); returns i0, i1, lambda0, lambda1.
__device__ __forceinline__ void resize_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = __fsub_rn(s, (float)i0);
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  l0 = __fsub_rn(1.f, l1);
}

__device__ __forceinline__ float resize_sample(const float* __restrict__ p, int W, int y0, int y1, int x0, int x1,
                                               float ly0, float ly1, float lx0, float lx1) {
  const float top = __fadd_rn(__fmul_rn(lx0, p[(size_t)y0 * W + x0]), __fmul_rn(lx1, p[(size_t)y0 * W + x1]));
  const float bot = __fadd_rn(__fmul_rn(lx0, p[(size_t)y1 * W + x0]), __fmul_rn(lx1, p[(size_t)y1 * W + x1]));
  return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

// F.interpolate(mode="bilinear", align_corners=False) of BC planes, optional per-channel rescale (the SceneFlow flow resize)
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int BC, int Hs,
                                                              int Ws, int Hd, int Wd, const float* __restrict__ chan_scale, int C) {
  vst::pdl_grid_sync();
  const size_t total = (size_t)BC * Hd * Wd;
  const float sh = (float)Hs / (float)Hd, sw = (float)Ws / (float)Wd;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;
    const int px = iu % Wd, py = (iu / Wd) % Hd, bc = iu / ((unsigned)Wd * Hd);
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    resize_src(py, sh, Hs, y0, y1, ly0, ly1);
    resize_src(px, sw, Ws, x0, x1, lx0, lx1);
    float v = resize_sample(src + (size_t)bc * Hs * Ws, Ws, y0, y1, x0, x1, ly0, ly1, lx0, lx1);
    if (chan_scale) v = __fmul_rn(v, chan_scale[bc % C]);
    dst[i] = v;
  }
}
__global__ void __launch_bounds__(256) motion_mask_kernel(float* __restrict__ mask, const float* __restrict__ motion, size_t n) {
  vst::pdl_grid_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    mask[i] = motion[i] != 0.f ? 0.f : mask[i];
}

__global__ void __launch_bounds__(256) feature_temporal_kernel(
    const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ flow,
    const float* __restrict__ mask, float* __restrict__ out, float* __restrict__ scratch, int B, int C, int Hf,
    int Wf, int H, int W) {
  vst::pdl_grid_sync();
  const size_t HWf = (size_t)Hf * Wf, HW = (size_t)H * W, total = (size_t)B * HWf;
  const float sh = (float)H / (float)Hf, sw = (float)W / (float)Wf;
  const float mu = (float)((double)Wf / (double)W), mv = (float)((double)Hf / (double)H);
  float v[2] = {0.f, 0.f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;   // B*Hf*Wf < 2^32 (checked by the launcher)
    const int px = iu % Wf, py = (iu / Wf) % Hf, b = iu / (unsigned)HWf;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    resize_src(py, sh, H, y0, y1, ly0, ly1);
    resize_src(px, sw, W, x0, x1, lx0, lx1);
    const float m = resize_sample(mask + (size_t)b * HW, W, y0, y1, x0, x1, ly0, ly1, lx0, lx1) > 0.f ? 1.f : 0.f;
    v[1] += m;
    if (m == 0.f) continue;
    const float* fl = flow + (size_t)b * 2 * HW;
    const float u = __fmul_rn(resize_sample(fl, W, y0, y1, x0, x1, ly0, ly1, lx0, lx1), mu);
    const float w = __fmul_rn(resize_sample(fl + HW, W, y0, y1, x0, x1, ly0, ly1, lx0, lx1), mv);
    const Bilin bl = bilin_setup(px, py, u, w, Wf, Hf);
    const float* p1 = f1 + (size_t)b * C * HWf;
    const float* p2 = f2 + (size_t)b * C * HWf + (size_t)py * Wf + px;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) {
      const float d = p2[c * HWf] - bilin_sample(p1 + c * HWf, bl, Wf, Hf);
      acc = fmaf(d, d, acc);
    }
    v[0] += acc;
  }
  v[1] *= (float)C;
  grid_reduce_finish<2>(v, out, scratch);
}

__global__ void __launch_bounds__(256) output_temporal_kernel(
    const float* __restrict__ s1, const float* __restrict__ s2, const float* __restrict__ i1,
    const float* __restrict__ i2, const float* __restrict__ flow, const float* __restrict__ mask,
    float* __restrict__ out, float* __restrict__ scratch, int B, int H, int W, int luminance) {
  vst::pdl_grid_sync();
  const size_t HW = (size_t)H * W, total = (size_t)B * HW;
  float v[2] = {0.f, 0.f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;   // B*H*W < 2^32 (checked by the launcher)
    const int px = iu % W, py = (iu / W) % H, b = iu / (unsigned)HW;
    const size_t pix = (size_t)py * W + px;
    const float m = mask[(size_t)b * HW + pix];
    v[1] += m * 3.f;
    if (m == 0.f) continue;
    const float* f = flow + (size_t)b * 2 * HW + pix;
    const Bilin bl = bilin_setup(px, py, f[0], f[HW], W, H);
    const size_t base = (size_t)b * 3 * HW;
    float o[3], lum = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = s2[base + c * HW + pix] - bilin_sample(s1 + base + c * HW, bl, W, H);
    if (luminance) {
      const float k[3] = {0.2126f, 0.7152f, 0.0722f};
#pragma unroll
      for (int c = 0; c < 3; ++c)
        lum += k[c] * (i2[base + c * HW + pix] - bilin_sample(i1 + base + c * HW, bl, W, H));
    }
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = o[c] - lum;
      acc = fmaf(d, d, acc);
    }
    v[0] += m * acc;
  }
  grid_reduce_finish<2>(v, out, scratch);
}

__global__ void __launch_bounds__(256) sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                         float* __restrict__ out, float* __restrict__ scratch, size_t n) {
  vst::pdl_grid_sync();
  float v[1] = {0.f};
  const size_t n4 = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) ? 0 : n / 4;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 p = a4[i], q = b4[i];
    const float d0 = p.x - q.x, d1 = p.y - q.y, d2 = p.z - q.z, d3 = p.w - q.w;
    v[0] += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (size_t i = n4 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    v[0] += d * d;
  }
  grid_reduce_finish<1>(v, out, scratch);
}

__global__ void __launch_bounds__(256) frame_diff_sqsum_kernel(const float* __restrict__ x0, const float* __restrict__ x1,
                                                               const float* __restrict__ y0, const float* __restrict__ y1, float lo,
                                                               float hi, float* __restrict__ out, float* __restrict__ scratch, size_t n) {
  vst::pdl_grid_sync();
  float v[1] = {0.f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float dx = x1[i] - x0[i];
    const float dy = fminf(fmaxf(y1[i], lo), hi) - fminf(fmaxf(y0[i], lo), hi);
    const float d = dx - dy;
    v[0] = fmaf(d, d, v[0]);
  }
  grid_reduce_finish<1>(v, out, scratch);
}

__global__ void __launch_bounds__(256) tv_kernel(const float* __restrict__ x, float* __restrict__ out,
                                                 float* __restrict__ scratch, int BC, int H, int W, int mode) {
  vst::pdl_grid_sync();
  const int Hm = H - 1, Wm = W - 1;
  const size_t total = (size_t)BC * Hm * Wm;
  float v[1] = {0.f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;   // BC*(H-1)*(W-1) < 2^32 (checked by the launcher)
    const int px = iu % Wm, py = (iu / Wm) % Hm;
    const size_t p = iu / ((unsigned)Wm * Hm);
    const float* c = x + (p * H + py) * W + px;
    const float dx = c[1] - c[0], dy = c[W] - c[0];
    const float s = dx * dx + dy * dy;
    v[0] += mode == 0 ? s : sqrtf(fmaxf(s, 1e-8f));
  }
  grid_reduce_finish<1>(v, out, scratch);
}

// fp32 Gram: G[b] += scale * F_tile F_tile^T over a slice of HW (split-K, fp32 atomics).
constexpr int GR_T = 64, GR_K = 16;
__global__ void __launch_bounds__(256) gram_f32_kernel(const float* __restrict__ y, float* __restrict__ out, int C,
                                                       int HW, int chunk, float scale) {
  vst::pdl_grid_sync();
  __shared__ float sa[GR_K][GR_T + 4], sb[GR_K][GR_T + 4];
  const int b = blockIdx.z, ti = blockIdx.y * GR_T, tj = blockIdx.x * GR_T;
  // blockIdx.z encodes (batch, split)
  const int splits = cdiv(HW, chunk);
  const int bb = b / splits, sp = b % splits;
  const int k_begin = sp * chunk, k_end = min(HW, k_begin + chunk);
  const float* F = y + (size_t)bb * C * HW;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += GR_K) {
    for (int idx = threadIdx.x; idx < GR_T * GR_K; idx += 256) {
      const int kk = idx % GR_K, r = idx / GR_K;
      const int k = k0 + kk;
      sa[kk][r] = (ti + r < C && k < k_end) ? F[(size_t)(ti + r) * HW + k] : 0.f;
      sb[kk][r] = (tj + r < C && k < k_end) ? F[(size_t)(tj + r) * HW + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GR_K; ++kk) {
      float a[4], bq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sa[kk][ty * 4 + i], bq[i] = sb[kk][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bq[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = ti + ty * 4 + i, c = tj + tx * 4 + j;
      if (r < C && c < C) atomicAdd(&out[((size_t)bb * C + r) * C + c], acc[i][j] * scale);
    }
}

static inline int red_grid(size_t total) {
  size_t g = (total + 255) / 256;
  return (int)(g < (size_t)kRedMaxBlocks ? (g ? g : 1) : (size_t)kRedMaxBlocks);
}

}  // namespace vst

using namespace vst;

extern "C" {

size_t vst_reduce_scratch_floats(void) { return (size_t)kRedMaxBlocks * kRedMaxVals + 4; }

int vst_warp_f32(const float* x, const float* flo, float* out, int32_t* corner_out, int B, int C, int H, int W,
                 void* stream) {
  VST_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "warp: empty shape");
  VST_DEVPTR(x); VST_DEVPTR(flo); VST_DEVPTR(out);
  VST_CHECK_ARG(H <= 65535 && B <= 65535, "warp: H and B must be <= 65535");
  // measured (tools/warp_probe.py, 2048^2 x 4): tiled 0.89 TB/s for white AND smooth flow against 2.3 / 3.1 TB/s of the direct
  // gather - three load / sync / gather phases per tile at 3 blocks per SM cost more than the L1 wavefronts they save - so the
  // tiled kernel stays opt-in (VST_WARP_TILED=1); results are bit-identical either way
  static const bool tiled = [] { const char* e = getenv("VST_WARP_TILED"); return e && atoi(e) != 0; }();
  if (tiled && W >= WT_X && H >= WT_Y && cdiv(H, WT_Y) <= 65535) {
    vst::launch(warp_f32_tiled_kernel, dim3(cdiv(W, WT_X), cdiv(H, WT_Y), B), 256, 0, (cudaStream_t)stream, x, flo, out, corner_out, B, C, H, W);
    VST_LAUNCH_CHECK();
    return VST_OK;
  }
  const int threads = W >= 256 ? 256 : (W >= 128 ? 128 : 64);
  static const bool lean = [] { const char* e = getenv("VST_WARP_LEAN"); return e ? atoi(e) != 0 : true; }();
  if (lean && (size_t)H * W < ((size_t)1 << 31) && (reinterpret_cast<uintptr_t>(corner_out) & 7) == 0) {
    const float dW = (float)(W > 1 ? W - 1 : 1), dH = (float)(H > 1 ? H - 1 : 1);
    const float rW = 1.0f / dW, rH = 1.0f / dH;   // IEEE single division on the host: RN(1 / d), what div_invariant needs
    const dim3 grid(cdiv(W, 32), cdiv(H, 8 * WL_ROWS), B);
    if (C == 3) vst::launch(warp_f32_lean_kernel<3>, grid, 256, 0, (cudaStream_t)stream, x, flo, out, corner_out, C, H, W, dW, rW, dH, rH);
    else vst::launch(warp_f32_lean_kernel<0>, grid, 256, 0, (cudaStream_t)stream, x, flo, out, corner_out, C, H, W, dW, rW, dH, rH);
    VST_LAUNCH_CHECK();
    return VST_OK;
  }
  vst::launch(warp_f32_kernel, dim3(cdiv(W, threads), H, B), threads, 0, (cudaStream_t)stream, x, flo, out, corner_out, B, C, H, W);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_flow_warp_mask_f32(const float* flo01, const float* flo10, float* mask, int B, int H, int W, float threshold,
                           void* stream) {
  VST_CHECK_ARG(B > 0 && H > 0 && W > 0 && (size_t)B * H * W < ((size_t)1 << 32), "flow_warp_mask: empty shape");
  VST_DEVPTR(flo01); VST_DEVPTR(flo10); VST_DEVPTR(mask);
  vst::launch(flow_warp_mask_kernel, red_grid((size_t)B * H * W), 256, 0, (cudaStream_t)stream, flo01, flo10, mask, B, H, W, threshold);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_resize_bilinear_f32(const float* src, float* dst, int BC, int Hs, int Ws, int Hd, int Wd, const float* chan_scale, int C,
                            void* stream) {
  VST_CHECK_ARG(BC > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && (size_t)BC * Hd * Wd < ((size_t)1 << 32), "resize_bilinear: bad shape");
  VST_CHECK_ARG(!chan_scale || C > 0, "resize_bilinear: chan_scale needs C > 0");
  VST_DEVPTR(src); VST_DEVPTR(dst);
  if (chan_scale) VST_DEVPTR(chan_scale);
  vst::launch(resize_bilinear_kernel, red_grid((size_t)BC * Hd * Wd), 256, 0, (cudaStream_t)stream, src, dst, BC, Hs, Ws, Hd, Wd, chan_scale, C);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_motion_mask_f32(float* mask, const float* motion, size_t n, void* stream) {
  VST_CHECK_ARG(n > 0, "motion_mask: empty");
  VST_DEVPTR(mask); VST_DEVPTR(motion);
  vst::launch(motion_mask_kernel, red_grid(n), 256, 0, (cudaStream_t)stream, mask, motion, n);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_gram_f32(const float* y, float* out, int B, int C, int HW, float scale, void* stream) {
  VST_CHECK_ARG(B > 0 && C > 0 && HW > 0, "gram: empty shape");
  VST_DEVPTR(y); VST_DEVPTR(out);
  cudaStream_t st = (cudaStream_t)stream;
  VST_CUDA(cudaMemsetAsync(out, 0, (size_t)B * C * C * sizeof(float), st));
  const int tiles = cdiv(C, GR_T);
  // enough split-K slices to fill the machine, each a multiple of GR_K
  int splits = cdiv(kNumSMs * 2, tiles * tiles * B);
  int chunk = cdiv(cdiv(HW, splits), GR_K) * GR_K;
  if (chunk < GR_K * 8) chunk = GR_K * 8;
  splits = cdiv(HW, chunk);
  dim3 grid(tiles, tiles, B * splits);
  vst::launch(gram_f32_kernel, grid, 256, 0, st, y, out, C, HW, chunk, scale);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_feature_temporal_f32(const float* f1, const float* f2, const float* flow, const float* mask, float* out,
                             float* scratch, int B, int C, int Hf, int Wf, int H, int W, void* stream) {
  VST_CHECK_ARG(B > 0 && C > 0 && Hf > 0 && Wf > 0 && H > 0 && W > 0 && (size_t)B * H * W < ((size_t)1 << 32), "feature_temporal: empty shape");
  VST_DEVPTR(f1); VST_DEVPTR(f2); VST_DEVPTR(flow); VST_DEVPTR(mask); VST_DEVPTR(out); VST_DEVPTR(scratch);
  vst::launch(feature_temporal_kernel, red_grid((size_t)B * Hf * Wf), 256, 0, (cudaStream_t)stream, f1, f2, flow, mask, out, scratch,
                                                                                       B, C, Hf, Wf, H, W);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_output_temporal_f32(const float* s1, const float* s2, const float* i1, const float* i2, const float* flow,
                            const float* mask, float* out, float* scratch, int B, int H, int W, int luminance,
                            void* stream) {
  VST_CHECK_ARG(B > 0 && H > 0 && W > 0 && (size_t)B * H * W < ((size_t)1 << 32), "output_temporal: empty shape");
  VST_DEVPTR(s1); VST_DEVPTR(s2); VST_DEVPTR(flow); VST_DEVPTR(mask); VST_DEVPTR(out); VST_DEVPTR(scratch);
  if (luminance) { VST_DEVPTR(i1); VST_DEVPTR(i2); }
  vst::launch(output_temporal_kernel, red_grid((size_t)B * H * W), 256, 0, (cudaStream_t)stream, s1, s2, i1, i2, flow, mask, out,
                                                                                    scratch, B, H, W, luminance);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_sqdiff_sum_f32(const float* a, const float* b, float* out, float* scratch, size_t n, void* stream) {
  VST_CHECK_ARG(n > 0, "sqdiff_sum: empty");
  VST_DEVPTR(a); VST_DEVPTR(b); VST_DEVPTR(out); VST_DEVPTR(scratch);
  vst::launch(sqdiff_sum_kernel, red_grid(n / 4 + 1), 256, 0, (cudaStream_t)stream, a, b, out, scratch, n);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_frame_diff_sqsum_f32(const float* x0, const float* x1, const float* y0, const float* y1, float lo, float hi, float* out,
                             float* scratch, size_t n, void* stream) {
  VST_CHECK_ARG(n > 0 && lo <= hi, "frame_diff_sqsum: bad arguments");
  VST_DEVPTR(x0); VST_DEVPTR(x1); VST_DEVPTR(y0); VST_DEVPTR(y1); VST_DEVPTR(out); VST_DEVPTR(scratch);
  vst::launch(frame_diff_sqsum_kernel, red_grid(n), 256, 0, (cudaStream_t)stream, x0, x1, y0, y1, lo, hi, out, scratch, n);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tv_f32(const float* x, float* out, float* scratch, int BC, int H, int W, int mode, void* stream) {
  VST_CHECK_ARG((size_t)BC * H * W < ((size_t)1 << 32), "tv: tensor too large for 32-bit indexing");
  VST_CHECK_ARG(BC > 0 && H > 1 && W > 1, "tv: bad shape");
  VST_DEVPTR(x); VST_DEVPTR(out); VST_DEVPTR(scratch);
  vst::launch(tv_kernel, red_grid((size_t)BC * (H - 1) * (W - 1)), 256, 0, (cudaStream_t)stream, x, out, scratch, BC, H, W, mode);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

}  // extern "C"
