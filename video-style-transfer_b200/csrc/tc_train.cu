// Element-wise / reduction kernels of the bf16 training step on channels-last buffers (HBM-bound):
// InstanceNorm + ReLU backward with the padding fold of the incoming data gradient fused in, the VGG
// body's ReLU / max-pool adjoints, the content / style loss helpers, and the table-driven gather used
// for weight packing and weight-gradient unpacking.  Thread mapping everywhere: one thread owns an
// 8-channel group (one 16-byte bf16 vector) of a pixel; consecutive threads walk consecutive channel
// groups, so every global access is a full coalesced line.
#include "tc_layout.cuh"

namespace vst {

static inline int tt_grid(size_t total, int block = 256) {
  size_t g = (total + block - 1) / block;
  const size_t cap = (size_t)kNumSMs * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

struct F8 {
  float v[8];
};
__device__ __forceinline__ F8 ld8(const __nv_bfloat16* p) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
  F8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    r.v[2 * j] = f.x;
    r.v[2 * j + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const F8& a) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(a.v[2 * j], a.v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}

// ---- table-driven gather-sum ----------------------------------------------------------------------
__global__ void gather_sum_kernel(const float* __restrict__ src, const int* __restrict__ idx, int terms, void* __restrict__ dst,
                                  size_t n, int dst_bf16) {
  vst::pdl_grid_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < terms; ++j) {
      const int k = idx[i * terms + j];
      if (k >= 0) a += src[k];
    }
    if (dst_bf16) reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(a);
    else reinterpret_cast<float*>(dst)[i] = a;
  }
}

// ---- in-place fold of the halo of a padded data gradient onto its interior (rows, then columns) ------
// After both passes the interior [p, p+H) x [p, p+W) of G holds the gradient w.r.t. the unpadded tensor
// (aten::reflection_pad2d_backward / replication_pad2d_backward); the hot kernels below then stream it linearly.
__device__ __forceinline__ int fold_dst(int halo, int n, int p, int kind) {   // halo index in [0,p) or [n+p, n+2p) -> interior padded index
  if (kind == PADK_REFLECT) return halo < p ? 2 * p - halo : 2 * (n - 1 + p) - halo;
  return halo < p ? p : p + n - 1;
}
__global__ void __launch_bounds__(256) fold_rows_kernel(__nv_bfloat16* __restrict__ G, ActLayout L, int N) {
  vst::pdl_grid_sync();
  const int p = L.pad, Hp = L.H + 2 * p, Wp = L.W + 2 * p, groups = L.C >> 3;
  const size_t total = (size_t)N * Wp * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = i % groups, xp = (i / groups) % Wp, n = i / ((size_t)groups * Wp);
    __nv_bfloat16* base = G + ((size_t)n * Hp * Wp + xp) * L.C + g * 8;
    for (int k = 0; k < 2 * p; ++k) {
      const int halo = k < p ? k : L.H + k;                 // top rows 0..p-1, bottom rows H+p..H+2p-1
      const int dst = fold_dst(halo, L.H, p, L.kind);
      F8 d = ld8(base + (size_t)dst * Wp * L.C);
      const F8 sv = ld8(base + (size_t)halo * Wp * L.C);
#pragma unroll
      for (int j = 0; j < 8; ++j) d.v[j] += sv.v[j];
      st8(base + (size_t)dst * Wp * L.C, d);
    }
  }
}
__global__ void __launch_bounds__(256) fold_cols_kernel(__nv_bfloat16* __restrict__ G, ActLayout L, int N) {
  vst::pdl_grid_sync();
  const int p = L.pad, Hp = L.H + 2 * p, Wp = L.W + 2 * p, groups = L.C >> 3;
  const size_t total = (size_t)N * L.H * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = i % groups, y = (i / groups) % L.H, n = i / ((size_t)groups * L.H);
    __nv_bfloat16* base = G + ((size_t)n * Hp + y + p) * Wp * L.C + g * 8;
    for (int k = 0; k < 2 * p; ++k) {
      const int halo = k < p ? k : L.W + k;
      const int dst = fold_dst(halo, L.W, p, L.kind);
      F8 d = ld8(base + (size_t)dst * L.C);
      const F8 sv = ld8(base + (size_t)halo * L.C);
#pragma unroll
      for (int j = 0; j < 8; ++j) d.v[j] += sv.v[j];
      st8(base + (size_t)dst * L.C, d);
    }
  }
}

// ---- InstanceNorm (+ReLU) backward ------------------------------------------------------------------
// grid (row bands, N); thread -> fixed 8-channel group (its per-channel constants live in registers) and a pixel
// lane; two pixels per iteration with all loads issued before use.  G is read at interior positions only (the halo
// has been folded in place).
__device__ __forceinline__ F8 cvt8(const uint4& q) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
  F8 r;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(h[j]);
    r.v[2 * j] = f.x;
    r.v[2 * j + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ uint4 ldq(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <bool APPLY, int PX, int MINB>
__global__ void __launch_bounds__(256, MINB) in_bwd_kernel(const __nv_bfloat16* __restrict__ G, ActLayout GL,
                                                        const __nv_bfloat16* __restrict__ skip, const __nv_bfloat16* __restrict__ raw,
                                                        const double* __restrict__ stats, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ red,
                                                        __nv_bfloat16* __restrict__ draw, ActLayout DL,
                                                        __nv_bfloat16* __restrict__ gsum, int N, float eps, int relu,
                                                        int rows_per_block) {
  vst::pdl_grid_sync();
  // Per-channel constants live in shared memory as float4 rows (two LDS.128 per array and vector): keeping them in
  // registers cost 128 registers per thread = 25 % occupancy, and the kernel is bound by loads in flight (ncu: 3.9 warps
  // per scheduler, 55 % of the stall cycles on the L1TEX scoreboard).  A 4-channel-per-thread form at 5-6 blocks/SM was
  // measured slower (12.4-12.6 vs 11.9 ms per training step): twice the load instructions for the same bytes.
  extern __shared__ __align__(16) float sh[];  // cA[C], cB[C], c0[C], c1[C]; reduce: + s1[C], s2[C]
  const int n = blockIdx.y, C = GL.C, H = GL.H, W = GL.W, p = GL.pad, Wp = W + 2 * p, Hp = H + 2 * p;
  const float inv_cnt = 1.f / (float)(H * W);
  float* s_cA = sh; float* s_cB = sh + C; float* s_c0 = sh + 2 * C; float* s_c1 = sh + 3 * C;
  float* s_r1 = sh + 4 * C; float* s_r2 = sh + 5 * C;
  // mask z = A*r + Bc;  reduce: xhat = c0*r + c1 (c0 = rstd, c1 = -mean*rstd);  apply: o = A*gg + c0 + c1*r
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s1 = stats[((size_t)n * C + c) * 2], s2 = stats[((size_t)n * C + c) * 2 + 1];
    const double mean_d = s1 * (double)inv_cnt;
    const float mean = (float)mean_d;
    const float var = fmaxf((float)(s2 * (double)inv_cnt - mean_d * mean_d), 0.f);   // same arithmetic as the forward apply
    const float rstd = rsqrtf(var + eps);
    const float A = gamma[c] * rstd;
    s_cA[c] = A;
    s_cB[c] = beta[c] - mean * A;
    if (APPLY) {
      const float m1 = red[((size_t)n * C + c) * 2] * inv_cnt, m2 = red[((size_t)n * C + c) * 2 + 1] * inv_cnt;
      const float k1 = -A * m2 * rstd;
      s_c1[c] = k1;
      s_c0[c] = -A * m1 - k1 * mean;
    } else {
      s_c0[c] = rstd;
      s_c1[c] = -mean * rstd;
      s_r1[c] = 0.f;
      s_r2[c] = 0.f;
    }
  }
  __syncthreads();
  const int groups = C >> 3;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups, step = blockDim.x / groups;
  float a1[8], a2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a1[j] = a2[j] = 0.f;
  const float4* vA = reinterpret_cast<const float4*>(s_cA + g * 8);
  const float4* vB = reinterpret_cast<const float4*>(s_cB + g * 8);
  const float4* v0 = reinterpret_cast<const float4*>(s_c0 + g * 8);
  const float4* v1 = reinterpret_cast<const float4*>(s_c1 + g * 8);
  auto body = [&](const uint4& gq, const uint4& sq, const uint4& rq, bool has_skip, size_t pix, int y, int x) {
    F8 gv = cvt8(gq);
    const F8 rv = cvt8(rq);
    if (has_skip) {
      const F8 sv = cvt8(sq);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv.v[j] += sv.v[j];
    }
    if (APPLY && gsum) st8(gsum + pix * C + g * 8, gv);
    F8 o;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 qa = vA[h], qb = vB[h], q0 = v0[h], q1 = v1[h];
      const float cA[4] = {qa.x, qa.y, qa.z, qa.w}, cB[4] = {qb.x, qb.y, qb.z, qb.w};
      const float c0[4] = {q0.x, q0.y, q0.z, q0.w}, c1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = h * 4 + k;
        float gg = gv.v[j];
        if (relu && fmaf(rv.v[j], cA[k], cB[k]) <= 0.f) gg = 0.f;
        if (APPLY) o.v[j] = fmaf(cA[k], gg, fmaf(c1[k], rv.v[j], c0[k]));
        else { a1[j] += gg; a2[j] = fmaf(gg, fmaf(c0[k], rv.v[j], c1[k]), a2[j]); }
      }
    }
    if (APPLY) st8(draw + act_offset(DL, N, n, y, x) + g * 8, o);
  };
  if (pl < step) {
    const int y_begin = blockIdx.x * rows_per_block, y_end = min(H, y_begin + rows_per_block);
    const bool has_skip = skip != nullptr;
    for (int y = y_begin; y < y_end; ++y) {
      const __nv_bfloat16* grow = G + (((size_t)n * Hp + y + p) * Wp + p) * C + g * 8;
      const size_t pix0 = ((size_t)n * H + y) * W;
      const __nv_bfloat16* rrow = raw + pix0 * C + g * 8;
      const __nv_bfloat16* srow = has_skip ? skip + pix0 * C + g * 8 : nullptr;
      int x = pl;
      if (PX == 2)
      for (; x + step < W; x += 2 * step) {
        const int x2 = x + step;
        const uint4 g0 = ldq(grow + (size_t)x * C), g1 = ldq(grow + (size_t)x2 * C);
        const uint4 r0 = ldq(rrow + (size_t)x * C), r1 = ldq(rrow + (size_t)x2 * C);
        uint4 s0 = make_uint4(0, 0, 0, 0), s1 = s0;
        if (has_skip) { s0 = ldq(srow + (size_t)x * C); s1 = ldq(srow + (size_t)x2 * C); }
        body(g0, s0, r0, has_skip, pix0 + x, y, x);
        body(g1, s1, r1, has_skip, pix0 + x2, y, x2);
      }
      for (; x < W; x += step) {
        const uint4 g0 = ldq(grow + (size_t)x * C), r0 = ldq(rrow + (size_t)x * C);
        uint4 s0 = make_uint4(0, 0, 0, 0);
        if (has_skip) s0 = ldq(srow + (size_t)x * C);
        body(g0, s0, r0, has_skip, pix0 + x, y, x);
      }
    }
  }
  if (!APPLY) {
    // Block reduction of the per-thread partials.  Threads of one channel group sit `groups` lanes apart, so a layer with few
    // channels has MANY threads per group (C = 16: 128 of 256) - adding all of them into one shared-memory word per channel
    // (a compare-and-swap loop for floats) serialised the tail of every block: on RTNSTV's 16..48-channel layers the reduce
    // pass cost 4x the apply pass that reads AND writes the same tensors.  Now: a strided shuffle reduction inside each warp
    // (lanes l, l + groups, l + 2 groups ... by doubling offsets; idle lanes hold zeros), the warp totals parked in shared
    // memory with plain stores, and one thread per (channel, moment) adds the eight warps up - no shared-memory atomics.
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (groups <= 32) {
      for (int off = groups; off < 32; off <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float t1 = __shfl_down_sync(0xffffffffu, a1[j], off & 31), t2 = __shfl_down_sync(0xffffffffu, a2[j], off & 31);
          if (lane + off < 32) { a1[j] += t1; a2[j] += t2; }
        }
      }
      float* scr = sh + 6 * C;   // [warps][2 C]
      if (lane < groups) {       // the first lane of each group in this warp (lanes 0..groups-1 cover every group once)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          scr[wid * 2 * C + g * 8 + j] = a1[j];
          scr[wid * 2 * C + C + g * 8 + j] = a2[j];
        }
      }
      __syncthreads();
      for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        float sum = 0.f;
        for (int w = 0; w < nw; ++w) sum += scr[w * 2 * C + c];
        const int ch = c < C ? c : c - C;
        atomicAdd(&red[((size_t)n * C + ch) * 2 + (c < C ? 0 : 1)], sum);
      }
      return;
    }
    if (pl < step) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&s_r1[g * 8 + j], a1[j]);
        atomicAdd(&s_r2[g * 8 + j], a2[j]);
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      atomicAdd(&red[((size_t)n * C + c) * 2], s_r1[c]);
      atomicAdd(&red[((size_t)n * C + c) * 2 + 1], s_r2[c]);
    }
  }
}

__global__ void in_param_grads_kernel(const float* __restrict__ red, float* __restrict__ dgamma, float* __restrict__ dbeta, int N,
                                      int C) {
  vst::pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float a = 0.f, b = 0.f;
  for (int n = 0; n < N; ++n) {
    b += red[((size_t)n * C + c) * 2];
    a += red[((size_t)n * C + c) * 2 + 1];
  }
  dgamma[c] = a;
  dbeta[c] = b;
}

// ---- VGG body ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N,
                                                            int H, int W, int C) {
  vst::pdl_grid_sync();
  const int Ho = H / 2, Wo = W / 2, groups = C >> 3;
  const size_t total = (size_t)N * Ho * Wo * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;   // launcher guarantees total < 2^32: 32-bit divisions
    const int g = iu % groups;
    unsigned r = iu / groups;
    const int ox = r % Wo; r /= Wo;
    const int oy = r % Ho;
    const int n = r / Ho;
    const __nv_bfloat16* s = x + (((size_t)n * H + 2 * oy) * W + 2 * ox) * C + g * 8;
    const F8 a = ld8(s), b = ld8(s + C), c = ld8(s + (size_t)W * C), d = ld8(s + (size_t)W * C + C);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = fmaxf(fmaxf(a.v[j], b.v[j]), fmaxf(c.v[j], d.v[j]));
    st8(y + (((size_t)n * Ho + oy) * Wo + ox) * C + g * 8, o);
  }
}

// gm = (g + add) * (y > 0): the un-pooled ReLU adjoint is its own kernel - inside the pooled one it inherited 80 registers
// (37 % occupancy) and ran at 3.7 TB/s; alone it needs < 32 and streams at full occupancy.
__global__ void __launch_bounds__(256) relu_bwd_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                                                       const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ gm,
                                                       size_t total) {
  vst::pdl_grid_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    F8 gv = ld8(g + i * 8);
    const F8 yv = ld8(y + i * 8);
    if (add) {
      const F8 av = ld8(add + i * 8);
#pragma unroll
      for (int j = 0; j < 8; ++j) gv.v[j] += av.v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) gv.v[j] = yv.v[j] > 0.f ? gv.v[j] : 0.f;
    st8(gm + i * 8, gv);
  }
}

// gm = (route(g) + add) * (y > 0), route = max-pool adjoint (one thread per 2x2 window)
__global__ void __launch_bounds__(256) relu_pool_bwd_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ y,
                                                            const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ gm,
                                                            int N, int H, int W, int C) {
  vst::pdl_grid_sync();
  const int groups = C >> 3;
  // pooled: windows cover rows/cols [0, 2*Ho) x [0, 2*Wo); an odd last row / column gets only `add`
  const int Hc = (H + 1) / 2, Wc = (W + 1) / 2, Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)N * Hc * Wc * groups;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;
    const int gq = iu % groups;
    unsigned r = iu / groups;
    const int ox = r % Wc; r /= Wc;
    const int oy = r % Hc;
    const int n = r / Hc;
    const bool in_win = oy < Ho && ox < Wo;
    F8 gv;
#pragma unroll
    for (int j = 0; j < 8; ++j) gv.v[j] = 0.f;
    if (in_win) gv = ld8(g + (((size_t)n * Ho + oy) * Wo + ox) * C + gq * 8);
    F8 yv[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = 2 * oy + (k >> 1), xx = 2 * ox + (k & 1);
      ok[k] = yy < H && xx < W;
      if (ok[k]) yv[k] = ld8(y + (((size_t)n * H + yy) * W + xx) * C + gq * 8);
    }
    int best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      best[j] = 0;
      if (in_win) {
        float bv = yv[0].v[j];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (yv[k].v[j] > bv) { bv = yv[k].v[j]; best[j] = k; }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!ok[k]) continue;
      const int yy = 2 * oy + (k >> 1), xx = 2 * ox + (k & 1);
      const size_t off = (((size_t)n * H + yy) * W + xx) * C + gq * 8;
      F8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = (in_win && best[j] == k) ? gv.v[j] : 0.f;
      if (add) {
        const F8 av = ld8(add + off);
#pragma unroll
        for (int j = 0; j < 8; ++j) o.v[j] += av.v[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = yv[k].v[j] > 0.f ? o.v[j] : 0.f;
      st8(gm + off, o);
    }
  }
}

// ---- loss helpers -------------------------------------------------------------------------------------
constexpr int kRedBlocks = 1024;
__global__ void __launch_bounds__(256) sqdiff_sum_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                                              float* __restrict__ out, float* __restrict__ scratch, size_t n8) {
  vst::pdl_grid_sync();
  __shared__ float red[32];
  __shared__ bool last;
  float v = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const F8 p = ld8(a + i * 8), q = ld8(b + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = p.v[j] - q.v[j];
      v = fmaf(d, d, v);
    }
  }
  v = block_sum(v, red);
  // deterministic finish: per-block partials, the last block adds them in index order
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + kRedBlocks);
  if (threadIdx.x == 0) {
    scratch[blockIdx.x] = v;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += scratch[i];
    t = block_sum(t, red);
    if (threadIdx.x == 0) {
      out[0] = t;
      *counter = 0u;
    }
  }
}

__global__ void sqdiff_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, float scale,
                                       __nv_bfloat16* __restrict__ da, size_t n8) {
  vst::pdl_grid_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const F8 p = ld8(a + i * 8), q = ld8(b + i * 8);
    F8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = 2.f * scale * (p.v[j] - q.v[j]);
    st8(da + i * 8, o);
  }
}

__global__ void gram_grad_weights_kernel(const float* __restrict__ G, const float* __restrict__ Gs, int gs_batch, float scale,
                                         __nv_bfloat16* __restrict__ S, int B, int C) {
  vst::pdl_grid_sync();
  const size_t total = (size_t)B * C * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = i % C, r = (i / C) % C, b = i / ((size_t)C * C);
    const float* g = G + (size_t)b * C * C;
    const float* s = Gs + (size_t)(gs_batch == 1 ? 0 : b) * C * C;
    const float d = (g[(size_t)r * C + j] - s[(size_t)r * C + j]) + (g[(size_t)j * C + r] - s[(size_t)j * C + r]);
    S[i] = __float2bfloat16_rn(scale * d);
  }
}

__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n8) {
  vst::pdl_grid_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    F8 a = ld8(y + i * 8);
    const F8 b = ld8(x + i * 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) a.v[j] += b.v[j];
    st8(y + i * 8, a);
  }
}

// E[n][y][x'][kx*Co + co] = dz[n][co][y][x' - kx]; one thread per (n, y, x'), KE = 32 channels = four 16-byte stores
template <int K, int CO>   // K > 0: compile-time kernel width / channel count (row[] stays in registers)
__global__ void __launch_bounds__(256) rowconv_expand_kernel(const float* __restrict__ dz, __nv_bfloat16* __restrict__ E, int N, int Co_rt,
                                                             int H, int W, int k_rt) {
  vst::pdl_grid_sync();
  const int k = K > 0 ? K : k_rt, Co = K > 0 ? CO : Co_rt;
  const int Wp = W + k - 1;
  const size_t total = (size_t)N * H * Wp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;
    const int xp = iu % Wp, y = (iu / Wp) % H, n = iu / ((unsigned)Wp * H);
    __align__(16) __nv_bfloat16 row[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) row[j] = __float2bfloat16_rn(0.f);
    if (K > 0) {
#pragma unroll
      for (int kx = 0; kx < (K > 0 ? K : 1); ++kx) {
        const int x = xp - kx;
        const bool ok = x >= 0 && x < W;
#pragma unroll
        for (int co = 0; co < (K > 0 ? CO : 1); ++co)
          if (ok) row[kx * CO + co] = __float2bfloat16_rn(__ldg(dz + (((size_t)n * CO + co) * H + y) * W + x));
      }
    } else {
      for (int kx = 0; kx < k; ++kx) {
        const int x = xp - kx;
        if (x < 0 || x >= W) continue;
        for (int co = 0; co < Co; ++co)
          row[kx * Co + co] = __float2bfloat16_rn(__ldg(dz + (((size_t)n * Co + co) * H + y) * W + x));
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(E + i * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = reinterpret_cast<const uint4*>(row)[j];
  }
}

__global__ void __launch_bounds__(256) prologue_x27_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H, int W) {
  vst::pdl_grid_sync();
  const size_t HW = (size_t)H * W, total = (size_t)N * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;
    const int px = iu % W, py = (iu / W) % H, n = iu / (unsigned)HW;
    __align__(16) __nv_bfloat16 row[32];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
      for (int c = 0; c < 3; ++c) row[t * 3 + c] = __float2bfloat16_rn(ok ? __ldg(x + ((size_t)n * 3 + c) * HW + (size_t)yy * W + xx) : 0.f);
    }
#pragma unroll
    for (int j = 27; j < 32; ++j) row[j] = __float2bfloat16_rn(0.f);
    uint4* dst = reinterpret_cast<uint4*>(out + i * 32);
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = reinterpret_cast<const uint4*>(row)[j];
  }
}

static inline ActLayout to_layout(const vst_act_desc& d) { return ActLayout{d.H, d.W, d.C, d.pad, d.kind, d.parity}; }

}  // namespace vst

using namespace vst;

extern "C" {

int vst_gather_sum_f32(const float* src, const int* idx, int terms, void* dst, size_t n, int dst_bf16, void* stream) {
  VST_CHECK_ARG(n > 0 && terms >= 1 && terms <= 16, "gather_sum: bad arguments");
  VST_DEVPTR(src); VST_DEVPTR(idx); VST_DEVPTR(dst);
  vst::launch(gather_sum_kernel, tt_grid(n), 256, 0, (cudaStream_t)stream, src, idx, terms, dst, n, dst_bf16);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

static int in_bwd_launch(bool apply, const void* G, vst_act_desc g_desc, const void* skip, const void* raw, const double* stats,
                         const float* gamma, const float* beta, float* red, void* draw, vst_act_desc draw_desc, void* gsum, int N,
                         float eps, int relu, void* stream) {
  const ActLayout GL = to_layout(g_desc), DL = to_layout(draw_desc);
  VST_CHECK_ARG(N > 0 && GL.C % 8 == 0 && GL.C <= 1024 && GL.H > 0 && GL.W > 0, "in_bwd: bad shape");
  VST_CHECK_ARG(GL.parity == 0, "in_bwd: the incoming gradient must be a plain padded tensor");
  VST_CHECK_ARG(GL.kind != PADK_REFLECT || (2 * GL.pad < GL.H && 2 * GL.pad < GL.W), "in_bwd: reflect pad too large");
  VST_CHECK_ARG(GL.pad <= 5, "in_bwd: pad <= 5");
  static const int bps = [] { const char* e = getenv("VST_INBWD_BPS"); return e ? atoi(e) : 8; }();
  int rpb = cdiv(GL.H * N, kNumSMs * bps);
  if (rpb < 1) rpb = 1;
  dim3 grid(cdiv(GL.H, rpb), N);
  // constants + (reduce pass, C <= 256) the eight warps' partial sums [8][2 C]
  const size_t sh = (6 + ((!apply && GL.C <= 256) ? 16 : 0)) * (size_t)GL.C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (!apply && GL.pad > 0 && GL.kind != PADK_ZERO) {
    // the reduce pass runs first: fold the halo of G onto its interior in place (G is consumed by this layer only)
    vst::launch(fold_rows_kernel, tt_grid((size_t)N * (GL.W + 2 * GL.pad) * (GL.C / 8)), 256, 0, st, (__nv_bfloat16*)const_cast<void*>(G), GL, N);
    vst::launch(fold_cols_kernel, tt_grid((size_t)N * GL.H * (GL.C / 8)), 256, 0, st, (__nv_bfloat16*)const_cast<void*>(G), GL, N);
  }
  static const int variant = [] { const char* e = getenv("VST_INBWD_VARIANT"); return e ? atoi(e) : 0; }();
#define VST_INBWD_GO(AP, PX, MINB)                                                                                         \
  vst::launch(in_bwd_kernel<AP, PX, MINB>, grid, 256, sh, st, (const __nv_bfloat16*)G, GL, (const __nv_bfloat16*)skip,              \
                                                     (const __nv_bfloat16*)raw, stats, gamma, beta, red, (__nv_bfloat16*)draw, \
                                                     DL, (__nv_bfloat16*)gsum, N, eps, relu, rpb)
  if (apply) {
    VST_CHECK_ARG(DL.H == GL.H && DL.W == GL.W && DL.C == GL.C && DL.pad == 0, "in_bwd_apply: draw layout must be pad 0, same size");
    switch (variant) {
      case 1: VST_INBWD_GO(true, 1, 4); break;
      default: VST_INBWD_GO(true, 2, 3); break;
    }
  } else {
    VST_CUDA(cudaMemsetAsync(red, 0, (size_t)N * GL.C * 2 * sizeof(float), st));
    switch (variant) {
      case 1: VST_INBWD_GO(false, 1, 4); break;
      default: VST_INBWD_GO(false, 2, 3); break;
    }
  }
#undef VST_INBWD_GO
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_in_bwd_reduce(const void* G, vst_act_desc g_desc, const void* skip, const void* raw, const double* stats,
                         const float* gamma, const float* beta, float* red, int N, float eps, int relu, void* stream) {
  VST_DEVPTR(G); VST_DEVPTR(raw); VST_DEVPTR(stats); VST_DEVPTR(gamma); VST_DEVPTR(beta); VST_DEVPTR(red);
  return in_bwd_launch(false, G, g_desc, skip, raw, stats, gamma, beta, red, nullptr, g_desc, nullptr, N, eps, relu, stream);
}

int vst_tc_in_bwd_apply(const void* G, vst_act_desc g_desc, const void* skip, const void* raw, const double* stats,
                        const float* gamma, const float* beta, const float* red, void* draw, vst_act_desc draw_desc, void* gsum,
                        int N, float eps, int relu, void* stream) {
  VST_DEVPTR(G); VST_DEVPTR(raw); VST_DEVPTR(stats); VST_DEVPTR(gamma); VST_DEVPTR(beta); VST_DEVPTR(red); VST_DEVPTR(draw);
  return in_bwd_launch(true, G, g_desc, skip, raw, stats, gamma, beta, const_cast<float*>(red), draw, draw_desc, gsum, N, eps, relu,
                       stream);
}

int vst_tc_in_param_grads(const float* red, float* dgamma, float* dbeta, int N, int C, void* stream) {
  VST_CHECK_ARG(N > 0 && C > 0, "in_param_grads: bad shape");
  VST_DEVPTR(red); VST_DEVPTR(dgamma); VST_DEVPTR(dbeta);
  vst::launch(in_param_grads_kernel, cdiv(C, 128), 128, 0, (cudaStream_t)stream, red, dgamma, dbeta, N, C);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_maxpool2(const void* x, void* y, int N, int H, int W, int C, void* stream) {
  VST_CHECK_ARG(N > 0 && H >= 2 && W >= 2 && C % 8 == 0, "tc_maxpool2: bad shape");
  VST_CHECK_ARG((size_t)N * H * W * (C / 8) < ((size_t)1 << 32), "tc_maxpool2: tensor too large for 32-bit indexing");
  VST_DEVPTR(x); VST_DEVPTR(y);
  vst::launch(maxpool2_nhwc_kernel, tt_grid((size_t)N * (H / 2) * (W / 2) * (C / 8)), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, N, H, W, C);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_relu_pool_bwd(const void* g, const void* y, const void* add, void* gm, int N, int H, int W, int C, int pooled,
                         void* stream) {
  VST_CHECK_ARG(N > 0 && H >= 1 && W >= 1 && C % 8 == 0, "tc_relu_pool_bwd: bad shape");
  VST_CHECK_ARG((size_t)N * H * W * (C / 8) < ((size_t)1 << 32), "tc_relu_pool_bwd: tensor too large for 32-bit indexing");
  VST_DEVPTR(g); VST_DEVPTR(y); VST_DEVPTR(gm);
  const size_t total = pooled ? (size_t)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8) : (size_t)N * H * W * (C / 8);
  if (!pooled)
    vst::launch(relu_bwd_kernel, tt_grid(total), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)g, (const __nv_bfloat16*)y,
                                                                     (const __nv_bfloat16*)add, (__nv_bfloat16*)gm, total);
  else
    vst::launch(relu_pool_bwd_kernel, tt_grid(total), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)g, (const __nv_bfloat16*)y,
                                                                          (const __nv_bfloat16*)add, (__nv_bfloat16*)gm, N, H, W, C);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_sqdiff_sum_bf16(const void* a, const void* b, float* out, float* scratch, size_t n, void* stream) {
  VST_CHECK_ARG(n > 0 && n % 8 == 0, "tc_sqdiff_sum: n must be a positive multiple of 8");
  VST_DEVPTR(a); VST_DEVPTR(b); VST_DEVPTR(out); VST_DEVPTR(scratch);
  int grid = tt_grid(n / 8);
  if (grid > kRedBlocks) grid = kRedBlocks;
  vst::launch(sqdiff_sum_bf16_kernel, grid, 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, out, scratch, n / 8);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_sqdiff_bwd_bf16(const void* a, const void* b, float scale, void* da, size_t n, void* stream) {
  VST_CHECK_ARG(n > 0 && n % 8 == 0, "tc_sqdiff_bwd: n must be a positive multiple of 8");
  VST_DEVPTR(a); VST_DEVPTR(b); VST_DEVPTR(da);
  vst::launch(sqdiff_bwd_bf16_kernel, tt_grid(n / 8), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, scale,
                                                                           (__nv_bfloat16*)da, n / 8);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_gram_grad_weights(const float* G, const float* Gs, int gs_batch, float scale, void* S, int B, int C, void* stream) {
  VST_CHECK_ARG(B > 0 && C > 0 && (gs_batch == 1 || gs_batch == B), "gram_grad_weights: bad shape");
  VST_DEVPTR(G); VST_DEVPTR(Gs); VST_DEVPTR(S);
  vst::launch(gram_grad_weights_kernel, tt_grid((size_t)B * C * C), 256, 0, (cudaStream_t)stream, G, Gs, gs_batch, scale, (__nv_bfloat16*)S, B, C);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_rowconv_expand(const float* dz, void* E, int N, int Co, int H, int W, int k, int KE, void* stream) {
  VST_CHECK_ARG(N > 0 && Co > 0 && H > 0 && W > 0 && k > 0 && KE == 32 && k * Co <= KE, "rowconv_expand: need k*Co <= KE == 32");
  VST_CHECK_ARG((size_t)N * H * (W + k - 1) < ((size_t)1 << 32), "rowconv_expand: tensor too large for 32-bit indexing");
  VST_DEVPTR(dz); VST_DEVPTR(E);
  if (k == 9 && Co == 3)
    vst::launch(rowconv_expand_kernel<9, 3>, tt_grid((size_t)N * H * (W + k - 1)), 256, 0, (cudaStream_t)stream, dz, (__nv_bfloat16*)E, N, Co, H, W, k);
  else
    vst::launch(rowconv_expand_kernel<0, 0>, tt_grid((size_t)N * H * (W + k - 1)), 256, 0, (cudaStream_t)stream, dz, (__nv_bfloat16*)E, N, Co, H, W, k);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_prologue_x27(const float* x, void* out, int N, int H, int W, void* stream) {
  VST_CHECK_ARG(N > 0 && H > 0 && W > 0 && (size_t)N * H * W < ((size_t)1 << 32), "prologue_x27: bad shape");
  VST_DEVPTR(x); VST_DEVPTR(out);
  vst::launch(prologue_x27_kernel, tt_grid((size_t)N * H * W), 256, 0, (cudaStream_t)stream, x, (__nv_bfloat16*)out, N, H, W);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tc_add_bf16(const void* x, void* y, size_t n, void* stream) {
  VST_CHECK_ARG(n > 0 && n % 8 == 0, "tc_add: n must be a positive multiple of 8");
  VST_DEVPTR(x); VST_DEVPTR(y);
  vst::launch(add_bf16_kernel, tt_grid(n / 8), 256, 0, (cudaStream_t)stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, n / 8);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

}  // extern "C"
