// Hand-written backward kernels of the training step (fp32, NCHW, CUDA cores) - SURVEY.md §10 B1-B17.
// The reference relies on autograd (`loss.backward()`, RC/train_single/train_starry-night.py:151); each
// kernel here is the adjoint of one forward entry point of include/vst_b200.h.
#include <algorithm>
#include "common.cuh"

namespace vst {

static inline int bw_grid(size_t total, int block = 256) {
  size_t g = (total + block - 1) / block;
  const size_t cap = (size_t)kNumSMs * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// ---- weight transform for stride-1 dgrad: wT[ci][co][a][b] = w[co][ci][k-1-a][k-1-b] ----------
__global__ void weight_flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin, int k) {
  vst::pdl_grid_sync();
  const size_t total = (size_t)Cout * Cin * k * k;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = i % k, a = (i / k) % k, co = (i / ((size_t)k * k)) % Cout, ci = i / ((size_t)k * k * Cout);
    wt[i] = w[(((size_t)co * Cin + ci) * k + (k - 1 - a)) * k + (k - 1 - b)];
  }
}

// ---- generic transposed convolution (gather): out[n,co,P,Q] = sum_{ci,ky,kx} in[n,ci,(P+pad-ky)/s,(Q+pad-kx)/s] w[ci][co][ky][kx]
// over the (ky,kx) with exact division and in-range source.  Used for ConvTranspose2d forward
// (k3,s2,p1) and as the data-gradient of strided convolutions (pad 0, output = padded input grad).
__global__ void __launch_bounds__(256) conv_transpose_gather_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
    int N, int Cin, int H, int W, int Cout, int Ho, int Wo, int k, int s, int pad) {
  vst::pdl_grid_sync();
  const int co_tiles = cdiv(Cout, 8);
  const size_t total = (size_t)N * co_tiles * Ho * Wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = i % Wo, oy = (i / Wo) % Ho;
    const int ct = (i / ((size_t)Wo * Ho)) % co_tiles, n = i / ((size_t)Wo * Ho * co_tiles);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int ny = oy + pad - ky;
      if (ny < 0 || ny % s) continue;
      const int iy = ny / s;
      if (iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int nx = ox + pad - kx;
        if (nx < 0 || nx % s) continue;
        const int ix = nx / s;
        if (ix >= W) continue;
        for (int ci = 0; ci < Cin; ++ci) {
          const float v = x[(((size_t)n * Cin + ci) * H + iy) * W + ix];
          const float* wp = w + (((size_t)ci * Cout + ct * 8) * k + ky) * k + kx;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (ct * 8 + j < Cout) acc[j] = fmaf(v, wp[(size_t)j * k * k], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = ct * 8 + j;
      if (co < Cout) y[(((size_t)n * Cout + co) * Ho + oy) * Wo + ox] = acc[j] + (bias ? bias[co] : 0.f);
    }
  }
}

// ---- fold the gradient of the padded / upsampled input back onto the source tensor -----------
// dx[n,c,ys,xs] = sum over the ups x ups logical pixels (Y,X) of the source pixel, over every padded
// position that the pad mode maps onto (Y,X): itself, plus its mirror images for reflection padding
// (aten::reflection_pad2d_backward) - SURVEY.md B13/B14.
__global__ void __launch_bounds__(256) fold_pad_kernel(const float* __restrict__ dxp, float* __restrict__ dx, int NC, int Hs,
                                                       int Ws, int ups, int pad, int pad_mode, int Hp, int Wp) {
  vst::pdl_grid_sync();
  const int Hl = Hs * ups, Wl = Ws * ups;
  const size_t total = (size_t)NC * Hs * Ws;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xs = i % Ws, ys = (i / Ws) % Hs;
    const size_t nc = i / ((size_t)Ws * Hs);
    const float* p = dxp + nc * Hp * Wp;
    float acc = 0.f;
    for (int dy = 0; dy < ups; ++dy) {
      const int Y = ys * ups + dy;
      int py[3], npy = 0;
      py[npy++] = Y + pad;
      if (pad_mode == VST_PAD_REFLECT) {
        if (Y >= 1 && Y <= pad) py[npy++] = pad - Y;                               // top mirror: index -Y
        if (Y <= Hl - 2 && Y >= Hl - 1 - pad) py[npy++] = 2 * (Hl - 1) - Y + pad;  // bottom mirror
      }
      for (int dx_ = 0; dx_ < ups; ++dx_) {
        const int X = xs * ups + dx_;
        int px[3], npx = 0;
        px[npx++] = X + pad;
        if (pad_mode == VST_PAD_REFLECT) {
          if (X >= 1 && X <= pad) px[npx++] = pad - X;
          if (X <= Wl - 2 && X >= Wl - 1 - pad) px[npx++] = 2 * (Wl - 1) - X + pad;
        }
        for (int a = 0; a < npy; ++a)
          for (int b = 0; b < npx; ++b)
            if (py[a] < Hp && px[b] < Wp) acc += p[(size_t)py[a] * Wp + px[b]];
      }
    }
    dx[i] = acc;
  }
}

// ---- weight gradient ----------------------------------------------------------------------------
// dw[co][ci][ky][kx] = sum_{n,oy,ox} dy[n,co,oy,ox] * xl[n,ci,map(oy*s+ky-pad),map(ox*s+kx-pad)]
// Block = 16 co x 16 ci threads; per 8x32 output tile the dy tile and the padded x patch go through
// smem; each thread keeps the k taps of ONE kernel row (blockIdx.z picks ky) in registers and slides
// a k-wide window along x.  Split over output tiles with fp32 atomics at the end.
constexpr int WG_TH = 8, WG_C = 16;
constexpr int wg_tw(int S) { return S == 1 ? 32 : 16; }

template <int K, int S>
__global__ void __launch_bounds__(256) conv2d_wgrad_kernel(  // WG_TW output columns per tile

    const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int N, int Cin, int Hs, int Ws,
    int Cout, int Ho, int Wo, int pad, int pad_mode, int ups, int tiles_x, int tiles_y, int tile_splits) {
  vst::pdl_grid_sync();
  constexpr int WG_TW = wg_tw(S);
  constexpr int PW = (WG_TW - 1) * S + K;          // patch width
  constexpr int PWP = PW + ((PW & 1) ? 0 : 1);     // odd pitch -> conflict-free across ci planes
  __shared__ float s_x[WG_C][WG_TH][PWP];          // one kernel row: only rows oy*S + ky are needed
  __shared__ float s_dy[WG_C][WG_TH][WG_TW + 1];
  const int ci_l = threadIdx.x & 15, co_l = threadIdx.x >> 4;
  const int co0 = blockIdx.x * WG_C, ci0 = blockIdx.y * WG_C;
  const int ky = blockIdx.z % K, split = blockIdx.z / K;
  const int H = Hs * ups, W = Ws * ups;
  float acc[K];
#pragma unroll
  for (int t = 0; t < K; ++t) acc[t] = 0.f;
  const int total_tiles = N * tiles_y * tiles_x;
  for (int tile = split; tile < total_tiles; tile += tile_splits) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
    const int oy0 = ty * WG_TH, ox0 = tx * WG_TW;
    __syncthreads();
    for (int idx = threadIdx.x; idx < WG_C * WG_TH * PW; idx += 256) {
      const int q = idx % PW, r = (idx / PW) % WG_TH, c = idx / (PW * WG_TH);
      int iy = (oy0 + r) * S + ky - pad, ix = ox0 * S + q - pad;
      float v = 0.f;
      if (ci0 + c < Cin && oy0 + r < Ho) {
        if (pad_mode == VST_PAD_REFLECT) { iy = reflect_idx(iy, H); ix = reflect_idx(ix, W); }
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((size_t)n * Cin + ci0 + c) * Hs + iy / ups) * Ws + ix / ups];
      }
      s_x[c][r][q] = v;
    }
    for (int idx = threadIdx.x; idx < WG_C * WG_TH * WG_TW; idx += 256) {
      const int q = idx % WG_TW, r = (idx / WG_TW) % WG_TH, c = idx / (WG_TW * WG_TH);
      float v = 0.f;
      if (co0 + c < Cout && oy0 + r < Ho && ox0 + q < Wo) v = dy[(((size_t)n * Cout + co0 + c) * Ho + oy0 + r) * Wo + ox0 + q];
      s_dy[c][r][q] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < WG_TH; ++r) {
      float win[K];
#pragma unroll
      for (int t = 0; t < K - 1; ++t) win[t + 1] = s_x[ci_l][r][t];
      for (int q = 0; q < WG_TW; ++q) {
        if (S == 1) {
#pragma unroll
          for (int t = 0; t < K - 1; ++t) win[t] = win[t + 1];
          win[K - 1] = s_x[ci_l][r][q + K - 1];
        } else {
#pragma unroll
          for (int t = 0; t < K; ++t) win[t] = s_x[ci_l][r][q * S + t];
        }
        const float g = s_dy[co_l][r][q];
#pragma unroll
        for (int t = 0; t < K; ++t) acc[t] = fmaf(g, win[t], acc[t]);
      }
    }
  }
  if (co0 + co_l < Cout && ci0 + ci_l < Cin) {
    float* o = dw + ((((size_t)(co0 + co_l)) * Cin + ci0 + ci_l) * K + ky) * K;
#pragma unroll
    for (int t = 0; t < K; ++t) atomicAdd(o + t, acc[t]);
  }
}

// per-channel sum over N and HW (bias gradients): grid (C, splits), one atomic per block into the zeroed output
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int C,
                                                          int HW) {
  vst::pdl_grid_sync();
  __shared__ float red[32];
  const int c = blockIdx.x;
  const size_t total = (size_t)N * HW;
  float s = 0.f;
  for (size_t i = blockIdx.y * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.y * blockDim.x) {
    const size_t n = i / HW, p = i % HW;
    s += x[(n * C + c) * HW + p];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(&out[c], s);
}

// ---- activation adjoints from the saved OUTPUT ----------------------------------------------------
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dz, size_t n,
                               int act) {
  vst::pdl_grid_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = dy[i], o = y[i];
    float r = g;
    if (act == VST_ACT_RELU) r = o > 0.f ? g : 0.f;
    else if (act == VST_ACT_TANH) r = g * (1.f - o * o);
    else if (act == VST_ACT_RECONET_OUT) {  // y = tanh(z/255)*150 + 127.5
      const float t = (o - 127.5f) / 150.f;
      r = g * (150.f / 255.f) * (1.f - t * t);
    } else if (act == VST_ACT_RT_OUT) {     // y = (tanh(z)+1)/2*255
      const float t = o / 127.5f - 1.f;
      r = g * 127.5f * (1.f - t * t);
    }
    dz[i] = r;
  }
}

// ---- InstanceNorm backward (affine, instance statistics) + activation adjoint -----------------------
// z = xhat*gamma + beta, y = act(z) (+ residual).  g = dL/dz.
// dx = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)); dgamma += sum g*xhat; dbeta += sum g.
__device__ __forceinline__ float in_act_grad(float g, float z, int act) {
  switch (act) {
    case VST_ACT_RELU: return z > 0.f ? g : 0.f;
    case VST_ACT_TANH: { const float t = tanhf(z); return g * (1.f - t * t); }
    case VST_ACT_RT_OUT: { const float t = tanhf(z); return g * 127.5f * (1.f - t * t); }
    default: return g;
  }
}
__global__ void __launch_bounds__(1024) instance_norm_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
    float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int C, int HW, int act) {
  vst::pdl_grid_sync();
  // one block or one cluster of blocks per (n, c) plane (common.cuh: cluster_sum); a CTA owns a contiguous slice
  __shared__ float red[32];
  __shared__ float slots[2];
  const uint32_t nb = cluster_nctarank(), rank = cluster_ctarank_();
  const int plane = blockIdx.x / nb, c = plane % C;
  const int chunk = (((HW + (int)nb - 1) / (int)nb) + 3) & ~3;
  const int i0 = min(HW, (int)rank * chunk), i1 = min(HW, i0 + chunk);
  const int B = blockDim.x;
  const float mu = mean[plane], rs = rstd[plane], ga = gamma[c], be = beta[c];
  const float* xp = x + (size_t)plane * HW;
  const float* gp = dy + (size_t)plane * HW;
  float s1 = 0.f, s2 = 0.f;
  {
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    int i = i0 + threadIdx.x;
    for (; i + B < i1; i += 2 * B) {
      const float xv0 = xp[i], xv1 = xp[i + B], gv0 = gp[i], gv1 = gp[i + B];
      const float xh0 = (xv0 - mu) * rs, xh1 = (xv1 - mu) * rs;
      const float g0 = in_act_grad(gv0, fmaf(xh0, ga, be), act), g1 = in_act_grad(gv1, fmaf(xh1, ga, be), act);
      a0 += g0; a1 += g1;
      b0 = fmaf(g0, xh0, b0); b1 = fmaf(g1, xh1, b1);
    }
    for (; i < i1; i += B) {
      const float xh = (xp[i] - mu) * rs;
      const float g = in_act_grad(gp[i], fmaf(xh, ga, be), act);
      a0 += g;
      b0 = fmaf(g, xh, b0);
    }
    s1 = a0 + a1;
    s2 = b0 + b1;
  }
  s1 = cluster_sum(s1, red, &slots[0], nb);
  s2 = cluster_sum(s2, red, &slots[1], nb);
  if (nb > 1) cluster_barrier();   // nobody leaves while a peer may still read its slots
  if (threadIdx.x == 0 && rank == 0) {
    atomicAdd(&dbeta[c], s1);
    atomicAdd(&dgamma[c], s2);
  }
  const float m1 = s1 / (float)HW, m2 = s2 / (float)HW, k = ga * rs;
  float* op = dx + (size_t)plane * HW;
  for (int i = i0 + threadIdx.x; i < i1; i += B) {
    const float xh = (xp[i] - mu) * rs;
    const float g = in_act_grad(gp[i], fmaf(xh, ga, be), act);
    op[i] = k * (g - m1 - xh * m2);
  }
}

// ---- max-pool 2x2 backward: route to the first maximum of each window (scan order h, w) -------------
__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int NC,
                                    int H, int W) {
  vst::pdl_grid_sync();
  const size_t total = (size_t)NC * H * W;
  const int Ho = H / 2, Wo = W / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xx = i % W, yy = (i / W) % H;
    const size_t p = i / ((size_t)W * H);
    const int oy = yy >> 1, ox = xx >> 1;
    float r = 0.f;
    if (oy < Ho && ox < Wo) {
      const float* s = x + (p * H + 2 * oy) * W + 2 * ox;
      int best = 0;
      float bv = s[0];
      if (s[1] > bv) { bv = s[1]; best = 1; }
      if (s[W] > bv) { bv = s[W]; best = 2; }
      if (s[W + 1] > bv) { bv = s[W + 1]; best = 3; }
      if (best == (yy & 1) * 2 + (xx & 1)) r = dy[(p * Ho + oy) * Wo + ox];
    }
    dx[i] = r;
  }
}

// per-channel scale: y[n,c,:] = x[n,c,:] * s[c]   (vgg_normalize backward: 1/(255*std_c))
__global__ void channel_scale_kernel(const float* __restrict__ x, float* __restrict__ y, size_t total, int C, int HW, float s0,
                                     float s1, float s2) {
  vst::pdl_grid_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (i / HW) % C;
    y[i] = x[i] * (c == 0 ? s0 : c == 1 ? s1 : s2);
  }
}

// ---- warp backward w.r.t. the sampled tensor: scatter-add dy * bilinear weights (safe_add_2d) -------
struct Bilin2 {
  int x0, y0;
  float w[4];
};
__device__ __forceinline__ Bilin2 bilin2_setup(int px, int py, float fx, float fy, int W, int H) {
  const float ix = warp_src_coord(px, fx, W), iy = warp_src_coord(py, fy, H);
  const float x0f = floorf(ix), y0f = floorf(iy);
  Bilin2 b;
  b.x0 = (int)x0f;
  b.y0 = (int)y0f;
  const float wx1 = ix - x0f, wy1 = iy - y0f, wx0 = (x0f + 1.f) - ix, wy0 = (y0f + 1.f) - iy;
  b.w[0] = wx0 * wy0; b.w[1] = wx1 * wy0; b.w[2] = wx0 * wy1; b.w[3] = wx1 * wy1;
  return b;
}
__device__ __forceinline__ void bilin2_scatter(float* __restrict__ p, const Bilin2& b, int W, int H, float g) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cx = b.x0 + (k & 1), cy = b.y0 + (k >> 1);
    if (cx >= 0 && cx < W && cy >= 0 && cy < H) atomicAdd(p + (size_t)cy * W + cx, g * b.w[k]);
  }
}
__device__ __forceinline__ float bilin2_sample(const float* __restrict__ p, const Bilin2& b, int W, int H) {
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cx = b.x0 + (k & 1), cy = b.y0 + (k >> 1);
    if (cx >= 0 && cx < W && cy >= 0 && cy < H) acc += p[(size_t)cy * W + cx] * b.w[k];
  }
  return acc;
}

__global__ void __launch_bounds__(256) warp_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ flo,
                                                       float* __restrict__ dx, int B, int C, int H, int W) {
  vst::pdl_grid_sync();
  const size_t HW = (size_t)H * W, total = (size_t)B * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;   // B*H*W < 2^32 (checked by the launcher)
    const int px = iu % W, py = (iu / W) % H, b = iu / (unsigned)HW;
    const float* f = flo + (size_t)b * 2 * HW + (size_t)py * W + px;
    const Bilin2 bl = bilin2_setup(px, py, f[0], f[HW], W, H);
    for (int c = 0; c < C; ++c)
      bilin2_scatter(dx + ((size_t)b * C + c) * HW, bl, W, H, dy[((size_t)b * C + c) * HW + (size_t)py * W + px]);
  }
}

// resize helpers shared with the forward feature-temporal kernel (warp_ops.cu keeps its own copy)
__device__ __forceinline__ void resize_src2(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  s = s < 0.f ? 0.f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = fminf(fmaxf(__fsub_rn(s, (float)i0), 0.f), 1.f);
  l0 = __fsub_rn(1.f, l1);
}
__device__ __forceinline__ float resize_sample2(const float* __restrict__ p, int W, int y0, int y1, int x0, int x1, float ly0,
                                                float ly1, float lx0, float lx1) {
  const float top = __fadd_rn(__fmul_rn(lx0, p[(size_t)y0 * W + x0]), __fmul_rn(lx1, p[(size_t)y0 * W + x1]));
  const float bot = __fadd_rn(__fmul_rn(lx0, p[(size_t)y1 * W + x0]), __fmul_rn(lx1, p[(size_t)y1 * W + x1]));
  return __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
}

// feature-temporal backward: g = 2*m*(f2 - warp(f1))*scale;  df2 = g (dense write), df1 += scatter(-g)
constexpr int FTB_CPT = 16;
__global__ void __launch_bounds__(256) feature_temporal_bwd_kernel(
    const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ flow,
    const float* __restrict__ mask, const float* __restrict__ scale_dev, float scale_host, float* __restrict__ df1,
    float* __restrict__ df2, int B, int C, int Hf, int Wf, int H, int W) {
  vst::pdl_grid_sync();
  const size_t HWf = (size_t)Hf * Wf, HW = (size_t)H * W, total = (size_t)B * HWf;
  const float sh = (float)H / (float)Hf, sw = (float)W / (float)Wf;
  const float mu = (float)((double)Wf / (double)W), mv = (float)((double)Hf / (double)H);
  const float scale = 2.f * scale_host * (scale_dev ? scale_dev[0] : 1.f);
  // blockIdx.y owns a chunk of FTB_CPT channels: a feature map has few pixels (B*Hf*Wf ~ 56k at 1024x436) and many
  // channels, so one thread per pixel looping over all of them left most of the GPU idle (1.2 TB/s)
  const int c_begin = blockIdx.y * FTB_CPT, c_end = min(C, c_begin + FTB_CPT);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;   // B*Hf*Wf < 2^32 (checked by the launcher)
    const int px = iu % Wf, py = (iu / Wf) % Hf, b = iu / (unsigned)HWf;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    resize_src2(py, sh, H, y0, y1, ly0, ly1);
    resize_src2(px, sw, W, x0, x1, lx0, lx1);
    const float m = resize_sample2(mask + (size_t)b * HW, W, y0, y1, x0, x1, ly0, ly1, lx0, lx1) > 0.f ? 1.f : 0.f;
    float* o2 = df2 + (size_t)b * C * HWf + (size_t)py * Wf + px;
    if (m == 0.f) {
      for (int c = c_begin; c < c_end; ++c) o2[c * HWf] = 0.f;
      continue;
    }
    const float* fl = flow + (size_t)b * 2 * HW;
    const float u = __fmul_rn(resize_sample2(fl, W, y0, y1, x0, x1, ly0, ly1, lx0, lx1), mu);
    const float w = __fmul_rn(resize_sample2(fl + HW, W, y0, y1, x0, x1, ly0, ly1, lx0, lx1), mv);
    const Bilin2 bl = bilin2_setup(px, py, u, w, Wf, Hf);
    const float* p1 = f1 + (size_t)b * C * HWf;
    const float* p2 = f2 + (size_t)b * C * HWf + (size_t)py * Wf + px;
    for (int c = c_begin; c < c_end; ++c) {
      const float g = scale * (p2[c * HWf] - bilin2_sample(p1 + c * HWf, bl, Wf, Hf));
      o2[c * HWf] = g;
      bilin2_scatter(df1 + ((size_t)b * C + c) * HWf, bl, Wf, Hf, -g);
    }
  }
}

// output-temporal backward: e_c = (s2_c - ws_c) - Y;  ds2_c = 2*m*e_c*scale;  ds1 += scatter(-ds2_c)
__global__ void __launch_bounds__(256) output_temporal_bwd_kernel(
    const float* __restrict__ s1, const float* __restrict__ s2, const float* __restrict__ i1,
    const float* __restrict__ i2, const float* __restrict__ flow, const float* __restrict__ mask, float scale_host,
    const float* __restrict__ scale_dev, float* __restrict__ ds1, float* __restrict__ ds2, int B, int H, int W, int luminance) {
  vst::pdl_grid_sync();
  const size_t HW = (size_t)H * W, total = (size_t)B * HW;
  const float scale = 2.f * scale_host * (scale_dev ? scale_dev[0] : 1.f);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned iu = (unsigned)i;   // B*H*W < 2^32 (checked by the launcher)
    const int px = iu % W, py = (iu / W) % H, b = iu / (unsigned)HW;
    const size_t pix = (size_t)py * W + px, base = (size_t)b * 3 * HW;
    const float m = mask[(size_t)b * HW + pix];
    if (m == 0.f) {
#pragma unroll
      for (int c = 0; c < 3; ++c) ds2[base + c * HW + pix] = 0.f;
      continue;
    }
    const float* f = flow + (size_t)b * 2 * HW + pix;
    const Bilin2 bl = bilin2_setup(px, py, f[0], f[HW], W, H);
    float lum = 0.f;
    if (luminance) {
      const float k[3] = {0.2126f, 0.7152f, 0.0722f};
#pragma unroll
      for (int c = 0; c < 3; ++c) lum += k[c] * (i2[base + c * HW + pix] - bilin2_sample(i1 + base + c * HW, bl, W, H));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float e = (s2[base + c * HW + pix] - bilin2_sample(s1 + base + c * HW, bl, W, H)) - lum;
      const float g = m * e * scale;
      ds2[base + c * HW + pix] = g;
      bilin2_scatter(ds1 + base + c * HW, bl, W, H, -g);
    }
  }
}

// da = 2*(a-b)*scale (MSE numerators); `db` optional (= -da)
__global__ void sqdiff_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float scale, float* __restrict__ da,
                                  float* __restrict__ db, size_t n) {
  vst::pdl_grid_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = 2.f * (a[i] - b[i]) * scale;
    da[i] = g;
    if (db) db[i] = -g;
  }
}

// TV backward (gather form).  window pixels: (y,x) with y < H-1, x < W-1; terms dx = x[y][x+1]-x[y][x], dy = x[y+1][x]-x[y][x].
// mode 0: L = sum dx^2+dy^2            -> dL/d(dx) = 2dx
// mode 1: L = sum sqrt(max(s,1e-8))    -> dL/d(dx) = dx / sqrt(s) where s > 1e-8, else 0
__device__ __forceinline__ void tv_terms(const float* __restrict__ p, int W, int mode, float& gdx, float& gdy) {
  const float dx = p[1] - p[0], dy = p[W] - p[0];
  if (mode == 0) { gdx = 2.f * dx; gdy = 2.f * dy; return; }
  const float s = dx * dx + dy * dy;
  if (s > 1e-8f) { const float r = rsqrtf(s); gdx = dx * r; gdy = dy * r; } else { gdx = 0.f; gdy = 0.f; }
}
__global__ void tv_bwd_kernel(const float* __restrict__ x, float scale, float* __restrict__ dx_out, int BC, int H, int W, int mode) {
  vst::pdl_grid_sync();
  const size_t total = (size_t)BC * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int xx = i % W, yy = (i / W) % H;
    const float* p = x + i;
    float g = 0.f, a, b;
    if (yy < H - 1 && xx < W - 1) { tv_terms(p, W, mode, a, b); g -= a + b; }          // centre of its own window
    if (yy < H - 1 && xx >= 1) { tv_terms(p - 1, W, mode, a, b); g += a; }              // right neighbour of (y, x-1)
    if (yy >= 1 && xx < W - 1) { tv_terms(p - W, W, mode, a, b); g += b; }              // bottom neighbour of (y-1, x)
    dx_out[i] = g * scale;
  }
}

// Gram backward: dF[b][i][p] = scale * sum_j (dG[b][i][j] + dG[b][j][i]) * F[b][j][p]
constexpr int GB_TI = 64, GB_TP = 64, GB_K = 16;
__global__ void __launch_bounds__(256) gram_bwd_kernel(const float* __restrict__ F, const float* __restrict__ dG,
                                                       float* __restrict__ dF, int C, int HW, float scale) {
  vst::pdl_grid_sync();
  __shared__ float sa[GB_K][GB_TI + 4], sb[GB_K][GB_TP + 4];
  const int b = blockIdx.z, i0 = blockIdx.y * GB_TI, p0 = blockIdx.x * GB_TP;
  const float* Fb = F + (size_t)b * C * HW;
  const float* Gb = dG + (size_t)b * C * C;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int j0 = 0; j0 < C; j0 += GB_K) {
    for (int idx = threadIdx.x; idx < GB_TI * GB_K; idx += 256) {
      const int jj = idx % GB_K, r = idx / GB_K;
      const int i = i0 + r, j = j0 + jj;
      sa[jj][r] = (i < C && j < C) ? Gb[(size_t)i * C + j] + Gb[(size_t)j * C + i] : 0.f;
    }
    for (int idx = threadIdx.x; idx < GB_TP * GB_K; idx += 256) {
      const int pp = idx % GB_TP, jj = idx / GB_TP;
      const int j = j0 + jj, p = p0 + pp;
      sb[jj][pp] = (j < C && p < HW) ? Fb[(size_t)j * HW + p] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < GB_K; ++jj) {
      float a[4], q[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = sa[jj][ty * 4 + r], q[r] = sb[jj][tx * 4 + r];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a[r], q[c], acc[r][c]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = i0 + ty * 4 + r, p = p0 + tx * 4 + c;
      if (i < C && p < HW) dF[((size_t)b * C + i) * HW + p] = acc[r][c] * scale;
    }
}

// Adam (torch.optim.Adam defaults: no weight decay, no amsgrad), single tensor, in place
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float step_size, float b1, float b2, float omb1, float omb2, float eps, float bc2_sqrt, float gscale,
                            const float* __restrict__ skip_flag) {
  vst::pdl_grid_sync();
  // a step whose loss divided by an empty occlusion mask (vst_loss_terms_f32 raised the flag) must leave weights and moments
  // untouched - the reference raises ZeroDivisionError before backward() (RC/...starry-night.py:105,122)
  if (skip_flag && *skip_flag != 0.f) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] = b1 * m[i] + omb1 * gi;
    const float vi = v[i] = b2 * v[i] + omb2 * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}

__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, float alpha, size_t n) {
  vst::pdl_grid_sync();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}

struct LossTermsParams {
  int num_idx[16], den_idx[16], group[16];
  float coef[16], den_eps[16];
  int n_entries, n_groups;
};
__global__ void loss_terms_kernel(const float* __restrict__ sums, LossTermsParams p, float* __restrict__ terms,
                                  float* __restrict__ scale_out) {
  vst::pdl_grid_sync();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float acc[17];
  for (int g = 0; g <= p.n_groups; ++g) acc[g] = 0.f;
  float bad = 0.f;
  for (int i = 0; i < p.n_entries; ++i) {
    const float den = p.den_idx[i] < 0 ? 1.f : sums[p.den_idx[i]] + p.den_eps[i];
    const bool zero = den == 0.f;       // strict count (den_eps == 0) of an empty mask: no inf / NaN may reach the sweep
    if (zero) bad = 1.f;
    const float sc = zero ? 0.f : p.coef[i] / den;
    const float v = sums[p.num_idx[i]] * sc;
    scale_out[i] = sc;
    acc[p.group[i]] += v;
  }
  float total = 0.f;
  for (int g = 0; g < p.n_groups; ++g) { terms[g] = acc[g]; total += acc[g]; }
  terms[p.n_groups] = total;
  terms[p.n_groups + 1] = bad;
}

}  // namespace vst

using namespace vst;

extern "C" {

int vst_weight_flip_transpose_f32(const float* w, float* wt, int Cout, int Cin, int k, void* stream) {
  VST_CHECK_ARG(Cout > 0 && Cin > 0 && k > 0, "weight_flip_transpose: bad shape");
  VST_DEVPTR(w); VST_DEVPTR(wt);
  vst::launch(weight_flip_transpose_kernel, bw_grid((size_t)Cout * Cin * k * k), 256, 0, (cudaStream_t)stream, w, wt, Cout, Cin, k);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_conv_transpose_gather_f32(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int H, int W,
                                  int Cout, int Ho, int Wo, int k, int stride, int pad, void* stream) {
  VST_CHECK_ARG(N > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0 && Ho > 0 && Wo > 0 && k > 0 && stride > 0, "conv_transpose_gather: bad shape");
  VST_DEVPTR(x); VST_DEVPTR(w); VST_DEVPTR(y);
  const size_t total = (size_t)N * cdiv(Cout, 8) * Ho * Wo;
  vst::launch(conv_transpose_gather_kernel, bw_grid(total), 256, 0, (cudaStream_t)stream, x, w, bias, y, N, Cin, H, W, Cout, Ho, Wo, k,
                                                                                 stride, pad);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_fold_pad_f32(const float* dxp, float* dx, int NC, int Hs, int Ws, int ups, int pad, int pad_mode, int Hp, int Wp,
                     void* stream) {
  VST_CHECK_ARG(NC > 0 && Hs > 0 && Ws > 0 && (ups == 1 || ups == 2) && pad >= 0, "fold_pad: bad shape");
  VST_CHECK_ARG(pad_mode != VST_PAD_REFLECT || (2 * pad < Hs * ups && 2 * pad < Ws * ups), "fold_pad: reflect pad too large for the tensor");
  VST_DEVPTR(dxp); VST_DEVPTR(dx);
  vst::launch(fold_pad_kernel, bw_grid((size_t)NC * Hs * Ws), 256, 0, (cudaStream_t)stream, dxp, dx, NC, Hs, Ws, ups, pad, pad_mode, Hp, Wp);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, int N, int Cin, int H, int W, int Cout, int k, int stride,
                         int pad, int pad_mode, int ups, void* stream) {
  VST_CHECK_ARG(N > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "conv2d_wgrad: empty shape");
  VST_CHECK_ARG(ups == 1 || ups == 2, "conv2d_wgrad: ups must be 1 or 2");
  VST_DEVPTR(x); VST_DEVPTR(dy); VST_DEVPTR(dw);
  cudaStream_t st = (cudaStream_t)stream;
  const int Hl = H * ups, Wl = W * ups;
  const int Ho = (Hl + 2 * pad - k) / stride + 1, Wo = (Wl + 2 * pad - k) / stride + 1;
  VST_CUDA(cudaMemsetAsync(dw, 0, (size_t)Cout * Cin * k * k * sizeof(float), st));
  const int tiles_x = cdiv(Wo, wg_tw(stride)), tiles_y = cdiv(Ho, WG_TH);
  const int total_tiles = N * tiles_x * tiles_y;
  const int blocks_base = cdiv(Cout, WG_C) * cdiv(Cin, WG_C) * k;
  int splits = cdiv(kNumSMs * 4, blocks_base);
  if (splits > total_tiles) splits = total_tiles;
  if (splits < 1) splits = 1;
  dim3 grid(cdiv(Cout, WG_C), cdiv(Cin, WG_C), k * splits);
#define LAUNCH(K, S) \
  vst::launch(conv2d_wgrad_kernel<K, S>, grid, 256, 0, st, x, dy, dw, N, Cin, H, W, Cout, Ho, Wo, pad, pad_mode, ups, tiles_x, tiles_y, splits)
  if (k == 3 && stride == 1) LAUNCH(3, 1);
  else if (k == 3 && stride == 2) LAUNCH(3, 2);
  else if (k == 9 && stride == 1) LAUNCH(9, 1);
  else {
    set_error("conv2d_wgrad: k=%d stride=%d not implemented", k, stride);
    return VST_EUNSUPPORTED;
  }
#undef LAUNCH
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_channel_sum_f32(const float* x, float* out, int N, int C, int HW, void* stream) {
  VST_CHECK_ARG(N > 0 && C > 0 && HW > 0, "channel_sum: empty shape");
  VST_DEVPTR(x); VST_DEVPTR(out);
  cudaStream_t st = (cudaStream_t)stream;
  VST_CUDA(cudaMemsetAsync(out, 0, C * sizeof(float), st));
  int splits = cdiv(kNumSMs * 4, C);
  const int max_splits = cdiv((int)std::min<size_t>((size_t)N * HW, (size_t)1 << 30), 256);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  vst::launch(channel_sum_kernel, dim3(C, splits), 256, 0, st, x, out, N, C, HW);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_act_bwd_f32(const float* dy, const float* y, float* dz, size_t n, int act, void* stream) {
  VST_CHECK_ARG(n > 0, "act_bwd: empty");
  VST_DEVPTR(dy); VST_DEVPTR(y); VST_DEVPTR(dz);
  vst::launch(act_bwd_kernel, bw_grid(n), 256, 0, (cudaStream_t)stream, dy, y, dz, n, act);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_instance_norm_bwd_f32(const float* x, const float* dy, const float* gamma, const float* beta, const float* mean,
                              const float* rstd, float* dx, float* dgamma, float* dbeta, int N, int C, int HW, int act,
                              void* stream) {
  VST_CHECK_ARG(N > 0 && C > 0 && HW > 0, "instance_norm_bwd: empty shape");
  VST_DEVPTR(x); VST_DEVPTR(dy); VST_DEVPTR(gamma); VST_DEVPTR(beta); VST_DEVPTR(mean); VST_DEVPTR(rstd); VST_DEVPTR(dx);
  VST_DEVPTR(dgamma); VST_DEVPTR(dbeta);
  cudaStream_t st = (cudaStream_t)stream;
  VST_CUDA(cudaMemsetAsync(dgamma, 0, C * sizeof(float), st));
  VST_CUDA(cudaMemsetAsync(dbeta, 0, C * sizeof(float), st));
  const int threads = HW >= 4096 ? 1024 : (HW >= 512 ? 256 : 64);
  const int nb = plane_cluster_size(N * C, HW);
  VST_CUDA(launch_cluster(instance_norm_bwd_kernel, N * C * nb, nb, threads, st, x, dy, gamma, beta, mean, rstd, dx, dgamma, dbeta, C, HW, act));
  return VST_OK;
}

int vst_maxpool2_bwd_f32(const float* x, const float* dy, float* dx, int NC, int H, int W, void* stream) {
  VST_CHECK_ARG(NC > 0 && H >= 2 && W >= 2, "maxpool2_bwd: bad shape");
  VST_DEVPTR(x); VST_DEVPTR(dy); VST_DEVPTR(dx);
  vst::launch(maxpool2_bwd_kernel, bw_grid((size_t)NC * H * W), 256, 0, (cudaStream_t)stream, x, dy, dx, NC, H, W);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_vgg_normalize_bwd_f32(const float* dy, float* dx, int N, int HW, void* stream) {
  VST_CHECK_ARG(N > 0 && HW > 0, "vgg_normalize_bwd: empty shape");
  VST_DEVPTR(dy); VST_DEVPTR(dx);
  vst::launch(channel_scale_kernel, bw_grid((size_t)N * 3 * HW), 256, 0, (cudaStream_t)stream, dy, dx, (size_t)N * 3 * HW, 3, HW, 1.f / (255.f * 0.229f), 1.f / (255.f * 0.224f), 1.f / (255.f * 0.225f));
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_warp_bwd_f32(const float* dy, const float* flo, float* dx, int B, int C, int H, int W, void* stream) {
  VST_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && (size_t)B * H * W < ((size_t)1 << 32), "warp_bwd: empty shape");
  VST_DEVPTR(dy); VST_DEVPTR(flo); VST_DEVPTR(dx);
  cudaStream_t st = (cudaStream_t)stream;
  VST_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * C * H * W * sizeof(float), st));
  vst::launch(warp_bwd_kernel, bw_grid((size_t)B * H * W), 256, 0, st, dy, flo, dx, B, C, H, W);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_feature_temporal_bwd_f32(const float* f1, const float* f2, const float* flow, const float* mask, float scale,
                                 const float* scale_dev, float* df1, float* df2, int B, int C, int Hf, int Wf, int H, int W, void* stream) {
  VST_CHECK_ARG(B > 0 && C > 0 && Hf > 0 && Wf > 0 && H > 0 && W > 0 && (size_t)B * H * W < ((size_t)1 << 32), "feature_temporal_bwd: empty shape");
  VST_DEVPTR(f1); VST_DEVPTR(f2); VST_DEVPTR(flow); VST_DEVPTR(mask); VST_DEVPTR(df1); VST_DEVPTR(df2);
  cudaStream_t st = (cudaStream_t)stream;
  VST_CUDA(cudaMemsetAsync(df1, 0, (size_t)B * C * Hf * Wf * sizeof(float), st));
  vst::launch(feature_temporal_bwd_kernel, dim3(bw_grid((size_t)B * Hf * Wf), cdiv(C, FTB_CPT)), 256, 0, st, f1, f2, flow, mask, scale_dev, scale,
                                                                                                    df1, df2, B, C, Hf, Wf, H, W);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_output_temporal_bwd_f32(const float* s1, const float* s2, const float* i1, const float* i2, const float* flow,
                                const float* mask, float scale, const float* scale_dev, float* ds1, float* ds2, int B, int H, int W,
                                int luminance, void* stream) {
  VST_CHECK_ARG(B > 0 && H > 0 && W > 0 && (size_t)B * H * W < ((size_t)1 << 32), "output_temporal_bwd: empty shape");
  VST_DEVPTR(s1); VST_DEVPTR(s2); VST_DEVPTR(flow); VST_DEVPTR(mask); VST_DEVPTR(ds1); VST_DEVPTR(ds2);
  if (luminance) { VST_DEVPTR(i1); VST_DEVPTR(i2); }
  cudaStream_t st = (cudaStream_t)stream;
  VST_CUDA(cudaMemsetAsync(ds1, 0, (size_t)B * 3 * H * W * sizeof(float), st));
  vst::launch(output_temporal_bwd_kernel, bw_grid((size_t)B * H * W), 256, 0, st, s1, s2, i1, i2, flow, mask, scale, scale_dev, ds1, ds2, B, H, W,
                                                                         luminance);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_sqdiff_bwd_f32(const float* a, const float* b, float scale, float* da, float* db, size_t n, void* stream) {
  VST_CHECK_ARG(n > 0, "sqdiff_bwd: empty");
  VST_DEVPTR(a); VST_DEVPTR(b); VST_DEVPTR(da);
  vst::launch(sqdiff_bwd_kernel, bw_grid(n), 256, 0, (cudaStream_t)stream, a, b, scale, da, db, n);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_tv_bwd_f32(const float* x, float scale, float* dx, int BC, int H, int W, int mode, void* stream) {
  VST_CHECK_ARG(BC > 0 && H > 1 && W > 1, "tv_bwd: bad shape");
  VST_DEVPTR(x); VST_DEVPTR(dx);
  vst::launch(tv_bwd_kernel, bw_grid((size_t)BC * H * W), 256, 0, (cudaStream_t)stream, x, scale, dx, BC, H, W, mode);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_gram_bwd_f32(const float* y, const float* dG, float* dy, int B, int C, int HW, float scale, void* stream) {
  VST_CHECK_ARG(B > 0 && C > 0 && HW > 0, "gram_bwd: empty shape");
  VST_DEVPTR(y); VST_DEVPTR(dG); VST_DEVPTR(dy);
  dim3 grid(cdiv(HW, GB_TP), cdiv(C, GB_TI), B);
  vst::launch(gram_bwd_kernel, grid, 256, 0, (cudaStream_t)stream, y, dG, dy, C, HW, scale);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_axpy_f32(const float* x, float* y, float alpha, size_t n, void* stream) {
  VST_CHECK_ARG(n > 0, "axpy: empty");
  VST_DEVPTR(x); VST_DEVPTR(y);
  vst::launch(axpy_kernel, bw_grid(n), 256, 0, (cudaStream_t)stream, x, y, alpha, n);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_loss_terms_f32(const float* sums, const int* num_idx_host, const int* den_idx_host, const float* coef_host,
                       const float* den_eps_host, const int* group_host, int n_entries, int n_groups, float* terms_out,
                       float* scale_out, void* stream) {
  VST_CHECK_ARG(n_entries > 0 && n_entries <= 16 && n_groups > 0 && n_groups <= 16, "loss_terms: 1..16 entries/groups");
  VST_DEVPTR(sums); VST_DEVPTR(terms_out); VST_DEVPTR(scale_out);
  LossTermsParams p;
  for (int i = 0; i < n_entries; ++i) {
    VST_CHECK_ARG(group_host[i] >= 0 && group_host[i] < n_groups && num_idx_host[i] >= 0, "loss_terms: bad index");
    p.num_idx[i] = num_idx_host[i]; p.den_idx[i] = den_idx_host[i]; p.group[i] = group_host[i];
    p.coef[i] = coef_host[i]; p.den_eps[i] = den_eps_host[i];
  }
  p.n_entries = n_entries; p.n_groups = n_groups;
  vst::launch(loss_terms_kernel, 1, 32, 0, (cudaStream_t)stream, sums, p, terms_out, scale_out);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_adam_f32(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2, float eps, int step,
                 float grad_scale, const float* skip_flag, void* stream) {
  VST_CHECK_ARG(n > 0 && step >= 1, "adam: bad arguments");
  VST_DEVPTR(p); VST_DEVPTR(g); VST_DEVPTR(m); VST_DEVPTR(v);
  // scalar prefactors in double like torch.optim.Adam's Python floats (b1, b2 arrive as the nearest floats of 0.9 / 0.999)
  const double b1d = b1 == 0.9f ? 0.9 : (double)b1, b2d = b2 == 0.999f ? 0.999 : (double)b2;
  const double bc1 = 1.0 - pow(b1d, (double)step), bc2 = 1.0 - pow(b2d, (double)step);
  vst::launch(adam_kernel, bw_grid(n), 256, 0, (cudaStream_t)stream, p, g, m, v, n, (float)((double)lr / bc1), (float)b1d, (float)b2d, (float)(1.0 - b1d),
                                                            (float)(1.0 - b2d), eps, (float)sqrt(bc2), grad_scale, skip_flag);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

}  // extern "C"
