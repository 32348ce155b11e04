// fp32 reference-semantics kernels (CUDA cores): direct convolution with fused reflect/zero
// padding and nearest-x2 read, ConvTranspose2d(k3,s2,p1,op1), InstanceNorm(+act,+residual),
// max-pool, vgg_normalize.  These follow the reference's maths op for op; they are the fp32
// path of BASELINE.json (<=1e-4 rel-L2) and the building blocks of the training step.
#include <stdarg.h>
#include <stdlib.h>
#include <mutex>
#include "common.cuh"

namespace vst {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("VST_PDL"); return e ? atoi(e) != 0 : true; }();
  return on;
}

int require_device_ptr(const void* p, const char* name) {
  if (p == nullptr) {
    set_error("%s is NULL", name);
    return VST_EINVAL;
  }
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("%s: cudaPointerGetAttributes -> %s", name, cudaGetErrorString(e));
    return VST_EDEVICE;
  }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) {
    set_error("%s is not device memory (this library has no CPU fallback)", name);
    return VST_EDEVICE;
  }
  return VST_OK;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case VST_ACT_RELU: return fmaxf(v, 0.f);
    case VST_ACT_TANH: return tanhf(v);
    case VST_ACT_RECONET_OUT: return tanhf(v / 255.f) * 150.f + 127.5f;
    case VST_ACT_RT_OUT: return (tanhf(v) + 1.f) / 2.f * 255.f;
    default: return v;
  }
}

// ------------------------------------------------------------------------------------------
// Direct convolution.  Block = 256 threads = 32x8 output pixels, 16 output channels each.
// Input patches (with padding resolved at load time) and the weight slab go through smem.
// ------------------------------------------------------------------------------------------
constexpr int CV_TW = 32, CV_TH = 8, CV_CO = 16;

template <int K, int S, int CI>
__global__ void __launch_bounds__(256) conv2d_f32_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    float* __restrict__ y, int Cin, int Hs, int Ws, int Cout, int Ho, int Wo, int pad, int pad_mode,
    int ups, int act) {
  vst::pdl_grid_sync();
  constexpr int PH = (CV_TH - 1) * S + K, PW = (CV_TW - 1) * S + K;
  __shared__ float s_in[CI][PH][PW + 1];
  __shared__ __align__(16) float s_w[CI][K * K][CV_CO];

  const int co_tiles = cdiv(Cout, CV_CO);
  const int n = blockIdx.z / co_tiles, co0 = (blockIdx.z % co_tiles) * CV_CO;
  const int ox0 = blockIdx.x * CV_TW, oy0 = blockIdx.y * CV_TH;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int H = Hs * ups, W = Ws * ups;  // logical (post-upsample) input size
  const int iy0 = oy0 * S - pad, ix0 = ox0 * S - pad;

  // Two-level accumulation: the taps of one channel group go into `part`, which joins `acc` through a compensated (Kahan)
  // add.  A single running sum over Cin*K*K <= 3888 products drifts ~3x further from the exact result than the blocked sums
  // of the reference's CPU convolution; on trained checkpoints (weights up to 9.5, InstanceNorm gains up to 4) that drift,
  // not the algorithm, was the whole fp32 error (2.4e-4 centred on the SD2 checkpoint against the 1e-4 bar).
  float acc[CV_CO], comp[CV_CO], part[CV_CO];
#pragma unroll
  for (int i = 0; i < CV_CO; ++i) acc[i] = comp[i] = part[i] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CI) {
    for (int idx = threadIdx.x; idx < CI * PH * PW; idx += 256) {
      const int c = idx / (PH * PW), r = (idx / PW) % PH, q = idx % PW;
      int iy = iy0 + r, ix = ix0 + q;
      float v = 0.f;
      if (c0 + c < Cin) {
        bool ok = true;
        if (pad_mode == VST_PAD_REFLECT) {
          // tiles past the image edge may overshoot by more than the pad: clamp after reflecting
          iy = reflect_idx(iy, H);
          ix = reflect_idx(ix, W);
          ok = (iy >= 0 && iy < H && ix >= 0 && ix < W);
        } else {
          ok = (iy >= 0 && iy < H && ix >= 0 && ix < W);
        }
        if (ok) v = x[(((size_t)n * Cin + c0 + c) * Hs + iy / ups) * Ws + ix / ups];
      }
      s_in[c][r][q] = v;
    }
    for (int idx = threadIdx.x; idx < CI * K * K * CV_CO; idx += 256) {
      const int co = idx % CV_CO, t = (idx / CV_CO) % (K * K), c = idx / (CV_CO * K * K);
      float v = 0.f;
      if (c0 + c < Cin && co0 + co < Cout) v = w[(((size_t)(co0 + co)) * Cin + c0 + c) * (K * K) + t];
      s_w[c][t][co] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CI; ++c) {
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const float v = s_in[c][ty * S + ky][tx * S + kx];
          const float4* wp = reinterpret_cast<const float4*>(&s_w[c][ky * K + kx][0]);
#pragma unroll
          for (int j = 0; j < CV_CO / 4; ++j) {
            const float4 ww = wp[j];
            part[4 * j + 0] = fmaf(v, ww.x, part[4 * j + 0]);
            part[4 * j + 1] = fmaf(v, ww.y, part[4 * j + 1]);
            part[4 * j + 2] = fmaf(v, ww.z, part[4 * j + 2]);
            part[4 * j + 3] = fmaf(v, ww.w, part[4 * j + 3]);
          }
        }
      }
      if (K >= 5 || c == CI - 1) {   // 9x9: one channel (81 products) per partial sum; 3x3 / 1x1: the CI-channel group
#pragma unroll
        for (int i = 0; i < CV_CO; ++i) {
          const float yk = __fsub_rn(part[i], comp[i]);
          const float tk = __fadd_rn(acc[i], yk);
          comp[i] = __fsub_rn(__fsub_rn(tk, acc[i]), yk);
          acc[i] = tk;
          part[i] = 0.f;
        }
      }
    }
    __syncthreads();
  }
  const int ox = ox0 + tx, oy = oy0 + ty;
  if (ox < Wo && oy < Ho) {
#pragma unroll
    for (int co = 0; co < CV_CO; ++co) {
      if (co0 + co < Cout) {
        float v = acc[co] + (bias ? bias[co0 + co] : 0.f);
        y[(((size_t)n * Cout + co0 + co) * Ho + oy) * Wo + ox] = apply_act(v, act);
      }
    }
  }
}

// ConvTranspose2d(k=3, s=2, p=1, op=1) as a gather: out[oy][ox] += in[(oy+1-ky)/2][(ox+1-kx)/2] * w[ci][co][ky][kx]
// for the (ky,kx) that make both numerators even and in range.  One thread per output pixel x 8 couts.
__global__ void __launch_bounds__(256) conv_transpose2d_f32_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    float* __restrict__ y, int N, int Cin, int H, int W, int Cout) {
  vst::pdl_grid_sync();
  const int Ho = 2 * H, Wo = 2 * W;
  const int co_tiles = cdiv(Cout, 8);
  const size_t total = (size_t)N * co_tiles * Ho * Wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = i % Wo, oy = (i / Wo) % Ho;
    const int ct = (i / ((size_t)Wo * Ho)) % co_tiles, n = i / ((size_t)Wo * Ho * co_tiles);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int ny = oy + 1 - ky;
      if (ny < 0 || (ny & 1)) continue;
      const int iy = ny >> 1;
      if (iy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int nx = ox + 1 - kx;
        if (nx < 0 || (nx & 1)) continue;
        const int ix = nx >> 1;
        if (ix >= W) continue;
        for (int ci = 0; ci < Cin; ++ci) {
          const float v = x[(((size_t)n * Cin + ci) * H + iy) * W + ix];
          const float* wp = w + (((size_t)ci * Cout + ct * 8) * 3 + ky) * 3 + kx;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (ct * 8 + j < Cout) acc[j] = fmaf(v, wp[(size_t)j * 9], acc[j]);
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

// InstanceNorm: one block - or one cluster of blocks (common.cuh) - per (n,c) plane, two-pass statistics (mean, then centred
// variance); a CTA of the cluster owns a contiguous slice of the plane, four loads per thread in flight.
__global__ void __launch_bounds__(1024) instance_norm_f32_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
    const float* __restrict__ residual, float* __restrict__ y, float* __restrict__ mean_out,
    float* __restrict__ rstd_out, int C, int HW, float eps, int act) {
  vst::pdl_grid_sync();
  __shared__ float red[32];
  __shared__ float slots[2];
  const uint32_t nb = cluster_nctarank(), rank = cluster_ctarank_();
  const int plane = blockIdx.x / nb, c = plane % C;
  const int chunk = (((HW + (int)nb - 1) / (int)nb) + 3) & ~3;
  const int i0 = min(HW, (int)rank * chunk), i1 = min(HW, i0 + chunk);
  const int B = blockDim.x;
  const float* xp = x + (size_t)plane * HW;
  float s = 0.f;
  {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = i0 + threadIdx.x;
    for (; i + 3 * B < i1; i += 4 * B) { s0 += xp[i]; s1 += xp[i + B]; s2 += xp[i + 2 * B]; s3 += xp[i + 3 * B]; }
    for (; i < i1; i += B) s0 += xp[i];
    s = (s0 + s1) + (s2 + s3);
  }
  const float mean = cluster_sum(s, red, &slots[0], nb) / (float)HW;
  float q = 0.f;
  {
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
    int i = i0 + threadIdx.x;
    for (; i + 3 * B < i1; i += 4 * B) {
      const float d0 = xp[i] - mean, d1 = xp[i + B] - mean, d2 = xp[i + 2 * B] - mean, d3 = xp[i + 3 * B] - mean;
      q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
    }
    for (; i < i1; i += B) { const float d = xp[i] - mean; q0 = fmaf(d, d, q0); }
    q = (q0 + q1) + (q2 + q3);
  }
  const float var = cluster_sum(q, red, &slots[1], nb) / (float)HW;
  if (nb > 1) cluster_barrier();   // nobody leaves (or reuses its slots) while a peer may still read them
  const float rstd = rsqrtf(var + eps);
  if (threadIdx.x == 0 && rank == 0) {
    if (mean_out) mean_out[plane] = mean;
    if (rstd_out) rstd_out[plane] = rstd;
  }
  const float g = gamma[c], b = beta[c];
  float* yp = y + (size_t)plane * HW;
  const float* rp = residual ? residual + (size_t)plane * HW : nullptr;
  for (int i = i0 + threadIdx.x; i < i1; i += B) {
    float v = apply_act((xp[i] - mean) * rstd * g + b, act);
    if (rp) v += rp[i];
    yp[i] = v;
  }
}

__global__ void maxpool2_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int NC, int H, int W) {
  vst::pdl_grid_sync();
  const int Ho = H / 2, Wo = W / 2;
  const size_t total = (size_t)NC * Ho * Wo;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = i % Wo, oy = (i / Wo) % Ho;
    const size_t p = i / ((size_t)Wo * Ho);
    const float* s = x + (p * H + 2 * oy) * W + 2 * ox;
    y[i] = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[W], s[W + 1]));
  }
}

__global__ void vgg_normalize_f32_kernel(float* __restrict__ x, float* __restrict__ y, int N, int HW, int inplace_div) {
  vst::pdl_grid_sync();
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  const size_t total = (size_t)N * 3 * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (i / HW) % 3;
    const float d = __fdiv_rn(x[i], 255.0f);
    if (inplace_div) x[i] = d;
    y[i] = __fdiv_rn(__fsub_rn(d, mean[c]), stdv[c]);
  }
}

// Byte frame of the `Inference` iterators (RC/utilities.py:219-224, RT/utilities.py:318-326): clamp(0, 255), HWC, RGB -> BGR,
// astype(uint8) truncation.  One thread per pixel: three coalesced plane reads, one 3-byte write.
__global__ void __launch_bounds__(256) pack_bgr_u8_kernel(const float* __restrict__ img, uint8_t* __restrict__ out, int N, int HW) {
  vst::pdl_grid_sync();
  const size_t total = (size_t)N * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, p = i - n * HW;
    const float* src = img + n * 3 * (size_t)HW + p;
    uint8_t* dst = out + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[2 - c] = (uint8_t)fminf(fmaxf(src[(size_t)c * HW], 0.f), 255.f);
  }
}

static inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = (size_t)kNumSMs * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace vst

using namespace vst;

extern "C" {

int vst_abi_version(void) { return VST_ABI_VERSION; }
const char* vst_last_error(void) { return vst::g_err; }

int vst_device_arch(void) {
  int dev = 0, maj = 0, mnr = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, dev);
  return maj * 10 + mnr;
}

int vst_conv2d_f32(const float* x, const float* w, const float* bias, float* y, int N, int Cin, int H,
                   int W, int Cout, int k, int stride, int pad, int pad_mode, int ups, int act,
                   void* stream) {
  VST_CHECK_ARG(N > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "conv2d: empty shape");
  VST_CHECK_ARG(ups == 1 || ups == 2, "conv2d: ups must be 1 or 2");
  VST_CHECK_ARG(stride == 1 || stride == 2, "conv2d: stride must be 1 or 2");
  VST_CHECK_ARG(pad_mode == VST_PAD_ZERO || pad_mode == VST_PAD_REFLECT, "conv2d: bad pad_mode");
  VST_CHECK_ARG(pad_mode != VST_PAD_REFLECT || (pad < H * ups && pad < W * ups), "conv2d: reflect pad >= size");
  VST_DEVPTR(x); VST_DEVPTR(w); VST_DEVPTR(y);
  const int Hl = H * ups, Wl = W * ups;
  const int Ho = (Hl + 2 * pad - k) / stride + 1, Wo = (Wl + 2 * pad - k) / stride + 1;
  VST_CHECK_ARG(Ho > 0 && Wo > 0, "conv2d: empty output");
  dim3 grid(cdiv(Wo, CV_TW), cdiv(Ho, CV_TH), N * cdiv(Cout, CV_CO));
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(K, S, CI) \
  vst::launch(conv2d_f32_kernel<K, S, CI>, grid, 256, 0, st, x, w, bias, y, Cin, H, W, Cout, Ho, Wo, pad, pad_mode, ups, act)
  if (k == 3 && stride == 1) LAUNCH(3, 1, 8);
  else if (k == 3 && stride == 2) LAUNCH(3, 2, 8);
  else if (k == 9 && stride == 1) LAUNCH(9, 1, 4);
  else if (k == 1 && stride == 1) LAUNCH(1, 1, 8);
  else {
    set_error("conv2d: k=%d stride=%d not implemented (3/s1, 3/s2, 9/s1, 1/s1)", k, stride);
    return VST_EUNSUPPORTED;
  }
#undef LAUNCH
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_pack_bgr_u8(const float* img, unsigned char* out, int N, int H, int W, void* stream) {
  VST_CHECK_ARG(N > 0 && H > 0 && W > 0, "pack_bgr_u8: empty shape");
  VST_DEVPTR(img); VST_DEVPTR(out);
  vst::launch(pack_bgr_u8_kernel, grid_for((size_t)N * H * W, 256), 256, 0, (cudaStream_t)stream, img, out, N, H * W);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_conv_transpose2d_f32(const float* x, const float* w, const float* bias, float* y, int N, int Cin,
                             int H, int W, int Cout, void* stream) {
  VST_CHECK_ARG(N > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "conv_transpose2d: empty shape");
  VST_DEVPTR(x); VST_DEVPTR(w); VST_DEVPTR(y);
  const size_t total = (size_t)N * cdiv(Cout, 8) * 4 * H * W;
  vst::launch(conv_transpose2d_f32_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, x, w, bias, y, N, Cin, H, W, Cout);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_instance_norm_f32(const float* x, const float* gamma, const float* beta, const float* residual,
                          float* y, float* mean_out, float* rstd_out, int N, int C, int HW, float eps,
                          int act, void* stream) {
  VST_CHECK_ARG(N > 0 && C > 0 && HW > 0, "instance_norm: empty shape");
  VST_DEVPTR(x); VST_DEVPTR(gamma); VST_DEVPTR(beta); VST_DEVPTR(y);
  const int threads = HW >= 4096 ? 1024 : (HW >= 512 ? 256 : 64);
  const int nb = plane_cluster_size(N * C, HW);
  VST_CUDA(launch_cluster(instance_norm_f32_kernel, N * C * nb, nb, threads, (cudaStream_t)stream, x, gamma, beta, residual, y, mean_out,
                          rstd_out, C, HW, eps, act));
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_maxpool2_f32(const float* x, float* y, int NC, int H, int W, void* stream) {
  VST_CHECK_ARG(NC > 0 && H >= 2 && W >= 2, "maxpool2: bad shape");
  VST_DEVPTR(x); VST_DEVPTR(y);
  const size_t total = (size_t)NC * (H / 2) * (W / 2);
  vst::launch(maxpool2_f32_kernel, grid_for(total, 256), 256, 0, (cudaStream_t)stream, x, y, NC, H, W);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

int vst_vgg_normalize_f32(float* x, float* y, int N, int HW, int inplace_div, void* stream) {
  VST_CHECK_ARG(N > 0 && HW > 0, "vgg_normalize: empty shape");
  VST_DEVPTR(x); VST_DEVPTR(y);
  vst::launch(vgg_normalize_f32_kernel, grid_for((size_t)N * 3 * HW, 256), 256, 0, (cudaStream_t)stream, x, y, N, HW, inplace_div);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

}  // extern "C"
