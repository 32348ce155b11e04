// Shared helpers for the vst_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/vst_b200.h"

namespace vst {

void set_error(const char* fmt, ...);

#define VST_CHECK_ARG(cond, ...)                \
  do {                                          \
    if (!(cond)) {                              \
      vst::set_error(__VA_ARGS__);              \
      return VST_EINVAL;                        \
    }                                           \
  } while (0)

#define VST_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      vst::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return VST_ECUDA;                                                            \
    }                                                                              \
  } while (0)

// Launch epilogue: surface launch-configuration errors without synchronising.
#define VST_LAUNCH_CHECK()                                                         \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      vst::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return VST_ECUDA;                                                            \
    }                                                                              \
  } while (0)

// No CPU fallback: host pointers are rejected.
int require_device_ptr(const void* p, const char* name);
#define VST_DEVPTR(p)                                   \
  do {                                                  \
    int r__ = vst::require_device_ptr((p), #p);         \
    if (r__ != VST_OK) return r__;                      \
  } while (0)

constexpr int kNumSMs = 148;

// ---- programmatic dependent launch -----------------------------------------------------------------------------------
// Every kernel of the library is launched through vst::launch() with the programmatic-stream-serialisation attribute and
// executes griddepcontrol.wait before its first global-memory access: the next kernel of a stream (or of a captured graph)
// is scheduled while its predecessor drains, its CTAs become resident as the predecessor's exit and run their prologue
// (barrier init, TMEM allocation, descriptor prefetch), and only then block until the predecessor's memory is visible.
// Because EVERY attributed kernel waits, completion is transitive along the stream (B done => B passed its wait => A done).
// launch_dependents is issued right after the wait: the dependent grid is released once every CTA of this grid has started,
// i.e. never before the last wave.  VST_PDL=0 launches without the attribute (the device instructions are then no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef VST_PDL_EARLY
#define VST_PDL_EARLY 0
#endif
__device__ __forceinline__ void pdl_trigger() {
#if VST_PDL_EARLY
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_grid_sync() { pdl_wait(); pdl_trigger(); }

bool pdl_enabled();   // api.cu-level switch, read once (VST_PDL, default on)

template <typename... KArgs, typename... Args>
inline cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ReflectionPad2d index: -i -> i, n-1+i -> n-1-i (edge not repeated), valid for |overshoot| < n.
__host__ __device__ inline int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in every thread. `red` = 32 floats of shared memory.
__device__ inline float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// ---- few large planes: a thread-block CLUSTER per (n, c) plane --------------------------------------------------------------
// The fp32 InstanceNorm kernels own one plane per block; a 3-channel 640x360 x 8 output layer is then 24 blocks on 148 SMs
// (172 us for 22 MB).  With a cluster of up to 8 CTAs per plane every CTA takes a slice of the plane and the block totals meet
// through distributed shared memory, in rank order (deterministic).  Launched without the attribute the cluster is 1 x 1 x 1.
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_ctarank_() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Sum of `v` over all threads of all CTAs of the cluster; `red` = 32 floats, `slot` = one float of shared memory that is not
// reused before the NEXT cluster_barrier of the caller (each reduction of a kernel takes its own slot).
__device__ inline float cluster_sum(float v, float* red, float* slot, uint32_t nb) {
  v = block_sum(v, red);
  if (nb == 1) return v;
  if (threadIdx.x == 0) *slot = v;
  cluster_barrier();
  float t = 0.f;
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(slot);
  for (uint32_t r = 0; r < nb; ++r) {
    uint32_t ra;
    float x;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(sa), "r"(r));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(x) : "r"(ra) : "memory");
    t += x;
  }
  return t;
}
inline int plane_cluster_size(int planes, int HW) {
  if (HW < 32768) return 1;
  int nb = 1;
  while (nb < 8 && planes * nb < 2 * kNumSMs) nb <<= 1;
  return nb;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster(void (*kern)(KArgs...), int grid, int cluster, int block, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = cluster; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = cluster > 1 ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// The reference's sampling coordinate for `warp` (RC/utilities.py:50-54 + ATen
// grid_sampler_unnormalize, align_corners=False), each step rounded to fp32, no FMA contraction:
//   v = p + f;  n = 2*v/max(S-1,1) - 1;  i = ((n + 1)*S - 1)/2
__device__ __forceinline__ float warp_src_coord(int p, float f, int S) {
  const float v = __fadd_rn((float)p, f);
  const float d = (float)(S > 1 ? S - 1 : 1);
  const float n = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, v), d), 1.0f);
  return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(n, 1.0f), (float)S), 1.0f), 2.0f);
}

}  // namespace vst
