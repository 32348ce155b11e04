// Micro-benchmark: how fast can one SM pull operand tiles L2 -> smem?
//   mode 0: TMA 2-D tensor boxes (64 bf16 = 128 B rows, SWIZZLE_128B), R rows per box
//   mode 1: 1-D bulk copies (cp.async.bulk) of the same byte count
//   mode 2: TMA 2-D boxes with 64 B rows (SWIZZLE_64B)
// 148 CTAs, one issuing thread each, ring of S buffers; source is small (L2 resident).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_rate tma_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t par) {
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(s32(b)), "r"(par) : "memory");
}

struct Params { CUtensorMap tm; const uint8_t* src; int mode, rows, iters, S, row_bytes; size_t src_bytes; };

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ Params p, unsigned long long* clk_out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  const int bytes = p.rows * p.row_bytes;
  const int slot = (bytes + 1023) & ~1023;
  uint64_t* bar = (uint64_t*)(sm + (size_t)p.S * slot);
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.S; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t0 = clock64();
    int s = 0; uint32_t ph = 0;
    const int nrow_total = (int)(p.src_bytes / p.row_bytes);
    for (int it = 0; it < p.iters + p.S; ++it) {
      if (it >= p.S) mwait(&bar[s], ph);          // data of iteration it-S landed
      if (it < p.iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(bytes) : "memory");
        const int r0 = ((it * 37 + blockIdx.x * 101) * p.rows) % (nrow_total - p.rows);
        if (p.mode == 1) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(s32(sm + (size_t)s * slot)), "l"(p.src + (size_t)r0 * p.row_bytes), "r"(bytes), "r"(s32(&bar[s])) : "memory");
        } else {
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(s32(sm + (size_t)s * slot)), "l"((uint64_t)&p.tm), "r"(s32(&bar[s])), "r"(0), "r"(r0) : "memory");
        }
      }
      if (++s == p.S) { s = 0; if (it >= p.S) ph ^= 1; }
    }
    clk_out[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  Enc enc = (Enc)f;
  const size_t src_bytes = 8u << 20;   // 8 MiB: L2 resident
  uint8_t* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  unsigned long long* clk; cudaMalloc(&clk, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 2000;
  printf("mode rows row_bytes S | clk/load  clk/row  B/clk/SM  chip TB/s@1.9GHz\n");
  for (int mode = 0; mode < 3; ++mode)
    for (int rows : {48, 96, 128, 192, 256})
      for (int S : {2, 4, 6}) {
        Params p{}; p.mode = mode; p.rows = rows; p.iters = iters; p.S = S; p.src = src; p.src_bytes = src_bytes;
        p.row_bytes = mode == 2 ? 64 : 128;
        const int elems = p.row_bytes / 2;
        if (mode != 1) {
          cuuint64_t dims[2] = {(cuuint64_t)elems, (cuuint64_t)(src_bytes / p.row_bytes)};
          cuuint64_t str[1] = {(cuuint64_t)p.row_bytes};
          cuuint32_t box[2] = {(cuuint32_t)elems, (cuuint32_t)rows}, es[2] = {1, 1};
          CUresult r = enc(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           mode == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        }
        const int slot = (rows * p.row_bytes + 1023) & ~1023;
        const size_t smem = (size_t)S * slot + 2048;
        if (smem > 220 * 1024) continue;
        k<<<148, 128, smem>>>(p, clk);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d rows %d S %d: %s\n", mode, rows, S, cudaGetErrorString(e)); return 1; }
        std::vector<unsigned long long> h(148);
        cudaMemcpy(h.data(), clk, 148 * 8, cudaMemcpyDeviceToHost);
        double mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        const double cpl = mx / iters, bpc = rows * p.row_bytes / cpl;
        printf("%d %4d %4d %d | %8.1f %7.2f %8.1f %8.2f\n", mode, rows, p.row_bytes, S, cpl, cpl / rows, bpc, bpc * 148 * 1.9e9 / 1e12);
      }
  return 0;
}
