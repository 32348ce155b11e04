// Channels-last bf16 activation layouts shared by the inference engine and the training primitives.
#pragma once
#include "common.cuh"

namespace vst {

enum PadKind : int { PADK_REFLECT = 0, PADK_REPLICATE = 1, PADK_ZERO = 2 };

__device__ __forceinline__ int map_pad(int i, int n, int kind, bool& ok) {
  ok = true;
  if (i >= 0 && i < n) return i;
  if (kind == PADK_REFLECT) return reflect_idx(i, n);
  if (kind == PADK_REPLICATE) return i < 0 ? 0 : n - 1;
  ok = false;
  return 0;
}

struct ActLayout {
  int H, W, C;     // interior size, channels (C % 8 == 0)
  int pad, kind;   // halo and how it is filled
  int parity;      // 1: stored as 4 parity planes of the padded tensor
};

__host__ __device__ inline size_t act_elems(const ActLayout& L, int N) {
  return (size_t)N * (L.H + 2 * L.pad) * (L.W + 2 * L.pad) * L.C;
}
// element offset of padded pixel (n, yp, xp)
__host__ __device__ inline size_t act_offset(const ActLayout& L, int N, int n, int yp, int xp) {
  const int Hp = L.H + 2 * L.pad, Wp = L.W + 2 * L.pad;
  if (L.parity) {
    const int pl = (yp & 1) * 2 + (xp & 1), H2 = Hp / 2, W2 = Wp / 2;
    return ((((size_t)pl * N + n) * H2 + (yp >> 1)) * W2 + (xp >> 1)) * L.C;
  }
  return (((size_t)n * Hp + yp) * Wp + xp) * L.C;
}

}  // namespace vst
