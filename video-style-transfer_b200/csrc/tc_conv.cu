// tcgen05 / TMEM / TMA tap-GEMM convolution kernel (see tc_conv.cuh for the scheme).
//
// Warp roles (384 threads, 1 CTA per SM, persistent over tiles):
//   warp 0        A producer   - one elected lane issues the activation boxes (TMA)
//   warp 6        B producer   - one elected lane issues the weight boxes (TMA)
//   warp 1 (+11)  MMA issuer   - allocates TMEM, one elected lane issues tcgen05.mma + commits
//   warps 2..5    epilogue set 0, warps 7..10 epilogue set 1 (narrow layers; ping-pong tiles for the row convolution):
//                 tcgen05.ld the accumulators (lane group = warp_id % 4), convert, stage, store, fused IN statistics
// Pipelines: smem full/empty mbarriers between TMA and MMA (S stages), TMEM full/empty mbarriers between MMA and
// epilogue (2..16 accumulator stages, so the epilogue of tile i overlaps the MMAs of the following tiles).
// Main-loop modes (chosen per layer by tapgemm_plan): per-tap boxes, dy-sharing boxes, shared-memory row ring,
// accumulator-ring row streaming.
#include <stdio.h>
#include <stdlib.h>
#include <mutex>
#include <algorithm>
#include <vector>
#include "tc_conv.cuh"
#include "tc_ptx.cuh"

namespace vst {

int tg_smem_budget();

struct TileCoord {
  int ph, nt, n, y0, x0;
  int valid;   // 0: padding tile of a CTA pair (odd tile count) - operands are loaded, nothing is stored
};
__device__ __forceinline__ TileCoord decode_tile(const TapGemmParams& p, int tile) {
  // phase is the FASTEST index: the 4 output phases of an x2-upsample conv (and the 4 parity phases of a stride-2
  // data gradient) read the same input tile, so neighbouring CTAs share it through L2 instead of re-reading HBM
  TileCoord t;
  t.ph = tile % p.n_phase;
  tile /= p.n_phase;
  const int tx = tile % p.tiles_x;
  tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  tile /= p.tiles_y;
  t.n = tile % p.n_img;
  t.nt = tile / p.n_img;
  t.x0 = tx * p.tile_step_x;
  t.y0 = ty * p.TH;
  t.valid = 1;
  return t;
}

// Row-streaming mode (p.stream): the convolution's taps differ only by a row offset (dy = s_dy0 + t, same dx / plane) -
// conv1 over X9, the ConvTanh row convolution, its data gradient over E.  A CTA owns a 128-pixel column strip and a
// chunk of output rows; input rows enter a shared-memory RING once each (instead of once per tap) and tap t of
// output row y reads ring entry (y - r0) + t; the tap weights stay resident in shared memory for the whole kernel.
struct StreamUnit {
  int n, x0, r0, rows;
};
__device__ __forceinline__ int stream_units(const TapGemmParams& p) { return p.n_img * p.tiles_x * p.s_chunks; }
__device__ __forceinline__ StreamUnit decode_unit(const TapGemmParams& p, int u) {
  StreamUnit s;
  const int chunk = u % p.s_chunks;
  u /= p.s_chunks;
  s.x0 = (u % p.tiles_x) * p.tile_step_x;
  s.n = u / p.tiles_x;
  s.r0 = chunk * p.s_rpc;
  s.rows = min(p.s_rpc, p.Ho - s.r0);
  return s;
}

// Tile enumeration shared by the epilogue of both modes.
struct TileWalk {
  int tile, unit, j, rows;
  StreamUnit su;
};
__device__ __forceinline__ void walk_init(const TapGemmParams& p, TileWalk& w) {
  w.tile = blockIdx.x - gridDim.x;
  w.unit = blockIdx.x - gridDim.x;
  w.j = 0;
  w.rows = 0;
}
__device__ __forceinline__ bool walk_next(const TapGemmParams& p, TileWalk& w, TileCoord& tc, int total_tiles) {
  if (!p.stream) {
    w.tile += gridDim.x;
    if (p.cta2) {   // both CTAs of a pair run the same number of tiles; the odd one out is a padding tile
      if ((w.tile & ~1) >= total_tiles) return false;
      tc = decode_tile(p, min(w.tile, total_tiles - 1));
      tc.valid = w.tile < total_tiles;
      return true;
    }
    if (w.tile >= total_tiles) return false;
    tc = decode_tile(p, w.tile);
    return true;
  }
  if (++w.j >= w.rows) {
    do {
      w.unit += gridDim.x;
      if (w.unit >= stream_units(p)) return false;
      w.su = decode_unit(p, w.unit);
      w.rows = w.su.rows;
    } while (w.rows <= 0);
    w.j = 0;
  }
  tc.ph = 0; tc.nt = 0; tc.n = w.su.n; tc.x0 = w.su.x0; tc.y0 = w.su.r0 + w.j; tc.valid = 1;
  return true;
}

// 256-bit global store (STG.E.256 on sm_100): one full 32-byte sector per lane
__device__ __forceinline__ void stg256(void* ptr, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// InstanceNorm statistics of one 16-channel chunk (direct epilogue: a lane holds 16 channels of ONE pixel per sub-tile and
// has summed them / their squares over the tile's sub-tiles in registers, so the per-channel total runs ACROSS the warp's lanes).  Halving butterfly: at each step a lane
// keeps half of its channels and hands the other half to its partner, so 16 channels x 32 pixels collapse in 8+4+2+1
// shuffles; sums and squares each, plus one exchange that leaves the total SUM of channel (lane >> 1) in even lanes and its
// total SUM OF SQUARES in odd lanes - i.e. lane l holds stats[...][c0 + (l >> 1)][l & 1], 32 consecutive outputs per warp.
// The order of the additions is fixed, so the result is bit-reproducible.
__device__ __forceinline__ float chunk_stats2(const float (&sum)[16], const float (&sq)[16], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
  float res[2];
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    float a[8], b[4], c[2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float lo = pass ? sq[j] : sum[j], hi = pass ? sq[j + 8] : sum[j + 8];
      a[j] = (h16 ? hi : lo) + __shfl_xor_sync(0xffffffffu, h16 ? lo : hi, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = (h8 ? a[j + 4] : a[j]) + __shfl_xor_sync(0xffffffffu, h8 ? a[j] : a[j + 4], 8);
#pragma unroll
    for (int j = 0; j < 2; ++j) c[j] = (h4 ? b[j + 2] : b[j]) + __shfl_xor_sync(0xffffffffu, h4 ? b[j] : b[j + 2], 4);
    res[pass] = (h2 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, h2 ? c[0] : c[1], 2);
  }
  const float other = __shfl_xor_sync(0xffffffffu, h1 ? res[0] : res[1], 1);
  return (h1 ? res[1] : res[0]) + other;
}

// Fused input normalisation, one ring slot: y = max(a x + b, 0) in place on this thread's 16-byte chunk of 8 rows.  Kept as a
// ROLLED loop (two rows per trip for ILP): fully unrolled with both storage formats inlined it was ~700 SASS instructions per
// slot and the transform warps stalled on instruction fetch (ncu: stall_no_inst on every line) in a kernel whose other roles
// already fill the instruction cache.
template <bool HALF>
__device__ __noinline__ void xform_rows(uint32_t base, int tw, int rsub, int lc, const float (&ca)[8], const float (&cb)[8], int relu) {
  const float lo = relu ? 0.f : -3.0e38f;
#pragma unroll 4
  for (int it = 0; it < 8; ++it) {
    const int row = tw * 32 + it * 4 + rsub;
    const uint32_t addr = base + (uint32_t)row * 128u + (uint32_t)((lc ^ (row & 7)) << 4);
    uint4 q = lds128(addr);
    uint32_t* qw = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 f;
      if (HALF) f = __half22float2(*reinterpret_cast<const __half2*>(&qw[j]));
      else f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&qw[j]));
      const float a = fmaxf(fmaf(f.x, ca[2 * j], cb[2 * j]), lo), b = fmaxf(fmaf(f.y, ca[2 * j + 1], cb[2 * j + 1]), lo);
      if (HALF) {
        __half2 h = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));
        qw[j] = *reinterpret_cast<uint32_t*>(&h);
      } else {
        qw[j] = pack_bf16x2(a, b);
      }
    }
    sts128(addr, q);
  }
}

__device__ __forceinline__ void epi_bar_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }
__device__ __forceinline__ void epi_bar_sync_id(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---------------------------------------------------------------------------------------------------------------------
// TMA-store epilogue of the bf16-NHWC layers (p.epi_tma), run by the eight epilogue warps.  A separate, non-inlined function
// per mode (PP: ping-pong groups of the narrow layers) so that each mode gets its own register allocation under the kernel's
// 128-register cap.
template <bool CTA2, bool PP>
__device__ __forceinline__ void epilogue_tma(const TapGemmParams& p, uint32_t stg_s, uint64_t* tfull, uint64_t* tempty, uint32_t tmem_base) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int MT = p.MT;
  const int total_tiles = p.n_phase * p.n_ntile * p.n_img * p.tiles_y * p.tiles_x;
  const int acc_cols = MT * p.N_mma;
  const int AS = p.acc_stages, as_sh = 31 - __clz(AS);
  const uint32_t tempty_s = smem_u32(tempty);
  const int eset = warp >= 7 ? 1 : 0;
  const int ETH = 256;
  const int lg = warp & 3;
  const int row = lg * 32 + lane;
  const int col_half = ((p.N_mma >> 1) + 31) & ~31;
  const int col_begin = eset ? min(col_half, p.N_mma) : 0;
  const int col_end = !eset ? min(col_half, p.N_mma) : p.N_mma;
  const int lw = 31 - __clz(p.TW);
  uint32_t tl = 0;
  TileWalk walk;
  TileCoord tc;
  walk_init(p, walk);
  (void)ETH; (void)tempty_s;
    // ============ TMA-store epilogue (default for bf16 NHWC outputs) ============
    // TMEM -> registers -> 16-bit values in a SWIZZLED shared-memory tile -> cp.async.bulk.tensor stores (one per channel
    // box; the tensor map clips the tile at the frame edge) - no per-thread global stores, no address arithmetic per pixel.
    // The tile is cut into channel boxes of 64 / 32 / 16 channels (SWIZZLE_128B / 64B / 32B), pixel-major inside a box, so a
    // warp writing one 16-byte chunk per lane (lane = pixel) touches eight distinct bank groups per quarter-warp.
    // InstanceNorm statistics come from the staged tile through the WARP-LEVEL tensor-core path: for a block of 8 channels
    // and 16 pixels, ldmatrix.trans delivers X^T (channels x pixels) as both the A and the B fragment of one
    // mma.m16n8k16 whose A rows 8..15 are ones: D rows 0..7 = the 8x8 Gram block (its diagonal = sum of squares),
    // D rows 8..15 = the channel sums.  1 ldmatrix.x4 + 2 MMAs per 8 channels x 32 pixels instead of ~25 scalar
    // instructions per 8 values; fixed order, so the statistics stay bit-reproducible.
    // The epilogue is a latency chain (tcgen05.ld -> pack -> barrier -> store / statistics -> barrier) run by few warps, so
    //  * everything loop-invariant is hoisted (a thread's pixel row is fixed, hence its swizzle terms; the second 16-byte
    //    chunk of a 16-channel group is the first one's address ^ 16);
    //  * narrow layers (N <= 64, p.epi_pp) run the two warp sets as independent GROUPS on alternate sub-tiles - each with its
    //    own staging buffer, named barrier, store-issuing lane and statistics accumulators - so two chains overlap; wider
    //    layers keep one group of eight warps (columns split between the sets), double-buffered where shared memory allows.
    constexpr bool pingpong = PP;
    const int NB = p.N_mma, n64 = NB >> 6, rem = NB & 63;
    const uint32_t buf_bytes = (uint32_t)NB * 256u;
    const int wi = warp - (eset ? 7 : 2);                   // warp inside its set
    const int GW = pingpong ? 4 : 8;                        // warps per group
    const int gw = pingpong ? wi : eset * 4 + wi;           // warp inside its group
    const int gthreads = GW * 32, bar_id = pingpong ? 1 + eset : 1;
    const int cb = pingpong ? 0 : col_begin, ce = pingpong ? NB : col_end;
    const bool two_buf = !pingpong && p.epi_nbuf == 2;
    const uint32_t gbase = stg_s + ((pingpong && eset) ? buf_bytes : 0u);
    const bool do_stats = p.stats && !(p.dbg & 1);
    const int c64_end = n64 * 64, c32_end = c64_end + (rem & 32);
    const uint32_t base32 = (uint32_t)n64 << 14, base16 = base32 + ((rem & 32) ? 8192u : 0u);
    // byte offset of (channel block c8, pixel r) inside a staging buffer (set-up and statistics addressing only)
    auto stg_off = [&](int c8, int r) -> uint32_t {
      const int c = c8 * 8;
      if (c < c64_end) return ((uint32_t)(c >> 6) << 14) + ((uint32_t)r << 7) + ((uint32_t)((c8 & 7) ^ (r & 7)) << 4);
      if (c < c32_end) return base32 + ((uint32_t)r << 6) + ((uint32_t)((((c - c64_end) >> 3) ^ (r >> 1)) & 3) << 4);
      return base16 + ((uint32_t)r << 5) + ((uint32_t)((((c - c32_end) >> 3) ^ (r >> 2)) & 1) << 4);
    };
    // this thread's staging row: per box kind the row base and the swizzle XOR term
    const uint32_t rb64 = (uint32_t)row << 7, sx64 = (uint32_t)(row & 7) << 4;
    const uint32_t rb32 = base32 + ((uint32_t)row << 6), sx32 = (uint32_t)((row >> 1) & 3) << 4;
    const uint32_t rb16 = base16 + ((uint32_t)row << 5) + ((uint32_t)((row >> 2) & 1) << 4);
    // Statistics: warp gw of a group owns the 8-channel blocks gw, gw + GW, ... (all 128 pixels of the sub-tile: four
    // ldmatrix.x4.trans + eight MMAs per block); partial sums of the two groups / CTAs meet in the fp64 atomics.  Two-level
    // accumulation: the MMA chain (whose adder truncates) is cut after 32 links and summed on in plain fp32.
    // ping-pong groups use at most two blocks per warp: slots k + 2 then hold a SECOND, independent MMA chain of block k
    // (the chain's latency, not its issue rate, is what a sub-tile waits for)
    float acc[4][4], td[4], ts0[4], ts1[4];
    auto comb = [&](int k, int j) -> float { return (pingpong && k < 2) ? acc[k][j] + acc[(k + 2) & 3][j] : acc[k][j]; };
    uint32_t ldm_off[4], ldm_step[4];
    int slot_c8[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
      td[k] = ts0[k] = ts1[k] = 0.f;
      const int c8 = gw + GW * k;
      slot_c8[k] = (c8 * 8 < NB) ? c8 : -1;
      ldm_off[k] = ldm_step[k] = 0;
      if (slot_c8[k] >= 0) { ldm_off[k] = stg_off(c8, lane); ldm_step[k] = stg_off(c8, 32) - stg_off(c8, 0); }
    }
    const bool odd_g = (lane >> 2) & 1;
    int st_n = -1, st_c = 0, chain = 0;
    auto flush = [&]() {   // per warp, no barrier
      if (st_n >= 0) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (slot_c8[k] >= 0) {
            const int cq = st_c + slot_c8[k] * 8 + g;     // sum of squares: D[g][g] sits in lane 4g + (g >> 1), register g & 1
            if (t == (g >> 1) && cq < p.Cout)
              atomicAdd(p.stats + ((size_t)st_n * p.Cout + cq) * 2 + 1, (double)(td[k] + (odd_g ? comb(k, 1) : comb(k, 0))));
            const int cs = st_c + slot_c8[k] * 8 + 2 * t;  // sums: D[8][2t], D[8][2t + 1] in lanes 0..3
            if (g == 0) {
              if (cs < p.Cout) atomicAdd(p.stats + ((size_t)st_n * p.Cout + cs) * 2, (double)(ts0[k] + comb(k, 2)));
              if (cs + 1 < p.Cout) atomicAdd(p.stats + ((size_t)st_n * p.Cout + cs + 1) * 2, (double)(ts1[k] + comb(k, 3)));
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
        td[k] = ts0[k] = ts1[k] = 0.f;
      }
      chain = 0;
    };
    const uint32_t ONES = 0x3F803F80u;
    // the bulk stores are issued by one lane of the group's last warp (it owns the fewest statistics blocks)
    const bool issuer = (warp == 10 || (pingpong && warp == 5)) && elect_one();
    const int sub_shift = 7 - lw;   // sub-tile m starts at tile row (m * 128) >> lw
    for (; walk_next(p, walk, tc, total_tiles); ++tl) {
      const uint32_t acc_i = tl & (AS - 1), accph = (tl >> as_sh) & 1;
      mbar_wait(&tfull[acc_i], accph);
      tc_fence_after();
      const int cbase = tc.nt * p.N_mma;
      if (do_stats && (tc.n != st_n || cbase != st_c)) {
        flush();
        st_n = tc.n;
        st_c = cbase;
      }
      const int vy = tc.valid ? min(p.TH, p.Ho - tc.y0) : 0, vx = min(p.TW, p.Wo - tc.x0);
      const bool full_tile = (vy == p.TH) && (vx == p.TW);
      const uint32_t taddr0 = tmem_base + ((uint32_t)(lg * 32) << 16) + acc_i * acc_cols;
      const uint32_t it0 = tl * (uint32_t)MT;    // running sub-tile index of this tile's first sub-tile
      // last sub-tile of this tile this group reads (ping-pong: sub-tiles of the group's parity; -1: none)
      const int last_m = !pingpong ? MT - 1 : ((((it0 + MT - 1) & 1u) == (uint32_t)eset) ? MT - 1 : MT - 2);
      if (last_m < 0) {   // nothing to read from this accumulator: release it straight away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc_i]);
      }
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const uint32_t it = it0 + (uint32_t)m;
        if (pingpong && (it & 1u) != (uint32_t)eset) continue;
        const uint32_t sbuf = gbase + ((two_buf && (it & 1u)) ? buf_bytes : 0u);
        const uint32_t taddr = taddr0 + m * p.N_mma;
        bool zero_px = false;
        if (!full_tile) {
          const int tr = m * 128 + row, ty = tr >> lw, tx = tr & (p.TW - 1);
          zero_px = !(ty < vy && tx < vx);   // pixels past the frame edge: zeros (clipped by the store, neutral for the statistics)
        }
        if (pingpong && !(p.dbg & 64)) {
          // N <= 64: the whole pixel row fits in registers, so the accumulators are read and packed BEFORE the buffer
          // barrier - the tcgen05.ld latency and the conversion overlap the previous bulk store's shared-memory read
          uint4 q[8];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (i * 32 < NB && !(p.dbg & 4)) {
              uint32_t r[32];
              const bool two = (i * 32 + 16 < NB);
              tmem_ld16(taddr + i * 32, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
              if (two) tmem_ld16(taddr + i * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
              tmem_ld_wait();
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (h == 1 && !two) break;
                const int c = i * 32 + h * 16;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[h * 16 + j]);
                if (p.bias) {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (cbase + c + j < p.Cout) v[j] += p.bias[cbase + c + j];
                }
                if (p.relu) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                uint4 q0, q1;
                q0.x = pack_bf16x2(v[0], v[1]);   q0.y = pack_bf16x2(v[2], v[3]);
                q0.z = pack_bf16x2(v[4], v[5]);   q0.w = pack_bf16x2(v[6], v[7]);
                q1.x = pack_bf16x2(v[8], v[9]);   q1.y = pack_bf16x2(v[10], v[11]);
                q1.z = pack_bf16x2(v[12], v[13]); q1.w = pack_bf16x2(v[14], v[15]);
                if (zero_px) { q0 = make_uint4(0u, 0u, 0u, 0u); q1 = q0; }
                q[i * 4 + h * 2] = q0; q[i * 4 + h * 2 + 1] = q1;
              }
            }
          }
          if (m == last_m) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc_i]);
          }
          if (issuer) bulk_wait_read0();
          epi_bar_sync_id(bar_id, gthreads);   // the group's buffer is free: its last store has read it, its statistics are done
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = i * 16;
            if (c < NB && !(p.dbg & 4)) {
              uint32_t a0;
              if (c < c64_end) a0 = rb64 + ((((uint32_t)c & 48u) << 1) ^ sx64);
              else if (c < c32_end) a0 = rb32 + (((uint32_t)(c - c64_end) << 1) ^ sx32);
              else a0 = rb16;
              a0 += sbuf;
              sts128(a0, q[2 * i]);
              sts128(a0 ^ 16u, q[2 * i + 1]);
            }
          }
        } else {
        if (!two_buf) {   // one buffer per group: the previous store has read it, and every warp is done with its statistics
          if (issuer) bulk_wait_read0();
          epi_bar_sync_id(bar_id, gthreads);
        }
#pragma unroll 1
        for (int c0 = cb; c0 < ((p.dbg & 4) ? 0 : ce); c0 += 32) {
          uint32_t r[32];
          const bool two = (c0 + 16 < ce);
          tmem_ld16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
          if (two) tmem_ld16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 1 && !two) break;
            const int c = c0 + h * 16;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[h * 16 + j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cbase + c + j < p.Cout) v[j] += p.bias[cbase + c + j];
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            uint4 q0, q1;
            q0.x = pack_bf16x2(v[0], v[1]);   q0.y = pack_bf16x2(v[2], v[3]);
            q0.z = pack_bf16x2(v[4], v[5]);   q0.w = pack_bf16x2(v[6], v[7]);
            q1.x = pack_bf16x2(v[8], v[9]);   q1.y = pack_bf16x2(v[10], v[11]);
            q1.z = pack_bf16x2(v[12], v[13]); q1.w = pack_bf16x2(v[14], v[15]);
            if (zero_px) { q0 = make_uint4(0u, 0u, 0u, 0u); q1 = q0; }
            uint32_t a0;
            if (c < c64_end) a0 = ((uint32_t)(c >> 6) << 14) + rb64 + ((((uint32_t)c & 48u) << 1) ^ sx64);
            else if (c < c32_end) a0 = rb32 + (((uint32_t)(c - c64_end) << 1) ^ sx32);
            else a0 = rb16;
            a0 += sbuf;
            sts128(a0, q0);
            sts128(a0 ^ 16u, q1);
          }
        }
        if (m == last_m) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {   // accumulators drained: MMA may reuse them (pair: the barrier lives in the rank-0 CTA)
            if constexpr (CTA2) mbar_arrive_cluster(mapa_u32(tempty_s + acc_i * 8, 0));
            else mbar_arrive(&tempty[acc_i]);
          }
        }
        }
        fence_proxy_async_smem();                       // this thread's tile writes -> visible to the bulk store
        if (two_buf && issuer) bulk_wait_read0();       // the store issued one sub-tile ago has read the OTHER buffer
        epi_bar_sync_id(bar_id, gthreads);
        if (issuer && tc.valid && !(p.dbg & 2)) {
          const int sx = tc.x0 + ((m << 7) & (p.TW - 1)), sy = tc.y0 + (sub_shift >= 0 ? (m << sub_shift) : (m >> -sub_shift));
          const CUtensorMap* mo = &p.tmO[tc.ph][0];
#pragma unroll 1
          for (int b = 0; b < n64; ++b)
            if (cbase + 64 * b < p.Cout) tma_store_4d(mo, sbuf + ((uint32_t)b << 14), cbase + 64 * b, sx, sy, tc.n);
          if ((rem & 32) && cbase + c64_end < p.Cout) tma_store_4d(mo + 1, sbuf + base32, cbase + c64_end, sx, sy, tc.n);
          if ((rem & 16) && cbase + c32_end < p.Cout) tma_store_4d(mo + 2, sbuf + base16, cbase + c32_end, sx, sy, tc.n);
          bulk_commit();
        }
        if (do_stats) {
          if (pingpong) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              if (slot_c8[k] >= 0) {
                uint32_t a = sbuf + ldm_off[k];
#pragma unroll
                for (int pq = 0; pq < 4; ++pq, a += ldm_step[k]) {
                  uint32_t x0, x1, x2, x3;
                  ldmatrix_x4_trans(a, x0, x1, x2, x3);
                  mma16816_bf16(acc[k], x0, ONES, x1, ONES, x0, x1);
                  mma16816_bf16(acc[k + 2], x2, ONES, x3, ONES, x2, x3);
                }
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (slot_c8[k] >= 0) {
                uint32_t a = sbuf + ldm_off[k];
#pragma unroll
                for (int pq = 0; pq < 4; ++pq, a += ldm_step[k]) {
                  uint32_t x0, x1, x2, x3;
                  ldmatrix_x4_trans(a, x0, x1, x2, x3);
                  mma16816_bf16(acc[k], x0, ONES, x1, ONES, x0, x1);
                  mma16816_bf16(acc[k], x2, ONES, x3, ONES, x2, x3);
                }
              }
            }
          }
          if (++chain == (pingpong ? 8 : 4)) {   // second level: plain fp32 adds after every 32 chained MMAs
            chain = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              td[k] += odd_g ? comb(k, 1) : comb(k, 0);
              ts0[k] += comb(k, 2);
              ts1[k] += comb(k, 3);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
          }
        }
      }
    }
    if (do_stats) flush();
    if (issuer) bulk_wait0();   // the last stores have left shared memory (and completed) before the CTA retires
}

// warps: 0 A-producer, 1 MMA, 2..5 epilogue set 0, 6 B-producer, 7..10 epilogue set 1 (bf16 NHWC epilogue only: the narrow
// layers are bound by the epilogue's instruction stream, so two warps share each TMEM lane group and split the columns)
constexpr int TG_THREADS = 384;   // + warp 11: second MMA issuer (layers with MT >= 2 sub-tiles, see p.mma2)
// register cap: launch bounds of 512 threads make ptxas keep the kernel within 128 registers per thread (49 K of the 64 K
// registers for its 384 threads), so blocks of the HBM-bound companion kernels of ANOTHER stream can co-reside on the SM
#ifndef TG_REGCAP_THREADS
#define TG_REGCAP_THREADS 512
#endif
constexpr int RC_LD = 33;        // row pitch (floats) of the row-conv staging tile
constexpr int TG_DIRECT_SCRATCH = 8192;   // direct epilogue: per-thread statistics slots [chunks][epilogue threads] (16 x 128 or 4 x 256 floats)

// ---------------------------------------------------------------------------------------------------------------------
// Apply rider (warps 12..15 of a launch with p.rider.on): the InstanceNorm apply pass of ANOTHER half-batch streamed beside this
// kernel's MMAs - see ApplyRider in tc_conv.cuh.  Work unit = one padded destination row; CTA b takes rows b, b + grid, ...
// (the 148 CTAs sweep 148 consecutive rows at a time).  A thread owns one 8-channel group for the whole job, so its 16
// constants are derived straight from the statistics when the image changes (no shared memory, no barrier) and live in
// registers; four pixels (16-byte vectors, plus the residual's) are in flight per thread.  Arithmetic order is
// apply_lds_kernel's: fmaf(x, a, b) -> ReLU -> + residual -> round.
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {   // bytes: multiple of 16
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

template <bool RES>
__device__ __forceinline__ uint4 rider_vec(const uint4 q, const uint4 r, const float (&ca)[8], const float (&cb)[8], int relu) {
  const uint32_t qi[4] = {q.x, q.y, q.z, q.w}, ri[4] = {r.x, r.y, r.z, r.w};
  uint32_t oo[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = bf16x2_to_f2(qi[k]);
    float a = fmaf(f.x, ca[2 * k], cb[2 * k]), b = fmaf(f.y, ca[2 * k + 1], cb[2 * k + 1]);
    if (relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
    if (RES) {
      const float2 rr = bf16x2_to_f2(ri[k]);
      a += rr.x; b += rr.y;
    }
    oo[k] = pack_bf16x2(a, b);
  }
  return make_uint4(oo[0], oo[1], oo[2], oo[3]);
}

// One warp per scheduler has nobody to hide its latencies behind, so the row loop is software-pipelined by hand (the next
// PX vectors are in flight while the current PX are converted and stored) and kept short: the interior of a row needs no
// padding arithmetic, the few halo columns are a separate tail.
template <int PX, bool RES>
__device__ __forceinline__ void apply_rider_rows(const ApplyRider& jp, int t, int cta, int n_cta) {
  // every field into a local first: `jp` lives in the kernel's parameter space behind a generic reference, and the compiler
  // must assume the global stores below may alias it - it would re-load layout fields after every store
  const ActLayout DL = jp.DL, RL = jp.RL;
  const int N = jp.N, relu = jp.relu;
  const float eps = jp.eps;
  const double* __restrict__ stats = jp.stats;
  const float* __restrict__ gamma = jp.gamma;
  const float* __restrict__ beta = jp.beta;
  const __nv_bfloat16* __restrict__ raw = reinterpret_cast<const __nv_bfloat16*>(jp.raw);
  const __nv_bfloat16* __restrict__ residual = reinterpret_cast<const __nv_bfloat16*>(jp.residual);
  __nv_bfloat16* __restrict__ dst = reinterpret_cast<__nv_bfloat16*>(jp.dst);
  const int C = DL.C, groups = C >> 3, H = DL.H, W = DL.W, pad = DL.pad, kind = DL.kind;
  const int Hp = H + 2 * pad;
  const int dpar = DL.parity, rpar = RL.parity, rpad = RL.pad, rC = RL.C;
  const int g = t % groups, pl = t / groups, step = 128 / groups;
  const float inv_cnt = 1.f / (float)(H * W);
  const uint32_t row_bytes = (uint32_t)(W * C * 2);
  float ca[8], cb[8];
  int cur_n = -1;
  const int total = Hp * N;
  auto prefetch_row = [&](int row) {   // one thread: the whole source row (and the residual's) towards L2
    if (row >= total) return;
    const int n = row / Hp, yp = row - n * Hp;
    bool ok;
    const int sy = map_pad(yp - pad, H, kind, ok);
    if (!ok) return;
    l2_prefetch_bulk(raw + ((size_t)n * H + sy) * W * C, row_bytes);
    if (RES) {
      const int wp_r = (RL.W + 2 * rpad) >> rpar;   // pixels per stored row (per parity plane)
      const uint32_t rb = (uint32_t)(wp_r * rC * 2);
      l2_prefetch_bulk(residual + act_offset(RL, N, n, sy + rpad, 0), rb);
      if (rpar) l2_prefetch_bulk(residual + act_offset(RL, N, n, sy + rpad, 1), rb);
    }
  };
  if (t == 0) { prefetch_row(cta); prefetch_row(cta + n_cta); }
  if (pl >= step) return;
  const int xstride = PX * step;
  for (int row = cta; row < total; row += n_cta) {
    if (t == 0) prefetch_row(row + 2 * n_cta);
    const int n = row / Hp, yp = row - n * Hp;
    if (n != cur_n) {
      cur_n = n;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = g * 8 + k;
        const double s1 = stats[((size_t)n * C + c) * 2], s2 = stats[((size_t)n * C + c) * 2 + 1];
        const double mean_d = s1 * (double)inv_cnt;
        const float mean = (float)mean_d;
        const float var = fmaxf((float)(s2 * (double)inv_cnt - mean_d * mean_d), 0.f);
        const float a = gamma[c] * rsqrtf(var + eps);
        ca[k] = a;
        cb[k] = beta[c] - mean * a;
      }
    }
    bool oky;
    const int sy = map_pad(yp - pad, H, kind, oky);
    __nv_bfloat16* d0 = dst + act_offset(DL, N, n, yp, 0) + g * 8;
    __nv_bfloat16* d1 = dst + act_offset(DL, N, n, yp, dpar) + g * 8;
    auto dptr = [&](int x) { return reinterpret_cast<uint4*>(((x & dpar) ? d1 : d0) + (size_t)(x >> dpar) * C); };
    if (!oky) {   // zero-padded halo row
      for (int x = pl; x < W + 2 * pad; x += step) *dptr(x) = make_uint4(0, 0, 0, 0);
      continue;
    }
    const __nv_bfloat16* rrow = raw + ((size_t)n * H + sy) * W * C + g * 8;
    const __nv_bfloat16* res0 = nullptr;
    const __nv_bfloat16* res1 = nullptr;
    if (RES) {
      res0 = residual + act_offset(RL, N, n, sy + rpad, 0) + g * 8;
      res1 = residual + act_offset(RL, N, n, sy + rpad, rpar) + g * 8;
    }
    auto rptr = [&](int sx) {
      const int rx = sx + rpad;
      return reinterpret_cast<const uint4*>(((rx & rpar) ? res1 : res0) + (size_t)(rx >> rpar) * rC);
    };
    // ---- interior: source pixel sx -> destination pixel sx + pad
    uint4 q[PX], r[PX], qn[PX], rn[PX];
#pragma unroll
    for (int u = 0; u < PX; ++u) {
      const int sx = pl + u * step;
      qn[u] = rn[u] = make_uint4(0, 0, 0, 0);
      if (sx < W) {
        qn[u] = __ldg(reinterpret_cast<const uint4*>(rrow + (size_t)sx * C));
        if (RES) rn[u] = __ldg(rptr(sx));
      }
    }
    for (int sx0 = pl; sx0 < W; sx0 += xstride) {
#pragma unroll
      for (int u = 0; u < PX; ++u) { q[u] = qn[u]; r[u] = rn[u]; }
      const int nx0 = sx0 + xstride;
      if (nx0 < W) {
#pragma unroll
        for (int u = 0; u < PX; ++u) {
          const int sx = nx0 + u * step;
          if (sx < W) {
            qn[u] = __ldg(reinterpret_cast<const uint4*>(rrow + (size_t)sx * C));
            if (RES) rn[u] = __ldg(rptr(sx));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < PX; ++u) {
        const int sx = sx0 + u * step;
        if (sx < W) *dptr(sx + pad) = rider_vec<RES>(q[u], r[u], ca, cb, relu);
      }
    }
    // ---- halo columns: 2 * pad pixels per row, mirrored / replicated / zero
    for (int e = pl; e < 2 * pad; e += step) {
      const int x = e < pad ? e : W + e;
      bool okx;
      const int sx = map_pad(x - pad, W, kind, okx);
      uint4 o = make_uint4(0, 0, 0, 0);
      if (okx) {
        const uint4 qq = __ldg(reinterpret_cast<const uint4*>(rrow + (size_t)sx * C));
        uint4 rr = make_uint4(0, 0, 0, 0);
        if (RES) rr = __ldg(rptr(sx));
        o = rider_vec<RES>(qq, rr, ca, cb, relu);
      }
      *dptr(x) = o;
    }
  }
}

__device__ __noinline__ void apply_rider_run(const ApplyRider& j, int t, int cta, int n_cta) {
  if (j.residual) apply_rider_rows<4, true>(j, t, cta, n_cta);
  else apply_rider_rows<8, false>(j, t, cta, n_cta);
}

// smem carve-up (host mirrors this in launch_tapgemm):
//   [S stages x G k-blocks x (A MT*128 x BK | B N_mma x BK, 1024-aligned)] [epilogue staging] [barriers]
// MODE / EPI >= 0 SPECIALISE the kernel for one main-loop mode / one epilogue: the other modes' code is not compiled in.  This
// matters more than it looks: the generic kernel is ~15 K instructions, its roles run different regions of it at the same
// time, and its hot loops then stall on instruction fetch and share one register allocation (ncu: stall_no_inst on the
// epilogue's lines; adding an unrelated epilogue made the OLD one 20 % slower).  -1 = decided at run time (generic kernel).
enum : int { TGM_PLAIN = 0, TGM_DYSH = 1, TGM_RING = 2, TGM_ACCRING = 3 };
enum : int { TGE_STAGED = 0, TGE_DIRECT = 1, TGE_ROWCONV = 2, TGE_F32 = 3, TGE_TMA = 4, TGE_TMAPP = 5 };
// DUO: the two-CTAs-per-SM instantiation.  The narrow layers (conv1, conv2, deconv2: 48 / 96 output channels) leave every
// unit of the SM half idle - tensor pipe 45 %, issue slots 44 %, DRAM 28 % (profiles/r02_conv1_tapgemm_ncu_details.txt) -
// because ONE chain of tile -> accumulator -> epilogue round trips paces the CTA.  Like the narrow Gram (tc_pcgemm.cu), the
// fix is a second, independent chain: two CTAs per SM, each with half the shared memory, 256 of the 512 TMEM columns and
// <= 80 registers per thread (384 threads x 2 x 80 fits the register file; the 128-register build does not).
template <int BK, bool CTA2, int MODE, int EPI, bool DUO = false>   // CTA2: the CTA-pair instantiation (cluster launch only); the plain one holds no cta_group::2 code
__global__ void __launch_bounds__(DUO ? 384 : TG_REGCAP_THREADS, DUO ? 2 : 1) tapgemm_kernel(const __grid_constant__ TapGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int SUB_BYTES = 128 * BK * 2;
  const int MT = p.MT;
  const int a_bytes = MT * SUB_BYTES;
  // CTA pair (p.cta2): M = 256 MMAs over the two SMs of a cluster; this CTA holds its own 128 pixels of A and HALF of the
  // weight rows, the rank-0 CTA issues every MMA and owns the operand-full / accumulator-empty barriers
  constexpr bool cta2 = CTA2;
  uint32_t crank = 0u;
  if constexpr (CTA2) crank = cluster_ctarank();
  const int b_bytes = (cta2 ? p.N_mma / 2 : p.N_mma) * BK * 2;
  const int kb_bytes = a_bytes + ((b_bytes + 1023) & ~1023);
  const int G = p.group;
  const bool stream = MODE < 0 ? p.stream != 0 : (MODE == TGM_RING || MODE == TGM_ACCRING);
  const bool acc_ring = MODE < 0 ? p.stream == 2 : MODE == TGM_ACCRING;
  const bool e_nhwc = EPI < 0 ? p.epi_mode == TG_EPI_BF16_NHWC : (EPI == TGE_STAGED || EPI == TGE_DIRECT || EPI == TGE_TMA || EPI == TGE_TMAPP);
  const bool e_rowconv = EPI < 0 ? p.epi_mode == TG_EPI_ROWCONV : EPI == TGE_ROWCONV;
  const bool e_direct = EPI < 0 ? p.epi_direct != 0 : EPI == TGE_DIRECT;
  const bool e_tma = EPI < 0 ? p.epi_tma != 0 : (EPI == TGE_TMA || EPI == TGE_TMAPP);
  const bool e_pp = EPI < 0 ? p.epi_pp != 0 : EPI == TGE_TMAPP;
  const bool fuse_in = ((MODE < 0 || MODE == TGM_ACCRING) && (EPI < 0 || EPI == TGE_ROWCONV)) ? p.fuse_in != 0 : false;
  const int b_al = (b_bytes + 1023) & ~1023;
  // stream mode: "stage" = one ring slot holding one input row (kb_per_tap k-blocks of 128 pixels); weights follow the ring
  const bool dysh = MODE < 0 ? p.dyshare != 0 : MODE == TGM_DYSH;
  const int a_box_bytes = p.box_rows * p.TW * BK * 2;                 // dy-sharing: one box of TH + dy_max - 1 rows
  const int a_box_al = (a_box_bytes + 1023) & ~1023;
  const bool wres = dysh && p.w_res != 0;   // dy-sharing with the weights resident behind the ring (loaded once, like the stream modes)
  const int stage_bytes = stream ? p.kb_per_tap * SUB_BYTES : dysh ? a_box_al + (wres ? 0 : p.dy_max * b_al) : G * kb_bytes;
  const int S = p.stages;
  const int w_region = (stream || wres) ? p.n_taps * p.kb_per_tap * b_al : 0;
  uint8_t* stg = smem + (size_t)S * stage_bytes + w_region;  // epilogue staging tile
  const int stg_pitch = p.N_mma * 2 + 16;         // bytes per staged bf16 row
  const int stg_one = max(128 * stg_pitch, TG_DIRECT_SCRATCH);   // one staged sub-tile
  const int stg_bytes = e_nhwc ? (e_tma ? p.epi_nbuf * p.N_mma * 256
                                                           : e_direct ? TG_DIRECT_SCRATCH : (p.epi_spp ? 2 : 1) * stg_one)
                        : e_rowconv ? 2 * 128 * RC_LD * 4 : 0;   // one tile per epilogue warp set
  uint64_t* full = reinterpret_cast<uint64_t*>(stg + ((stg_bytes + 15) & ~15));
  uint64_t* empty = full + S;
  uint64_t* tfull = empty + S;
  uint64_t* tempty = tfull + 16;
  uint64_t* wfull = tempty + 16;   // stream mode: resident weights have landed
  uint64_t* xfull = wfull + 1;     // fused input normalisation: ring slot s has been transformed in place (16 slots)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.n_phase * p.n_ntile * p.n_img * p.tiles_y * p.tiles_x;
  const int groups = (p.n_taps * p.kb_per_tap) / G;
  const int acc_cols = MT * p.N_mma;
  uint32_t tmem_cols = 32;
  const int AS = p.acc_stages, as_sh = 31 - __clz(AS);   // TMEM accumulator stages (power of two, 2..8)
  while ((int)tmem_cols < AS * acc_cols) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], (stream || wres) ? 1 : 2);   // A-producer + B-producer (each arrive.expect_tx); resident weights: A only
      mbar_init(&empty[s], p.mma2 ? 2 : 1);  // tcgen05.commit of each MMA-issuing warp
    }
    for (int a = 0; a < AS; ++a) {
      mbar_init(&tfull[a], p.mma2 ? 2 : 1);
      mbar_init(&tempty[a], (p.epi8 ? 8 : 4) * (cta2 ? 2 : 1));   // pair: the peer's epilogue warps arrive remotely
    }
    mbar_init(wfull, 1);
    for (int s = 0; s < 16; ++s) mbar_init(&xfull[s], 4);   // one arrival per transform warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CTA2) tmem_alloc2(tmem_slot, tmem_cols);
    else tmem_alloc(tmem_slot, tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: barriers, TMEM and descriptor prefetch above overlap the previous kernel's tail; nothing
  // below may touch global memory before the predecessor's writes are visible
  pdl_wait();
  pdl_trigger();

  // Role loops run on ONE lane each with everything loop-invariant hoisted into registers and all
  // shared-memory objects addressed by 32-bit shared addresses: the issue loops are serial code, so
  // their instruction count per k-block bounds the whole pipeline.
  const uint32_t smem_s = smem_u32(smem), full_s = smem_u32(full), empty_s = smem_u32(empty);
  const uint32_t tfull_s = smem_u32(tfull), tempty_s = smem_u32(tempty);
  const int n_taps = p.n_taps, kbpt = p.kb_per_tap, N_mma = p.N_mma, n_ntile = p.n_ntile;

  if (p.rider.on == 3 && warp < 12) {
    // timing experiment (VST_RIDER_DBG=2): the rider alone, every tap-GEMM role idle
  } else if (warp == 0) {
    // ================================ A producer (activations) ====================
    const bool leader = elect_one();
    if (leader && stream) {
      // one TMA box per input row and k-block: (BK channels) x (128 pixels) x (1 row); rows outside the tensor zero-fill
      const int tp = p.tap_packed[0];
      const int dx0 = (int)(signed char)(tp & 0xff), pl0 = tp >> 16;
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = (uint32_t)stage_bytes;
      for (int u = blockIdx.x; u < stream_units(p); u += gridDim.x) {
        const StreamUnit su = decode_unit(p, u);
        if (su.rows <= 0) continue;
        const int y_first = su.r0 + p.s_dy0, y_last = su.r0 + su.rows - 1 + p.s_dy0 + n_taps - 1;
        for (int y = y_first; y <= y_last; ++y) {
          mbar_wait_a(empty_s + s * 8, ph ^ 1);
          const uint32_t bar = full_s + s * 8;
          mbar_expect_tx_a(bar, tx_bytes);
          uint32_t sa = smem_s + s * stage_bytes;
          const int ysrc = fuse_in ? reflect_idx(y, p.in_H) : y;   // fused input: no halo in memory, rows mirror here
          for (int kb = 0; kb < kbpt; ++kb, sa += SUB_BYTES) tma_load_5d_a(sa, &p.tmA, bar, kb * BK, su.x0 + dx0, ysrc, su.n, pl0);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    } else if (leader && cta2) {
      // both CTAs load their own pixels; the bytes are counted on the rank-0 CTA's full barrier, which expects both halves
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = (uint32_t)(2 * G * a_bytes);
      const uint32_t full0 = mapa_u32(full_s, 0);
      for (int tile = blockIdx.x; (tile & ~1) < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, min(tile, total_tiles - 1));
        const int* tapp = &p.tap_packed[tc.ph * n_taps];
        int tp = tapp[0], t = 0, kb = 0;
        for (int g = 0; g < groups; ++g) {
          mbar_wait_a(empty_s + s * 8, ph ^ 1);
          if (crank == 0) mbar_expect_tx_a(full_s + s * 8, tx_bytes);
          uint32_t sa = smem_s + s * stage_bytes;
          for (int j = 0; j < G; ++j) {
            tma_load_5d_2sm(sa, &p.tmA, full0 + s * 8, kb * BK, tc.x0 + (int)(signed char)(tp & 0xff),
                            tc.y0 + (int)(signed char)((tp >> 8) & 0xff), tc.n, tp >> 16);
            sa += kb_bytes;
            if (++kb == kbpt) {
              kb = 0;
              if (++t < n_taps) tp = tapp[t];
            }
          }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    } else if (leader && dysh) {
      // one box per (column, k-block): TH + dy_max - 1 rows starting at the column's first tap row
      int s = 0;
      uint32_t ph = 0;
      const int n_cols = p.n_cols;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int cb = tc.ph * n_cols;
        for (int c = 0; c < n_cols; ++c) {
          const int cx = tc.x0 + p.col_dx[cb + c], cy = tc.y0 + p.col_dy0[cb + c], cpl = p.col_pl[cb + c];
          for (int kb = 0; kb < kbpt; ++kb) {
            mbar_wait_a(empty_s + s * 8, ph ^ 1);
            const uint32_t bar = full_s + s * 8;
            mbar_expect_tx_a(bar, (uint32_t)a_box_bytes);
            tma_load_5d_a(smem_s + s * stage_bytes, &p.tmA, bar, kb * BK, cx, cy, tc.n, cpl);
            if (++s == S) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (leader) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = (uint32_t)(G * a_bytes);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int* tapp = &p.tap_packed[tc.ph * n_taps];
        int tp = tapp[0], t = 0, kb = 0;
        for (int g = 0; g < groups; ++g) {
          mbar_wait_a(empty_s + s * 8, ph ^ 1);
          const uint32_t bar = full_s + s * 8;
          uint32_t sa = smem_s + s * stage_bytes;
          mbar_expect_tx_a(bar, tx_bytes);
          for (int j = 0; j < G; ++j) {
            tma_load_5d_a(sa, &p.tmA, bar, kb * BK, tc.x0 + (int)(signed char)(tp & 0xff),
                          tc.y0 + (int)(signed char)((tp >> 8) & 0xff), tc.n, tp >> 16);
            sa += kb_bytes;
            if (++kb == kbpt) {
              kb = 0;
              if (++t < n_taps) tp = tapp[t];
            }
          }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 6) {
    // ================================ B producer (weights) ========================
    const bool leader = elect_one();
    if (leader && (stream || wres)) {
      // all taps x k-blocks once: they stay resident behind the ring
      const uint32_t wbar = smem_u32(wfull);
      mbar_expect_tx_a(wbar, (uint32_t)(n_taps * kbpt * b_bytes));
      uint32_t sb = smem_s + S * stage_bytes;
      if (p.merge_taps) {   // kbpt == 1: tap t -> block n_taps-1-t (see the tap-merged MMAs)
        for (int t = 0; t < n_taps; ++t) tma_load_2d_a(sb + (uint32_t)(n_taps - 1 - t) * b_al, &p.tmB, wbar, t * BK, 0);
      } else {
        for (int i = 0; i < n_taps * kbpt; ++i, sb += b_al) tma_load_2d_a(sb, &p.tmB, wbar, i * BK, 0);
      }
    } else if (leader && cta2) {
      // this CTA's half of the weight rows (the tensor map's box is N_mma / 2 rows)
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = (uint32_t)(2 * G * b_bytes);
      const uint32_t full0 = mapa_u32(full_s, 0);
      for (int tile = blockIdx.x; (tile & ~1) < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, min(tile, total_tiles - 1));
        const int brow = tc.nt * N_mma + (int)crank * (N_mma / 2);
        int kc = 0;
        for (int g = 0; g < groups; ++g) {
          mbar_wait_a(empty_s + s * 8, ph ^ 1);
          if (crank == 0) mbar_expect_tx_a(full_s + s * 8, tx_bytes);
          uint32_t sb = smem_s + s * stage_bytes + a_bytes;
          for (int j = 0; j < G; ++j) {
            tma_load_2d_2sm(sb, &p.tmB, full0 + s * 8, kc, brow);
            kc += BK;
            sb += kb_bytes;
          }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    } else if (leader && dysh) {
      int s = 0;
      uint32_t ph = 0;
      const int n_cols = p.n_cols;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int brow = (tc.ph * n_ntile + tc.nt) * N_mma + tc.n * p.b_img_rows;
        const int cb = tc.ph * n_cols;
        for (int c = 0; c < n_cols; ++c) {
          const int cn = p.col_n[cb + c], t0 = p.col_t0[cb + c], ts = p.col_ts[cb + c];
          for (int kb = 0; kb < kbpt; ++kb) {
            mbar_wait_a(empty_s + s * 8, ph ^ 1);
            const uint32_t bar = full_s + s * 8;
            mbar_expect_tx_a(bar, (uint32_t)(cn * b_bytes));
            uint32_t sb = smem_s + s * stage_bytes + a_box_al;
            for (int j = 0; j < cn; ++j, sb += b_al) tma_load_2d_a(sb, &p.tmB, bar, ((t0 + j * ts) * kbpt + kb) * BK, brow);
            if (++s == S) { s = 0; ph ^= 1; }
          }
        }
      }
    } else if (leader) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx_bytes = (uint32_t)(G * b_bytes);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord tc = decode_tile(p, tile);
        const int brow = (tc.ph * n_ntile + tc.nt) * N_mma + tc.n * p.b_img_rows;
        int kc = 0;
        for (int g = 0; g < groups; ++g) {
          mbar_wait_a(empty_s + s * 8, ph ^ 1);
          const uint32_t bar = full_s + s * 8;
          uint32_t sb = smem_s + s * stage_bytes + a_bytes;
          mbar_expect_tx_a(bar, tx_bytes);
          for (int j = 0; j < G; ++j) {
            tma_load_2d_a(sb, &p.tmB, bar, kc, brow);
            kc += BK;
            sb += kb_bytes;
          }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || (warp == 11 && p.mma2)) {
    // ================================ MMA issuer(s) ===============================
    // The narrow layers are bound by this serial issue stream (~100 clk per MMA against 16-48 clk of math), so layers whose
    // CTA tile has MT >= 2 sub-tiles split them over TWO issuing warps (1 and 11): each owns half of the sub-tiles (their own
    // TMEM accumulators), waits on the same full / tempty barriers and commits separately (empty / tfull count 2).
    const int m_begin = (p.mma2 && warp == 11) ? (MT >> 1) : 0;
    const int m_end = (p.mma2 && warp == 1) ? (MT >> 1) : MT;
    // elect.sync (not `lane == 0`) lets the compiler issue UTCHMMA straight from the uniform datapath; with a data-dependent
    // lane predicate it wraps EVERY MMA in an elect/branch loop.  The issue loop is serial code and its instruction count
    // per MMA bounds the whole kernel (12 instructions per MMA before, ~4 now), so descriptors are RUNNING 64-bit values
    // advanced by in-place adds.
    if (elect_one()) {
      const uint32_t idesc = make_idesc(128, N_mma, p.half);
      // descriptor of smem offset 0; all offsets are multiples of 16 B and stay below 256 KB, so a
      // plain add on the (addr >> 4) field never carries out of it
      const uint64_t desc0 = smem_desc_hi(BK * 2) | (uint64_t)((smem_s & 0x3FFFFu) >> 4);
      const uint32_t stage_d = stage_bytes >> 4, kb_d = kb_bytes >> 4, ab_d = a_bytes >> 4;
      constexpr uint32_t sub_d = SUB_BYTES >> 4;
      if (acc_ring) {
        // Accumulator-ring streaming: an input row is consumed the moment it lands - it feeds tap t of output row (i - t)
        // for every t, i.e. up to n_taps accumulators that live side by side in TMEM (16 slots) - and its ring slot is
        // released straight away.  The shared-memory ring is then a plain prefetch queue (no n_taps-row window to hold),
        // so TMA runs many rows ahead and every input byte crosses L2 -> smem once.
        mbar_wait(wfull, 0);
        tc_fence_after();
        const uint64_t descw0 = desc0 + (uint64_t)((S * stage_bytes) >> 4);
        const uint32_t tapw_d = (uint32_t)(kbpt * b_al) >> 4, bal_d = (uint32_t)b_al >> 4;
        uint32_t tl = 0, sl = 0, ph = 0;
        uint64_t da = desc0;
        for (int u = blockIdx.x; u < stream_units(p); u += gridDim.x) {
          const StreamUnit su = decode_unit(p, u);
          if (su.rows <= 0) continue;
          const int n_in = su.rows + n_taps - 1;
          for (int i = 0; i < n_in; ++i) {
            if (i < su.rows) {   // output row i starts with this input row: its accumulator slot must be drained
              const uint32_t tj = tl + (uint32_t)i;
              mbar_wait_a(tempty_s + (tj & (AS - 1)) * 8, ((tj >> as_sh) & 1) ^ 1);
            }
            mbar_wait_a((fuse_in ? smem_u32(xfull) : full_s) + sl * 8, ph);
            tc_fence_after();
            const int t_lo = max(0, i - su.rows + 1), t_hi = min(i, n_taps - 1);
            // accumulator of output row (i - t): slot (tl + i - t) mod AS; AS * acc_cols is a power of two (16 x 16|32), so
            // the TMEM column offset just steps down by acc_cols per tap under a mask
            uint32_t off = ((tl + (uint32_t)(i - t_lo)) & (AS - 1)) * acc_cols;
            const uint32_t off_mask = (uint32_t)(AS * acc_cols - 1);
            uint64_t dw = descw0 + (uint64_t)t_lo * tapw_d;
            if (p.merge_taps) {
              // Tap-merged MMAs: all taps read the SAME A operand (this input row) and differ only in weights and in the
              // accumulator they feed, so consecutive taps are issued as ONE MMA whose N dimension spans several accumulator
              // slots: the weights sit in shared memory in REVERSE tap order (block b = n_taps-1-t), so that block b of the
              // merged B tile lands on slot (g - t) = (g - n_taps + 1 + b) - ascending with b.  An MMA covers at most 256
              // columns and may not wrap around the 16-slot ring; the t = 0 block (a fresh accumulator) is its own MMA because
              // it alone overwrites.  2-3 MMAs per 16-wide k-step instead of 9: the A row is read from shared memory 2-3 times
              // instead of 9, and the serial issue stream shrinks accordingly.
              const uint32_t g = tl + (uint32_t)i;
              const int max_blk = 256 / N_mma;
              int b = n_taps - 1 - t_hi, left = t_hi - t_lo + 1 - (t_lo == 0 ? 1 : 0);
              uint32_t slot = (g - (uint32_t)t_hi) & (AS - 1);
              while (left > 0) {
                int nb = min(left, min(max_blk, (int)(AS - slot)));
                const uint32_t idm = make_idesc(128, nb * N_mma, p.half);
                const uint32_t d_tmem = tmem_base + slot * acc_cols;
                const uint64_t wb = descw0 + (uint64_t)b * bal_d;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16_acc(d_tmem, da + 2 * k, wb + 2 * k, idm);
                b += nb; left -= nb; slot = (slot + nb) & (AS - 1);
              }
              if (t_lo == 0) {   // tap 0: block n_taps-1, slot g, overwrites
                const uint32_t d_tmem = tmem_base + (g & (AS - 1)) * acc_cols;
                const uint64_t wb = descw0 + (uint64_t)(n_taps - 1) * bal_d;
                umma_bf16(d_tmem, da, wb, idesc, 0u);
#pragma unroll
                for (int k = 1; k < BK / 16; ++k) umma_bf16_acc(d_tmem, da + 2 * k, wb + 2 * k, idesc);
              }
            } else if (kbpt == 1) {
              // lean form of the serial issue loop (it paces this layer): one k-block per tap, descriptors advance by adds
              int t = t_lo;
              if (t == 0) {   // first tap of a fresh accumulator overwrites it
                umma_bf16(tmem_base + off, da, dw, idesc, 0u);
#pragma unroll
                for (int k = 1; k < BK / 16; ++k) umma_bf16_acc(tmem_base + off, da + 2 * k, dw + 2 * k, idesc);
                dw += tapw_d;
                off = (off - acc_cols) & off_mask;
                ++t;
              }
#pragma unroll 1
              for (; t <= t_hi; ++t) {
                const uint32_t d_tmem = tmem_base + off;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16_acc(d_tmem, da + 2 * k, dw + 2 * k, idesc);
                dw += tapw_d;
                off = (off - acc_cols) & off_mask;
              }
            } else {
              for (int t = t_lo; t <= t_hi; ++t) {
                const uint32_t d_tmem = tmem_base + off;
                uint64_t a = da, w = dw;
                for (int kb = 0; kb < kbpt; ++kb) {
                  uint64_t ak = a, wk = w;
                  umma_bf16(d_tmem, ak, wk, idesc, (t | kb) != 0 ? 1u : 0u);
#pragma unroll
                  for (int k = 1; k < BK / 16; ++k) {
                    ak += 2; wk += 2;
                    umma_bf16_acc(d_tmem, ak, wk, idesc);
                  }
                  a += sub_d;
                  w += bal_d;
                }
                dw += tapw_d;
                off = (off - acc_cols) & off_mask;
              }
            }
            umma_commit_a(empty_s + sl * 8);
            if (i >= n_taps - 1) umma_commit_a(tfull_s + ((tl + (uint32_t)(i - n_taps + 1)) & (AS - 1)) * 8);
            if (++sl == (uint32_t)S) { sl = 0; ph ^= 1; da = desc0; } else da += stage_d;
          }
          tl += (uint32_t)su.rows;
        }
      } else if (stream) {
        // Serial issue loop: no divisions / modulos, ring slots and their descriptors advance incrementally.
        mbar_wait(wfull, 0);
        tc_fence_after();
        const uint64_t descw0 = desc0 + (uint64_t)((S * stage_bytes) >> 4);
        const uint32_t tapw_d = (uint32_t)(kbpt * b_al) >> 4, bal_d = (uint32_t)b_al >> 4;
        const uint64_t desc_end = desc0 + (uint64_t)S * stage_d;     // one past the last ring slot
        uint32_t tl = 0;
        uint32_t ws = 0, wph = 0;        // next ring entry to wait for: slot, parity
        uint32_t s0 = 0;                 // slot of the oldest row of the current window
        uint64_t da0 = desc0;            // its descriptor
        for (int u = blockIdx.x; u < stream_units(p); u += gridDim.x) {
          const StreamUnit su = decode_unit(p, u);
          if (su.rows <= 0) continue;
          int waited = 0;
          for (int j = 0; j < su.rows; ++j, ++tl) {
            const uint32_t acc = tl & (AS - 1), accph = (tl >> as_sh) & 1;
            mbar_wait_a(tempty_s + acc * 8, accph ^ 1);
            for (const int need = j + n_taps; waited < need; ++waited) {
              mbar_wait_a(full_s + ws * 8, wph);
              if (++ws == (uint32_t)S) { ws = 0; wph ^= 1; }
            }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * acc_cols;
            uint64_t da = da0, dw = descw0;
            for (int t = 0; t < n_taps; ++t) {
              uint64_t a = da, w = dw;
              for (int kb = 0; kb < kbpt; ++kb) {
                uint64_t ak = a, wk = w;
                umma_bf16(d_tmem, ak, wk, idesc, (t | kb) != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 1; k < BK / 16; ++k) {
                  ak += 2; wk += 2;
                  umma_bf16_acc(d_tmem, ak, wk, idesc);
                }
                a += sub_d;
                w += bal_d;
              }
              dw += tapw_d;
              da += stage_d;
              if (da == desc_end) da = desc0;
            }
            umma_commit_a(empty_s + s0 * 8);     // the oldest row of this window is no longer needed
            umma_commit_a(tfull_s + acc * 8);
            if (++s0 == (uint32_t)S) { s0 = 0; da0 = desc0; } else da0 += stage_d;
          }
          // the last n_taps-1 rows of the unit are never the oldest row of a window: release them now
          for (int t = 1; t < n_taps; ++t) {
            umma_commit_a(empty_s + s0 * 8);
            if (++s0 == (uint32_t)S) { s0 = 0; da0 = desc0; } else da0 += stage_d;
          }
        }
      }
      if (cta2) {
        // M = 256 MMAs over the pair: issued by the rank-0 CTA only; commits arrive on the barriers of both CTAs
        if (crank == 0) {
          const uint32_t idesc2 = make_idesc(256, N_mma, p.half);
          int s = 0;
          uint32_t ph = 0, tl = 0;
          for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
            const uint32_t acc = tl & (AS - 1), accph = (tl >> as_sh) & 1;
            mbar_wait_a(tempty_s + acc * 8, accph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * acc_cols;
            for (int g = 0; g < groups; ++g) {
              mbar_wait_a(full_s + s * 8, ph);
              tc_fence_after();
              uint64_t da = desc0 + (uint64_t)(s * stage_d);
              for (int j = 0; j < G; ++j) {
                uint64_t a = da;
                uint32_t dm = d_tmem;
                for (int m = 0; m < MT; ++m) {
                  uint64_t ak = a, bk = da + ab_d;
                  umma2_bf16(dm, ak, bk, idesc2, (g | j) != 0 ? 1u : 0u);
#pragma unroll
                  for (int k = 1; k < BK / 16; ++k) {
                    ak += 2; bk += 2;
                    umma2_bf16_acc(dm, ak, bk, idesc2);
                  }
                  a += sub_d;
                  dm += N_mma;
                }
                da += kb_d;
              }
              umma2_commit_mc(empty_s + s * 8);
              if (++s == S) { s = 0; ph ^= 1; }
            }
            umma2_commit_mc(tfull_s + acc * 8);
          }
        }
      } else if (dysh) {
        // dy-sharing: stage = one A box (all rows the column's taps touch) + the column's weight tiles; tap j reads the box
        // from row j on (descriptor offset j * TW * BK * 2 bytes - a whole number of swizzle atoms since TW >= 8)
        const int n_cols = p.n_cols;
        const uint32_t row_d = (uint32_t)(p.TW * BK * 2) >> 4, aal_d = (uint32_t)a_box_al >> 4, bal_d = (uint32_t)b_al >> 4;
        const uint64_t descw0 = desc0 + (uint64_t)((S * stage_bytes) >> 4);   // resident weights: tile (tap t, k-block kb) at (t * kbpt + kb) * b_al
        if (wres) {
          mbar_wait(wfull, 0);
          tc_fence_after();
        }
        int s = 0;
        uint32_t ph = 0, tl = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tl) {
          const uint32_t acc = tl & (AS - 1), accph = (tl >> as_sh) & 1;
          mbar_wait_a(tempty_s + acc * 8, accph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * acc_cols;
          const int cb = (tile % p.n_phase) * n_cols;
          uint32_t first = 0;   // 0 until the first MMA of every sub-tile has been issued
          for (int c = 0; c < n_cols; ++c) {
            const int cn = p.col_n[cb + c];
            for (int kb = 0; kb < kbpt; ++kb) {
              mbar_wait_a(full_s + s * 8, ph);
              tc_fence_after();
              const uint64_t da = desc0 + (uint64_t)(s * stage_d);
              const uint32_t bstep = wres ? (uint32_t)(p.col_ts[cb + c] * kbpt) * bal_d : bal_d;
              uint64_t aj = da, bj = wres ? descw0 + (uint64_t)((p.col_t0[cb + c] * kbpt + kb) * (int)bal_d) : da + aal_d;
              for (int j = 0; j < cn; ++j) {
                uint64_t a = aj;
                uint32_t dm = d_tmem;
                for (int m = 0; m < MT; ++m) {
                  uint64_t ak = a, bk = bj;
                  umma_bf16(dm, ak, bk, idesc, first);
#pragma unroll
                  for (int k = 1; k < BK / 16; ++k) {
                    ak += 2; bk += 2;
                    umma_bf16_acc(dm, ak, bk, idesc);
                  }
                  a += sub_d;
                  dm += N_mma;
                }
                first = 1;
                aj += row_d;
                bj += bstep;
              }
              umma_commit_a(empty_s + s * 8);
              if (++s == S) { s = 0; ph ^= 1; }
            }
          }
          umma_commit_a(tfull_s + acc * 8);
        }
      }
      int s = 0;
      uint32_t ph = 0, tl = 0;
      for (int tile = blockIdx.x; !stream && !dysh && !cta2 && tile < total_tiles; tile += gridDim.x, ++tl) {
        const uint32_t acc = tl & (AS - 1), accph = (tl >> as_sh) & 1;
        mbar_wait_a(tempty_s + acc * 8, accph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_cols;
        for (int g = 0; g < groups; ++g) {
          mbar_wait_a(full_s + s * 8, ph);
          tc_fence_after();
          uint64_t da = desc0 + (uint64_t)(s * stage_d);
          for (int j = 0; j < G; ++j) {
            uint64_t a = da + (uint64_t)(m_begin * sub_d);
            uint32_t dm = d_tmem + m_begin * N_mma;
            for (int m = m_begin; m < m_end; ++m) {
              uint64_t ak = a, bk = da + ab_d;
              umma_bf16(dm, ak, bk, idesc, (g | j) != 0 ? 1u : 0u);
#pragma unroll
              for (int k = 1; k < BK / 16; ++k) {
                ak += 2; bk += 2;
                umma_bf16_acc(dm, ak, bk, idesc);
              }
              a += sub_d;
              dm += N_mma;
            }
            da += kb_d;
          }
          umma_commit_a(empty_s + s * 8);
          if (++s == S) { s = 0; ph ^= 1; }
        }
        umma_commit_a(tfull_s + acc * 8);
      }
    }
  } else if (warp >= 12 && fuse_in) {
    // ================================ fused input normalisation (warps 12..15, launched only when p.fuse_in) ================
    // The A operand is the PRODUCER's raw convolution output (16-bit NHWC, no halo): each ring slot (one input row, 128
    // pixels x 64 channels, SWIZZLE_128B) is normalised in place - y = max(a_c x + b_c, 0) with a, b from the producer's
    // InstanceNorm statistics - between the TMA landing (full) and the MMAs (xfull), so the separate apply pass over the
    // full-resolution tensor (read 2 B + write 2 B per element of HBM traffic) does not exist.  Every input row is loaded once
    // in this mode, so it is also transformed once.  Reflection padding: rows are mirrored by the producer's row index,
    // the <= 4 out-of-range pixels at the left / right frame edge arrive as TMA zero-fill and are overwritten here with their
    // mirror pixels (of the already transformed row).  A thread always owns the same logical 16-byte chunk (8 channels), so its
    // 16 constants live in registers; the physical chunk is lc ^ (row & 7), and a warp covers four whole 128-byte rows per
    // access - conflict-free.
    const int tw = warp - 12, lc = lane & 7, rsub = lane >> 3;
    const uint32_t xfull_s = smem_u32(xfull);
    const int dx0 = (int)(signed char)(p.tap_packed[0] & 0xff);
    float ca[8], cb[8];
    int cur_n = -1, s = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < stream_units(p); u += gridDim.x) {
      const StreamUnit su = decode_unit(p, u);
      if (su.rows <= 0) continue;
      if (su.n != cur_n) {
        cur_n = su.n;
        const double inv_cnt = 1.0 / ((double)p.in_H * (double)p.in_W);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = lc * 8 + j;
          ca[j] = cb[j] = 0.f;                      // channels past in_C are TMA zero-fill and stay zero
          if (c < p.in_C) {
            const double s1 = p.in_stats[((size_t)su.n * p.in_C + c) * 2], s2 = p.in_stats[((size_t)su.n * p.in_C + c) * 2 + 1];
            const double mean_d = s1 * inv_cnt;
            const float var = fmaxf((float)(s2 * inv_cnt - mean_d * mean_d), 0.f);
            ca[j] = p.in_gamma[c] * rsqrtf(var + p.in_eps);
            cb[j] = p.in_beta[c] - (float)mean_d * ca[j];
          }
        }
      }
      const int xs = su.x0 + dx0;                   // frame x of the slot's first pixel
      const int n_in = su.rows + n_taps - 1;
      for (int i = 0; i < n_in; ++i) {
        mbar_wait_a(full_s + s * 8, ph);
        const uint32_t base = smem_s + s * stage_bytes;
        if (lc * 8 < p.in_C) {
          if (p.half) xform_rows<true>(base, tw, rsub, lc, ca, cb, p.in_relu);
          else xform_rows<false>(base, tw, rsub, lc, ca, cb, p.in_relu);
        }
        const bool edge = xs < 0 || xs + 128 > p.in_W;
        if (edge) epi_bar_sync_id(3, 128);          // edge strip: the mirror copy below reads rows other warps transformed
        if (tw == 0 && edge) {
          // mirror pixels: frame x = -k  <- x = k;  x = W-1+k <- x = W-1-k  (k = 1..4), one 16-byte chunk per lane
          const int k = rsub + 1;
          if (xs < 0 && -k - xs >= 0) {
            const int rd = -k - xs, rs = k - xs;
            if (rs < 128) sts128(base + rd * 128u + ((lc ^ (rd & 7)) << 4), lds128(base + rs * 128u + ((lc ^ (rs & 7)) << 4)));
          }
          if (xs + 128 > p.in_W) {
            const int rd = p.in_W - 1 + k - xs, rs = p.in_W - 1 - k - xs;
            if (rd < 128 && rs >= 0) sts128(base + rd * 128u + ((lc ^ (rd & 7)) << 4), lds128(base + rs * 128u + ((lc ^ (rs & 7)) << 4)));
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&xfull[s]);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
    (void)xfull_s;
  } else if (warp >= 12) {
    // ================================ apply rider (warps 12..15, launched only when p.rider.on) ========
    if (!fuse_in && (p.rider.on & 1)) apply_rider_run(p.rider, (int)threadIdx.x - 384, (int)blockIdx.x, (int)gridDim.x);
  } else if (warp < 7 || ((p.epi8 || e_rowconv) && warp <= 10)) {
    // ================================ epilogue (warps 2..5, and 7..10 for bf16 NHWC) ====
    const int eset = warp >= 7 ? 1 : 0;      // column half this warp converts out of TMEM
    const int ETH = p.epi8 ? 256 : 128;   // epilogue threads
    const int lg = warp & 3;                 // TMEM lane group this warp may read
    const int row = lg * 32 + lane;          // sub-tile row == TMEM lane
    const int et = eset * 128 + (warp - (eset ? 7 : 2)) * 32 + lane;   // index among the epilogue threads
    const int col_half = ((p.N_mma >> 1) + 31) & ~31;             // set 0: [0, col_half), set 1: [col_half, N_mma)
    const int col_begin = (ETH == 256 && eset) ? min(col_half, p.N_mma) : 0;
    const int col_end = (ETH == 256 && !eset) ? min(col_half, p.N_mma) : p.N_mma;
    const int lw = 31 - __clz(p.TW);         // TW is a power of two
    const uint32_t stg_s = smem_u32(stg);    // staging tile, shared-space address
    // Fused InstanceNorm statistics: accumulated in the STORE phase, from the 16-byte vector each thread has just read back
    // from the staging tile for its global store (no second pass over the tile, no extra shared-memory loads).  A thread
    // always owns the same 8-channel chunk, so 8 sums + 8 sums of squares persist in registers across tiles; they are
    // reduced over the block through the (then idle) staging tile when the image changes and at the end.
    float sa[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sa[j] = sq[j] = 0.f;
    int st_n = -1, st_c = 0, st_ch = -1, st_cw = 0, st_lpr = 8;
    auto flush_stats = [&]() {    // called by ALL epilogue threads while the staging tile is free
      // fixed-order block reduction (no shared-memory atomics): every thread parks its 8 partials, then thread c < cw adds up
      // the threads that own channel c in thread order; the per-CTA result meets the other CTAs' in an fp64 atomic.  With the
      // static tile -> CTA assignment this makes the statistics - and every frame - bit-reproducible, like the direct epilogue.
      if (st_n >= 0) {
        float* scr = reinterpret_cast<float*>(stg);
        for (int pass = 0; pass < 2; ++pass) {
          epi_bar_sync(ETH);
#pragma unroll
          for (int j = 0; j < 8; ++j) scr[et * 8 + j] = pass ? sq[j] : sa[j];
          epi_bar_sync(ETH);
          for (int c = et; c < st_cw; c += ETH) {   // (a layer wider than the epilogue has threads: several channels per thread)
            const int ch = c >> 3, j = c & 7;
            float sum = 0.f;
            for (int t = ch; t < ETH; t += st_lpr) sum += scr[t * 8 + j];
            atomicAdd(p.stats + ((size_t)st_n * p.Cout + st_c + c) * 2 + pass, (double)sum);
          }
        }
        epi_bar_sync(ETH);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sa[j] = sq[j] = 0.f;
    };
    uint32_t tl = 0;
    TileWalk walk;
    TileCoord tc;
    walk_init(p, walk);
    if (e_nhwc && e_direct) {
      // ============ direct epilogue: TMEM -> registers -> 32-byte global stores; no staging tile, no block barrier per tile
      // A lane owns one pixel (TMEM lane) and walks its warp set's column range 32 columns per tcgen05.ld wait; every 16
      // channels leave as ONE 256-bit store (a full sector per lane).  Statistics: chunk_stats() above, accumulated per
      // thread in shared-memory slots; the four warps of a set are combined in fixed order when the image changes and the
      // per-CTA partial meets the other CTAs' in an fp64 atomic (order-independent to ~1e-16, i.e. the same fp32 statistics
      // whichever CTA arrives first - two runs of a frame, or a frame alone / inside a batch, give the same bits).
      float* accs = reinterpret_cast<float*>(stg);
      const int nch = (col_end - col_begin + 15) >> 4;
      const int wi = warp - (eset ? 7 : 2);
      const bool do_stats = p.stats && !(p.dbg & 1);
      for (int k = 0; k < nch; ++k) accs[k * ETH + et] = 0.f;
      int st_n = -1, st_c = 0;
      auto flush = [&]() {   // called by ALL epilogue threads (uniform)
        if (st_n >= 0) {
          epi_bar_sync(ETH);
          if (wi == 0) {
            for (int k = 0; k < nch; ++k) {
              const float* a4 = accs + k * ETH + eset * 128 + lane;
              const float s = ((a4[0] + a4[32]) + a4[64]) + a4[96];
              const int c = st_c + col_begin + k * 16 + (lane >> 1);
              if (c < p.Cout) atomicAdd(p.stats + ((size_t)st_n * p.Cout + c) * 2 + (lane & 1), (double)s);
            }
          }
          epi_bar_sync(ETH);
        }
        for (int k = 0; k < nch; ++k) accs[k * ETH + et] = 0.f;
      };
      __nv_bfloat16* const out_base = reinterpret_cast<__nv_bfloat16*>(p.out0);
      const bool al32 = ((reinterpret_cast<uintptr_t>(out_base) & 31) == 0) && (((p.out_cstride * 2) & 31) == 0);
      for (; walk_next(p, walk, tc, total_tiles); ++tl) {
        const uint32_t acc = tl & (AS - 1), accph = (tl >> as_sh) & 1;
        mbar_wait(&tfull[acc], accph);
        tc_fence_after();
        const int cbase = tc.nt * p.N_mma;
        if (do_stats && (tc.n != st_n || cbase != st_c)) {
          flush();
          st_n = tc.n;
          st_c = cbase;
        }
        const int vy = tc.valid ? min(p.TH, p.Ho - tc.y0) : 0, vx = min(p.TW, p.Wo - tc.x0);
        const int cw = min(p.N_mma, p.Cout - cbase);   // channels this tile owns (multiple of 8)
        const int oyb = p.ph_oy[tc.ph], oxb = p.ph_ox[tc.ph];
        const uint32_t taddr0 = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * acc_cols;
        const bool full_tile = (vy == p.TH) && (vx == p.TW);
        // Loop nest: 16-channel chunk OUTER, the tile's MT sub-tiles INNER - the chunk's sums / sums of squares run in
        // registers over the sub-tiles, so the cross-lane butterfly (the bulk of the statistics' instructions) is paid once
        // per chunk and TILE, not once per sub-tile.
        // (chunk, sub-tile) pairs flattened into one software-pipelined sequence: the tcgen05.ld of step i + 1 is in flight
        // while step i is packed, stored and accumulated
        const int mt_sh = 31 - __clz(MT);
        const int steps = (p.dbg & 4) ? 0 : nch * MT;
        float ssum[16], ssq[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) ssum[j] = ssq[j] = 0.f;
        uint32_t r[2][16];
        if (steps > 0) tmem_ld16(taddr0 + col_begin, r[0]);
#pragma unroll 1
        for (int i0 = 0; i0 < steps; i0 += 2) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = i0 + u;
            if (i >= steps) break;
            const int mm = i & (MT - 1), c0 = col_begin + ((i >> mt_sh) << 4);
            tmem_ld_wait();
            if (i + 1 < steps) {
              const int i1 = i + 1;
              tmem_ld16(taddr0 + (i1 & (MT - 1)) * p.N_mma + col_begin + ((i1 >> mt_sh) << 4), r[u ^ 1]);
            }
            const int tr = mm * 128 + row, ty = tr >> lw, tx = tr & (p.TW - 1);
            const bool valid = full_tile || (ty < vy && tx < vx);
            const int left = cw - c0;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[u][j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cbase + c0 + j < p.Cout) v[j] += p.bias[cbase + c0 + j];
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (p.out_f32) {
              // fp32 NHWC output (the "fp16" plan's residual-block convs): 16 floats = two 256-bit stores, statistics of the
              // unrounded values
              if (valid) {
                if (!(p.dbg & 2)) {
                  const int oy = (tc.y0 + ty) * p.out_mul + oyb, ox = (tc.x0 + tx) * p.out_mul + oxb;
                  float* const o32 = reinterpret_cast<float*>(p.out0) +
                                     (size_t)((uint32_t)((tc.n * p.Hout + oy) * p.Wout + ox)) * (uint32_t)p.out_cstride + cbase + c0;
                  const uint32_t* vr = reinterpret_cast<const uint32_t*>(v);
                  if (left >= 8) stg256(o32, make_uint4(vr[0], vr[1], vr[2], vr[3]), make_uint4(vr[4], vr[5], vr[6], vr[7]));
                  if (left >= 16) stg256(o32 + 8, make_uint4(vr[8], vr[9], vr[10], vr[11]), make_uint4(vr[12], vr[13], vr[14], vr[15]));
                }
                if (do_stats) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) { ssum[j] += v[j]; ssq[j] = fmaf(v[j], v[j], ssq[j]); }
                }
              }
            } else {
            uint4 q0, q1;
            q0.x = pack16x2(v[0], v[1], p.half);   q0.y = pack16x2(v[2], v[3], p.half);
            q0.z = pack16x2(v[4], v[5], p.half);   q0.w = pack16x2(v[6], v[7], p.half);
            q1.x = pack16x2(v[8], v[9], p.half);   q1.y = pack16x2(v[10], v[11], p.half);
            q1.z = pack16x2(v[12], v[13], p.half); q1.w = pack16x2(v[14], v[15], p.half);
            if (valid) {
              if (!(p.dbg & 2)) {
                const int oy = (tc.y0 + ty) * p.out_mul + oyb, ox = (tc.x0 + tx) * p.out_mul + oxb;
                __nv_bfloat16* const optr =
                    out_base + (size_t)((uint32_t)((tc.n * p.Hout + oy) * p.Wout + ox)) * (uint32_t)p.out_cstride + cbase + c0;
                if (left >= 16) {
                  if (al32) stg256(optr, q0, q1);
                  else { *reinterpret_cast<uint4*>(optr) = q0; *reinterpret_cast<uint4*>(optr + 8) = q1; }
                } else if (left >= 8) {
                  *reinterpret_cast<uint4*>(optr) = q0;
                }
              }
              if (do_stats) {
                // statistics of the STORED (16-bit-rounded) values; pixels outside the tile's valid extent do not count
                const uint32_t qq[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float2 f = unpack16x2(qq[j], p.half);
                  ssum[2 * j] += f.x;     ssq[2 * j] = fmaf(f.x, f.x, ssq[2 * j]);
                  ssum[2 * j + 1] += f.y; ssq[2 * j + 1] = fmaf(f.y, f.y, ssq[2 * j + 1]);
                }
              }
            }
            }
            if (mm == MT - 1) {   // last sub-tile of this chunk: one cross-lane butterfly per chunk and tile
              if (do_stats) {
                accs[(i >> mt_sh) * ETH + et] += chunk_stats2(ssum, ssq, lane);
#pragma unroll
                for (int j = 0; j < 16; ++j) ssum[j] = ssq[j] = 0.f;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {   // accumulators drained: MMA may reuse them (pair: the barrier lives in the rank-0 CTA)
          if constexpr (CTA2) mbar_arrive_cluster(mapa_u32(tempty_s + acc * 8, 0));
          else mbar_arrive(&tempty[acc]);
        }
      }
      if (do_stats) flush();
    }
    if (e_nhwc && e_tma) {
      bool done = false;
      if constexpr (!CTA2) {
        if (e_pp) { epilogue_tma<false, true>(p, stg_s, tfull, tempty, tmem_base); done = true; }
      }
      if (!done) epilogue_tma<CTA2, false>(p, stg_s, tfull, tempty, tmem_base);
    }
    // ---- staged epilogue (VST_EPI_DIRECT=0), row-conv and fp32 epilogues: coalesced store mapping below
    const bool rc_pingpong = e_rowconv;   // the two warp sets take alternate tiles (own staging tile)
    for (; !(e_nhwc && (e_direct || e_tma)) && walk_next(p, walk, tc, total_tiles); ++tl) {
      if (rc_pingpong && (int)(tl & 1) != eset) continue;
      const uint32_t acc = tl & (AS - 1), accph = (tl >> as_sh) & 1;
      mbar_wait(&tfull[acc], accph);
      tc_fence_after();
      const int cbase = tc.nt * p.N_mma;
      const int vy = tc.valid ? min(p.TH, p.Ho - tc.y0) : 0, vx = min(p.TW, p.Wo - tc.x0);
      const bool full_tile = (vy == p.TH) && (vx == p.TW);
      const bool do_stats = p.stats && e_nhwc && !(p.dbg & 1);
      if (do_stats && (tc.n != st_n || cbase != st_c)) {   // uniform over the epilogue threads; staging is free here
        flush_stats();
        st_n = tc.n;
        st_c = cbase;
      }

      // Staged ping-pong (p.epi_spp): the two warp sets take ALTERNATE sub-tiles, each with all N_mma columns, its own staging
      // tile and its own named barrier - two independent TMEM -> staging -> store chains per CTA instead of one chain that all
      // 256 threads walk in lock step (two block barriers per sub-tile, every phase's latency exposed).
      const bool spp = e_nhwc && p.epi_spp != 0;
      const int ETHs = spp ? 128 : ETH;                       // threads that share one sub-tile
      const int el = spp ? (et & 127) : et;                   // index among them
      const int cbeg = spp ? 0 : col_begin, cend = spp ? p.N_mma : col_end;
      const uint32_t stg_me = stg_s + ((spp && eset) ? (uint32_t)stg_one : 0u);
      for (int m = 0; m < MT; ++m) {
        if (spp && (m & 1) != eset) continue;
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + acc * acc_cols + m * p.N_mma;
        const bool last_m = spp ? (m + 2 >= MT) : (m == MT - 1);
        if (e_nhwc) {
          // ---- TMEM -> bf16 staging tile (row-major, padded pitch); 32 columns per wait
          const uint32_t srow = stg_me + (uint32_t)row * stg_pitch;
          for (int c0 = cbeg; c0 < ((p.dbg & 4) ? 0 : cend); c0 += 32) {
            uint32_t r[32];
            const bool two = (c0 + 16 < cend);
            tmem_ld16(taddr + c0, *reinterpret_cast<uint32_t(*)[16]>(&r[0]));
            if (two) tmem_ld16(taddr + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&r[16]));
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 1 && !two) break;
              float v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[h * 16 + j]);
              if (p.bias) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (cbase + c0 + h * 16 + j < p.Cout) v[j] += p.bias[cbase + c0 + h * 16 + j];
              }
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              uint4 q0, q1;
              q0.x = pack_bf16x2(v[0], v[1]);   q0.y = pack_bf16x2(v[2], v[3]);
              q0.z = pack_bf16x2(v[4], v[5]);   q0.w = pack_bf16x2(v[6], v[7]);
              q1.x = pack_bf16x2(v[8], v[9]);   q1.y = pack_bf16x2(v[10], v[11]);
              q1.z = pack_bf16x2(v[12], v[13]); q1.w = pack_bf16x2(v[14], v[15]);
              sts128(srow + (c0 + h * 16) * 2, q0);
              sts128(srow + (c0 + h * 16) * 2 + 16, q1);
            }
          }
          if (last_m) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {   // accumulators drained: MMA may reuse them (pair: the barrier lives in the rank-0 CTA)
              if constexpr (CTA2) mbar_arrive_cluster(mapa_u32(tempty_s + acc * 8, 0));
              else mbar_arrive(&tempty[acc]);
            }
          }
          if (spp) epi_bar_sync_id(2 + eset, 128); else epi_bar_sync(ETH);
          const int rbase = m * 128;   // first tile row of this sub-tile
          // ---- coalesced 16-byte stores: LPR lanes per pixel, 128/LPR pixels per pass
          {
            const int cw = min(p.N_mma, p.Cout - cbase);   // channels this tile owns (multiple of 8)
            const int cpr = cw >> 3;                       // 16-byte chunks per pixel
            const int lsh = cpr <= 8 ? 3 : cpr <= 16 ? 4 : 5, lpr = 1 << lsh;
            const int ch = el & (lpr - 1), r0 = el >> lsh, rstep = ETHs >> lsh;
            st_ch = ch < cpr ? ch : -1;
            st_cw = cw;
            st_lpr = lpr;
            if (ch < cpr && !(p.dbg & 2)) {
              __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out0) + cbase + ch * 8;
              const int oyb = p.ph_oy[tc.ph], oxb = p.ph_ox[tc.ph];
              const int pix_n = tc.n * p.Hout;
              const uint32_t sbase = stg_me + ch * 16;
#pragma unroll 4
              for (int r = r0; r < 128; r += rstep) {
                const int tr = rbase + r, ty = tr >> lw, tx = tr & (p.TW - 1);
                if (full_tile || (ty < vy && tx < vx)) {
                  const int oy = (tc.y0 + ty) * p.out_mul + oyb, ox = (tc.x0 + tx) * p.out_mul + oxb;
                  const uint32_t pix = (uint32_t)((pix_n + oy) * p.Wout + ox);
                  const uint4 q = lds128(sbase + (uint32_t)r * stg_pitch);
                  *reinterpret_cast<uint4*>(obase + (size_t)pix * (uint32_t)p.out_cstride) = q;
                  if (do_stats) {
                    const float2 f0 = bf16x2_to_f2(q.x), f1 = bf16x2_to_f2(q.y), f2 = bf16x2_to_f2(q.z), f3 = bf16x2_to_f2(q.w);
                    sa[0] += f0.x; sq[0] = fmaf(f0.x, f0.x, sq[0]); sa[1] += f0.y; sq[1] = fmaf(f0.y, f0.y, sq[1]);
                    sa[2] += f1.x; sq[2] = fmaf(f1.x, f1.x, sq[2]); sa[3] += f1.y; sq[3] = fmaf(f1.y, f1.y, sq[3]);
                    sa[4] += f2.x; sq[4] = fmaf(f2.x, f2.x, sq[4]); sa[5] += f2.y; sq[5] = fmaf(f2.y, f2.y, sq[5]);
                    sa[6] += f3.x; sq[6] = fmaf(f3.x, f3.x, sq[6]); sa[7] += f3.y; sq[7] = fmaf(f3.y, f3.y, sq[7]);
                  }
                }
              }
            }
          }
          if (spp) epi_bar_sync_id(2 + eset, 128); else epi_bar_sync(ETH);  // staging tile free for the next sub-tile
        } else if (e_rowconv) {
          // ---- D[x'][(kx,co)] -> staging (fp32), then out[x][co] = bias + sum_kx D[x+kx][kx*rc_co+co]
          // The per-pixel tail (27 shared loads, tanh, 6 stores) is a long dependent instruction stream: with one warp per
          // scheduler it, not the MMAs, paced this layer - so warp sets 0 / 1 work on alternate tiles.
          const int el = et & 127;                                   // index inside this warp set
          const uint32_t rstg = stg_s + (uint32_t)eset * (128 * RC_LD * 4);
          for (int c0 = 0; c0 < p.N_mma; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) sts_f32(rstg + (uint32_t)(row * RC_LD + c0 + j) * 4, __uint_as_float(r[j]));
          }
          if (last_m) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          epi_bar_sync_id(1 + eset, 128);
          const int x = tc.x0 + el, y = tc.y0 + m;   // sub-tile m = output row y0 + m
          if (el < p.tile_step_x && x < p.Wo && y < p.Ho && !(p.dbg & 8)) {
            const size_t plane = (size_t)p.Hout * p.Wout, pix = (size_t)y * p.Wout + x;
            float* o = reinterpret_cast<float*>(p.out0);
            if (p.rc_k == 9 && p.rc_co == 3) {
              float a[3];
#pragma unroll
              for (int co = 0; co < 3; ++co) a[co] = p.bias ? p.bias[co] : 0.f;
#pragma unroll
              for (int kx = 0; kx < 9; ++kx) {
                const uint32_t rowa = rstg + (uint32_t)((el + kx) * RC_LD + kx * 3) * 4;
#pragma unroll
                for (int co = 0; co < 3; ++co) a[co] += lds_f32(rowa + co * 4);
              }
#pragma unroll
              for (int co = 0; co < 3; ++co) {
                if (p.dbg & 16) { float th; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(a[co] * (1.f / 255.f))); a[co] = fmaf(th, 150.f, 127.5f); }
                else a[co] = epi_act(a[co], p.act);
                if (o) o[((size_t)tc.n * 3 + co) * plane + pix] = a[co];
              }
              if (p.out_u8) {
                uint8_t* u = p.out_u8 + ((size_t)tc.n * plane + pix) * 3;
#pragma unroll
                for (int co = 0; co < 3; ++co) u[2 - co] = (uint8_t)fminf(fmaxf(a[co], 0.f), 255.f);
              }
            } else {
              for (int co = 0; co < p.rc_co; ++co) {
                float a = p.bias ? p.bias[co] : 0.f;
                for (int kx = 0; kx < p.rc_k; ++kx) a += lds_f32(rstg + (uint32_t)((el + kx) * RC_LD + kx * p.rc_co + co) * 4);
                a = epi_act(a, p.act);
                if (o) o[((size_t)tc.n * p.rc_co + co) * plane + pix] = a;
                if (p.out_u8 && co < 3) p.out_u8[((size_t)tc.n * plane + pix) * 3 + (2 - co)] = (uint8_t)fminf(fmaxf(a, 0.f), 255.f);
              }
            }
          }
          epi_bar_sync_id(1 + eset, 128);
        } else {
          // ---- TG_EPI_F32_NCHW: direct per-thread stores (coalesced along x across lanes)
          const int tr = m * 128 + row, r_ty = tr >> lw, r_tx = tr & (p.TW - 1);
          const int y = tc.y0 + r_ty, x = tc.x0 + r_tx;
          const bool valid = (y < p.Ho) && (x < p.Wo);
          const int oy = y * p.out_mul + p.ph_oy[tc.ph], ox = x * p.out_mul + p.ph_ox[tc.ph];
          float* o = reinterpret_cast<float*>(p.out0);
          const size_t plane = (size_t)p.Hout * p.Wout, pix = (size_t)oy * p.Wout + ox;
          for (int c0 = 0; c0 < p.N_mma; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(taddr + c0, r);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int c = cbase + c0 + j;
                if (c < p.Cout) {
                  float a = __uint_as_float(r[j]) + (p.bias ? p.bias[c] : 0.f);
                  a = epi_act(a, p.act);
                  if (o) o[((size_t)tc.n * p.Cout + c) * plane + pix] = a;
                  // Inference byte path: clamp(0,255), astype(uint8) truncation, RGB->BGR (RC/utilities.py:219-224)
                  if (p.out_u8 && c < 3) p.out_u8[((size_t)tc.n * plane + pix) * 3 + (2 - c)] = (uint8_t)fminf(fmaxf(a, 0.f), 255.f);
                }
              }
            }
          }
          if (last_m) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
        }
      }
    }
    if (p.stats && e_nhwc && !e_direct && !e_tma && !(p.dbg & 1)) flush_stats();
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CTA2) cluster_sync_all();   // neither CTA frees TMEM or exits while the other may still read / signal it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc2(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int BK) {
  return BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : BK == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

int make_tmap_act(CUtensorMap* out, const void* base, int C, int X, int Y, int N, int P, size_t pix_stride_elems,
                  size_t row_stride_elems, size_t img_stride_elems, size_t plane_stride_elems, int BK, int TW, int TH) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return VST_ECUDA;
  }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)N, (cuuint64_t)P};
  cuuint64_t strides[4] = {pix_stride_elems * 2, row_stride_elems * 2, img_stride_elems * 2, plane_stride_elems * 2};
  cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 4; ++i)
    if (strides[i] % 16 != 0) {
      set_error("tensor map stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)strides[i]);
      return VST_EINVAL;
    }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(BK), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(act C=%d X=%d Y=%d N=%d P=%d BK=%d TW=%d TH=%d) -> %d", C, X, Y, N, P, BK, TW, TH, (int)r);
    return VST_ECUDA;
  }
  return VST_OK;
}

int make_tmap_act_generic(CUtensorMap* out, const void* base, int d0, int d1, int d2, int d3, int d4, size_t s1, size_t s2,
                          size_t s3, size_t s4, int b0, int b1, int b2, int b3) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return VST_ECUDA;
  }
  cuuint64_t dims[5] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2, (cuuint64_t)d3, (cuuint64_t)d4};
  cuuint64_t strides[4] = {s1 * 2, s2 * 2, s3 * 2, s4 * 2};
  cuuint32_t box[5] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2, (cuuint32_t)b3, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < 4; ++i)
    if (strides[i] % 16 != 0) {
      set_error("tensor map stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)strides[i]);
      return VST_EINVAL;
    }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(b0), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(generic dims %d %d %d %d %d box %d %d %d %d) -> %d", d0, d1, d2, d3, d4, b0, b1, b2, b3, (int)r);
    return VST_ECUDA;
  }
  return VST_OK;
}

int make_tmap_wgt(CUtensorMap* out, const void* base, int K, int rows, int BK, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return VST_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(BK), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(wgt K=%d rows=%d BK=%d box_rows=%d) -> %d", K, rows, BK, box_rows, (int)r);
    return VST_ECUDA;
  }
  return VST_OK;
}

// Output map of the TMA-store epilogue: 4-D (C, X, Y, N) 16-bit view, box (bc, bx, by, 1), swizzle by the box's channel width
int make_tmap_out(CUtensorMap* out, const void* base, int C, int X, int Y, int N, size_t pix_stride_elems, size_t row_stride_elems,
                  size_t img_stride_elems, int bc, int bx, int by) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return VST_ECUDA;
  }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)X, (cuuint64_t)Y, (cuuint64_t)N};
  cuuint64_t strides[3] = {pix_stride_elems * 2, row_stride_elems * 2, img_stride_elems * 2};
  cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bx, (cuuint32_t)by, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i)
    if (strides[i] % 16 != 0) {
      set_error("output tensor map stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)strides[i]);
      return VST_EINVAL;
    }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bc), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(out C=%d X=%d Y=%d N=%d box %d %d %d) -> %d", C, X, Y, N, bc, bx, by, (int)r);
    return VST_ECUDA;
  }
  return VST_OK;
}

// Dynamic shared memory a tap-GEMM CTA may plan with (pipeline stages + staging + barriers).  VST_TG_SMEM_KB lowers it so
// that blocks of another stream's HBM-bound kernels (each needs ~2.5 KB + the 1 KB per-block reservation) fit next to the
// persistent CTA in the SM's 228 KB.
int tg_smem_budget() {
  static const int kb = [] { const char* e = getenv("VST_TG_SMEM_KB"); const int v = e ? atoi(e) : 212; return v < 96 ? 96 : v > 220 ? 220 : v; }();
  return kb * 1024;
}

int choose_mt(int N_mma) {
  // 2 accumulator stages x MT x N_mma fp32 columns must fit the 512 TMEM columns
  int mt = 1;
  while (mt < 4 && 2 * (mt * 2) * N_mma <= 512) mt *= 2;
  return mt;
}

void choose_tile(int Ho, int Wo, int MT, int* TW, int* TH) {
  // candidates in order of preference; a later one wins only with strictly less overhang
  const int cand[6] = {32, 64, 16, 128, 8, 256};
  const int M = 128 * MT;
  long best = -1;
  for (int i = 0; i < 6; ++i) {
    const int tw = cand[i], th = M / tw;
    if (th < 1 || th > 256) continue;
    const long cover = (long)cdiv(Wo, tw) * tw * (long)cdiv(Ho, th) * th;
    if (best < 0 || cover < best) {
      best = cover;
      *TW = tw;
      *TH = th;
    }
  }
}

// Row-streaming eligibility: one phase, one N tile, shared weights, and taps that are a pure row stencil
// (same dx / plane, dy increasing by one).  VST_STREAM selects the policy:
//   0  never stream
//   2  (default) accumulator-ring streaming where the layer fits it (n_taps + slack accumulators in 16 TMEM slots: the
//      ConvTanh row convolution and its data gradient) - measured 0.87 -> 0.61 ms on deconv3 at 4x1080p
//   1  shared-memory row ring for every eligible layer (EXPERIMENTAL: correct, but the ring holds the n_taps-row window so
//      TMA runs only 1-2 rows ahead, and one 128-pixel row per tile does not amortise the epilogue - slower on conv1)
//   3  accumulator ring where possible, else the row ring
static int epi_staging_bytes(const TapGemmParams& p) {
  return p.epi_mode == TG_EPI_BF16_NHWC ? (p.epi_tma ? p.epi_nbuf * p.N_mma * 256
                                           : p.epi_direct ? TG_DIRECT_SCRATCH : (p.epi_spp ? 2 : 1) * std::max(128 * (p.N_mma * 2 + 16), TG_DIRECT_SCRATCH))
         : p.epi_mode == TG_EPI_ROWCONV ? 2 * 128 * 33 * 4 : 0;
}
static int stream_mode_env() {
  static const int m = [] { const char* e = getenv("VST_STREAM"); return e ? atoi(e) : 2; }();
  return m;
}
bool tapgemm_stream_enabled() { return stream_mode_env() != 0; }
bool tapgemm_try_stream(TapGemmParams& p, int BK) {
  const bool off = !tapgemm_stream_enabled();
  p.stream = 0; p.merge_taps = 0;
  if (off || p.n_phase != 1 || p.n_ntile > 1 || p.b_img_rows != 0 || p.n_taps < 5 || p.n_taps > 16) return false;
  if (p.epi_mode == TG_EPI_F32_NCHW) return false;
  for (int t = 1; t < p.n_taps; ++t)
    if (p.tap_dx[t] != p.tap_dx[0] || p.tap_pl[t] != p.tap_pl[0] || p.tap_dy[t] != p.tap_dy[0] + t) return false;
  // mode 2 (accumulator ring): all n_taps live accumulators + slack must fit 16 TMEM slots
  const bool acc_ring = stream_mode_env() >= 2 && 16 * p.N_mma <= 512 && p.n_taps <= 12;
  if (!acc_ring && stream_mode_env() == 2) return false;
  if (!acc_ring) {  // the ring must hold the tap window plus at least two rows in flight
    const int stg_bytes = epi_staging_bytes(p);
    const int b_al = (p.N_mma * BK * 2 + 1023) & ~1023;
    const int slot = p.kb_per_tap * 128 * BK * 2, w_region = p.n_taps * p.kb_per_tap * b_al;
    const int ring = (tg_smem_budget() - stg_bytes - 2048 - w_region) / slot;
    if (ring < p.n_taps + 2) return false;
  }
  p.stream = acc_ring ? 2 : 1;
  {  // tap-merged MMAs (accumulator ring, one k-block per tap, weight tiles that tile the swizzle atoms exactly)
    static const bool on = [] { const char* e = getenv("VST_MERGE_TAPS"); return e ? atoi(e) != 0 : true; }();
    p.merge_taps = (on && acc_ring && p.kb_per_tap == 1 && BK == 64 && p.N_mma % 32 == 0 && 16 * p.N_mma <= 512) ? 1 : 0;
  }
  p.s_dy0 = p.tap_dy[0];
  p.MT = 1; p.TW = 128; p.TH = 1; p.group = 1;
  if (p.tile_step_x <= 0) p.tile_step_x = 128;
  p.tiles_x = cdiv(p.Wo, p.tile_step_x);
  p.tiles_y = p.Ho;
  // row chunks per strip: minimise rounds x (rows per chunk + window overlap) over one-CTA-per-SM waves
  const int strips = p.n_img * p.tiles_x;
  long best = -1;
  for (int c = 1; c <= 64 && c <= p.Ho; ++c) {
    const int rpc = cdiv(p.Ho, c);
    const int units = strips * cdiv(p.Ho, rpc);
    const long cost = (long)cdiv(units, kNumSMs) * (rpc + p.n_taps - 1);
    if (best < 0 || cost < best) { best = cost; p.s_rpc = rpc; }
  }
  p.s_chunks = cdiv(p.Ho, p.s_rpc);
  return true;
}

// dy-sharing eligibility and column tables (see TapGemmParams::dyshare).  VST_DYSHARE=0 disables it.
static bool try_dyshare(TapGemmParams& p, int BK) {
  static const int mode = [] { const char* e = getenv("VST_DYSHARE"); return e ? atoi(e) : 1; }();
  p.dyshare = 0; p.n_cols = 0; p.dy_max = 0; p.box_rows = 0; p.w_res = 0;
  if (!mode || p.stream || p.epi_mode == TG_EPI_ROWCONV || p.n_taps < 2 || p.n_taps > 48) return false;
  if (p.tile_step_x > 0 && p.tile_step_x != p.TW) return false;
  if (p.TW < 8 || p.MT <= 0 || p.TW * p.TH != 128 * p.MT) return false;
  int n_cols = -1, dy_max = 1;
  for (int ph = 0; ph < p.n_phase; ++ph) {
    const int tb = ph * p.n_taps;
    bool used[TG_MAX_TAPS] = {false};
    int nc = 0;
    for (int t = 0; t < p.n_taps; ++t) {
      if (used[t]) continue;
      used[t] = true;
      int n = 1, stride = 0, last = t;
      for (;;) {
        int t2 = -1;
        if (n == 1) {
          for (int u = last + 1; u < p.n_taps; ++u)
            if (!used[u] && p.tap_dx[tb + u] == p.tap_dx[tb + t] && p.tap_pl[tb + u] == p.tap_pl[tb + t] &&
                p.tap_dy[tb + u] == p.tap_dy[tb + t] + n) { t2 = u; break; }
          if (t2 >= 0) stride = t2 - t;
        } else {
          const int u = t + n * stride;
          if (u < p.n_taps && !used[u] && p.tap_dx[tb + u] == p.tap_dx[tb + t] && p.tap_pl[tb + u] == p.tap_pl[tb + t] &&
              p.tap_dy[tb + u] == p.tap_dy[tb + t] + n) t2 = u;
        }
        if (t2 < 0 || stride > 127) break;
        used[t2] = true;
        last = t2;
        ++n;
      }
      if (ph * (n_cols < 0 ? 0 : n_cols) + nc >= 48 || nc >= 48) return false;
      const int slot = (n_cols < 0 ? 0 : ph * n_cols) + nc;
      if (slot >= 48) return false;
      p.col_dx[slot] = p.tap_dx[tb + t]; p.col_dy0[slot] = p.tap_dy[tb + t]; p.col_pl[slot] = p.tap_pl[tb + t];
      p.col_n[slot] = (signed char)n; p.col_t0[slot] = (signed char)t; p.col_ts[slot] = (signed char)stride;
      if (n > dy_max) dy_max = n;
      ++nc;
    }
    if (n_cols < 0) n_cols = nc;
    else if (nc != n_cols) return false;
  }
  if (dy_max < 2 || n_cols * p.n_phase > 48) return false;
  // Re-tile for the mode: tall tiles share more rows per box.  Cost = L2 -> shared-memory bytes per covered output pixel
  // (boxes + weight tiles), inflated by the tile grid's overhang; at least 3 pipeline stages must fit.
  const int b_bytes = p.N_mma * BK * 2, b_al = (b_bytes + 1023) & ~1023;
  const long kb = p.kb_per_tap;
  // resident weights: one phase, one N tile, shared weights, and all tap tiles within 32 KB (conv1: 9 x 3 KB) - they are then
  // loaded once per CTA instead of once per tile, and the stages hold activations only
  static const bool wres_on = [] { const char* e = getenv("VST_WRES"); return e ? atoi(e) != 0 : true; }();
  const int w_all = p.n_taps * p.kb_per_tap * b_al;
  const bool wres = wres_on && p.n_phase == 1 && p.n_ntile == 1 && p.b_img_rows == 0 && w_all <= 32 * 1024;
  const int budget = (p.duo ? tg_smem_budget() / 2 : tg_smem_budget()) - epi_staging_bytes(p) - 2560 - (wres ? w_all : 0);
  double best = -1.;
  int best_tw = 0, best_mt = 0;
  int mt0 = p.MT;
  if (p.duo) while (mt0 > 1 && 2 * mt0 * p.N_mma > 256) mt0 >>= 1;   // two accumulator stages inside 256 TMEM columns
  for (int mt = mt0; mt >= 1; mt >>= 1) {
    for (int tw = 8; tw <= 256; tw <<= 1) {
      const int th = 128 * mt / tw;
      if (th < 1 || th > 200 || th * tw != 128 * mt) continue;
      const int rows = th + dy_max - 1;
      const int a_box_al = (rows * tw * BK * 2 + 1023) & ~1023;
      const int stage = a_box_al + (wres ? 0 : dy_max * b_al);
      if (budget / stage < 3) continue;
      const double cover = (double)cdiv(p.Wo, tw) * tw * (double)cdiv(p.Ho, th) * th / ((double)p.Wo * p.Ho);
      const double cost = cover * (double)(n_cols * kb * a_box_al + (wres ? 0L : (long)p.n_taps * kb * b_bytes)) / (128. * mt);
      if (best < 0. || cost < best * 0.97) { best = cost; best_tw = tw; best_mt = mt; }
    }
  }
  if (best < 0.) return false;
  // worth it only when it removes a good part of the L2 -> shared-memory traffic of a tile
  const double old_cost = (double)p.n_taps * kb * (p.MT * 128 * BK * 2 + b_bytes) / (128. * p.MT);
  if (best > 0.8 * old_cost) return false;
  p.MT = best_mt; p.TW = best_tw; p.TH = 128 * best_mt / best_tw;
  p.tiles_x = cdiv(p.Wo, p.TW); p.tiles_y = cdiv(p.Ho, p.TH);
  if (p.tile_step_x > 0) p.tile_step_x = p.TW;
  const int box_rows = p.TH + dy_max - 1;
  p.dyshare = 1; p.n_cols = n_cols; p.dy_max = dy_max; p.box_rows = box_rows;
  p.w_res = wres ? 1 : 0;
  p.group = 1;
  return true;
}

// CTA-pair eligibility: wide single-phase bf16-NHWC layers with one 128-pixel sub-tile per CTA (the 192-channel trunk, the
// >= 128-channel VGG layers).  Their main loop saturates shared-memory bandwidth at M = 128 per CTA (TMA fills + operand
// reads ~ 125 B/clk); as a pair each CTA stages only half of the weight tile.  VST_CTA2=0 disables it, 2 also admits
// layers with several sub-tiles / N tiles (measured neutral on the VGG layers, so not the default).
static bool try_cta2(TapGemmParams& p) {
  static const int mode = [] { const char* e = getenv("VST_CTA2"); return e ? atoi(e) : 1; }();
  p.cta2 = 0;
  if (!mode || p.stream || p.dyshare) return false;
  if (p.n_phase != 1 || p.b_img_rows != 0 || p.epi_mode != TG_EPI_BF16_NHWC) return false;
  if (p.N_mma < 128 || p.N_mma % 32 != 0) return false;
  // the two tiles of a pair (consecutive tile indices) must read the same weight rows: an N tile may not end on an odd tile
  if (p.n_ntile > 1 && ((long)p.n_img * p.tiles_y * p.tiles_x) % 2 != 0) return false;
  if (mode == 1 && (p.MT != 1 || p.n_ntile != 1)) return false;   // VST_CTA2=1: one sub-tile, one N tile only; 2: all of the above
  if (p.tile_step_x > 0 && p.tile_step_x != p.TW) return false;
  p.cta2 = 1;
  return true;
}

void tapgemm_plan(TapGemmParams& p, int BK) {
  static const bool verbose = [] { const char* e = getenv("VST_TG_VERBOSE"); return e && atoi(e) != 0; }();
  p.cta2 = 0;
  // bf16-NHWC epilogue flavour.  VST_EPI_DIRECT: 0 (default) staged - shared-memory tile, coalesced 16-byte stores, statistics in
  // the store loop; 1 direct - TMEM -> registers -> 256-bit stores, butterfly statistics; 2 per layer (direct where a tile has
  // >= 4 sub-tiles).  Same-box A/B at the default bench settings (sustained, power-capped clocks): staged 552 frames/s on one
  // lane, direct 518, per-layer 540 - the direct form executes ~1.6x the instructions (the butterflies), which costs more than the
  // staging tile's shared-memory round trip once the SM clock is what the power cap leaves; it wins only at burst clocks.  Both
  // give bit-reproducible statistics.  fp16 / fp32-output layers (the "fp16" plan) always take the direct path.
  { static const int direct = [] { const char* e = getenv("VST_EPI_DIRECT"); return e ? atoi(e) : 0; }();
    p.epi_direct = (direct == 1 || (direct == 2 && p.MT >= 4) || p.half || p.out_f32) ? 1 : 0; }
  // TMA-store epilogue with tensor-core statistics (VST_EPI_TMA=0: the staged epilogue with per-thread stores): every 16-bit
  // NHWC output whose pixel stride keeps the 16-byte alignment TMA needs
  { static const int tma = [] { const char* e = getenv("VST_EPI_TMA"); return e ? atoi(e) : 0; }();
    p.epi_tma = (tma && p.epi_mode == TG_EPI_BF16_NHWC && !p.epi_direct && p.out_cstride % 8 == 0 &&
                 (reinterpret_cast<uintptr_t>(p.out0) & 15) == 0) ? 1 : 0;
    p.epi_nbuf = 1;   // the pipeline mode is planned with ONE staging buffer; launch_tapgemm adds a second where it costs no stage
    p.tmo_for = nullptr; }
  p.w_res = 0;
  // two CTAs per SM (the DUO kernel instantiation): narrow bf16-NHWC layers with the staged epilogue; the dy-sharing tile is
  // then planned for half the shared memory and 256 TMEM columns (VST_TG_DUO = widest N_mma, 0 = off)
  // Measured (1080p x 4, same box): conv1 0.535 -> 0.612 ms, deconv2 0.565 -> 0.632 ms - SLOWER, so the default is off: the
  // chain that paces these layers is evidently not a per-CTA one (unlike the narrow Gram's), and the re-planned tiles are
  // half as tall (a 40-row box per 32 output rows instead of 72 per 64).  Kept as an experiment switch.
  // staged ping-pong epilogue (see the kernel): narrow bf16-NHWC layers with an even number of sub-tiles per tile
  static const int spp_max_n = [] { const char* e = getenv("VST_EPI_SPP"); return e ? atoi(e) : 64; }();   // widest N_mma; 0 = off
  p.epi_spp = (spp_max_n > 0 && p.epi_mode == TG_EPI_BF16_NHWC && !p.epi_tma && !p.epi_direct && !p.out_f32 && p.N_mma <= spp_max_n) ? 1 : 0;
  static const int duo_max_n = [] { const char* e = getenv("VST_TG_DUO"); return e ? atoi(e) : 0; }();
  p.duo = 0;
  const bool duo_try = duo_max_n > 0 && p.epi_mode == TG_EPI_BF16_NHWC && !p.epi_tma && !p.epi_direct && !p.out_f32 && !p.half &&
                       p.N_mma <= duo_max_n && 2 * p.N_mma <= 256 && (BK == 32 || BK == 64);
  if (tapgemm_try_stream(p, BK)) p.dyshare = 0;
  else {
    bool ok = false;
    if (duo_try) { p.duo = 1; ok = try_dyshare(p, BK); if (!ok) p.duo = 0; }
    if (!ok && !try_dyshare(p, BK)) try_cta2(p);
  }
  if (p.epi_tma && p.tile_step_x > 0 && p.tile_step_x != p.TW) p.epi_tma = 0;   // overlapping tiles: per-thread stores
  if (verbose)
    fprintf(stderr, "tapgemm_plan: N=%d taps=%dx%d kbpt=%d BK=%d MT=%d tile %dx%d grid %dx%d -> stream=%d dyshare=%d cols=%d dy_max=%d box_rows=%d cta2=%d w_res=%d epi_tma=%d\n",
            p.N_mma, p.n_phase, p.n_taps, p.kb_per_tap, BK, p.MT, p.TW, p.TH, p.Wo, p.Ho, p.stream, p.dyshare, p.n_cols, p.dy_max,
            p.box_rows, p.cta2, p.w_res, p.epi_tma);
}

typedef void (*TgKernel)(const TapGemmParams);

// Kernel instantiation for a planned layer: a specialised one for the (k-block, pair, main-loop mode, epilogue) combinations
// the networks use (VST_TG_GENERIC=1: always the generic kernel), else the generic kernel that decides at run time.
static TgKernel pick_kernel(const TapGemmParams& p, int BK) {
  static const bool generic_only = [] { const char* e = getenv("VST_TG_GENERIC"); return e && atoi(e) != 0; }();
  const int mode = p.stream == 2 ? TGM_ACCRING : p.stream ? TGM_RING : p.dyshare ? TGM_DYSH : TGM_PLAIN;
  const int epi = p.epi_mode == TG_EPI_ROWCONV ? TGE_ROWCONV : p.epi_mode == TG_EPI_F32_NCHW ? TGE_F32
                  : p.epi_tma ? (p.epi_pp ? TGE_TMAPP : TGE_TMA) : p.epi_direct ? TGE_DIRECT : TGE_STAGED;
  {
    static const bool list = [] { const char* e = getenv("VST_TG_VERBOSE"); return e && atoi(e) >= 2; }();
    if (list) {   // one line per distinct (k-block, pair, mode, epilogue, width) combination: what to specialise
      static std::mutex mu;
      static std::vector<long> seen;
      const long key = ((((long)BK * 2 + (p.cta2 != 0)) * 8 + mode) * 8 + epi) * 1024 + p.N_mma * 2 + (p.stats != nullptr);
      std::lock_guard<std::mutex> lk(mu);
      bool dup = false;
      for (long k : seen) dup |= (k == key);
      if (!dup) {
        seen.push_back(key);
        fprintf(stderr, "tapgemm combo: BK=%d cta2=%d mode=%d epi=%d N=%d MT=%d taps=%dx%d kbpt=%d stats=%d bias=%d relu=%d\n", BK, p.cta2, mode, epi,
                p.N_mma, p.MT, p.n_phase, p.n_taps, p.kb_per_tap, p.stats != nullptr, p.bias != nullptr, p.relu);
      }
    }
  }
  if (p.duo) {
    if (BK == 32 && !p.cta2 && mode == TGM_DYSH && epi == TGE_STAGED) return tapgemm_kernel<32, false, TGM_DYSH, TGE_STAGED, true>;
    if (BK == 64 && !p.cta2 && mode == TGM_DYSH && epi == TGE_STAGED) return tapgemm_kernel<64, false, TGM_DYSH, TGE_STAGED, true>;
    return nullptr;
  }
  if (!generic_only) {
#define TG_SPEC(bk, c2, m, e) if (BK == bk && (p.cta2 != 0) == c2 && mode == m && epi == e) return tapgemm_kernel<bk, c2, m, e>;
    TG_SPEC(64, true, TGM_PLAIN, TGE_STAGED)   TG_SPEC(64, true, TGM_PLAIN, TGE_TMA)
    TG_SPEC(32, false, TGM_DYSH, TGE_STAGED)   TG_SPEC(32, false, TGM_DYSH, TGE_TMA)   TG_SPEC(32, false, TGM_DYSH, TGE_TMAPP)
    TG_SPEC(64, false, TGM_DYSH, TGE_STAGED)   TG_SPEC(64, false, TGM_DYSH, TGE_TMA)   TG_SPEC(64, false, TGM_DYSH, TGE_TMAPP)
    TG_SPEC(64, false, TGM_ACCRING, TGE_ROWCONV)
    TG_SPEC(32, true, TGM_PLAIN, TGE_STAGED)   TG_SPEC(32, false, TGM_PLAIN, TGE_STAGED)
    TG_SPEC(64, false, TGM_PLAIN, TGE_STAGED)  TG_SPEC(64, false, TGM_PLAIN, TGE_F32)
#undef TG_SPEC
  }
  if (p.cta2) {
    switch (BK) {
      case 64: return tapgemm_kernel<64, true, -1, -1>;
      case 32: return tapgemm_kernel<32, true, -1, -1>;
      case 16: return tapgemm_kernel<16, true, -1, -1>;
    }
  } else {
    switch (BK) {
      case 64: return tapgemm_kernel<64, false, -1, -1>;
      case 32: return tapgemm_kernel<32, false, -1, -1>;
      case 16: return tapgemm_kernel<16, false, -1, -1>;
    }
  }
  return nullptr;
}

static int ensure_smem_attr(TgKernel kern, bool duo = false) {
  static std::mutex mu;
  static std::vector<TgKernel> done;
  std::lock_guard<std::mutex> lk(mu);
  for (TgKernel k : done) if (k == kern) return VST_OK;
  VST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  // the DUO instantiations need BOTH of their ~108 KB CTAs on an SM: ask for the max-shared split explicitly (the driver's
  // default only guarantees room for one)
  if (duo) VST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  done.push_back(kern);
  return VST_OK;
}

int launch_tapgemm(TapGemmParams& p, int BK, cudaStream_t st) {
  if (p.MT <= 0) p.MT = 1;
  VST_CHECK_ARG(p.TW * p.TH == 128 * p.MT && (p.TW & (p.TW - 1)) == 0, "tapgemm: TW*TH must be 128*MT, TW a power of two");
  VST_CHECK_ARG(p.N_mma % 16 == 0 && p.N_mma >= 16 && p.N_mma <= 256, "tapgemm: N_mma=%d invalid", p.N_mma);
  VST_CHECK_ARG(2 * p.MT * p.N_mma <= 512, "tapgemm: 2*MT*N_mma exceeds the 512 TMEM columns");
  VST_CHECK_ARG(p.n_phase * p.n_taps <= TG_MAX_TAPS, "tapgemm: too many taps");
  VST_CHECK_ARG(p.epi_mode != TG_EPI_BF16_NHWC || p.Cout % 8 == 0, "tapgemm: bf16 NHWC output needs Cout %% 8 == 0");
  if (p.tile_step_x <= 0) p.tile_step_x = p.TW;
  { const char* e = getenv("VST_TG_DBG"); p.dbg = e ? atoi(e) : 0; }
  // narrow layers are bound by the epilogue's instruction stream: give them the second epilogue warp set; the wide
  // (trunk, VGG >= 128-channel) layers are MMA/smem-bound and lose ~5 % to the extra warps' issue slots
  // TMEM accumulator stages: as many as fit the 512 columns (narrow layers hide the MMA<->epilogue handshake behind them)
  p.acc_stages = 2;
  { const char* e = getenv("VST_ACC_STAGES"); const int cap = e ? atoi(e) : 8;
    while (p.acc_stages < cap && 2 * p.acc_stages * p.MT * p.N_mma <= 512) p.acc_stages *= 2; }
  if (p.stream == 2) p.acc_stages = 16;
  // second MMA-issuing warp: measured neutral (the narrow layers are not issue-bound any more), opt-in with VST_MMA2=1
  { const char* e = getenv("VST_MMA2"); p.mma2 = (p.MT >= 2 && !p.stream && e && atoi(e) != 0) ? 1 : 0; }
  { const char* e = getenv("VST_EPI8"); const int lim = e ? atoi(e) : 256; p.epi8 = (p.epi_mode == TG_EPI_BF16_NHWC && (p.N_mma <= lim || p.epi_tma)) ? 1 : 0; }
  if (p.epi_tma && (p.tmo_for != p.out0 || !p.out0)) {
    // output tensor maps: per phase a view (Cout, Wo, Ho, n_img) of the NHWC tensor that starts at the phase's first pixel
    // and steps out_mul pixels / rows; boxes of 64 / 32 / 16 channels x one 128-pixel sub-tile.  Tiles that overhang the
    // frame are clipped by the store itself.
    VST_CHECK_ARG(p.out0, "tapgemm: NULL output");
    const int bw = p.TW < 128 ? p.TW : 128, bh = 128 / bw;
    const int n64 = p.N_mma >> 6, rem = p.N_mma & 63;
    for (int ph = 0; ph < p.n_phase; ++ph) {
      const int oy = p.ph_oy[ph], ox = p.ph_ox[ph];
      const int wo = std::min(p.Wo, (p.Wout - ox + p.out_mul - 1) / p.out_mul), ho = std::min(p.Ho, (p.Hout - oy + p.out_mul - 1) / p.out_mul);
      const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(p.out0) + ((size_t)oy * p.Wout + ox) * p.out_cstride;
      const size_t pix = (size_t)p.out_mul * p.out_cstride, rowst = (size_t)p.out_mul * p.Wout * p.out_cstride,
                   img = (size_t)p.Hout * p.Wout * p.out_cstride;
      const int widths[3] = {64, 32, 16};
      const bool used[3] = {n64 > 0, (rem & 32) != 0, (rem & 16) != 0};
      for (int k = 0; k < 3; ++k)
        if (used[k]) {
          int r = make_tmap_out(&p.tmO[ph][k], base, p.Cout, wo, ho, p.n_img, pix, rowst, img, widths[k], bw, bh);
          if (r != VST_OK) return r;
        }
    }
    p.tmo_for = p.out0;
  }
  for (int i = 0; i < p.n_phase * p.n_taps; ++i)
    p.tap_packed[i] = (p.tap_dx[i] & 0xff) | ((p.tap_dy[i] & 0xff) << 8) | ((int)p.tap_pl[i] << 16);
  const int a_bytes = p.MT * 128 * BK * 2;
  const int b_bytes = (tapgemm_b_box_rows(p) * BK * 2 + 1023) & ~1023;
  const int kb_bytes = a_bytes + b_bytes;
  if (p.cta2) p.mma2 = 0;
  if (p.epi_spp && (!p.epi8 || p.MT < 2 || (p.MT & 1) || p.cta2 || p.stream || p.epi_tma || p.epi_direct)) p.epi_spp = 0;
  const int kblocks = p.n_taps * p.kb_per_tap;
  p.epi_nbuf = 1; p.epi_pp = 0;
  const int stg_bytes = epi_staging_bytes(p);
  const int budget = tg_smem_budget() - stg_bytes - 2560;
  VST_CHECK_ARG(!p.fuse_in || (p.stream == 2 && BK == 64 && p.kb_per_tap == 1 && p.in_C <= 64 && p.in_stats && p.in_gamma && p.in_beta),
                "tapgemm: fused input normalisation needs the accumulator-ring mode with one 64-channel k-block");
  if (p.stream) {
    const int slot = p.kb_per_tap * 128 * BK * 2, w_region = p.n_taps * p.kb_per_tap * b_bytes;
    int ring = (budget - w_region) / slot;
    if (ring > 16) ring = 16;
    VST_CHECK_ARG(ring >= (p.stream == 2 ? 2 : p.n_taps + 1), "tapgemm(stream): ring of %d rows cannot hold %d taps + 1", ring, p.n_taps);
    p.stages = ring;
    int stg_bytes_st = stg_bytes;
    if (p.epi_tma && p.N_mma <= 96 && ring * slot + w_region + stg_bytes <= budget) {   // room for a second staging buffer
      p.epi_nbuf = 2; stg_bytes_st = 2 * stg_bytes;
    }
    p.epi_pp = (p.epi_tma && p.epi_nbuf == 2 && p.N_mma <= 64) ? 1 : 0;
    const size_t smem_st = (size_t)ring * slot + w_region + stg_bytes_st + 16 + 1024 + 1024;
    const int units = p.n_img * p.tiles_x * p.s_chunks;
    const int grid_st = units < kNumSMs ? units : kNumSMs;
    for (int i = 0; i < p.n_taps; ++i)
      p.tap_packed[i] = (p.tap_dx[i] & 0xff) | ((p.tap_dy[i] & 0xff) << 8) | ((int)p.tap_pl[i] << 16);
    TgKernel kern = pick_kernel(p, BK);
    if (!kern) { set_error("tapgemm: BK=%d unsupported", BK); return VST_EUNSUPPORTED; }
    int rs = ensure_smem_attr(kern);
    if (rs != VST_OK) return rs;
    vst::launch(kern, grid_st, (p.fuse_in || p.rider.on) ? 512 : TG_THREADS, smem_st, st, p);
    VST_LAUNCH_CHECK();
    return VST_OK;
  }
  if (p.group <= 0) {
    // group k-blocks so that one mbarrier round trip moves a few tens of KB
    int g = 1;
    for (int c = 1; c <= 4; ++c)
      if (kblocks % c == 0 && c * kb_bytes <= 48 * 1024 && budget / (c * kb_bytes) >= 3) g = c;
    p.group = g;
  }
  VST_CHECK_ARG(kblocks % p.group == 0, "tapgemm: group %d does not divide %d k-blocks", p.group, kblocks);
  if (p.dyshare) p.mma2 = 0;
  const int w_res_bytes = (p.dyshare && p.w_res) ? p.n_taps * p.kb_per_tap * b_bytes : 0;
  const int stage_bytes = p.dyshare ? ((p.box_rows * p.TW * BK * 2 + 1023) & ~1023) + (p.w_res ? 0 : p.dy_max * b_bytes) : p.group * kb_bytes;
  int stages = (budget - w_res_bytes) / stage_bytes;
  // two CTAs per SM (see the DUO instantiation): narrow bf16-NHWC dy-sharing layers with the staged epilogue whose ring still
  // holds >= 3 stages in half the shared memory and whose accumulators fit 256 TMEM columns twice over
  // (planned by tapgemm_plan; a launch that carries an apply rider needs the 512-thread form and runs one CTA per SM)
  if (p.duo && (!p.dyshare || p.cta2 || p.rider.on || 2 * p.MT * p.N_mma > 256)) p.duo = 0;
  if (p.duo) {
    const int budget2 = tg_smem_budget() / 2 - stg_bytes - 2560;
    const int st2 = (budget2 - w_res_bytes) / stage_bytes;
    if (st2 >= 2) {
      stages = st2;
      while (p.acc_stages > 2 && p.acc_stages * p.MT * p.N_mma > 256) p.acc_stages >>= 1;
    } else {
      p.duo = 0;
    }
  }
  if (stages > 8) stages = 8;
  VST_CHECK_ARG(stages >= 2, "tapgemm: stage of %d bytes leaves < 2 pipeline stages", stage_bytes);
  p.stages = stages;
  int stg_bytes_all = stg_bytes;
  if (p.epi_tma && p.N_mma <= 96 && stages * stage_bytes + w_res_bytes + stg_bytes <= budget) {   // room for a second staging buffer at the same depth
    p.epi_nbuf = 2; stg_bytes_all = 2 * stg_bytes;
  }
  p.epi_pp = (p.epi_tma && p.epi_nbuf == 2 && p.N_mma <= 64) ? 1 : 0;
  const size_t smem = (size_t)stages * stage_bytes + w_res_bytes + stg_bytes_all + 16 + 1024 /*align*/ + 1024 /*barriers*/;
  const int total_tiles = p.n_phase * p.n_ntile * p.n_img * p.tiles_y * p.tiles_x;
  const int n_cta = p.duo ? 2 * kNumSMs : kNumSMs;
  const int grid = total_tiles < n_cta ? total_tiles : n_cta;
  TgKernel kern = pick_kernel(p, BK);
  if (!kern) { set_error("tapgemm: BK=%d unsupported", BK); return VST_EUNSUPPORTED; }
  int ra = ensure_smem_attr(kern, p.duo != 0);
  if (ra != VST_OK) return ra;
  if (p.cta2) {
    // clusters of two CTAs (one SM pair each); an even grid so that every CTA has its partner
    cudaLaunchConfig_t cfg = {};
    int g2 = (total_tiles + 1) & ~1;
    if (g2 > (kNumSMs & ~1)) g2 = kNumSMs & ~1;
    cfg.gridDim = dim3(g2);
    cfg.blockDim = dim3(p.rider.on ? 512 : TG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    VST_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return VST_OK;
  }
  vst::launch(kern, grid, p.rider.on ? 512 : TG_THREADS, smem, st, p);
  VST_LAUNCH_CHECK();
  return VST_OK;
}

}  // namespace vst
