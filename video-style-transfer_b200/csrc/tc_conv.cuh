// Tap-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// A convolution over an NHWC bf16 activation tensor is evaluated as a sum over filter taps of
// plain GEMMs:   D[pixel, cout] += A_tap[pixel, cin] * W_tap[cout, cin]^T
// where A_tap is the activation tensor shifted by the tap offset.  Every A_tap tile is one TMA
// box load (the padding is already materialised in the tensor, or comes from TMA's zero fill for
// the VGG zero-pad convs), every W_tap tile is a TMA box of the K-major packed weight matrix,
// accumulation runs in TMEM (fp32) through tcgen05.mma kind::f16, and the epilogue reads the
// accumulators back with tcgen05.ld.
//
// Tile = 128 output pixels (TH rows x TW cols of one image, TH*TW = 128) x N_mma channels.
// Stride-2 convs read "parity planes" (the producer writes the padded tensor split by row/col
// parity), nearest-x2-upsample convs run as 4 output phases with pre-summed 2x2 weights.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "tc_layout.cuh"

namespace vst {

constexpr int TG_MAX_TAPS = 96;

// A "rider": an InstanceNorm-apply pass (y = act(IN(raw)) (+ residual) into the consumer's padded layout) that belongs to
// ANOTHER half-batch and is carried by a tap-GEMM launch.  The persistent tap-GEMM CTAs leave the SM's load/store path and a
// quarter of its registers idle, the apply pass needs nothing else: four extra warps per CTA stream it beside the MMAs, so
// the HBM-bound pass costs no time of its own (vst_plan_forward_pair walks two half-batch plans in lock step and gives
// every tap-GEMM of one the pending apply of the other).  Same arithmetic as apply_lds_kernel (engine.cu), bit for bit.
struct ApplyRider {
  int on;
  const void* raw;          // [N][H][W][C] 16-bit, the producer's raw output
  const double* stats;      // [N][C][2]
  const float* gamma;
  const float* beta;
  const void* residual;     // optional, layout RL
  void* dst;                // layout DL
  ActLayout RL, DL;
  int N, relu;
  float eps;
};

enum TgEpilogue : int {
  TG_EPI_BF16_NHWC = 0,  // raw accumulators (+bias, +relu) -> bf16 NHWC
  TG_EPI_F32_NCHW = 1,   // (+bias, act) -> fp32 NCHW, optional uint8 BGR HWC copy (Cout == 3)
  // k x k convolution with few output channels evaluated as a ROW convolution: the GEMM produces
  // D[x', (kx, co)] = sum_{ky, c} in[y+ky][x'][c] * w[co][c][ky][kx] for 128 consecutive padded input
  // pixels x' of one output row, and the epilogue sums the kx-shifted columns:
  //   out[x][co] = bias[co] + sum_kx D[x + kx][kx*rc_co + co]      (128 - k + 1 outputs per tile)
  TG_EPI_ROWCONV = 2,
};

struct TapGemmParams {
  CUtensorMap tmA;  // 5-D (C, X, Y, N, P) bf16, box (BK, TW, TH, 1, 1)
  CUtensorMap tmB;  // 2-D (K, rows)       bf16, box (BK, N_mma)
  int n_img, tiles_x, tiles_y, TW, TH;  // tile = TH x TW pixels = MT sub-tiles of 128 (TW a power of two)
  int MT;                               // 128-row MMA sub-tiles per CTA tile (1, 2, 4): one TMA box, MT accumulators
  int n_phase, n_ntile;
  int n_taps, kb_per_tap;  // taps per phase, BK-blocks per tap
  int N_mma, stages;
  int group;           // k-blocks per pipeline stage (one mbarrier round trip per group)
  int tile_step_x;     // x advance per tile (TW, or TW - k + 1 for TG_EPI_ROWCONV)
  int rc_k, rc_co;     // TG_EPI_ROWCONV: kernel width, real output channels
  int dbg;             // experiment switches (VST_TG_DBG): 1 skip stats, 2 skip global stores, 4 skip TMEM->staging
  int Ho, Wo;          // valid extent of the tile grid (per phase)
  int Cout;            // real output channels
  int out_mul;         // output pixel = (y*out_mul + ph_oy, x*out_mul + ph_ox)
  int Hout, Wout;      // full output extent
  int out_cstride;     // channel stride (elements per pixel) of an NHWC output
  int epi_mode, act, relu;
  void* out0;          // bf16 NHWC or fp32 NCHW
  uint8_t* out_u8;     // optional [N,Hout,Wout,3] BGR (TG_EPI_F32_NCHW, Cout == 3)
  const float* bias;   // optional [Cout]
  double* stats;       // optional [N][Cout][2] (sum, sum of squares of the bf16-rounded outputs): deterministic per-CTA fp32
                       // partials, one fp64 atomic per channel and CTA per image change (order-independent to ~1e-16)
  // fused input normalisation (accumulator-ring stream mode only; see the transform warps in tc_conv.cu): the A tensor map is
  // over the producer's RAW output [N][in_H][in_W][in_C] (16-bit, no halo), tap offsets are relative to the unpadded frame
  int fuse_in, in_relu, in_H, in_W, in_C;
  const double* in_stats;   // [N][in_C][2] of the producer
  const float* in_gamma;
  const float* in_beta;
  float in_eps;
  int half;            // 1: operands and 16-bit outputs are fp16 instead of bf16 (the "fp16" inference plan)
  int out_f32;         // 1: TG_EPI_BF16_NHWC stores fp32 NHWC (out_cstride in floats) - the residual blocks' second conv of the
                       // "fp16" plan, whose InstanceNorm + residual add run in fp32; statistics are then those of the fp32 values
  int epi_direct;      // bf16-NHWC epilogue without the shared-memory staging tile (set by tapgemm_plan; VST_EPI_DIRECT=0: staged)
  // TMA-store epilogue (set by tapgemm_plan, maps built by launch_tapgemm): the 16-bit tile is staged in swizzled shared-memory
  // boxes of 64 / 32 / 16 channels (SWIZZLE_128B / 64B / 32B: conflict-free for the writers), leaves through
  // cp.async.bulk.tensor stores, and the InstanceNorm statistics come from the staged tile through warp-level MMAs
  int epi_tma;         // 1: on
  int epi_nbuf;        // staging buffers (set by launch_tapgemm: 2 where shared memory allows and N_mma <= 96, else 1)
  int epi_pp;          // 1: N_mma <= 64 with two buffers - the two epilogue warp sets work on alternate sub-tiles (own buffer each)
  int w_res;           // dy-sharing mode: all weight tiles resident behind the ring (one phase, one N tile, <= 32 KB): loaded once
  const void* tmo_for; // output pointer the maps below were built for (cache key)
  CUtensorMap tmO[4][3];   // [phase][kind 0: 64-channel box, 1: 32, 2: 16], dims (Cout, Wo, Ho, n_img), box (w, min(TW,128), 128/min(TW,128), 1)
  signed char tap_dx[TG_MAX_TAPS], tap_dy[TG_MAX_TAPS], tap_pl[TG_MAX_TAPS];  // [phase*n_taps + t]
  int tap_packed[TG_MAX_TAPS];  // filled by launch_tapgemm: (dx & 0xff) | (dy & 0xff) << 8 | pl << 16
  signed char ph_oy[4], ph_ox[4];
  int b_img_rows;      // weight rows to skip per image (per-image 1x1 weights of the Gram backward), else 0
  // row-streaming mode (set by launch_tapgemm when the taps form a pure row stencil): see tc_conv.cu
  int mma2;            // 1: two MMA-issuing warps, each owning half of the MT sub-tiles (set by launch_tapgemm)
  int acc_stages;      // TMEM accumulator stages (2, 4 or 8; set by launch_tapgemm)
  int epi8;            // 1: eight epilogue warps (set by launch_tapgemm for narrow bf16-NHWC layers)
  // dy-sharing mode (set by tapgemm_plan): taps with the same dx / plane and consecutive dy ("a column") read ONE A box of
  // TH + n - 1 rows; tap j of the column is the same box shifted by j rows (a descriptor offset).  See tc_conv.cu.
  int dyshare, n_cols, dy_max, box_rows;   // columns per phase, longest column, rows of the A box
  signed char col_dx[48], col_dy0[48], col_pl[48], col_n[48], col_t0[48], col_ts[48];  // [phase*n_cols + c]; tap j = t0 + j*ts
  int cta2;            // 1: CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256): set by tapgemm_plan for wide single-phase layers
  int merge_taps;      // stream == 2: consecutive taps issued as one MMA spanning several accumulator slots (see tc_conv.cu)
  int stream;          // 1: ring of input rows + resident weights
  int s_dy0;           // row offset of tap 0 (taps are dy = s_dy0 + t)
  int s_chunks, s_rpc; // row chunks per column strip, output rows per chunk
  int epi_spp;         // 1: staged epilogue in ping-pong - the two warp sets take alternate sub-tiles (set by tapgemm_plan / launch_tapgemm)
  int duo;             // 1: two CTAs per SM (set by launch_tapgemm for epilogue-latency-bound narrow layers, see tc_conv.cu)
  ApplyRider rider;    // optional apply pass of another half-batch (warps 12..15; launch with TG_THREADS + 128)
};

// Host-side description of one tensor operand for cuTensorMapEncodeTiled.
int make_tmap_act(CUtensorMap* out, const void* base, int C, int X, int Y, int N, int P, size_t pix_stride_elems,
                  size_t row_stride_elems, size_t img_stride_elems, size_t plane_stride_elems, int BK, int TW, int TH);
// Generic 5-D bf16 map: dims (d0..d4), element strides of d1..d4, box (b0, b1, b2, b3, 1); swizzle from b0.
int make_tmap_act_generic(CUtensorMap* out, const void* base, int d0, int d1, int d2, int d3, int d4, size_t s1, size_t s2,
                          size_t s3, size_t s4, int b0, int b1, int b2, int b3);
int make_tmap_wgt(CUtensorMap* out, const void* base, int K, int rows, int BK, int box_rows);

// Switches `p` to row-streaming mode when its taps are a pure row stencil (call after the taps / epilogue / grid fields
// are set and BEFORE the A tensor map is built: the mode fixes TW = 128, TH = 1, MT = 1).
bool tapgemm_try_stream(TapGemmParams& p, int BK);
// Chooses the pipeline mode for `p` (row streaming, dy-sharing or plain per-tap boxes).  Call after the taps / tile / grid /
// epilogue fields are set and BEFORE the A tensor map is built; build the map with tapgemm_box_rows(p) rows per box.
void tapgemm_plan(TapGemmParams& p, int BK);
inline int tapgemm_box_rows(const TapGemmParams& p) { return p.dyshare ? p.box_rows : p.TH; }
// rows per box of the WEIGHT tensor map: a CTA pair loads half of the N_mma rows per CTA
inline int tapgemm_b_box_rows(const TapGemmParams& p) { return p.cta2 ? p.N_mma / 2 : p.N_mma; }
bool tapgemm_stream_enabled();
// Picks stages / smem and launches on `st`.  BK in {16, 32, 64}.
int launch_tapgemm(TapGemmParams& p, int BK, cudaStream_t st);

// Chooses (TW, TH) with TW*TH == 128*MT (TW a power of two <= 256) minimising overhang.
void choose_tile(int Ho, int Wo, int MT, int* TW, int* TH);
// Sub-tiles per CTA tile for an N_mma-wide layer (TMEM holds 2 x MT x N_mma fp32 columns).
int choose_mt(int N_mma);

}  // namespace vst
