/* vst_b200.h - C ABI of the B200-native ReCoNet / RTNSTV frame path.
 *
 * The reference (Maboroshi0327/Video-Style-Transfer) is pure Python/PyTorch and has no FFI of
 * its own: the boundary it offers is the Python module surface of
 * RC/network.py, RC/utilities.py, RT/network.py, RT/vgg19.py, RT/utilities.py.  This header is
 * the layer directly beneath that surface: every ATen call the reference's hot path makes is
 * replaced by one entry point here.  Each declaration cites the reference line it replaces
 * (RC/ = Real-time-Coherent-Video-Style-Transfer-Network-(ReCoNet)/, RT/ =
 * Real-Time-Neural-Style-Transfer-for-Videos-(RTNSTV)/).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless a name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, workspaces); nothing here allocates
 *     device memory or synchronises, except vst_plan_* creation which uploads packed weights
 *     into caller-provided storage;
 *   - every launch goes to `stream` (a cudaStream_t passed as void*; NULL = legacy stream);
 *   - return 0 on success, a negative VST_E* code otherwise; vst_last_error() gives the text;
 *   - fp32 tensors are NCHW contiguous exactly like the reference's; the bf16 tensor-core
 *     path uses its own internal NHWC layouts behind vst_plan_*;
 *   - there is no CPU fallback: a host pointer yields VST_EDEVICE.
 */
#ifndef VST_B200_H
#define VST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VST_OK 0
#define VST_EINVAL (-1)       /* bad shape / argument */
#define VST_EDEVICE (-2)      /* pointer is not device memory (no CPU fallback) */
#define VST_ECUDA (-3)        /* CUDA runtime / driver error, see vst_last_error() */
#define VST_EUNSUPPORTED (-4) /* combination not implemented */
#define VST_EWORKSPACE (-5)   /* workspace too small */

#define VST_ABI_VERSION 1

int vst_abi_version(void);
const char* vst_last_error(void);
/* Compute capability of the current device as major*10+minor (100 on B200); <0 on error. */
int vst_device_arch(void);

/* ---- activation / padding / epilogue enums ------------------------------------------- */
#define VST_PAD_ZERO 0    /* Conv2d(padding=1) in the VGG body (RC/network.py:12) */
#define VST_PAD_REFLECT 1 /* ReflectionPad2d(k//2) (RC/network.py:67-69, RT/network.py:13-14) */
#define VST_ACT_NONE 0
#define VST_ACT_RELU 1
#define VST_ACT_TANH 2         /* RT conv4: Tanh after IN (RT/network.py:76) */
#define VST_ACT_RECONET_OUT 3  /* tanh(y/255)*150 + 255/2 (RC/network.py:85) */
#define VST_ACT_RT_OUT 4       /* (tanh(y)+1)/2*255 (RT/network.py:76,90) */

/* ======================================================================================
 * fp32 reference-semantics path (CUDA cores).  NCHW fp32, same maths as the reference ops.
 * ====================================================================================== */

/* [ReflectionPad2d|zero pad] -> Conv2d(k, stride) -> optional bias -> act.
 * `ups` = 1, or 2 to read the input through a nearest x2 upsample (src = dst/2) first.
 * Replaces RC/network.py:72-75 (ConvLayer.forward), :114-120 (UpsampleConvLayer.forward),
 * :83-85 (ConvTanh, act = VST_ACT_RECONET_OUT), RT/network.py:19-21, and the VGG convs
 * RC/network.py:29-37 (pad zero, act relu).  x:[N,Cin,H,W] w:[Cout,Cin,k,k] y:[N,Cout,Ho,Wo],
 * Ho = (H*ups + 2*pad - k)/stride + 1. */
int vst_conv2d_f32(const float* x, const float* w, const float* bias, float* y,
                   int N, int Cin, int H, int W, int Cout, int k, int stride, int pad,
                   int pad_mode, int ups, int act, void* stream);

/* ConvTranspose2d(k=3, stride=2, padding=1, output_padding=1): out = 2*in (RT/network.py:51,56).
 * w:[Cin,Cout,3,3]. */
int vst_conv_transpose2d_f32(const float* x, const float* w, const float* bias, float* y,
                             int N, int Cin, int H, int W, int Cout, void* stream);

/* InstanceNorm2d(affine, eps, biased variance, instance statistics) -> act -> (+ residual).
 * Replaces RC/network.py:95-97, :146-149 (residual added AFTER the norm, no ReLU after the add),
 * RT/network.py:22-24.  `residual` may be NULL.  `mean_out`/`rstd_out` ([N*C], may be NULL) are
 * kept for the backward pass. */
int vst_instance_norm_f32(const float* x, const float* gamma, const float* beta,
                          const float* residual, float* y, float* mean_out, float* rstd_out,
                          int N, int C, int HW, float eps, int act, void* stream);

/* max_pool2d(2, 2) floor (VGG body, RC/network.py:12 via torchvision). */
int vst_maxpool2_f32(const float* x, float* y, int NC, int H, int W, void* stream);

/* vgg_normalize: y = (x/255 - mean_c)/std_c on [N,3,H,W].  If `inplace_div` != 0, x itself is
 * overwritten with x/255 first (RC/utilities.py:101-106 mutates its argument); RT's variant
 * (RT/utilities.py:163-169) passes 0. */
int vst_vgg_normalize_f32(float* x, float* y, int N, int HW, int inplace_div, void* stream);

/* The byte frame `Inference.__iter__` yields (RC/utilities.py:219-224, RT/utilities.py:318-326): img [N,3,H,W] fp32 RGB ->
 * out [N,H,W,3] uint8 BGR, clamp(0,255) then astype(uint8) truncation. */
int vst_pack_bgr_u8(const float* img, unsigned char* out, int N, int H, int W, void* stream);

/* warp(x, flo): bilinear grid_sample, zeros padding, align_corners=False on the reference's
 * (W-1)-normalised grid (RC/utilities.py:39-57 == RT/utilities.py:59-77).  x:[B,C,H,W]
 * flo:[B,2,H,W] (ch0 = dx along W).  If `corner_out` != NULL it receives the int32 north-west
 * corner indices [B,H,W,2] (x0,y0) - the bit-exact index target of BASELINE.json. */
int vst_warp_f32(const float* x, const float* flo, float* out, int32_t* corner_out,
                 int B, int C, int H, int W, void* stream);

/* flow_warp_mask batched: mask[b] = (|warp(grid+f01, f10) - grid|_1 < threshold) as 0/1 floats.
 * Replaces RC/utilities.py:60-90 (threshold 2) and RT/utilities.py:80-110.  flows [B,2,H,W]. */
int vst_flow_warp_mask_f32(const float* flo01, const float* flo10, float* mask,
                           int B, int H, int W, float threshold, void* stream);

/* SceneFlow sample contract on the device (RC/datasets.py:114-143).
 * vst_resize_bilinear_f32: dst[bc] = F.interpolate(src[bc], size=(Hd,Wd), mode="bilinear", align_corners=False)
 *   * chan_scale[bc % C] (chan_scale may be NULL) - the flow resize with its per-channel rescale (:116-134).
 * vst_motion_mask_f32: mask[i] *= (motion[i] != 0 ? 0 : 1) - the motion-boundary factor (:137-143). */
int vst_resize_bilinear_f32(const float* src, float* dst, int BC, int Hs, int Ws, int Hd, int Wd,
                            const float* chan_scale, int C, void* stream);
int vst_motion_mask_f32(float* mask, const float* motion, size_t n, void* stream);

/* gram_matrix: out[b] = F F^T * scale, F = y.view(B,C,HW); scale = 1/(C*H*W) for RC
 * (RC/utilities.py:93-98), 1/(H*W) for RT (RT/utilities.py:155-160). out:[B,C,C] fp32. */
int vst_gram_f32(const float* y, float* out, int B, int C, int HW, float scale, void* stream);

/* ---- loss reductions (RC/train_single/train_starry-night.py:91-145, RT/train.py:36-60,125-132)
 * Each writes raw fp32 sums into `out`; the host applies the reference's `*= 1/count`, `*= lambda`
 * in the reference's order.  `scratch` must hold vst_reduce_scratch_floats() floats. */
size_t vst_reduce_scratch_floats(void);

/* Feature-temporal: flow/mask are FULL resolution [B,2,H,W]/[B,H,W]; the bilinear resize to
 * (Hf,Wf) with the u*=Wf/W, v*=Hf/H scaling and the (mask>0) test are fused (:91-100).
 * out[0] = sum mask_f*(f2 - warp(f1,flow_f))^2 over [B,C,Hf,Wf]; out[1] = C * sum(mask_f)
 * (== torch.nonzero(mask_f.expand(C)).shape[0], :104). */
int vst_feature_temporal_f32(const float* f1, const float* f2, const float* flow, const float* mask,
                             float* out, float* scratch, int B, int C, int Hf, int Wf, int H, int W,
                             void* stream);

/* Output-temporal (:109-123): out[0] = sum mask*((s2 - warp(s1)) - Y(i2 - warp(i1)))^2 over
 * [B,3,H,W], Y = .2126 R + .7152 G + .0722 B; out[1] = 3*sum(mask).
 * `luminance` = 0 gives RT's temporal term sum mask*(s2 - warp(s1))^2 (RT/train.py:129-131). */
int vst_output_temporal_f32(const float* s1, const float* s2, const float* i1, const float* i2,
                            const float* flow, const float* mask, float* out, float* scratch,
                            int B, int H, int W, int luminance, void* stream);

/* out[0] = sum (a-b)^2 over n elements (content loss / Gram MSE numerators, :126-138). */
int vst_sqdiff_sum_f32(const float* a, const float* b, float* out, float* scratch, size_t n, void* stream);

/* Stability metric numerator of `calculate_mse` (RC/utilities.py:126-176): out[0] = sum ((x1 - x0) - (clamp(y1) - clamp(y0)))^2
 * over n elements, y clamped to [lo, hi] like the reference's `output_tensor.clamp(0, 255)`. */
int vst_frame_diff_sqsum_f32(const float* x0, const float* x1, const float* y0, const float* y1, float lo, float hi,
                             float* out, float* scratch, size_t n, void* stream);

/* TV term on [B,C,H,W] over the top-left (H-1)x(W-1) window.  mode 0 (RC :141-145):
 * out[0] = sum dx^2+dy^2.  mode 1 (RT/train.py:55-57): out[0] = sum sqrt(max(dx^2+dy^2, 1e-8)). */
int vst_tv_f32(const float* x, float* out, float* scratch, int BC, int H, int W, int mode, void* stream);

/* ======================================================================================
 * Training step, fp32 reference-semantics path: the adjoint of every forward entry above.  The
 * reference has no backward code of its own - it calls autograd (`loss.backward()`,
 * RC/train_single/train_starry-night.py:151, RT/train.py:142); SURVEY.md §10 B1-B17 lists the
 * adjoints that call expands to.  Gradient outputs are OVERWRITTEN unless stated otherwise.
 * `scale_dev` (nullable) is a device scalar multiplied into `scale` (the 1/count factors that
 * only exist on the device, vst_loss_terms_f32).
 * ====================================================================================== */

/* wT[ci][co][a][b] = w[co][ci][k-1-a][k-1-b]: the weights of the stride-1 data gradient (B13),
 * dx_padded = vst_conv2d_f32(dy, wT, pad = k-1-p zero). */
int vst_weight_flip_transpose_f32(const float* w, float* wt, int Cout, int Cin, int k, void* stream);

/* Generic transposed convolution (gather form): y[n,co,P,Q] = sum x[n,ci,(P+pad-ky)/s,(Q+pad-kx)/s]
 * * w[ci][co][ky][kx].  With pad = 0 and (Ho,Wo) = padded input size it is the data gradient of a
 * strided Conv2d (x = dy, w = the forward weights [Cout_f][Cin_f][k][k] read as [ci][co]). */
int vst_conv_transpose_gather_f32(const float* x, const float* w, const float* bias, float* y, int N,
                                  int Cin, int H, int W, int Cout, int Ho, int Wo, int k, int stride,
                                  int pad, void* stream);

/* Adjoint of [nearest x`ups` upsample ->] pad(`pad`, zero|reflect): folds the gradient of the padded
 * tensor dxp [NC,Hp,Wp] back onto the source [NC,Hs,Ws] (aten::reflection_pad2d_backward + the 2x2
 * sum of the nearest upsample, B13/B14). */
int vst_fold_pad_f32(const float* dxp, float* dx, int NC, int Hs, int Ws, int ups, int pad, int pad_mode,
                     int Hp, int Wp, void* stream);

/* Weight gradient of vst_conv2d_f32 (same argument meaning): dw[Cout][Cin][k][k] =
 * sum_{n,oy,ox} dy[n,co,oy,ox] * pad(ups(x))[n,ci,oy*s+ky,ox*s+kx]. */
int vst_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, int N, int Cin, int H, int W,
                         int Cout, int k, int stride, int pad, int pad_mode, int ups, void* stream);

/* out[c] = sum over n and HW of x[n,c,:] (conv bias gradients). */
int vst_channel_sum_f32(const float* x, float* out, int N, int C, int HW, void* stream);

/* dz = dy * act'(z) from the saved OUTPUT y = act(z) (ConvTanh B10, VGG in-place ReLU B8). */
int vst_act_bwd_f32(const float* dy, const float* y, float* dz, size_t n, int act, void* stream);

/* InstanceNorm2d(affine) backward fused with the adjoint of the activation that followed it (B11,
 * B12): x = the norm's input (conv output), mean/rstd from vst_instance_norm_f32.  dgamma/dbeta [C]
 * are overwritten with the sums over the batch. */
int vst_instance_norm_bwd_f32(const float* x, const float* dy, const float* gamma, const float* beta,
                              const float* mean, const float* rstd, float* dx, float* dgamma,
                              float* dbeta, int N, int C, int HW, int act, void* stream);

/* max_pool2d(2,2) backward: the gradient goes to the first maximum of each window (B8). */
int vst_maxpool2_bwd_f32(const float* x, const float* dy, float* dx, int NC, int H, int W, void* stream);

/* vgg_normalize backward: dx = dy / (255 * std_c) (B9). */
int vst_vgg_normalize_bwd_f32(const float* dy, float* dx, int N, int HW, void* stream);

/* warp backward w.r.t. the sampled tensor: scatter-add of dy * bilinear weights into the in-bounds
 * corners (TORCH/GridSampler.h safe_add_2d, B5).  dx is zeroed first. */
int vst_warp_bwd_f32(const float* dy, const float* flo, float* dx, int B, int C, int H, int W, void* stream);

/* Feature-temporal backward (B4+B5): g = 2*scale*m_f*(f2 - warp(f1, flow_f)); df2 = g, df1 = scatter(-g). */
int vst_feature_temporal_bwd_f32(const float* f1, const float* f2, const float* flow, const float* mask,
                                 float scale, const float* scale_dev, float* df1, float* df2, int B, int C,
                                 int Hf, int Wf, int H, int W, void* stream);

/* Output-temporal backward (B3+B5): ds2 = 2*scale*m*((s2 - warp(s1)) - Y), ds1 = scatter(-ds2). */
int vst_output_temporal_bwd_f32(const float* s1, const float* s2, const float* i1, const float* i2,
                                const float* flow, const float* mask, float scale, const float* scale_dev,
                                float* ds1, float* ds2, int B, int H, int W, int luminance, void* stream);

/* da = 2*scale*(a-b); db (nullable) = -da  (content / Gram MSE numerators, B6/B7). */
int vst_sqdiff_bwd_f32(const float* a, const float* b, float scale, float* da, float* db, size_t n, void* stream);

/* TV backward for both modes of vst_tv_f32 (B2); dx = scale * dTV/dx. */
int vst_tv_bwd_f32(const float* x, float scale, float* dx, int BC, int H, int W, int mode, void* stream);

/* Gram backward (B7): dy[b] = scale * (dG[b] + dG[b]^T) F[b]. */
int vst_gram_bwd_f32(const float* y, const float* dG, float* dy, int B, int C, int HW, float scale, void* stream);

/* y[i] += x[i] (gradient fan-in: residual adds B15, tap gradients B8). */
int vst_axpy_f32(const float* x, float* y, float alpha, size_t n, void* stream);

/* Loss assembly on the device (B1; RC/...starry-night.py:104-105,121-122,148, RT/train.py:128-139):
 * for entry i: v_i = coef[i] * sums[num_idx[i]] / den_i, den_i = den_idx[i] < 0 ? 1 : sums[den_idx[i]] + den_eps[i];
 * terms_out[group[i]] += v_i, terms_out[n_groups] = total; scale_out[i] = coef[i] / den_i (the factor the
 * backward kernels need).  Index/coef arrays are HOST arrays, n_entries <= 16.  terms_out holds n_groups + 2 floats:
 * the last one is 1 when some den_i was exactly zero (an empty occlusion mask, where the reference raises
 * ZeroDivisionError at `1 / non_zero_count` before backward()); that entry's scale is then 0 instead of inf. */
int vst_loss_terms_f32(const float* sums, const int* num_idx_host, const int* den_idx_host,
                       const float* coef_host, const float* den_eps_host, const int* group_host,
                       int n_entries, int n_groups, float* terms_out, float* scale_out, void* stream);

/* torch.optim.Adam defaults (no weight decay / amsgrad), in place over a flat buffer (B17):
 * g' = g*grad_scale; m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2;
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)   (RC/...starry-night.py:44, RT/train.py:82).
 * skip_flag (nullable, device): when *skip_flag != 0 the launch updates nothing (see vst_loss_terms_f32). */
int vst_adam_f32(float* p, const float* g, float* m, float* v, size_t n, float lr, float b1, float b2,
                 float eps, int step, float grad_scale, const float* skip_flag, void* stream);

/* ======================================================================================
 * bf16 tensor-core path: tcgen05.mma (kind::f16, bf16 in, fp32 TMEM accumulators) fed by TMA.
 * A "plan" owns the packed weights, TMA descriptors and launch geometry of one network at one
 * input shape; buffers live in a caller-provided arena.
 * ====================================================================================== */
typedef struct vst_plan vst_plan;

#define VST_NET_RECONET 0      /* RC/network.py:153-190 (also SD1/SD2 widths via `widths`) */

/* Layer widths: ReCoNet {48,96,192}; SD1 {32,64,64}; SD2 {16,32,64}; in_ch = 3*input_frame_num. */
typedef struct {
  int net;            /* VST_NET_* */
  int in_ch;          /* 3 * input_frame_num */
  int c1, c2, c3;     /* conv1/conv2/trunk widths */
  int d1, d2;         /* deconv1 / deconv2 output widths (96,48 for ReCoNet) */
  int N, H, W;        /* batch, frame size (H%4==0, W%4==0) */
  int flags;          /* VST_PLAN_* */
} vst_net_desc;

/* Plan flag: fp16 (instead of bf16) operands and activation storage, the residual trunk's stream and the raw output of each
 * block's second convolution in fp32.  Same tensor-core rate; 8x finer rounding plus an exact residual add - the mode that holds
 * 2e-2 on trained checkpoints whose `features` tensor has collapsed to ~1 % of the stream feeding it. */
#define VST_PLAN_FP16 1

/* Bytes of device arena the plan needs (activations + packed weights + stats). */
size_t vst_plan_arena_bytes(const vst_net_desc* d);

/* Build a plan.  `weights_host` is an array of `n_tensors` host fp32 pointers in the reference's
 * state_dict order (RC/network.py:157-169: conv1.conv2d.weight, conv1.conv2d.bias,
 * conv1.instance.weight, conv1.instance.bias, ...); they are repacked to bf16 K-major tap
 * matrices inside `arena` (device, `arena_bytes` long).  Synchronous; one-time. */
int vst_plan_create(const vst_net_desc* d, const float* const* weights_host, int n_tensors,
                    void* arena, size_t arena_bytes, void* stream, vst_plan** out);
void vst_plan_destroy(vst_plan* p);

/* ReCoNet.forward on the tensor-core path (replaces RC/network.py:171-190 for inference, the
 * call at RC/utilities.py:218).  x: fp32 NCHW [N,in_ch,H,W] in 0..255.
 * img_out: fp32 NCHW [N,3,H,W] (may be NULL); u8_out: uint8 HWC BGR [N,H,W,3] = the byte image
 * `Inference.__iter__` yields after clamp + cvtColor + astype(uint8) (RC/utilities.py:219-224)
 * (may be NULL); features_out: fp32 NCHW [N,c3,H/4,W/4] (may be NULL). */
int vst_plan_forward(vst_plan* p, const float* x, float* img_out, uint8_t* u8_out,
                     float* features_out, void* stream);

/* The same forward fed with decoder frames: frames_bgr = uint8 [N,H,W,3] in BGR order, i.e. what cv2.VideoCapture.read hands
 * to `cvframe_to_tensor` (RC/utilities.py:119-123, 226-231).  The BGR->RGB swap, the uint8 -> float conversion and the HWC -> CHW
 * permute that the reference runs on the host per frame happen inside the first kernel; results are bit-identical to
 * vst_plan_forward on cvframe_to_tensor's output, the upload is 4x smaller.  Single-frame networks (in_ch == 3) only. */
int vst_plan_forward_bgr8(vst_plan* p, const uint8_t* frames_bgr, float* img_out, uint8_t* u8_out,
                          float* features_out, void* stream);

/* The same forward for TWO half-batch plans (own arenas, same network / precision "bf16") walked in lock step on one
 * stream: every tap-GEMM launch of one plan carries the pending InstanceNorm-apply pass of the other on four extra warps,
 * so the HBM-bound passes between the convolutions (RC/network.py:94-98, 145-150: `self.instance(...)`, ReLU, residual add)
 * run beside the tensor-core work instead of between it.  Results are bit-identical to two vst_plan_forward calls.
 * VST_EUNSUPPORTED for fp16 plans or plans with the timing / stop-after hooks armed. */
int vst_plan_forward_pair(vst_plan* pa, vst_plan* pb, const float* xa, const float* xb, uint8_t* u8_a, uint8_t* u8_b,
                          float* img_a, float* img_b, void* stream);

/* Number of kernel launches one vst_plan_forward issues (for bench.py's gpu_launches). */
int vst_plan_launches(const vst_plan* p);

/* Per-launch timing of the 16 tap-GEMM convolution launches with CUDA events recorded on the
 * forward's stream (the roofline numbers in bench.py).  set_timing(1) starts recording (a ring of
 * the last 32 forwards); get_timing averages them into ms_out[16] (stage order: conv1, conv2, conv3,
 * res1.conv1 ... res5.conv2, deconv1, deconv2, deconv3).  The caller synchronises the stream first. */
int vst_plan_set_timing(vst_plan* p, int enable);
int vst_plan_get_timing(vst_plan* p, float* ms_out, int* n_forwards);

/* Debug/test hook: make vst_plan_forward return after stage `stage` (0 conv1 ... 14 deconv2); -1
 * restores the full forward.  Needed because activation buffers are recycled along the net. */
int vst_plan_set_stop_after(vst_plan* p, int stage);

/* Debug/test hook: copy an internal activation (by layer index, after IN/act) out as fp32 NCHW. */
int vst_plan_debug_activation(vst_plan* p, int layer, float* out_nchw, size_t out_elems, void* stream);

/* Stand-alone tensor-core tap-GEMM convolution on NHWC bf16 (tests + kernel sweep).
 * x: bf16 [N,Hp,Wp,Cin] already padded by `pad` on each side; w: fp32 [Cout,Cin,k,k] (device);
 * y: fp32 NCHW [N,Cout,Ho,Wo] raw conv output (no bias).  Workspace from
 * vst_tc_conv_workspace_bytes(). */
size_t vst_tc_conv_workspace_bytes(int N, int Cin, int H, int W, int Cout, int k);
int vst_tc_conv3x3_f32io(const float* x_nchw, const float* w, float* y_nchw, int N, int Cin, int H,
                         int W, int Cout, int pad_mode, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ======================================================================================
 * Tensor-core primitives on channels-last bf16 buffers: the building blocks of the bf16 training
 * step (forward with saved activations, data gradients, weight gradients, Gram matrices).  The
 * caller (vst_b200/tc.py) owns every buffer and describes geometry with plain structs.
 * ====================================================================================== */

/* A channels-last bf16 activation buffer: [N][H+2*pad][W+2*pad][C], or - with parity != 0 - the same
 * padded tensor split into 4 row/col-parity planes [4][N][(H+2pad)/2][(W+2pad)/2][C] (operand of
 * stride-2 taps).  kind says how a producer fills the halo: 0 reflect (RC/network.py:67-73),
 * 1 replicate (the low-res operand of the nearest-x2 convolutions), 2 zero. */
typedef struct { int H, W, C, pad, kind, parity; } vst_act_desc;

#define VST_TG_MAX_TAPS 96
#define VST_EPI_BF16_NHWC 0
#define VST_EPI_F32_NCHW 1
#define VST_EPI_ROWCONV 2

/* One tap-GEMM convolution launch (csrc/tc_conv.cu):
 *   out[n][(y*out_mul + ph_oy) ][(x*out_mul + ph_ox)][co] = epi( sum_{t < n_taps} sum_{ci}
 *        a[n][plane_t][y + tap_dy_t][x + tap_dx_t][ci] * b[(phase*n_ntile + ntile)*N_mma + co][(t*kb_per_tap)*BK + ci] )
 * for (y, x) in grid_h x grid_w, every phase.  Coordinates outside `a` read as zero (TMA fill). */
typedef struct {
  const void* a; int a_C, a_X, a_Y, a_N, a_P;   /* dense [P][N][Y][X][C] bf16 */
  const void* b; int b_K, b_rows;               /* packed weights [rows][K] bf16, K-major */
  int b_img_rows;                               /* rows to skip per image (per-image weights), else 0 */
  int BK, kb_per_tap, n_taps, n_phase, n_ntile, N_mma;
  int grid_h, grid_w, out_mul;
  int Hout, Wout, Cout, out_cstride;
  int epi_mode, act, relu;
  int rc_k, rc_co, tile_step_x, TW, TH, MT;     /* 0 = choose */
  void* out; unsigned char* out_u8; const float* bias; double* stats;
  signed char tap_dx[VST_TG_MAX_TAPS], tap_dy[VST_TG_MAX_TAPS], tap_pl[VST_TG_MAX_TAPS];
  signed char ph_oy[4], ph_ox[4];
} vst_tapgemm_desc;
int vst_tc_tapgemm(const vst_tapgemm_desc* d, void* stream);

/* Host-only: the pipeline mode and tiling vst_tc_tapgemm would pick for `d` (no pointer is dereferenced, nothing is
 * launched).  stream: 0 per-tap boxes / 1 shared-memory row ring / 2 accumulator ring; dyshare: taps grouped into n_cols
 * columns per phase - column c of phase p (index p*n_cols + c) starts at tap (col_dx, col_dy0, col_pl), has col_n taps
 * with consecutive dy whose weight-tap indices are col_t0 + j*col_ts, and reads one TMA box of box_rows rows. */
typedef struct {
  int stream, dyshare, n_cols, dy_max, box_rows, TW, TH, MT, tiles_x, tiles_y;
  int cta2;   /* 1: launched as clusters of two CTAs issuing M = 256 tcgen05.mma.cta_group::2 (each CTA stages half the weights) */
  signed char col_dx[48], col_dy0[48], col_pl[48], col_n[48], col_t0[48], col_ts[48];
} vst_tapgemm_plan_info;
int vst_tc_tapgemm_plan(const vst_tapgemm_desc* d, vst_tapgemm_plan_info* info);

/* Pixel-contraction GEMM (csrc/tc_pcgemm.cu) - weight gradients (B13) and Gram matrices (a13, B7):
 *   out[img?][t][m][n] += scale * sum_{img?, y < grid_h, x < grid_w}
 *        a[img][a_pl_t][y + a_dy_t][x + a_dx_t][m] * b[img][b_pl_t][y + b_dy_t][x + b_dx_t][n]
 * per_image != 0 keeps one output per image (Gram); otherwise images are summed (weight gradient).
 * `out` is fp32 and must be zeroed by the caller (K splits accumulate with atomics). */
typedef struct {
  const void* a; int a_C, a_X, a_Y, a_N, a_P;
  const void* b; int b_C, b_X, b_Y, b_N, b_P;
  int n_img, grid_h, grid_w, n_taps, M, N, per_image, k_splits;
  float scale; float* out;
  signed char a_dx[VST_TG_MAX_TAPS], a_dy[VST_TG_MAX_TAPS], a_pl[VST_TG_MAX_TAPS];
  signed char b_dx[VST_TG_MAX_TAPS], b_dy[VST_TG_MAX_TAPS], b_pl[VST_TG_MAX_TAPS];
} vst_pcgemm_desc;
int vst_tc_pcgemm(const vst_pcgemm_desc* d, void* stream);

/* dst[i] = sum_j src[idx[i*terms + j]] (idx < 0 = skip): weight packing (fp32 -> bf16 tap matrices, incl. the
 * pre-summed phase weights of the x2-upsample convs) and weight-gradient unpacking (fp32 -> fp32) are
 * both table-driven gathers; the tables are built once per layer on the host. */
int vst_gather_sum_f32(const float* src, const int* idx, int terms, void* dst, size_t n, int dst_bf16, void* stream);

/* fp32 NCHW [N,Cin,H,W] -> padded channels-last bf16 (channels Cin..C-1 zero), and back (interior only). */
int vst_tc_nchw_to_act(const float* x, int Cin, void* dst, vst_act_desc L, int N, void* stream);
int vst_tc_act_to_nchw(const void* act, vst_act_desc L, int N, float* out, void* stream);
/* The first `c_count` (1..8) channels only, as fp32 NCHW [N, c_count, H, W]: the 3 real channels of a stylising network's
 * 16-channel-padded output operand (RT/network.py:88-90 `conv4`), without converting and slicing the 13 zero channels. */
int vst_tc_act_to_nchw_first(const void* act, vst_act_desc L, int N, int c_count, float* out, void* stream);
/* conv1 operand: fp32 NCHW frame -> X9 [N][H+8][W][KR] (per pixel the 9 x Cin (kx, c) window; KR = 32 for Cin = 3). */
int vst_tc_prologue_x9(const float* x, void* x9, int N, int Cin, int H, int W, int KR, void* stream);

/* First VGG convolution operand: fp32 NCHW [N,3,H,W] -> [N][H][W][32] bf16 with channel (ky*3+kx)*3 + c = x[c][y+ky-1][x+kx-1]
 * (zero outside, channels 27..31 zero): the 3x3x3 -> 64 convolution becomes ONE K = 32 GEMM tap instead of nine K = 64 taps. */
int vst_tc_prologue_x27(const float* x, void* out, int N, int H, int W, void* stream);

/* Row-convolution operand of the k x k, few-output-channel layer's adjoints (ConvTanh, RC/network.py:78-85):
 * E[n][y][x'][kx*Co + co] = dz[n][co][y][x' - kx] (0 outside), x' over the W + k - 1 padded columns, KE channels
 * (k*Co <= KE, rest zero).  With it the weight gradient is a 9-tap pixel contraction and the data gradient a
 * 9-tap GEMM over E, mirroring the forward row convolution. */
int vst_tc_rowconv_expand(const float* dz, void* E, int N, int Co, int H, int W, int k, int KE, void* stream);

/* y = act(InstanceNorm(raw)) (+ residual) written into the consumer's padded layout (RC/network.py:95-97,146-149).
 * raw: [N][H][W][C] bf16; stats: [N][C][2] fp64 = sum, sum of squares (from the tap-GEMM epilogue:
 * deterministic per-CTA fp32 partials meeting in fp64 atomics). */
int vst_tc_in_apply(const void* raw, const double* stats, const float* gamma, const float* beta, const void* residual,
                    vst_act_desc res_desc, void* dst, vst_act_desc dst_desc, int N, float eps, int relu, void* stream);

/* InstanceNorm + ReLU backward on channels-last bf16 (B11, B12, B15), two passes.
 * The incoming gradient is g = fold(G) (+ skip): G is the data gradient of the consumer convolution over ITS padded
 * input domain [N][H+2gp][W+2gp][C] (gp = g_desc.pad, folded back per g_desc.kind: reflect / replicate / none);
 * skip (nullable) is an unpadded [N][H][W][C] gradient (residual fan-in, feature-temporal gradient).
 *   reduce: red[n][c] = { sum g', sum g' * xhat },  g' = g * relu'(xhat*gamma + beta)
 *   apply : draw = gamma*rstd*(g' - mean(g') - xhat*mean(g'*xhat)) -> `draw` in layout draw_desc (pad 0; plain or parity);
 *           gsum (nullable) <- g (unpadded), the gradient w.r.t. this layer's output for the residual skip. */
int vst_tc_in_bwd_reduce(const void* G, vst_act_desc g_desc, const void* skip, const void* raw, const double* stats,
                         const float* gamma, const float* beta, float* red, int N, float eps, int relu, void* stream);
int vst_tc_in_bwd_apply(const void* G, vst_act_desc g_desc, const void* skip, const void* raw, const double* stats,
                        const float* gamma, const float* beta, const float* red, void* draw, vst_act_desc draw_desc,
                        void* gsum, int N, float eps, int relu, void* stream);
/* dgamma[c] = sum_n red[n][c][1], dbeta[c] = sum_n red[n][c][0]. */
int vst_tc_in_param_grads(const float* red, float* dgamma, float* dbeta, int N, int C, void* stream);

/* VGG body on channels-last bf16 (unpadded [N][H][W][C]): max_pool2d(2,2) floor, and the fused adjoint of
 * [ReLU -> optional max-pool]: gm = (g_up + add) * (y > 0), where g_up = g (pooled == 0, same size as y) or the
 * max-pool routing of g (pooled != 0, g is [N][H/2][W/2][C]; first maximum in scan order wins). */
int vst_tc_maxpool2(const void* x, void* y, int N, int H, int W, int C, void* stream);
int vst_tc_relu_pool_bwd(const void* g, const void* y, const void* add, void* gm, int N, int H, int W, int C, int pooled,
                         void* stream);

/* bf16 loss helpers: out[0] = sum (a-b)^2 (content term on VGG taps); da = 2*scale*(a-b) as bf16. */
int vst_tc_sqdiff_sum_bf16(const void* a, const void* b, float* out, float* scratch, size_t n, void* stream);
int vst_tc_sqdiff_bwd_bf16(const void* a, const void* b, float scale, void* da, size_t n, void* stream);
/* Style-term gradient as per-image 1x1 weights: S[b][i][j] = scale*(D[b][i][j] + D[b][j][i]), D = G - Gs (Gs broadcast
 * over the batch when gs_batch == 1), written bf16 [B][C][C]; dF = S F is then one tap-GEMM launch (B7). */
int vst_tc_gram_grad_weights(const float* G, const float* Gs, int gs_batch, float scale, void* S, int B, int C, void* stream);
/* y += x on bf16 buffers (tap-gradient fan-in). */
int vst_tc_add_bf16(const void* x, void* y, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VST_B200_H */
