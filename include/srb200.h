/*
 * srb200 -- C ABI of the B200-native (sm_100a) super-resolution hot path.
 *
 * This is the drop-in boundary for watercore2001/BasicSR4RS' EDSR / RCAN / SwinIR
 * forward+backward.  Every entry point replaces a group of aten:: library calls the
 * reference issues from basicsr/archs/{arch_util,edsr_arch,rcan_arch,swinir_arch}.py
 * (file:line given per function).  Conventions:
 *   - plain device pointers + sizes; no torch types
 *   - the caller owns every buffer (workspaces included); the library never
 *     allocates, never synchronises, never changes the current device
 *   - work is enqueued on `stream` and is asynchronous w.r.t. the host
 *   - return 0 on success, a negative SRB200_E* code otherwise
 *
 * Activation layout inside a network: NHWC, bf16, channel count padded to a multiple
 * of 64 with zeros ("NHWC64").  Weights are fp32 nn.Parameters on the host side and are
 * repacked to bf16 tap-major GEMM operands by srb200_pack_weight(s) (derived cache).
 *
 * Kernel-variant selectors read from the environment at call time (tuning / A-B measurements only; every
 * variant computes the same result): SRB_TAPGEMM_1CTA (no cta_group::2 pairs), SRB_TAPGEMM_NO_TEAMS (one 8-warp
 * epilogue instead of two 4-warp teams), SRB_WGRAD_1CTA, SRB_WG_KPIX / SRB_WG2_KPIX = 64 | 128 | 256 (pixels per
 * weight-gradient pipeline stage), SRB_WG_SPLITS (split-K factor of the 1-CTA weight-gradient kernel), SRB_WG_FOLD = 0 | 1 (three taps of a stencil row per
 * weight-gradient unit for 64-wide input-channel tiles; default: 64-output-channel layers with >= 100k pixels).
 */
#ifndef SRB200_H_
#define SRB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* srb200_stream_t; /* cudaStream_t */

enum {
  SRB200_OK = 0,
  SRB200_EINVAL = -1,   /* bad shape / alignment / flag combination */
  SRB200_EARCH = -2,    /* device is not sm_100 */
  SRB200_EDRIVER = -3,  /* cuTensorMapEncodeTiled unavailable or failed */
  SRB200_ELAUNCH = -4   /* kernel launch failed */
};

enum { SRB200_ACT_NONE = 0, SRB200_ACT_RELU = 1, SRB200_ACT_LRELU = 2, SRB200_ACT_GELU = 3 };
enum {
  SRB200_OUT_NHWC = 0,     /* bf16 [B,H,W,ldo]                                         */
  SRB200_OUT_SHUFFLE = 1,  /* bf16 [B,H*r,W*r,Cout/r^2]: nn.PixelShuffle fused in store */
  SRB200_OUT_NCHW_F32 = 2  /* fp32 [B,out_c,H,W], v*out_scale + out_shift[c]            */
};
enum { SRB200_MASK_NONE = 0, SRB200_MASK_SIGN = 1, SRB200_MASK_DGELU = 2,
       SRB200_MASK_MUL = 3 /* v *= mask_src: the saved activation derivative (aux_mode 1) */ };

const char* srb200_version(void);
const char* srb200_strerror(int code);
/* 0 when device `dev` can run the kernels (compute capability 10.x). */
int srb200_check_device(int dev);

/* ------------------------------------------------------------------ index remaps
 * Bit-exact; any 2-byte or 4-byte element type (elem_bytes = 2 | 4).
 * pixel_shuffle: nn.PixelShuffle(r) of arch_util.py:131-139 on NCHW:
 *   out[b, c, y*r+i, x*r+j] = in[b, c*r*r + i*r + j, y, x];  bwd = its inverse.        */
int srb200_pixel_shuffle_nchw(const void* in, void* out, int B, int C_out, int H, int W, int r,
                              int elem_bytes, int inverse, srb200_stream_t stream);
/* Same map on NHWC tensors: in [B,H,W,C*r*r] (channel = c*r*r+i*r+j) -> out [B,H*r,W*r,C]. */
int srb200_pixel_shuffle_nhwc(const void* in, void* out, int B, int C_out, int H, int W, int r,
                              int elem_bytes, int inverse, srb200_stream_t stream);
/* window_partition (swinir_arch.py:63-75) fused with the cyclic shift torch.roll(x,(-s,-s),(1,2))
 * (swinir_arch.py:293-296):  win[(b*nWh+wy)*nWw+wx, iy, ix, :] = x[b, (wy*ws+iy+s)%H, (wx*ws+ix+s)%W, :]
 * inverse=1 is window_reverse (:78-92) followed by roll(+s,+s) (:313-316).               */
int srb200_window_remap(const void* in, void* out, int B, int H, int W, int C, int ws, int shift,
                        int elem_bytes, int inverse, srb200_stream_t stream);
/* torch.roll(x, (sy,sx), dims=(1,2)) on [B,H,W,C]. */
int srb200_roll_nhwc(const void* in, void* out, int B, int H, int W, int C, int sy, int sx,
                     int elem_bytes, srb200_stream_t stream);

/* ------------------------------------------------------------------ layout entry/exit
 * NCHW fp32 -> NHWC bf16 with C_pad channels (zeros beyond C):
 *   out[b,y,x,c] = bf16((in[b,c,y,x] - shift[c]) * scale)     (edsr_arch.py:53, mean/range)
 * shift may be NULL (=0).                                                                */
int srb200_nchw_to_nhwc(const float* in, void* out_bf16, int B, int C, int H, int W, int C_pad,
                        const float* shift, float scale, srb200_stream_t stream);
/* NHWC bf16 (C_pad channels) -> NCHW fp32, out = in*scale + shift[c] (edsr_arch.py:59).   */
int srb200_nhwc_to_nchw(const void* in_bf16, float* out, int B, int C, int H, int W, int C_pad,
                        const float* shift, float scale, srb200_stream_t stream);

/* Nearest-neighbour x2 on NHWC bf16 (C % 8 == 0): F.interpolate(scale_factor=2, mode='nearest') of SwinIR's
 * 'nearest+conv' reconstruction (swinir_arch.py:911-912).  inverse=0: in [B,H,W,C] -> out [B,2H,2W,C] (bit-exact
 * copy); inverse=1 (its gradient): in [B,2H,2W,C] -> out [B,H,W,C], each output = fp32 sum of its 2x2 block.     */
int srb200_nearest_up2(const void* in_bf16, void* out_bf16, int B, int H, int W, int C, int inverse,
                       srb200_stream_t stream);

/* ------------------------------------------------------------------ few-channel exit conv (conv_last, F -> C <= 7)
 * nn.Conv2d(F, C, 3, 1, 1) + `x / img_range + mean` (edsr_arch.py:48,58-59; rcan_arch.py:122,132-133;
 * swinir_arch.py:839,900,920) as ONE 1x1 tap-GEMM  T[p][tap*C + c] = X[p] . W[c][:][tap]  (reads X once instead of
 * nine times) followed by srb200_tap_stencil:
 *   out[b,c,y,x] = (sum_tap T[b, y+dy(tap), x+dx(tap), tap*C + c] + bias[c]) * scale + shift[c]     (NCHW fp32)
 * T is fp32 [B,H,W,Tp] (Tp >= 9*C); taps outside the image contribute zero (the conv's zero padding).            */
int srb200_tap_stencil(const float* t_f32, float* out, int B, int C, int H, int W, int Tp, const float* bias,
                       const float* shift, float scale, srb200_stream_t stream);
/* Backward twin: G[b,y,x, tap*C + c] = scale * g[b,c, y-dy(tap), x-dx(tap)] as NHWC bf16 with Gp channels (zero
 * beyond 9*C and outside the image).  Then dX = G . Wf (1x1 tap-GEMM) and dWf = G^T X (1x1 srb200_wgrad).       */
int srb200_tap_im2col(const float* g, void* out_bf16, int B, int C, int H, int W, int Gp, float scale,
                      srb200_stream_t stream);

/* ------------------------------------------------------------------ weight packing
 * conv / linear weight fp32 [Co, Ci, taps] (OIHW flattened; taps = 1 or 9) -> bf16
 *   transpose=0: out[t][n][k] = w[perm_out[n]][perm_in[k]][t]   (fprop operand, [taps][Np][Kp])
 *   transpose=1: out[t][k][n] = w[perm_out[n]][perm_in[k]][t]   (dgrad operand, [taps][Kp][Np])
 * perm_* are device int32 arrays mapping packed index -> original index, -1 = zero pad.   */
int srb200_pack_weight(const float* w, int Co, int Ci, int taps, const int32_t* perm_out, int Np,
                       const int32_t* perm_in, int Kp, int transpose, void* out_bf16,
                       srb200_stream_t stream);
/* inverse of pack for the weight gradient: gw[perm_out[n]][perm_in[k]][t] = alpha*acc[t][n][k]. */
int srb200_unpack_wgrad(const float* acc, float* gw, int Co, int Ci, int taps,
                        const int32_t* perm_out, int Np, const int32_t* perm_in, int Kp,
                        float alpha, srb200_stream_t stream);

/* Batched pack / unpack: one launch for every weight of a network (hundreds of layers).  `items_dev` is a
 * DEVICE array; item i covers 256-pair chunks [chunk_begin, chunk_begin + ceil(Np*Kp/256)) of the work list
 * (chunk_begin ascending, starting at 0; total_chunks = their sum).  Pack: src = fp32 weight, dst = bf16
 * operand (semantics of srb200_pack_weight).  Unpack: src = fp32 [taps][Np][Kp] accumulator, dst = fp32
 * gradient in the parameter layout, scaled by alpha (semantics of srb200_unpack_wgrad).                  */
typedef struct {
  const void* src;
  void* dst;
  const int32_t* perm_out; /* device, or NULL */
  const int32_t* perm_in;  /* device, or NULL */
  int64_t Co, Ci, taps, Np, Kp;
  int64_t transpose;       /* pack only: 0 / 1 = bf16 operand layouts of srb200_pack_weight, 2 = fp32 copy
                              out[n] = w[perm_out[n]] (zero padded): bias vectors (taps = Kp = Ci = 1) */
  int64_t chunk_begin;
  float alpha;             /* unpack: gradient scale.  pack, transpose = 2: the pad constant (see reserved) */
  int32_t reserved;        /* pack, transpose = 2: 1 + index of ONE pad element that receives `alpha` instead of zero
                              (0 = none) -- e.g. the bias v with GELU(v) = 1 that turns a pad channel of fc1's output
                              into a constant one, so that fc2's weight-gradient GEMM also yields its bias gradient */
} srb200_pack_item;
int srb200_pack_weights(const srb200_pack_item* items_dev, int n_items, int64_t total_chunks,
                        srb200_stream_t stream);
int srb200_unpack_wgrads(const srb200_pack_item* items_dev, int n_items, int64_t total_chunks,
                         srb200_stream_t stream);
/* Same as srb200_unpack_wgrads for up to SRB200_MAX_INLINE_ITEMS items given in HOST memory (copied into the
 * kernel parameters; chunk_begin is computed by the library): all weight and bias gradients of one layer's
 * backward in one launch, recordable into a CUDA graph.  A bias gradient is an item with taps = Kp = Ci = 1
 * whose src is the fp32 column-sum vector.                                                                */
#define SRB200_MAX_INLINE_ITEMS 8
int srb200_unpack_wgrads_inline(const srb200_pack_item* items_host, int n_items, srb200_stream_t stream);

/* ------------------------------------------------------------------ multi-tensor EMA
 * dst_i = a * dst_i + b * src_i for every fp32 tensor pair of a DEVICE table, one launch: the exponential moving
 * average of BaseModel.model_ema (basicsr/models/base_model.py:75-82; a = decay, b = 1 - decay) without its two
 * launches per parameter.  Item i covers 1024-float chunks [chunk_begin, chunk_begin + ceil(n/1024)).           */
typedef struct {
  const void* src;
  void* dst;
  int64_t n;
  int64_t chunk_begin;
} srb200_vec_item;
int srb200_multi_axpby(const srb200_vec_item* items_dev, int n_items, int64_t total_chunks, float a, float b,
                       srb200_stream_t stream);

/* ------------------------------------------------------------------ multi-tensor Adam
 * One launch = torch.optim.Adam's update (amsgrad off, L2 weight decay) of every fp32 tensor of a DEVICE table: the
 * optimizer step of SRModel.optimize_parameters (basicsr/models/sr_model.py:113) without torch's ~36-tensors-per-launch
 * argument packing.  Item i covers 1024-float chunks [chunk_begin, chunk_begin + ceil(n/1024)).  `step` = 1, 2, ... is the
 * update count (the hyper-parameters are doubles, as Python holds them: 1 - beta and the bias corrections 1 - beta^step
 * are formed on the host in double precision, like torch does, then rounded to fp32 once).                      */
typedef struct {
  void* p;        /* parameter, updated in place */
  const void* g;  /* gradient */
  void* m;        /* exp_avg */
  void* v;        /* exp_avg_sq */
  int64_t n;
  int64_t chunk_begin;
} srb200_adam_item;
int srb200_multi_adam(const srb200_adam_item* items_dev, int n_items, int64_t total_chunks, double lr, double beta1,
                      double beta2, double eps, double weight_decay, int64_t step, srb200_stream_t stream);

/* ------------------------------------------------------------------ tap-GEMM (conv3x3 / conv1x1 / Linear)
 * out[b,y,x,n] = epi( sum_{t,k} A[b, y+dy(t), x+dx(t), k] * Wp[t][n][k] )
 * TMA-fed tcgen05 implicit GEMM, fp32 accumulation in TMEM.  Replaces nn.Conv2d(…,3,1,1)
 * (arch_util.py:79-80,134,137; edsr_arch.py:44-48; rcan_arch.py:39-41,65,109-121;
 * swinir_arch.py:532,818,834-837) and nn.Linear (swinir_arch.py:51-53,135-137), forward
 * and the data gradient (flip=1 with the transposed pack).
 * epi(v): v += bias[n]; act; v *= alpha; mask (v *= act'(mask_src)); v += residual.       */
typedef struct {
  int32_t B, H, W;      /* GEMM rows = B*H*W output pixels                                 */
  int32_t Cin;          /* channels of A per source view, multiple of 64                   */
  int32_t src_r;        /* 1, or r: A is read through a pixel-unshuffle view of
                           in[B,H*r,W*r,Cin]; K per tap = r*r*Cin ordered (i*r+j, c)       */
  int32_t Cout;         /* packed N, multiple of 16                                        */
  int32_t ksize;        /* 1 or 3                                                          */
  int32_t flip;         /* 0: taps as correlation (fprop); 1: negated offsets (dgrad)      */
  int32_t act;          /* SRB200_ACT_*                                                    */
  float act_slope;      /* LeakyReLU negative slope                                        */
  float alpha;          /* scale after activation (res_scale)                              */
  int32_t mask_mode;    /* SRB200_MASK_*                                                   */
  float mask_slope;
  int32_t out_mode;     /* SRB200_OUT_*                                                    */
  int32_t out_r;        /* pixel-shuffle factor for SRB200_OUT_SHUFFLE                     */
  int32_t out_c;        /* SRB200_OUT_NCHW_F32: real channel count (<= Cout)               */
  float out_scale;      /* SRB200_OUT_NCHW_F32                                             */
} srb200_tapgemm_desc;

/* optional extras of the epilogue (pass NULL when unused) */
typedef struct {
  const float* residual_f32;     /* fp32 residual, layout of out: v += residual_f32 (fp32 skip stream)  */
  float* out_f32;                /* additionally store the fp32 result (layout of out)                  */
  const float* alpha_per_sample; /* [B] extra scale per batch sample (DropPath, swinir_arch.py:14-26)   */
  int32_t aux_mode;              /* what aux_out receives: 0 = the pre-activation (bias added), 1 = the DERIVATIVE of
                                    the activation at the pre-activation (SRB200_ACT_GELU: gelu'(a), computed from
                                    the same erf / exp as gelu(a)); the backward then needs SRB200_MASK_MUL only   */
  int32_t colsum_per_image;      /* 1: colsum is [B][Cout], one row per sample (RCAN's AdaptiveAvgPool2d(1),
                                    rcan_arch.py:19, for free from the conv that produces its input)             */
  float* colsum;                 /* fp32 [Cout], zeroed by the caller: += sum over pixels of the stored
                                    result (the bias gradient of the layer that consumes `out` as dY);
                                    SRB200_OUT_NHWC with Cout % 64 == 0 only                            */
  float colsum_scale;            /* the sums are multiplied by this when flushed (0 = 1.0; 1/HW for a mean)        */
  int32_t flags;                 /* SRB200_EXT_*                                                                    */
  /* LayerNorm folded into the GEMMs on either side of it (evaluation; swinir_arch.py:288-289,320-321: norm -> Linear):
   * the PRODUCER of a row tensor x (Cout = one N tile of 128 / 192 channels, plain NHWC output) also writes
   * ln_stats_out[pixel] = (mean, rstd) over the first ln_channels stored (bf16-rounded) channels, with ln_eps;
   * the CONSUMER Linear runs on the raw x with weights W' = W diag(gamma) and computes
   *     v = rstd * (acc - mean * ln_wsum[n]) + bias[n],  ln_wsum[n] = sum_c W'[n,c] (of the bf16-rounded packed
   * weights), bias = W beta + b  --  exactly W LN(x) + b, without a LayerNorm pass or a normalised copy of x.
   * Both sides need a layer whose N tile is 128 or 192 wide (Cout % 192 == 0, or % 128 == 0 and % 256 != 0) on the
   * TMA-store path; anything else returns SRB200_EINVAL.                                                            */
  const float* ln_stats_in;      /* fp32 [pixels][2] or NULL (with ln_wsum; not together with alpha_per_sample)     */
  const float* ln_wsum;          /* fp32 [Cout]                                                                     */
  float* ln_stats_out;           /* fp32 [pixels][2] or NULL                                                        */
  int32_t ln_channels;           /* real channel count of the rows ln_stats_out describes                           */
  float ln_eps;
} srb200_tapgemm_ext;
/* alpha_per_sample multiplies the BIAS only: v = acc + alpha[b] * bias.  For a Linear whose INPUT already carries the
 * per-sample DropPath factor (h' = alpha[b] * h stored by the producing epilogue, so that the weight-gradient GEMM
 * dW = dY^T h' needs no scaled copy of dY): alpha (h W^T + bias) = h' W^T + alpha bias.                              */
#define SRB200_EXT_ALPHA_ON_BIAS 1

int srb200_tapgemm(const srb200_tapgemm_desc* d, const void* in_bf16, const void* w_packed,
                   const float* bias,          /* [Cout] or NULL                           */
                   const void* mask_src,       /* bf16, layout of out, or NULL             */
                   const void* residual,       /* bf16, layout of out, or NULL             */
                   const float* out_shift,     /* [out_c] for NCHW_F32 or NULL             */
                   void* out,                  /* see out_mode                             */
                   void* aux_out,              /* optional bf16 copy of the pre-activation */
                   const srb200_tapgemm_ext* ext, /* optional, may be NULL                 */
                   srb200_stream_t stream);

/* ------------------------------------------------------------------ weight gradient
 * acc[t][n][k] += sum_{b,y,x} dY[b,y,x,n] * X[b, y+dy(t), x+dx(t), k]      (fp32 atomics)
 * tcgen05 GEMM with both operands MN-major, split over pixels.  acc must be zeroed by the
 * caller; finish with srb200_unpack_wgrad.  dy_r > 1 reads dY through the pixel-unshuffle
 * view (gradient of the fused PixelShuffle store).                                        */
int srb200_wgrad(const void* dy_bf16, const void* x_bf16, float* acc, int B, int H, int W,
                 int N /* dY channels (packed) */, int K /* X channels */, int ksize, int dy_r,
                 srb200_stream_t stream);
/* bias gradient: out[c] += sum over rows of dY[rows, C] (fp32 atomics; zero `out` first).
 * r > 1: dY is [B, H*r, Wf = W*r, C] and out[((Y%r)*r + X%r)*C + c] collects the r*r
 * pixel-unshuffle phases separately (bias gradient of a conv whose store fused PixelShuffle). */
int srb200_colsum(const void* dy_bf16, float* out, int64_t rows, int C, int r, int Wf,
                  srb200_stream_t stream);

/* ------------------------------------------------------------------ RCAN channel attention
 * ChannelAttention (rcan_arch.py:8-24) + RCAB tail (rcan_arch.py:44-46) on NHWC bf16 [B,HW,C]:
 *   channel_pool : out[b,c] += mean_hw t[b,hw,c]                     (AdaptiveAvgPool2d(1); zero out first)
 *   channel_dot  : out[b,c] += scale * sum_hw a[b,hw,c]*b[b,hw,c]    (backward: d s)
 *   ca_fc        : z = relu(W1 p + b1), s = sigmoid(W2 z + b2)       (the two 1x1 convs, fp32 params as stored)
 *   ca_apply     : y = x + res_scale * t * s[b,c]                    (x * y and res * res_scale + x fused)
 *   ca_fc_bwd    : parameter grads of the two 1x1 convs (summed over the batch) and gp = dL/dp
 *   ca_apply_bwd : gt = res_scale * g * s[b,c] + gp[b,c] / HW        (dL/dt incl. the pool branch)        */
int srb200_channel_pool(const void* t_bf16, float* out, int B, int HW, int C, srb200_stream_t stream);
int srb200_channel_dot(const void* a_bf16, const void* b_bf16, float* out, int B, int HW, int C,
                       float scale, srb200_stream_t stream);
int srb200_ca_fc(const float* p, const float* w1, const float* b1, const float* w2, const float* b2,
                 float* z, float* s, int B, int C, int Cr, srb200_stream_t stream);
/* x_f32 / y_f32 (optional): carry the skip stream in fp32 next to its bf16 copy -- RCAN stacks 200
 * res_scale=1 additions and needs it to stay inside the 1e-2 output bar.   */
int srb200_ca_apply(const void* t_bf16, const void* x_bf16, const float* x_f32, const float* s,
                    void* y_bf16, float* y_f32, int B, int HW, int C, float res_scale,
                    srb200_stream_t stream);
int srb200_ca_fc_bwd(const float* gs, const float* s, const float* z, const float* p, const float* w1,
                     const float* w2, float* gw1, float* gb1, float* gw2, float* gb2, float* gp,
                     int B, int C, int Cr, srb200_stream_t stream);
/* gt = res_scale*g*s[b,c] + gp[b,c]/HW; colsum (optional fp32 [C], zeroed by the caller) += column sums of gt,
 * the bias gradient of the conv that produced t.                                                          */
int srb200_ca_apply_bwd(const void* g_bf16, const float* s, const float* gp, void* gt_bf16, int B,
                        int HW, int C, float res_scale, float* colsum, srb200_stream_t stream);

/* Fused forms (one launch each way) of the four / three calls above, used by the RCAB Function:
 * srb200_ca_forward  = srb200_ca_fc + srb200_ca_apply (p = per-sample channel means, e.g. from srb200_tapgemm's
 *                      per-image column sums); writes z [B,Cr], s [B,C] for the backward.
 * srb200_ca_backward = srb200_channel_dot -> srb200_ca_fc_bwd -> srb200_ca_apply_bwd (+ column sums of gt) in one
 *                      persistent kernel whose phases are joined by a device-wide arrive barrier (the grid is
 *                      sized to at most half of what the device holds co-resident).  gs [B*C], colsum [C] and
 *                      sync [1 x uint32] must be zeroed by the caller; C/8 must be a power of two.               */
int srb200_ca_forward(const void* t_bf16, const void* x_bf16, const float* x_f32, const float* p, const float* w1,
                      const float* b1, const float* w2, const float* b2, float* z, float* s, void* y_bf16,
                      float* y_f32, int B, int HW, int C, int Cr, float res_scale, srb200_stream_t stream);
int srb200_ca_backward(const void* g_bf16, const void* t_bf16, const float* s, const float* z, const float* p,
                       const float* w1, const float* w2, float* gs, float* gw1, float* gb1, float* gw2, float* gb2,
                       void* gt_bf16, float* colsum, unsigned int* sync, int B, int HW, int C, int Cr,
                       float res_scale, srb200_stream_t stream);

/* ------------------------------------------------------------------ SwinIR token kernels
 * nn.LayerNorm(C, eps) over the C real channels of NHWC bf16 rows padded to Cp (pads stay 0)
 * (swinir_arch.py:240,251,288,321,602,886).  mean/rstd [T] fp32 are saved for the backward.
 * bwd: gx = LN'(gy) + gres (gres optional: fused add of the skip-path gradient);
 *      ggamma[c] += sum_t gy*xhat, gbeta[c] += sum_t gy (zero them first).                               */
int srb200_layernorm_fwd(const void* x_bf16, const float* gamma, const float* beta, void* y_bf16,
                         float* mean, float* rstd, int64_t T, int C, int Cp, float eps, int ones_channel,
                         srb200_stream_t stream);
/* ones_channel (-1 = none): a PAD channel (C <= ones_channel < Cp) of y that is set to 1.0 instead of 0.  The Linear
 * that consumes y has zero weights there, and its weight-gradient GEMM acc[n][ones_channel] = sum_tokens dY[n] * 1 is
 * that layer's bias gradient -- no separate column-sum pass.                                                        */
int srb200_layernorm_bwd(const void* gy_bf16, const void* x_bf16, const float* mean, const float* rstd,
                         const float* gamma, const void* gres_bf16, void* gx_bf16, float* ggamma,
                         float* gbeta, int64_t T, int C, int Cp, srb200_stream_t stream);
/* out[b, :] = g[b, :] * alpha[b]  (DropPath backward, swinir_arch.py:14-26)                              */
int srb200_scale_rows(const void* g_bf16, const float* alpha, void* out_bf16, int B,
                      int64_t elems_per_sample, srb200_stream_t stream);
/* Fused (shifted-)window attention, WindowAttention.forward minus the two Linear layers
 * (swinir_arch.py:151-172) together with roll / window_partition / window_reverse / roll
 * (:293-316) and the analytic 0/-100 mask (:262-281):
 *   qkv [B,H,W,3*Cp] bf16, channel = which*Cp + head*32 + d  (head_dim <= 32, zero padded)
 *   out [B,H,W,Cp]   bf16, channel = head*32 + d
 *   rpb_table fp32 [(2*ws-1)^2, num_heads] exactly as stored in the state dict
 * window_size 2..8; H, W multiples of it; shift in [0, window_size).  Window 8 with shift 0 / 4 and an even head
 * count runs on the tcgen05 / TMEM / TMA kernels (attention_tc.cu), everything else on the mma.sync kernels.
 * stats (optional, may be NULL): fp32 [B*(H/ws)*(W/ws)][num_heads][ws*ws] softmax statistics, the log2-domain
 * log-sum-exp of every (window, head, query) row in the kernel's own token order -- an opaque buffer that the
 * forward fills and the backward of the SAME geometry reads so that it need not redo the row maxima and sums
 * (without it the backward recomputes them on the mma.sync path).
 * bwd also accumulates the bias-table gradient into g_rpb_table (zero it first).
 * workspace: reserved, may be NULL (earlier builds merged the bias-table gradient across CTAs through it; every CTA
 * now bins its own sums in shared memory); without stats the backward runs on the mma.sync path.                  */
/* flags: SRB200_ATTN_ONES -- write 1.0 instead of 0 into out channel 31 (a pad lane of head 0; needs head_dim < 32): the
 * proj Linear that consumes `out` has zero weights there and its weight-gradient GEMM then yields its bias gradient in
 * column 31 (same trick as srb200_layernorm_fwd's ones_channel).                                                   */
#define SRB200_ATTN_ONES 1
/* out_alpha (optional, may be NULL; window 8 / even head count only, SRB200_EINVAL otherwise): fp32 [B], every output
 * row of sample b is multiplied by out_alpha[b] -- the per-sample DropPath factor of the attention branch
 * (swinir_arch.py:14-26,318), carried by `out` so that proj's weight-gradient GEMM needs no row-scaled copy of dY;
 * the SRB200_ATTN_ONES lane then holds out_alpha[b].                                                                  */
int srb200_window_attention_fwd(const void* qkv_bf16, const float* rpb_table, void* out_bf16, float* stats,
                                int B, int H, int W, int num_heads, int Cp, int window_size, int shift,
                                float scale, int flags, const float* out_alpha, srb200_stream_t stream);
int srb200_window_attention_bwd(const void* qkv_bf16, const void* gout_bf16, const float* rpb_table,
                                const float* stats, void* gqkv_bf16, float* g_rpb_table, float* workspace,
                                int B, int H, int W, int num_heads, int Cp, int window_size, int shift,
                                float scale, srb200_stream_t stream);

/* out = g * act'(y) for ReLU (slope 0) / LeakyReLU, y = forward output (bf16, n % 8 == 0). */
int srb200_act_bwd(const void* g_bf16, const void* y_bf16, void* out_bf16, int64_t n, float slope,
                   srb200_stream_t stream);

/* Programmatic dependent launch for the GEMM kernels (srb200_tapgemm, srb200_wgrad) launched from now on: their
 * prologue (barrier init, TMEM allocation, tensor-map fetch) overlaps the tail of the previous kernel of the stream.
 * Process-wide switch, returns the previous value; meant to bracket a CUDA-graph capture (the attribute is baked into
 * the captured launches).  The library's other kernels (channel attention, gradient finalize, LayerNorm, window
 * attention, column sums) execute griddepcontrol.launch_dependents on entry, so a GEMM that follows one of them
 * starts its prologue early as well.  Off by default for eager launches (nothing to gain when the host paces the
 * stream); the archs turn it on for their graph captures.  SRB_PDL=0|1 in the environment overrides.             */
int srb200_set_pdl(int on);

/* debug only: device buffer of 3*64 uint64 that receives CTA 0's per-warp-role clock64 timeline of the
 * following srb200_tapgemm launches (NULL switches it off). */
int srb200_debug_set_trace(void* dev_buf);
int srb200_debug_set_wgrad_trace(void* dev_buf); /* same for srb200_wgrad: 64 uint64 */
/* same for the tcgen05 window-attention kernels (window 8): [4 roles][64 stages][8 slots] uint64 */
int srb200_debug_set_attn_trace(void* dev_buf);

/* ------------------------------------------------------------------ image entry / exit (the callers' side)
 * One crop of a batched patch extraction: basicsr/data/transforms.py:28-96 (paired_random_crop), :166-225 (augment:
 * hflip, vflip, rot90 = transpose, in that order) and utils/img_util.py:9-37 (img2tensor: BGR->RGB, HWC->CHW, float)
 * in one pass over uint8 HWC images that are already in device memory.                                          */
typedef struct srb200_patch_item {
  const void* src;   /* uint8 HWC image, device memory                          */
  int64_t pitch;     /* bytes per image row                                     */
  int32_t top, left; /* crop origin (pixels)                                    */
  int32_t flags;     /* bit 0 hflip, bit 1 vflip, bit 2 rot90 (transpose)       */
  int32_t reserved;
} srb200_patch_item;
/* out fp32 [n, C, oh, ow] = augmented crop / div (255 for uint8 images: bit-equal to the reference's float32 img / 255.),
 * (oh, ow) = rot90 ? (pw, ph) : (ph, pw); all n items share ph x pw.
 * bgr2rgb swaps channels 0 and 2 of 3-channel images.  items_dev: device array of n srb200_patch_item.            */
int srb200_patch_from_u8(const void* items_dev, int n, int C, int ph, int pw, int bgr2rgb, float div, float* out,
                         srb200_stream_t stream);
/* acc[c, y0+y, x0+x] += w(y, x) * sr_tile[c, y, x]: one super-resolved tile into the fp32 scene accumulator
 * [C, acc_h, acc_w] with separable linear-ramp weights over its overlaps with the neighbouring tiles (ov_* pixels,
 * 0 = image border; the two tiles' ramps sum to 1; the outer `guard` pixels of a tile carry no weight).  Pixels
 * outside the accumulator are clipped (a rank's band of a larger scene).  Tiles that overlap each other must be
 * ordered on the stream (the adds are plain read-modify-writes).                                                 */
int srb200_tile_blend_add(const float* sr_tile, float* acc, int C, int th, int tw, int acc_h, int acc_w, int y0,
                          int x0, int ov_top, int ov_bottom, int ov_left, int ov_right, int guard,
                          srb200_stream_t stream);
/* tensor2img (basicsr/utils/img_util.py:40-96): fp32 CHW -> uint8 HWC, clamp to [lo, hi], normalise, x255, round half
 * to even (numpy .round()), optional RGB->BGR.                                                                    */
int srb200_tensor2img_u8(const float* src_chw, void* dst_hwc_u8, int C, int H, int W, float lo, float hi, int rgb2bgr,
                         srb200_stream_t stream);

/* fp32 mode (north_star: "1e-4 in the TF32/fp32 mode"): x fp32 [rows, C] -> bf16 [rows, 3*C] = [hi | lo | hi] with
 * hi = bf16(x), lo = bf16(x - hi).  Fed to srb200_tapgemm with Cin = 3*C against weights packed as
 * [w_hi ; w_hi ; w_lo] along K, the fp32 accumulator receives x_hi w_hi + x_lo w_hi + x_hi w_lo: fp32-class results
 * (2^-16 relative per product) on the bf16 tensor-core kernel.  C % 8 == 0.                                        */
int srb200_split3_bf16(const float* x_f32, void* out_bf16, int64_t rows, int C, srb200_stream_t stream);

/* fp32 mode of the SwinIR token path (evaluation only, no backward): the Linear layers run as srb200_tapgemm over
 * split operands (above); the two token kernels between them compute in plain fp32 and can emit their result already
 * split, so no separate srb200_split3_bf16 pass is needed in front of the next GEMM.
 * srb200_layernorm_f32: nn.LayerNorm(C, eps) (swinir_arch.py:240,251,602,886) over the C real channels of fp32 rows
 *   padded to Cp (pads are written as 0).  y_f32 [T, Cp] and / or y_split_bf16 [T, 3*Cp] = [hi | lo | hi]; either may
 *   be NULL, not both.  Cp % 8 == 0, Cp <= 1024.
 * srb200_window_attention_f32: WindowAttention.forward minus the two Linears (swinir_arch.py:151-172) with roll /
 *   window_partition / window_reverse (:293-316) and the analytic 0 / -100 mask (:262-281), same tensor layout as
 *   srb200_window_attention_fwd but fp32: qkv [B,H,W,3*Cp], channel = which*Cp + head*32 + d (head_dim <= 32, zero
 *   padded); out_f32 [B,H,W,Cp] and / or out_split_bf16 [B,H,W,3*Cp]; rpb_table fp32 [(2*ws-1)^2, num_heads].
 *   window_size 2..8, any head count with num_heads*32 <= Cp.                                                        */
int srb200_layernorm_f32(const float* x_f32, const float* gamma, const float* beta, float* y_f32, void* y_split_bf16,
                         int64_t T, int C, int Cp, float eps, srb200_stream_t stream);
int srb200_window_attention_f32(const float* qkv_f32, const float* rpb_table, float* out_f32, void* out_split_bf16,
                                int B, int H, int W, int num_heads, int Cp, int window_size, int shift, float scale,
                                srb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SRB200_H_ */
