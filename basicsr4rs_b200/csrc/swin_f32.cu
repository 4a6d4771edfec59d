// fp32 mode of the SwinIR token path (north_star: "max-abs 1e-4 in the TF32/fp32 mode"): LayerNorm and the fused
// (shifted-)window attention in plain fp32 arithmetic on fp32 NHWC activations.  The Linear layers and convs around
// them stay on the bf16 tcgen05 tap-GEMM with error-compensated operands (imageio.cu: split3), so both kernels can
// hand their result over already split into [hi | lo | hi] bf16 channels -- the GEMM's A operand -- instead of
// writing fp32 and paying a separate split pass.  Evaluation only (no backward).
//
// Reference: swinir_arch.py:144-175 (WindowAttention.forward), :262-281 (calculate_mask), :293-316 (roll /
// window_partition / window_reverse), nn.LayerNorm at :240,251,602,886.
#include <cuda_bf16.h>

#include "host_util.h"

namespace srb {

// 4 consecutive channels -> bf16 hi and lo halves, packed for one 8-byte store each
__device__ __forceinline__ void split4(const float (&v)[4], uint2& hi, uint2& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat16 hb = __float2bfloat16_rn(v[e]);
    const __nv_bfloat16 lb = __float2bfloat16_rn(v[e] - __bfloat162float(hb));
    h[e] = __bfloat16_as_ushort(hb);
    l[e] = __bfloat16_as_ushort(lb);
  }
  hi = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
  lo = make_uint2(l[0] | (l[1] << 16), l[2] | (l[3] << 16));
}

__device__ __forceinline__ void store_split4(__nv_bfloat16* row, int c, int Cp, const float (&v)[4]) {
  uint2 hi, lo;
  split4(v, hi, lo);
  *reinterpret_cast<uint2*>(row + c) = hi;
  *reinterpret_cast<uint2*>(row + Cp + c) = lo;
  *reinterpret_cast<uint2*>(row + 2 * Cp + c) = hi;
}

// ------------------------------------------------------------------ LayerNorm, one warp per token
// A lane owns the float4 groups lane, lane + 32, ... of its token's row (MAXV of them): the row is read once, both
// statistics passes run over registers (mean first, then the centred second moment: no E[x^2] - E[x]^2 cancellation).
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y32,
                                                            __nv_bfloat16* __restrict__ ysplit, int64_t T, int C, int Cp,
                                                            float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nv = Cp >> 2;
  const float inv_c = 1.0f / static_cast<float>(C);
  for (int64_t t = warp; t < T; t += nwarps) {
    const float4* row = reinterpret_cast<const float4*>(x + t * Cp);
    float v[MAXV][4];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int q = lane + 32 * i;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < nv) f = __ldg(row + q);
      const int c = 4 * q;
      v[i][0] = (c + 0 < C) ? f.x : 0.0f;
      v[i][1] = (c + 1 < C) ? f.y : 0.0f;
      v[i][2] = (c + 2 < C) ? f.z : 0.0f;
      v[i][3] = (c + 3 < C) ? f.w : 0.0f;
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = 4 * (lane + 32 * i);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = (c + e < C) ? v[i][e] - mean : 0.0f;
        v[i][e] = d;
        s2 = fmaf(d, d, s2);
      }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    const float rstd = 1.0f / sqrtf(s2 * inv_c + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int q = lane + 32 * i;
      if (q >= nv) continue;
      const int c = 4 * q;
      float y[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        y[e] = (c + e < C) ? fmaf(v[i][e] * rstd, __ldg(gamma + c + e), __ldg(beta + c + e)) : 0.0f;
      if (y32 != nullptr) *reinterpret_cast<float4*>(y32 + t * Cp + c) = make_float4(y[0], y[1], y[2], y[3]);
      if (ysplit != nullptr) store_split4(ysplit + t * 3 * Cp, c, Cp, y);
    }
  }
}

// ------------------------------------------------------------------ window attention, one CTA per (window, head)
// Thread i owns query token i of the window (ws^2 <= 64 tokens): q row and the output row live in registers, K / V of
// the head in shared memory (every thread reads the same key at the same time: broadcasts), the scores of a row in a
// [key][query] shared sheet between the max pass and the exp / PV pass.  The cyclic shift, window partition and
// reverse are address arithmetic; the SW-MSA mask is the analytic region test of calculate_mask.
struct AttnF32Params {
  const float* qkv;
  const float* table;
  float* out32;
  __nv_bfloat16* osplit;
  int B, H, W, heads, Ca, ws, shift;
  float scale;
};

__global__ void __launch_bounds__(64) window_attention_f32_kernel(const AttnF32Params p) {
  __shared__ float4 kS[64 * 8];
  __shared__ float4 vS[64 * 8];
  __shared__ float sS[64 * 64];
  __shared__ float tS[225];
  __shared__ int rS[64];
  const int i = threadIdx.x;
  const int ws = p.ws, n = ws * ws;
  const int head = blockIdx.x % p.heads;
  const int win = blockIdx.x / p.heads;
  const int wpr = p.W / ws;
  const int wy = win / wpr, wx = win - wy * wpr;
  const int b = blockIdx.y;
  const int span = 2 * ws - 1;

  for (int t = i; t < span * span; t += 64) tS[t] = __ldg(p.table + t * p.heads + head);

  const int iy = i / ws, ix = i - iy * ws;
  float q[32];
  size_t pix = 0;
  int region = 0;
  if (i < n) {
    const int ys = wy * ws + iy, xs = wx * ws + ix;  // position in the shifted image (swinir_arch.py:293-297)
    int y = ys + p.shift, x = xs + p.shift;          // roll(-shift): shifted[ys] = x[(ys + shift) % H]
    if (y >= p.H) y -= p.H;
    if (x >= p.W) x -= p.W;
    pix = (static_cast<size_t>(b) * p.H + y) * p.W + x;
    if (p.shift > 0) {  // calculate_mask (:262-281): slices (0,-ws), (-ws,-shift), (-shift,None) of the shifted image
      const int rh = ys < p.H - ws ? 0 : (ys < p.H - p.shift ? 1 : 2);
      const int rw = xs < p.W - ws ? 0 : (xs < p.W - p.shift ? 1 : 2);
      region = 3 * rh + rw;
    }
    rS[i] = region;
    const float4* src = reinterpret_cast<const float4*>(p.qkv + pix * (3 * static_cast<size_t>(p.Ca)) + head * 32);
    const int c4 = p.Ca >> 2;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float4 f = __ldg(src + e);
      q[4 * e + 0] = f.x * p.scale;  // q = q * self.scale (:155)
      q[4 * e + 1] = f.y * p.scale;
      q[4 * e + 2] = f.z * p.scale;
      q[4 * e + 3] = f.w * p.scale;
      kS[i * 8 + e] = __ldg(src + c4 + e);
      vS[i * 8 + e] = __ldg(src + 2 * c4 + e);
    }
  }
  __syncthreads();
  if (i >= n) return;

  // pass 1: s = q k^T + relative position bias (+ mask), row maximum
  float m = -INFINITY;
  {
    int j = 0;
    for (int jy = 0; jy < ws; ++jy) {
      const int trow = (iy - jy + ws - 1) * span + ix + ws - 1;
      for (int jx = 0; jx < ws; ++jx, ++j) {
        float s = 0.0f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 kk = kS[j * 8 + e];
          s = fmaf(q[4 * e + 0], kk.x, s);
          s = fmaf(q[4 * e + 1], kk.y, s);
          s = fmaf(q[4 * e + 2], kk.z, s);
          s = fmaf(q[4 * e + 3], kk.w, s);
        }
        s += tS[trow - jx];
        if (p.shift > 0 && rS[j] != region) s += -100.0f;
        sS[j * 64 + i] = s;
        m = fmaxf(m, s);
      }
    }
  }
  // pass 2: softmax numerators, their sum, P V
  float o[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) o[d] = 0.0f;
  float l = 0.0f;
  for (int j = 0; j < n; ++j) {
    const float pj = expf(sS[j * 64 + i] - m);
    l += pj;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float4 vv = vS[j * 8 + e];
      o[4 * e + 0] = fmaf(pj, vv.x, o[4 * e + 0]);
      o[4 * e + 1] = fmaf(pj, vv.y, o[4 * e + 1]);
      o[4 * e + 2] = fmaf(pj, vv.z, o[4 * e + 2]);
      o[4 * e + 3] = fmaf(pj, vv.w, o[4 * e + 3]);
    }
  }
  const float inv = 1.0f / l;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float y[4] = {o[4 * e + 0] * inv, o[4 * e + 1] * inv, o[4 * e + 2] * inv, o[4 * e + 3] * inv};
    const int c = head * 32 + 4 * e;
    if (p.out32 != nullptr)
      *reinterpret_cast<float4*>(p.out32 + pix * p.Ca + c) = make_float4(y[0], y[1], y[2], y[3]);
    if (p.osplit != nullptr) store_split4(p.osplit + pix * 3 * static_cast<size_t>(p.Ca), c, p.Ca, y);
  }
}

}  // namespace srb

extern "C" int srb200_layernorm_f32(const float* x_f32, const float* gamma, const float* beta, float* y_f32,
                                    void* y_split_bf16, int64_t T, int C, int Cp, float eps, srb200_stream_t stream) {
  if (!x_f32 || !gamma || !beta || (!y_f32 && !y_split_bf16) || T <= 0 || C <= 0 || Cp < C || Cp % 4 != 0 || Cp > 1024)
    return SRB200_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x_f32) | reinterpret_cast<uintptr_t>(y_f32) |
       reinterpret_cast<uintptr_t>(y_split_bf16)) & 15u)
    return SRB200_EINVAL;
  if (y_split_bf16 != nullptr && Cp % 8 != 0) return SRB200_EINVAL;  // the three 2*Cp-byte parts stay 16-byte aligned
  int64_t blocks = (T + 7) / 8;
  const int64_t cap = static_cast<int64_t>(srb::num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  auto* ys = static_cast<__nv_bfloat16*>(y_split_bf16);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int g = static_cast<int>(blocks);
  if (Cp <= 256)
    srb::layernorm_f32_kernel<2><<<g, 256, 0, st>>>(x_f32, gamma, beta, y_f32, ys, T, C, Cp, eps);
  else if (Cp <= 512)
    srb::layernorm_f32_kernel<4><<<g, 256, 0, st>>>(x_f32, gamma, beta, y_f32, ys, T, C, Cp, eps);
  else
    srb::layernorm_f32_kernel<8><<<g, 256, 0, st>>>(x_f32, gamma, beta, y_f32, ys, T, C, Cp, eps);
  return srb::launch_status();
}

extern "C" int srb200_window_attention_f32(const float* qkv_f32, const float* rpb_table, float* out_f32,
                                           void* out_split_bf16, int B, int H, int W, int num_heads, int Cp,
                                           int window_size, int shift, float scale, srb200_stream_t stream) {
  if (!qkv_f32 || !rpb_table || (!out_f32 && !out_split_bf16) || B <= 0 || H <= 0 || W <= 0 || num_heads <= 0)
    return SRB200_EINVAL;
  if (window_size < 2 || window_size > 8 || H % window_size || W % window_size || shift < 0 || shift >= window_size)
    return SRB200_EINVAL;
  if (Cp < num_heads * 32 || Cp % 8 != 0) return SRB200_EINVAL;
  if ((reinterpret_cast<uintptr_t>(qkv_f32) | reinterpret_cast<uintptr_t>(out_f32) |
       reinterpret_cast<uintptr_t>(out_split_bf16)) & 15u)
    return SRB200_EINVAL;
  const long long wins = static_cast<long long>(H / window_size) * (W / window_size);
  if (wins * num_heads > 0x7fffffffLL || B > 65535) return SRB200_EINVAL;
  srb::AttnF32Params p;
  p.qkv = qkv_f32;
  p.table = rpb_table;
  p.out32 = out_f32;
  p.osplit = static_cast<__nv_bfloat16*>(out_split_bf16);
  p.B = B;
  p.H = H;
  p.W = W;
  p.heads = num_heads;
  p.Ca = Cp;
  p.ws = window_size;
  p.shift = shift;
  p.scale = scale;
  const dim3 grid(static_cast<unsigned>(wins * num_heads), static_cast<unsigned>(B));
  srb::window_attention_f32_kernel<<<grid, 64, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return srb::launch_status();
}
