// RCAN channel attention (rcan_arch.py:8-24) and the RCAB tail (rcan_arch.py:44-46) on NHWC bf16:
//   pool    p[b,c]  = mean_hw t[b,:,:,c]                     (warp-shuffle + smem reduction, 1 pass)
//   fc      s[b,c]  = sigmoid(W2 relu(W1 p + b1) + b2)       (one CTA per image, warp-shuffle dots)
//   apply   y       = x + res_scale * t * s[b,c]             (one fused pass: 2 reads, 1 write)
// and their backward.  All HBM-bound; 128-bit accesses; fp32 math.
#include <cuda_bf16.h>

#include "host_util.h"
#include "ptx.cuh"

namespace srb {

__device__ __forceinline__ void unpack8(const uint4& m, float (&v)[8]) {
  v[0] = bf16_lo(m.x);
  v[1] = bf16_hi(m.x);
  v[2] = bf16_lo(m.y);
  v[3] = bf16_hi(m.y);
  v[4] = bf16_lo(m.z);
  v[5] = bf16_hi(m.z);
  v[6] = bf16_lo(m.w);
  v[7] = bf16_hi(m.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                    pack_bf16x2(v[6], v[7]));
}

// out[b, c] += scale * sum_hw a[b,hw,c] * (b2 ? b2[b,hw,c] : 1).  grid = (slabs, B).
template <bool HAS_MUL>
__global__ void channel_reduce_kernel(const __nv_bfloat16* __restrict__ a,
                                      const __nv_bfloat16* __restrict__ m, float* __restrict__ out,
                                      int HW, int C, float scale, int rows_per_block) {
  extern __shared__ float s_sum[];  // [C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_sum[i] = 0.0f;
  __syncthreads();
  const int groups = C / 8;
  const int g = threadIdx.x % groups;
  const int rl = threadIdx.x / groups;
  const int lanes = blockDim.x / groups;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, HW);
  if (rl < lanes) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t base = static_cast<size_t>(b) * HW * C + g * 8;
    for (int row = r0 + rl; row < r1; row += lanes) {
      float va[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(a + base + static_cast<size_t>(row) * C)), va);
      if (HAS_MUL) {
        float vm[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(m + base + static_cast<size_t>(row) * C)), vm);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += va[e] * vm[e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += va[e];
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&s_sum[g * 8 + e], acc[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x)
    atomicAdd(out + static_cast<size_t>(b) * C + i, s_sum[i] * scale);
}

// one CTA per image: z = relu(W1 p + b1) [Cr], s = sigmoid(W2 z + b2) [C]
__global__ void ca_fc_kernel(const float* __restrict__ p, const float* __restrict__ w1,
                             const float* __restrict__ b1, const float* __restrict__ w2,
                             const float* __restrict__ b2, float* __restrict__ z_out,
                             float* __restrict__ s_out, int C, int Cr) {
  extern __shared__ float sm[];  // p[C], z[Cr]
  float* sp = sm;
  float* sz = sm + C;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sp[i] = p[static_cast<size_t>(b) * C + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int j = warp; j < Cr; j += nwarps) {
    float acc = 0.0f;
    for (int i = lane; i < C; i += 32) acc += w1[static_cast<size_t>(j) * C + i] * sp[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float zz = fmaxf(acc + b1[j], 0.0f);
      sz[j] = zz;
      z_out[static_cast<size_t>(b) * Cr + j] = zz;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = b2[c];
    for (int j = 0; j < Cr; ++j) acc += w2[static_cast<size_t>(c) * Cr + j] * sz[j];
    s_out[static_cast<size_t>(b) * C + c] = 1.0f / (1.0f + __expf(-acc));
  }
}

// y = x + res_scale * t * s[b,c]; the skip stream may be carried in fp32 (x32 in, y32 out) so that the
// 200 stacked res_scale=1 additions of RCAN do not accumulate bf16 rounding.
__global__ void ca_apply_kernel(const uint4* __restrict__ t, const uint4* __restrict__ x,
                                const float* __restrict__ x32, const float* __restrict__ s,
                                uint4* __restrict__ y, float* __restrict__ y32, size_t nvec, int HW,
                                int C, float res_scale) {
  const int groups = C / 8;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < nvec;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    const size_t b = idx / (static_cast<size_t>(groups) * HW);
    float vt[8], vx[8], o[8];
    unpack8(__ldg(t + idx), vt);
    if (x32 != nullptr) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x32 + idx * 8));
      const float4 c = __ldg(reinterpret_cast<const float4*>(x32 + idx * 8 + 4));
      vx[0] = a.x; vx[1] = a.y; vx[2] = a.z; vx[3] = a.w;
      vx[4] = c.x; vx[5] = c.y; vx[6] = c.z; vx[7] = c.w;
    } else {
      unpack8(__ldg(x + idx), vx);
    }
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(s + b * C + g * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(s + b * C + g * 8 + 4));
    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = vx[e] + res_scale * vt[e] * sv[e];
    y[idx] = pack8(o);
    if (y32 != nullptr) {
      *reinterpret_cast<float4*>(y32 + idx * 8) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(y32 + idx * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// backward of the tiny FC, all images in one CTA (parameter grads are sums over the batch):
//   ga2 = gs * s(1-s); gW2 += ga2 (x) z; gb2 += ga2; gz = W2^T ga2; ga1 = gz * (z>0);
//   gW1 += ga1 (x) p; gb1 += ga1; gp = W1^T ga1
__global__ void ca_fc_bwd_kernel(const float* __restrict__ gs, const float* __restrict__ s,
                                 const float* __restrict__ z, const float* __restrict__ p,
                                 const float* __restrict__ w1, const float* __restrict__ w2,
                                 float* __restrict__ gw1, float* __restrict__ gb1,
                                 float* __restrict__ gw2, float* __restrict__ gb2,
                                 float* __restrict__ gp, int B, int C, int Cr) {
  // Everything (a few KB) is staged in shared memory by ONE round of global loads; the five dependent phases
  // then run out of smem -- as separate global-load phases the kernel paid ~6 DRAM/L2 latencies (17 us).
  extern __shared__ float sm[];  // ga2[B*C], ga1[B*Cr], z[B*Cr], p[B*C], w1[Cr*C], w2[C*Cr]
  float* ga2 = sm;
  float* ga1 = ga2 + static_cast<size_t>(B) * C;
  float* sz = ga1 + static_cast<size_t>(B) * Cr;
  float* sp = sz + static_cast<size_t>(B) * Cr;
  float* sw1 = sp + static_cast<size_t>(B) * C;
  float* sw2 = sw1 + static_cast<size_t>(C) * Cr;
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
    const float sv = __ldg(s + i);
    ga2[i] = __ldg(gs + i) * sv * (1.0f - sv);
    sp[i] = __ldg(p + i);
  }
  for (int i = threadIdx.x; i < B * Cr; i += blockDim.x) sz[i] = __ldg(z + i);
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {
    sw1[i] = __ldg(w1 + i);
    sw2[i] = __ldg(w2 + i);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B * Cr; i += blockDim.x) {
    const int b = i / Cr, j = i - b * Cr;
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) acc += sw2[c * Cr + j] * ga2[b * C + c];
    ga1[i] = sz[i] > 0.0f ? acc : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {  // gW2 [C][Cr], gW1 [Cr][C]
    const int c = i / Cr, j = i - c * Cr;
    float a2 = 0.0f, a1 = 0.0f;
    for (int b = 0; b < B; ++b) {
      a2 += ga2[b * C + c] * sz[b * Cr + j];
      a1 += ga1[b * Cr + j] * sp[b * C + c];
    }
    gw2[static_cast<size_t>(c) * Cr + j] = a2;
    gw1[static_cast<size_t>(j) * C + c] = a1;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.0f;
    for (int b = 0; b < B; ++b) a += ga2[b * C + c];
    gb2[c] = a;
  }
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
    float a = 0.0f;
    for (int b = 0; b < B; ++b) a += ga1[b * Cr + j];
    gb1[j] = a;
  }
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
    const int b = i / C, c = i - b * C;
    float acc = 0.0f;
    for (int j = 0; j < Cr; ++j) acc += sw1[j * C + c] * ga1[b * Cr + j];
    gp[i] = acc;
  }
}

// gt = res_scale * g * s[b,c] + gp[b,c] / HW
__global__ void ca_apply_bwd_kernel(const uint4* __restrict__ g, const float* __restrict__ s,
                                    const float* __restrict__ gp, uint4* __restrict__ gt, size_t nvec,
                                    int HW, int C, float res_scale, float* __restrict__ colsum) {
  extern __shared__ float s_cs[];  // [C] column sums of gt (only when colsum != NULL)
  const int groups = C / 8;
  const float inv_hw = 1.0f / static_cast<float>(HW);
  if (colsum != nullptr) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_cs[i] = 0.0f;
    __syncthreads();
  }
  // the grid stride is a multiple of `groups`, so a thread stays on one 8-channel group
  float cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int my_group = -1;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < nvec;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int gg = static_cast<int>(idx % groups);
    const size_t b = idx / (static_cast<size_t>(groups) * HW);
    float vg[8], o[8];
    unpack8(__ldg(g + idx), vg);
    const float* sp = s + b * C + gg * 8;
    const float* pp = gp + b * C + gg * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = res_scale * vg[e] * __ldg(sp + e) + __ldg(pp + e) * inv_hw;
    gt[idx] = pack8(o);
    my_group = gg;
#pragma unroll
    for (int e = 0; e < 8; ++e) cs[e] += o[e];
  }
  if (colsum != nullptr) {
    if (my_group >= 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&s_cs[my_group * 8 + e], cs[e]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(colsum + i, s_cs[i]);
  }
}

static inline int grid1d(size_t work, int block) {
  size_t g = (work + block - 1) / block;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  return static_cast<int>(g > cap ? cap : (g == 0 ? 1 : g));
}

}  // namespace srb

using namespace srb;

static int channel_reduce(const void* a, const void* m, float* out, int B, int HW, int C,
                          float scale, cudaStream_t st) {
  if (!a || !out || B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0 || C / 8 > 256) return SRB200_EINVAL;
  int rpb = HW / ((num_sms() * 4 + B - 1) / B);
  if (rpb < 32) rpb = 32;
  if (rpb > 512) rpb = 512;
  dim3 grid((HW + rpb - 1) / rpb, B);
  const __nv_bfloat16* pa = static_cast<const __nv_bfloat16*>(a);
  const __nv_bfloat16* pm = static_cast<const __nv_bfloat16*>(m);
  if (m)
    channel_reduce_kernel<true><<<grid, 256, C * sizeof(float), st>>>(pa, pm, out, HW, C, scale, rpb);
  else
    channel_reduce_kernel<false><<<grid, 256, C * sizeof(float), st>>>(pa, pm, out, HW, C, scale, rpb);
  return launch_status();
}

extern "C" int srb200_channel_pool(const void* t_bf16, float* out, int B, int HW, int C,
                                   srb200_stream_t stream) {
  return channel_reduce(t_bf16, nullptr, out, B, HW, C, 1.0f / static_cast<float>(HW > 0 ? HW : 1),
                        static_cast<cudaStream_t>(stream));
}

extern "C" int srb200_channel_dot(const void* a_bf16, const void* b_bf16, float* out, int B, int HW,
                                  int C, float scale, srb200_stream_t stream) {
  if (!b_bf16) return SRB200_EINVAL;
  return channel_reduce(a_bf16, b_bf16, out, B, HW, C, scale, static_cast<cudaStream_t>(stream));
}

extern "C" int srb200_ca_fc(const float* p, const float* w1, const float* b1, const float* w2,
                            const float* b2, float* z, float* s, int B, int C, int Cr,
                            srb200_stream_t stream) {
  if (!p || !w1 || !b1 || !w2 || !b2 || !z || !s || B <= 0 || C <= 0 || Cr <= 0) return SRB200_EINVAL;
  ca_fc_kernel<<<B, 128, (C + Cr) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      p, w1, b1, w2, b2, z, s, C, Cr);
  return launch_status();
}

extern "C" int srb200_ca_apply(const void* t_bf16, const void* x_bf16, const float* x_f32,
                               const float* s, void* y_bf16, float* y_f32, int B, int HW, int C,
                               float res_scale, srb200_stream_t stream) {
  if (!t_bf16 || (!x_bf16 && !x_f32) || !s || !y_bf16 || B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0)
    return SRB200_EINVAL;
  const size_t nvec = static_cast<size_t>(B) * HW * (C / 8);
  ca_apply_kernel<<<grid1d(nvec, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(t_bf16), static_cast<const uint4*>(x_bf16), x_f32, s,
      static_cast<uint4*>(y_bf16), y_f32, nvec, HW, C, res_scale);
  return launch_status();
}

extern "C" int srb200_ca_fc_bwd(const float* gs, const float* s, const float* z, const float* p,
                                const float* w1, const float* w2, float* gw1, float* gb1, float* gw2,
                                float* gb2, float* gp, int B, int C, int Cr, srb200_stream_t stream) {
  if (!gs || !s || !z || !p || !w1 || !w2 || !gw1 || !gb1 || !gw2 || !gb2 || !gp) return SRB200_EINVAL;
  if (B <= 0 || C <= 0 || Cr <= 0) return SRB200_EINVAL;
  const size_t smem = (2 * static_cast<size_t>(B) * C + 2 * static_cast<size_t>(B) * Cr +
                       2 * static_cast<size_t>(C) * Cr) * sizeof(float);
  if (smem > 200 * 1024) return SRB200_EINVAL;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    if (cudaFuncSetAttribute(ca_fc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return SRB200_ELAUNCH;
    configured = 200 * 1024;
  }
  ca_fc_bwd_kernel<<<1, 256, smem, static_cast<cudaStream_t>(stream)>>>(gs, s, z, p, w1, w2, gw1, gb1,
                                                                        gw2, gb2, gp, B, C, Cr);
  return launch_status();
}

extern "C" int srb200_ca_apply_bwd(const void* g_bf16, const float* s, const float* gp, void* gt_bf16,
                                   int B, int HW, int C, float res_scale, float* colsum,
                                   srb200_stream_t stream) {
  if (!g_bf16 || !s || !gp || !gt_bf16 || B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0)
    return SRB200_EINVAL;
  if (colsum != nullptr && 256 % (C / 8) != 0) return SRB200_EINVAL;  // a thread must stay on one channel group
  const size_t nvec = static_cast<size_t>(B) * HW * (C / 8);
  // with column sums: ~2 blocks per SM (every block ends with C global atomics onto the same addresses)
  int grid = grid1d(nvec, 256);
  if (colsum != nullptr && grid > 2 * num_sms()) grid = 2 * num_sms();
  ca_apply_bwd_kernel<<<grid, 256, colsum ? C * sizeof(float) : 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(g_bf16), s, gp, static_cast<uint4*>(gt_bf16), nvec, HW, C, res_scale, colsum);
  return launch_status();
}
