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

// ------------------------------------------------------------------ fused channel attention (one launch each way)
// forward:  s = sigmoid(W2 relu(W1 p + b1) + b2) for ALL images in every CTA's shared memory (B*C*Cr MACs: nothing),
//           then y = x + res_scale * t * s[b,c].  CTA 0 also writes z, s for the backward.
__global__ void __launch_bounds__(256) ca_forward_fused_kernel(
    const uint4* __restrict__ t, const uint4* __restrict__ x, const float* __restrict__ x32,
    const float* __restrict__ p, const float* __restrict__ w1, const float* __restrict__ b1,
    const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ z_out,
    float* __restrict__ s_out, uint4* __restrict__ y, float* __restrict__ y32, size_t nvec, int B, int HW, int C,
    int Cr, float res_scale) {
  extern __shared__ float sm[];  // s[B*C], z[B*Cr]
  pdl_trigger();  // the next (programmatically serialized) GEMM may start its prologue now
  float* ss = sm;
  float* sz = sm + static_cast<size_t>(B) * C;
  for (int i = threadIdx.x; i < B * Cr; i += blockDim.x) {
    const int b = i / Cr, j = i - b * Cr;
    float acc = __ldg(b1 + j);
    for (int c = 0; c < C; ++c) acc += __ldg(w1 + static_cast<size_t>(j) * C + c) * __ldg(p + static_cast<size_t>(b) * C + c);
    sz[i] = fmaxf(acc, 0.0f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
    const int b = i / C, c = i - b * C;
    float acc = __ldg(b2 + c);
    for (int j = 0; j < Cr; ++j) acc += __ldg(w2 + static_cast<size_t>(c) * Cr + j) * sz[b * Cr + j];
    ss[i] = 1.0f / (1.0f + __expf(-acc));
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < B * Cr; i += blockDim.x) z_out[i] = sz[i];
    for (int i = threadIdx.x; i < B * C; i += blockDim.x) s_out[i] = ss[i];
  }
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
    const float* sv = ss + b * C + g * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = vx[e] + res_scale * vt[e] * sv[e];
    y[idx] = pack8(o);
    if (y32 != nullptr) {
      *reinterpret_cast<float4*>(y32 + idx * 8) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(y32 + idx * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// backward, ONE persistent COOPERATIVE launch (the driver guarantees all CTAs co-resident), phases joined by a
// device-wide arrive barrier instead of three kernel boundaries.  Every CTA owns one contiguous slab of 16-byte vectors in phases 1 and 3:
//   1. gs[b,c] += res_scale * sum_hw g*t   (registers -> warp shuffle -> smem -> global atomics), then arrive
//   2. after ALL CTAs arrived: every CTA redoes the tiny FC backward in shared memory (B*C*Cr MACs) to get
//      gp[b,c] for the <= 2 images its slab touches; CTA 0 also writes gW1, gb1, gW2, gb2
//   3. gt = res_scale * g * s[b,c] + gp[b,c]/HW, plus its column sums (conv2's bias gradient)
// sync[0] = arrival counter, zeroed by the caller.
constexpr int kCaBatch = 4;  // independent 16-byte load pairs in flight per thread

__device__ __forceinline__ void group_reduce8(float (&v)[8], int groups) {
  // lanes l, l+groups, l+2*groups, ... hold the same channel group (groups is a power of two)
  for (int off = groups; off < 32; off <<= 1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += __shfl_xor_sync(0xffffffffu, v[e], off);
  }
}

__global__ void __launch_bounds__(256, 4) ca_backward_fused_kernel(
    const uint4* __restrict__ g, const uint4* __restrict__ t, const float* __restrict__ s,
    const float* __restrict__ z, const float* __restrict__ p, const float* __restrict__ w1,
    const float* __restrict__ w2, float* __restrict__ gs, float* __restrict__ gw1, float* __restrict__ gb1,
    float* __restrict__ gw2, float* __restrict__ gb2, uint4* __restrict__ gt, float* __restrict__ colsum,
    unsigned int* __restrict__ sync, size_t nvec, size_t slab, int B, int HW, int C, int Cr, float res_scale) {
  extern __shared__ float sm[];  // ga2[B*C] (phase 1: partial sums), ga1[B*Cr], gp2[2*C], cs[C]
  pdl_trigger();  // the next (programmatically serialized) GEMM may start its prologue now
  float* ga2 = sm;
  float* ga1 = sm + static_cast<size_t>(B) * C;
  float* gp2 = ga1 + static_cast<size_t>(B) * Cr;
  float* s_cs = gp2 + 2 * C;
  const int groups = C / 8;
  const int grp = threadIdx.x % groups;  // blockDim and slab are multiples of groups: fixed per thread
  const bool lead = (threadIdx.x & 31) < groups;
  const size_t per_img = static_cast<size_t>(groups) * HW;
  const size_t lo = blockIdx.x * slab;
  const size_t hi = (lo + slab < nvec) ? lo + slab : nvec;
  const int b_first = (lo < nvec) ? static_cast<int>(lo / per_img) : 0;
  // ---- phase 1
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) sm[i] = 0.0f;
  __syncthreads();
  {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int cur_b = b_first;
    auto flush = [&]() {  // (warp-uniform: every lane of a warp sees the same image index)
      group_reduce8(acc, groups);
      if (lead) {
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&ga2[cur_b * C + grp * 8 + e], acc[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    };
    for (size_t base = lo + threadIdx.x; base < hi; base += static_cast<size_t>(kCaBatch) * blockDim.x) {
      uint4 rg[kCaBatch], rt[kCaBatch];
#pragma unroll
      for (int u = 0; u < kCaBatch; ++u) {
        const size_t idx = base + static_cast<size_t>(u) * blockDim.x;
        if (idx < hi) {
          rg[u] = __ldg(g + idx);
          rt[u] = __ldg(t + idx);
        }
      }
#pragma unroll
      for (int u = 0; u < kCaBatch; ++u) {
        // a warp's 32 consecutive vectors never straddle two images (per_img % 32 == 0 is checked on the host)
        const size_t idx = base + static_cast<size_t>(u) * blockDim.x;
        const size_t idx0 = idx - (threadIdx.x & 31);
        if (idx0 < hi) {
          const int bb = static_cast<int>(idx0 / per_img);
          if (bb != cur_b) {
            flush();
            cur_b = bb;
          }
          if (idx < hi) {
            float vg[8], vt[8];
            unpack8(rg[u], vg);
            unpack8(rt[u], vt);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += vg[e] * vt[e];
          }
        }
      }
    }
    flush();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < B * C; i += blockDim.x)
    if (sm[i] != 0.0f) atomicAdd(gs + i, sm[i] * res_scale);
  __threadfence();
  __syncthreads();
  // ---- device-wide barrier
  if (threadIdx.x == 0) {
    atomicAdd(&sync[0], 1u);
    unsigned int spins = 0;
    while (*reinterpret_cast<volatile unsigned int*>(&sync[0]) < gridDim.x) {
      __nanosleep(32);
      if (++spins > (1u << 25)) __trap();  // a protocol bug traps instead of hanging the GPU
    }
    __threadfence();
  }
  __syncthreads();
  // ---- phase 2 (every CTA, redundantly; only CTA 0 writes the parameter gradients)
  for (int i = threadIdx.x; i < B * C; i += blockDim.x) {
    const float sv = __ldg(s + i);
    ga2[i] = __ldcg(gs + i) * sv * (1.0f - sv);  // (L2: written by other CTAs' atomics)
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) s_cs[i] = 0.0f;
  __syncthreads();
  for (int i = threadIdx.x; i < B * Cr; i += blockDim.x) {
    const int b = i / Cr, j = i - b * Cr;
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) acc += __ldg(w2 + static_cast<size_t>(c) * Cr + j) * ga2[b * C + c];
    ga1[i] = __ldg(z + i) > 0.0f ? acc : 0.0f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {  // gp for the (at most two) images of this slab
    const int which = i / C, c = i - which * C;
    const int b = b_first + which;
    float acc = 0.0f;
    if (b < B)
      for (int j = 0; j < Cr; ++j) acc += __ldg(w1 + static_cast<size_t>(j) * C + c) * ga1[b * Cr + j];
    gp2[i] = acc / static_cast<float>(HW);
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {  // gW2 [C][Cr], gW1 [Cr][C]
      const int c = i / Cr, j = i - c * Cr;
      float a2 = 0.0f, a1 = 0.0f;
      for (int b = 0; b < B; ++b) {
        a2 += ga2[b * C + c] * __ldg(z + b * Cr + j);
        a1 += ga1[b * Cr + j] * __ldg(p + b * C + c);
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
  }
  __syncthreads();
  // ---- phase 3 (g comes back from L2)
  float cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t base = lo + threadIdx.x; base < hi; base += static_cast<size_t>(kCaBatch) * blockDim.x) {
    uint4 rg[kCaBatch];
#pragma unroll
    for (int u = 0; u < kCaBatch; ++u) {
      const size_t idx = base + static_cast<size_t>(u) * blockDim.x;
      if (idx < hi) rg[u] = __ldg(g + idx);
    }
#pragma unroll
    for (int u = 0; u < kCaBatch; ++u) {
      const size_t idx = base + static_cast<size_t>(u) * blockDim.x;
      if (idx < hi) {
        const int b = static_cast<int>(idx / per_img);
        float vg[8], o[8];
        unpack8(rg[u], vg);
        const float* sp = s + static_cast<size_t>(b) * C + grp * 8;
        const float* pp = gp2 + (b - b_first) * C + grp * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[e] = res_scale * vg[e] * __ldg(sp + e) + pp[e];
          cs[e] += o[e];
        }
        gt[idx] = pack8(o);
      }
    }
  }
  group_reduce8(cs, groups);
  if (lead) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(&s_cs[grp * 8 + e], cs[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(colsum + i, s_cs[i]);
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
  static PerDeviceOnce configured;
  if (smem > 48 * 1024 && configured.ensure(ca_fc_bwd_kernel, 200 * 1024) != SRB200_OK) return SRB200_ELAUNCH;
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

extern "C" int srb200_ca_forward(const void* t_bf16, const void* x_bf16, const float* x_f32, const float* p,
                                 const float* w1, const float* b1, const float* w2, const float* b2, float* z,
                                 float* s, void* y_bf16, float* y_f32, int B, int HW, int C, int Cr,
                                 float res_scale, srb200_stream_t stream) {
  if (!t_bf16 || (!x_bf16 && !x_f32) || !p || !w1 || !b1 || !w2 || !b2 || !z || !s || !y_bf16) return SRB200_EINVAL;
  if (B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0 || Cr <= 0) return SRB200_EINVAL;
  const size_t smem = (static_cast<size_t>(B) * C + static_cast<size_t>(B) * Cr) * sizeof(float);
  if (smem > 48 * 1024) return SRB200_EINVAL;
  const size_t nvec = static_cast<size_t>(B) * HW * (C / 8);
  int grid = grid1d(nvec, 256);
  if (grid > 4 * num_sms()) grid = 4 * num_sms();  // every CTA repeats the FC: keep them few and fat
  ca_forward_fused_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(t_bf16), static_cast<const uint4*>(x_bf16), x_f32, p, w1, b1, w2, b2, z, s,
      static_cast<uint4*>(y_bf16), y_f32, nvec, B, HW, C, Cr, res_scale);
  return launch_status();
}

extern "C" int srb200_ca_backward(const void* g_bf16, const void* t_bf16, const float* s, const float* z,
                                  const float* p, const float* w1, const float* w2, float* gs, float* gw1, float* gb1,
                                  float* gw2, float* gb2, void* gt_bf16, float* colsum, unsigned int* sync, int B,
                                  int HW, int C, int Cr, float res_scale, srb200_stream_t stream) {
  if (!g_bf16 || !t_bf16 || !s || !z || !p || !w1 || !w2 || !gs || !gw1 || !gb1 || !gw2 || !gb2 || !gt_bf16 ||
      !colsum || !sync)
    return SRB200_EINVAL;
  const int groups = C / 8;
  if (B <= 0 || HW <= 0 || C <= 0 || C % 8 != 0 || Cr <= 0) return SRB200_EINVAL;
  if ((groups & (groups - 1)) != 0 || groups > 256) return SRB200_EINVAL;  // channel groups: a power of two
  const size_t smem = (static_cast<size_t>(B) * C + static_cast<size_t>(B) * Cr + 3 * static_cast<size_t>(C)) *
                      sizeof(float);
  if (smem > 48 * 1024) return SRB200_EINVAL;
  const size_t nvec = static_cast<size_t>(B) * HW * groups;
  const size_t per_img = static_cast<size_t>(groups) * HW;
  // all CTAs must be co-resident (device-wide barrier): use at most half of what the device can hold
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ca_backward_fused_kernel, 256, smem) != cudaSuccess || occ < 1)
    return SRB200_ELAUNCH;
  const int ctas_per_sm = occ >= 6 ? 3 : (occ >= 2 ? occ / 2 : 1);
  size_t grid = static_cast<size_t>(num_sms()) * ctas_per_sm;
  // slab: a multiple of the block size (so of `groups` and of a warp), no longer than one image (a slab then
  // touches at most two images), warps never straddling an image
  size_t slab = ((nvec + grid - 1) / grid + 255) / 256 * 256;
  if (per_img % 32 != 0 || slab > per_img) {
    // small or ragged feature maps: one CTA per image keeps every assumption trivially true
    if (B > static_cast<int>(grid)) return SRB200_EINVAL;
    slab = per_img;
  }
  grid = (nvec + slab - 1) / slab;
  // COOPERATIVE launch: the driver either places all CTAs at once or rejects the launch -- the device-wide arrive
  // barrier inside can then never spin on CTAs that were not scheduled (NCCL kernels of DDP, other streams or
  // early-resident PDL CTAs may hold SMs when this kernel starts).  Works under stream capture (cooperative kernel
  // nodes); the occupancy-derived grid above stays <= half of the co-resident capacity.
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, ca_backward_fused_kernel, static_cast<const uint4*>(g_bf16),
                         static_cast<const uint4*>(t_bf16), s, z, p, w1, w2, gs, gw1, gb1, gw2, gb2,
                         static_cast<uint4*>(gt_bf16), colsum, sync, nvec, slab, B, HW, C, Cr,
                         res_scale) != cudaSuccess) {
    cudaGetLastError();
    return SRB200_ELAUNCH;
  }
  return launch_status();
}
