// SwinIR token-wise kernels on NHWC bf16 (tokens == pixels, C padded to a multiple of 64 with zeros):
//   LayerNorm forward / backward over the C real channels (swinir_arch.py:240,251,288,321,602,886),
//   warp-per-token, fp32 statistics, 128-bit accesses, gamma/beta gradients reduced warp -> CTA -> atomics;
//   per-sample row scaling (DropPath backward, swinir_arch.py:14-26).
// HBM-bound: fwd moves 2*T*C*2 B, bwd 3*T*C*2 B (+ residual-gradient add fused).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace srb {

constexpr int LN_MAX_VEC = 4;  // up to 4 x (32 lanes x 8 ch) = 1024 padded channels

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void ln_unpack(const uint4& m, float* v) {
  v[0] = bf16_lo(m.x);
  v[1] = bf16_hi(m.x);
  v[2] = bf16_lo(m.y);
  v[3] = bf16_hi(m.y);
  v[4] = bf16_lo(m.z);
  v[5] = bf16_hi(m.z);
  v[6] = bf16_lo(m.w);
  v[7] = bf16_hi(m.w);
}
__device__ __forceinline__ uint4 ln_pack(const float* v) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                    pack_bf16x2(v[6], v[7]));
}

template <int NVEC>
__global__ void layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, long long T, int C, int Cp,
                                     float eps, int ones_ch) {
  pdl_trigger();  // a following programmatically serialized GEMM may start its prologue
  constexpr int TOK = (NVEC == 1) ? 2 : 1;  // tokens per warp iteration (independent loads in flight)
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int nv = Cp / 8;
  // per-lane affine parameters live in registers for the whole kernel
  float gm[NVEC][8], bt[NVEC][8];
#pragma unroll
  for (int i = 0; i < NVEC; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (lane + 32 * i) * 8 + e;
      gm[i][e] = (c < C) ? __ldg(gamma + c) : 0.0f;
      bt[i][e] = (c < C) ? __ldg(beta + c) : 0.0f;
    }
  const float inv_c = 1.0f / static_cast<float>(C);
  for (long long t0 = warp_global * TOK; t0 < T; t0 += nwarps * TOK) {
    float v[TOK][NVEC][8];
#pragma unroll
    for (int k = 0; k < TOK; ++k)
#pragma unroll
      for (int i = 0; i < NVEC; ++i) {
        const int vec = lane + 32 * i;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[k][i][e] = 0.0f;
        if (vec < nv && t0 + k < T) ln_unpack(__ldg(reinterpret_cast<const uint4*>(x + (t0 + k) * Cp) + vec), v[k][i]);
      }
#pragma unroll
    for (int k = 0; k < TOK; ++k) {
      if (t0 + k >= T) break;
      float sum = 0.0f;
#pragma unroll
      for (int i = 0; i < NVEC; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) sum += v[k][i][e];  // pad channels are zero by the layout invariant
      const float mean = warp_sum(sum) * inv_c;
      float sq = 0.0f;
#pragma unroll
      for (int i = 0; i < NVEC; ++i)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float d = v[k][i][e] - mean;
          sq += ((lane + 32 * i) * 8 + e < C) ? d * d : 0.0f;
        }
      const float rstd = rsqrtf(warp_sum(sq) * inv_c + eps);
#pragma unroll
      for (int i = 0; i < NVEC; ++i) {
        const int vec = lane + 32 * i;
        if (vec < nv) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = (v[k][i][e] - mean) * rstd * gm[i][e] + bt[i][e];
          // optional constant-one PAD channel: the weight-gradient GEMM of the Linear that consumes this tensor then
          // delivers sum_tokens dY * 1 = that layer's bias gradient in column ones_ch for free
          if (ones_ch >= 0 && (ones_ch >> 3) == vec) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (e == (ones_ch & 7)) o[e] = 1.0f;
          }
          *(reinterpret_cast<uint4*>(y + (t0 + k) * Cp) + vec) = ln_pack(o);
        }
      }
      if (lane == 0) {
        if (mean_out) mean_out[t0 + k] = mean;
        if (rstd_out) rstd_out[t0 + k] = rstd;
      }
    }
  }
}


// Forward, 8 lanes per token: a warp handles 4 tokens per iteration, every lane NV 16-byte vectors of its token
// (vector index (lane & 7) + 8 i), so all 32 lanes move data whatever Cp is (the warp-per-token kernel above keeps
// only Cp/8 of them busy: 24 of 32 at Cp = 192) and a reduction is 3 shuffles shared by four tokens instead of 5 per
// token.  gamma / beta live in shared memory (read as float4 when a row is written).
template <int NV>
__global__ void __launch_bounds__(256, 2) layernorm_fwd8_kernel(const __nv_bfloat16* __restrict__ x,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta,
                                                                __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                                                                float* __restrict__ rstd_out, long long T, int C, int Cp,
                                                                float eps, int ones_ch) {
  pdl_trigger();
  const int lane = threadIdx.x & 31, sub = lane & 7, slot = lane >> 3;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int nv = Cp / 8;
  const float inv_c = 1.0f / static_cast<float>(C);
  const float npad = static_cast<float>(NV * 64 - C);  // zero channels this token's 8 lanes sum over (incl. vec >= nv)
  // A lane always works on the same channels (vectors sub + 8 i): its gamma / beta live in registers.  (Read from
  // shared memory as four 16-byte loads per vector and token they made the kernel wait on the shared-memory
  // scoreboard -- ncu: 5.0 warps stalled on short_scoreboard per issue, 0.42 of the HBM roofline on 1M tokens.)
  float gm[NV][8], bt[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = (sub + 8 * i) * 8 + e;
      gm[i][e] = c < C ? __ldg(gamma + c) : 0.0f;
      bt[i][e] = c < C ? __ldg(beta + c) : 0.0f;
      if (c == ones_ch) bt[i][e] = 1.0f;  // constant-one PAD channel (gamma = 0 there): see the warp-per-token kernel
    }
  auto load = [&](long long t, uint4 (&raw)[NV]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vec = sub + 8 * i;
      raw[i] = (t < T && vec < nv) ? __ldg(reinterpret_cast<const uint4*>(x + t * Cp) + vec) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  // software pipeline: the loads of the next token group are in flight while this one is reduced and written (the
  // plain loop was latency-bound: ncu showed 33 % warp occupancy and neither DRAM nor the issue slots saturated)
  uint4 nxt[NV];
  long long t = warp_global * 4 + slot;
  load(t, nxt);
  for (; t < T + slot; t += nwarps * 4) {  // (whole warps stay in the shuffles)
    const bool tv = t < T;
    uint4 raw[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) raw[i] = nxt[i];
    load(t + nwarps * 4, nxt);
    float v[NV][8];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      ln_unpack(raw[i], v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[i][e];  // pad channels are zero by the layout invariant
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    const float mean = sum * inv_c;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        sq = fmaf(d, d, sq);
      }
    sq += __shfl_xor_sync(0xffffffffu, sq, 1);
    sq += __shfl_xor_sync(0xffffffffu, sq, 2);
    sq += __shfl_xor_sync(0xffffffffu, sq, 4);
    // every zero pad channel added mean^2 to the sum: take them out instead of testing each element
    const float rstd = rsqrtf(fmaxf(sq - npad * mean * mean, 0.0f) * inv_c + eps);
    if (tv) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vec = sub + 8 * i;
        if (vec < nv) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf((v[i][e] - mean) * rstd, gm[i][e], bt[i][e]);  // (gamma = 0 on pads)
          *(reinterpret_cast<uint4*>(y + t * Cp) + vec) = ln_pack(o);
        }
      }
      if (sub == 0) {
        if (mean_out) mean_out[t] = mean;
        if (rstd_out) rstd_out[t] = rstd;
      }
    }
  }
}

// gx = LN'(gy) (+ gres);  ggamma[c] += sum_t gy*xhat;  gbeta[c] += sum_t gy
template <int NVEC, int TOK>
__global__ void __launch_bounds__(256, NVEC == 1 ? 3 : 1) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ gy,
                                     const __nv_bfloat16* __restrict__ x,
                                     const float* __restrict__ mean_in,
                                     const float* __restrict__ rstd_in,
                                     const float* __restrict__ gamma,
                                     const __nv_bfloat16* __restrict__ gres,
                                     __nv_bfloat16* __restrict__ gx, float* __restrict__ ggamma,
                                     float* __restrict__ gbeta, long long T, int C, int Cp) {
  pdl_trigger();  // a following programmatically serialized GEMM may start its prologue
  extern __shared__ float s_red[];  // [2][Cp]
  for (int i = threadIdx.x; i < 2 * Cp; i += blockDim.x) s_red[i] = 0.0f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int nv = Cp / 8;
  // lm = 1 on real channels, 0 on pads.  gy is zero on pad channels (it is the data gradient of a Linear whose packed
  // weights are zero there) and so is gamma, so the sums need no per-element test; only the stored gx is masked (ncu:
  // this kernel is issue-bound, 63-66 % issue-slot use, and the three selects per element were a fifth of it)
  float ag[NVEC][8], ab[NVEC][8], gm[NVEC][8];
  uint4 lm[NVEC];  // bit mask of the packed bf16 output vector: all ones on real channels, zero on pads
#pragma unroll
  for (int i = 0; i < NVEC; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ag[i][e] = 0.0f;
      ab[i][e] = 0.0f;
      const int c = (lane + 32 * i) * 8 + e;
      gm[i][e] = (c < C) ? __ldg(gamma + c) : 0.0f;
    }
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int c0 = (lane + 32 * i) * 8;
    auto half_mask = [&](int c) { return (c < C ? 0x0000FFFFu : 0u) | (c + 1 < C ? 0xFFFF0000u : 0u); };
    lm[i] = make_uint4(half_mask(c0), half_mask(c0 + 2), half_mask(c0 + 4), half_mask(c0 + 6));
  }
  // TOK tokens per warp iteration: all their 16-byte loads (gy, x, gres) are issued before the first reduction, so
  // a warp keeps 3 * TOK * NVEC requests in flight instead of 2 (the one-token loop ran at 1.9 TB/s)
  for (long long t0 = warp_global * TOK; t0 < T; t0 += nwarps * TOK) {
    uint4 rg[TOK][NVEC], rx[TOK][NVEC], rr[TOK][NVEC];
    float mean[TOK], rstd[TOK];
#pragma unroll
    for (int k = 0; k < TOK; ++k) {
      const long long t = t0 + k;
      const bool tv = t < T;
      mean[k] = tv ? __ldg(mean_in + t) : 0.0f;
      rstd[k] = tv ? __ldg(rstd_in + t) : 0.0f;
#pragma unroll
      for (int i = 0; i < NVEC; ++i) {
        const int vec = lane + 32 * i;
        rg[k][i] = rx[k][i] = rr[k][i] = make_uint4(0u, 0u, 0u, 0u);
        if (tv && vec < nv) {
          rg[k][i] = __ldg(reinterpret_cast<const uint4*>(gy + t * Cp) + vec);
          rx[k][i] = __ldg(reinterpret_cast<const uint4*>(x + t * Cp) + vec);
          if (gres != nullptr) rr[k][i] = __ldg(reinterpret_cast<const uint4*>(gres + t * Cp) + vec);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < TOK; ++k) {
      const long long t = t0 + k;
      if (t >= T) break;
      float g[NVEC][8], xh[NVEC][8];
      float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
      for (int i = 0; i < NVEC; ++i) {
        const int vec = lane + 32 * i;
        if (vec < nv) {
          float xv[8];
          ln_unpack(rg[k][i], g[i]);
          ln_unpack(rx[k][i], xv);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            xh[i][e] = (xv[e] - mean[k]) * rstd[k];
            ag[i][e] = fmaf(g[i][e], xh[i][e], ag[i][e]);
            ab[i][e] += g[i][e];
            g[i][e] *= gm[i][e];  // d xhat (zero on pads)
            s1 += g[i][e];
            s2 = fmaf(g[i][e], xh[i][e], s2);
          }
        }
      }
      s1 = warp_sum(s1) / static_cast<float>(C);
      s2 = warp_sum(s2) / static_cast<float>(C);
#pragma unroll
      for (int i = 0; i < NVEC; ++i) {
        const int vec = lane + 32 * i;
        if (vec < nv) {
          float o[8], r[8];
          ln_unpack(rr[k][i], r);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf(rstd[k], g[i][e] - s1 - xh[i][e] * s2, r[e]);
          uint4 pk = ln_pack(o);
          pk.x &= lm[i].x;
          pk.y &= lm[i].y;
          pk.z &= lm[i].z;
          pk.w &= lm[i].w;
          *(reinterpret_cast<uint4*>(gx + t * Cp) + vec) = pk;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int vec = lane + 32 * i;
    if (vec < nv) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        atomicAdd(&s_red[vec * 8 + e], ag[i][e]);
        atomicAdd(&s_red[Cp + vec * 8 + e], ab[i][e]);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(ggamma + c, s_red[c]);
    atomicAdd(gbeta + c, s_red[Cp + c]);
  }
}

// Backward, 8 lanes per token (the forward's layout): a warp handles 4 tokens per iteration, lane (slot, sub) owns the
// 16-byte vectors sub + 8 i of token slot, so every lane moves data at Cp = 192 (the warp-per-token kernel keeps 24 of
// 32 busy), a row reduction is 3 shuffles shared by four tokens, and a lane's gamma / beta partial sums stay on the
// same channels for the whole kernel.  The gy / x loads of the next token group are in flight while this one is
// reduced (the residual gradient is only needed for the final store and is fetched at the top of its own iteration).
template <int NV>
__global__ void __launch_bounds__(NV <= 2 ? 256 : 384, NV <= 2 ? 2 : 1) layernorm_bwd8_kernel(const __nv_bfloat16* __restrict__ gy,
                                                                const __nv_bfloat16* __restrict__ x,
                                                                const float* __restrict__ mean_in,
                                                                const float* __restrict__ rstd_in,
                                                                const float* __restrict__ gamma,
                                                                const __nv_bfloat16* __restrict__ gres,
                                                                __nv_bfloat16* __restrict__ gx, float* __restrict__ ggamma,
                                                                float* __restrict__ gbeta, long long T, int C, int Cp) {
  pdl_trigger();
  extern __shared__ float s_red[];  // gamma[Cp] | sum gy*xhat [Cp] | sum gy [Cp]
  for (int i = threadIdx.x; i < Cp; i += blockDim.x) {
    s_red[i] = i < C ? __ldg(gamma + i) : 0.0f;
    s_red[Cp + i] = 0.0f;
    s_red[2 * Cp + i] = 0.0f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane & 7, slot = lane >> 3;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  const int nv = Cp / 8;
  const float inv_c = 1.0f / static_cast<float>(C);
  float ag[NV][8], ab[NV][8];
  uint4 lm[NV];  // bit mask of the packed output vector: ones on real channels, zero on pads (gy, gamma are zero there)
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int e = 0; e < 8; ++e) ag[i][e] = ab[i][e] = 0.0f;
    const int c0 = (sub + 8 * i) * 8;
    auto half_mask = [&](int c) { return (c < C ? 0x0000FFFFu : 0u) | (c + 1 < C ? 0xFFFF0000u : 0u); };
    lm[i] = make_uint4(half_mask(c0), half_mask(c0 + 2), half_mask(c0 + 4), half_mask(c0 + 6));
  }
  auto load = [&](long long t, uint4 (&rg)[NV], uint4 (&rx)[NV], float& mean, float& rstd) {
    const bool tv = t < T;
    mean = tv ? __ldg(mean_in + t) : 0.0f;
    rstd = tv ? __ldg(rstd_in + t) : 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vec = sub + 8 * i;
      const bool ok = tv && vec < nv;
      rg[i] = ok ? __ldg(reinterpret_cast<const uint4*>(gy + t * Cp) + vec) : make_uint4(0u, 0u, 0u, 0u);
      rx[i] = ok ? __ldg(reinterpret_cast<const uint4*>(x + t * Cp) + vec) : make_uint4(0u, 0u, 0u, 0u);
    }
  };
  uint4 ng[NV], nx[NV];
  float nmean, nrstd;
  long long t = warp_global * 4 + slot;
  load(t, ng, nx, nmean, nrstd);
  for (; t < T + slot; t += nwarps * 4) {  // (whole warps stay in the shuffles)
    const bool tv = t < T;
    uint4 rr[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vec = sub + 8 * i;
      rr[i] = (gres != nullptr && tv && vec < nv) ? __ldg(reinterpret_cast<const uint4*>(gres + t * Cp) + vec)
                                                  : make_uint4(0u, 0u, 0u, 0u);
    }
    float g[NV][8], xh[NV][8];
    const float mean = nmean, rstd = nrstd;
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float xv[8];
      ln_unpack(ng[i], g[i]);
      ln_unpack(nx[i], xv);
      const float4 g0 = *reinterpret_cast<const float4*>(s_red + (sub + 8 * i) * 8),
                   g1 = *reinterpret_cast<const float4*>(s_red + (sub + 8 * i) * 8 + 4);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        xh[i][e] = (xv[e] - mean) * rstd;  // (rstd = 0 on tokens beyond T: nothing accumulates)
        ag[i][e] = fmaf(g[i][e], xh[i][e], ag[i][e]);
        ab[i][e] += g[i][e];
        g[i][e] *= gm[e];  // d xhat (zero on pads)
        s1 += g[i][e];
        s2 = fmaf(g[i][e], xh[i][e], s2);
      }
    }
    load(t + nwarps * 4, ng, nx, nmean, nrstd);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 4);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
    s1 *= inv_c;
    s2 *= inv_c;
    if (tv) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vec = sub + 8 * i;
        if (vec < nv) {
          float o[8], r[8];
          ln_unpack(rr[i], r);
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf(rstd, g[i][e] - s1 - xh[i][e] * s2, r[e]);
          uint4 pk = ln_pack(o);
          pk.x &= lm[i].x;
          pk.y &= lm[i].y;
          pk.z &= lm[i].z;
          pk.w &= lm[i].w;
          *(reinterpret_cast<uint4*>(gx + t * Cp) + vec) = pk;
        }
      }
    }
  }
  // column sums: the four token slots of a warp fold with two shuffles, warps meet in shared memory, CTAs in L2
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ag[i][e] += __shfl_xor_sync(0xffffffffu, ag[i][e], 8);
      ab[i][e] += __shfl_xor_sync(0xffffffffu, ab[i][e], 8);
      ag[i][e] += __shfl_xor_sync(0xffffffffu, ag[i][e], 16);
      ab[i][e] += __shfl_xor_sync(0xffffffffu, ab[i][e], 16);
    }
  if (slot == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vec = sub + 8 * i;
      if (vec < nv) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          atomicAdd(&s_red[Cp + vec * 8 + e], ag[i][e]);
          atomicAdd(&s_red[2 * Cp + vec * 8 + e], ab[i][e]);
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(ggamma + c, s_red[Cp + c]);
    atomicAdd(gbeta + c, s_red[2 * Cp + c]);
  }
}

__global__ void scale_rows_kernel(const uint4* __restrict__ g, const float* __restrict__ alpha,
                                  uint4* __restrict__ out, size_t nvec, size_t vec_per_sample) {
  pdl_trigger();  // a following programmatically serialized GEMM may start its prologue
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < nvec;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float a = __ldg(alpha + idx / vec_per_sample);
    float v[8];
    ln_unpack(__ldg(g + idx), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= a;
    out[idx] = ln_pack(v);
  }
}

}  // namespace srb

using namespace srb;

extern "C" int srb200_layernorm_fwd(const void* x_bf16, const float* gamma, const float* beta,
                                    void* y_bf16, float* mean, float* rstd, int64_t T, int C, int Cp,
                                    float eps, int ones_channel, srb200_stream_t stream) {
  if (!x_bf16 || !gamma || !beta || !y_bf16 || T <= 0 || C <= 0 || Cp < C || Cp % 8 != 0)
    return SRB200_EINVAL;
  if (ones_channel >= 0 && (ones_channel < C || ones_channel >= Cp)) return SRB200_EINVAL;  // a PAD channel only
  const int oc = ones_channel;
  if (Cp > 256 * LN_MAX_VEC) return SRB200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int block = 256;
  long long blocks = (T + 7) / 8;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_bf16);
  __nv_bfloat16* y = static_cast<__nv_bfloat16*>(y_bf16);
  if (Cp <= 256 && SRB_ENV("SRB_LN_WARP_PER_TOKEN") == nullptr) {
    // 8 lanes per token: 32 tokens per 256-thread block iteration, a few blocks per SM, grid-stride
    long long b8 = (T + 31) / 32;
    const int nvl0 = (Cp / 8 + 7) / 8;
    const long long cap8 = static_cast<long long>(num_sms()) * (nvl0 >= 3 ? 2 : 4);  // resident blocks: one wave
    if (b8 > cap8) b8 = cap8;
    const size_t smem = 0;
    const int g8 = static_cast<int>(b8);
    const int nvl = (Cp / 8 + 7) / 8;
    if (nvl == 1) layernorm_fwd8_kernel<1><<<g8, block, smem, st>>>(x, gamma, beta, y, mean, rstd, T, C, Cp, eps, oc);
    else if (nvl == 2) layernorm_fwd8_kernel<2><<<g8, block, smem, st>>>(x, gamma, beta, y, mean, rstd, T, C, Cp, eps, oc);
    else if (nvl == 3) layernorm_fwd8_kernel<3><<<g8, block, smem, st>>>(x, gamma, beta, y, mean, rstd, T, C, Cp, eps, oc);
    else layernorm_fwd8_kernel<4><<<g8, block, smem, st>>>(x, gamma, beta, y, mean, rstd, T, C, Cp, eps, oc);
    return launch_status();
  }
  const int g = static_cast<int>(blocks);
  if (Cp <= 256) layernorm_fwd_kernel<1><<<g, block, 0, st>>>(x, gamma, beta, y, mean, rstd, T, C, Cp, eps, oc);
  else if (Cp <= 512) layernorm_fwd_kernel<2><<<g, block, 0, st>>>(x, gamma, beta, y, mean, rstd, T, C, Cp, eps, oc);
  else layernorm_fwd_kernel<4><<<g, block, 0, st>>>(x, gamma, beta, y, mean, rstd, T, C, Cp, eps, oc);
  return launch_status();
}

extern "C" int srb200_layernorm_bwd(const void* gy_bf16, const void* x_bf16, const float* mean,
                                    const float* rstd, const float* gamma, const void* gres_bf16,
                                    void* gx_bf16, float* ggamma, float* gbeta, int64_t T, int C,
                                    int Cp, srb200_stream_t stream) {
  if (!gy_bf16 || !x_bf16 || !mean || !rstd || !gamma || !gx_bf16 || !ggamma || !gbeta)
    return SRB200_EINVAL;
  if (T <= 0 || C <= 0 || Cp < C || Cp % 8 != 0 || Cp > 256 * LN_MAX_VEC) return SRB200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int block = 256;
  long long blocks = (T + 63) / 64;  // >= 8 tokens per warp so the column atomics amortise
  // persistent: as many blocks as stay resident (per-block prologue / column-sum epilogue is not free)
  long long cap = static_cast<long long>(num_sms()) * (Cp <= 256 ? 3 : 1);
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const size_t smem = 2 * static_cast<size_t>(Cp) * sizeof(float);
  const __nv_bfloat16* gy = static_cast<const __nv_bfloat16*>(gy_bf16);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_bf16);
  const __nv_bfloat16* gr = static_cast<const __nv_bfloat16*>(gres_bf16);
  __nv_bfloat16* gx = static_cast<__nv_bfloat16*>(gx_bf16);
  if (Cp <= 256 && SRB_ENV("SRB_LN_WARP_PER_TOKEN") == nullptr) {
    // 8 lanes per token; Cp <= 128: two resident 256-thread blocks per SM, above (48 accumulator registers per lane)
    // one 384-thread block with up to 170 registers per thread
    const int nvl = (Cp / 8 + 7) / 8;
    const int block8 = nvl <= 2 ? 256 : 384;
    long long b8 = (T * 8 + block8 - 1) / block8;
    const long long cap8 = static_cast<long long>(num_sms()) * (nvl <= 2 ? 2 : 1);
    if (b8 > cap8) b8 = cap8;
    const size_t smem8 = 3 * static_cast<size_t>(Cp) * sizeof(float);
    const int g8 = static_cast<int>(b8);
    if (nvl == 1) layernorm_bwd8_kernel<1><<<g8, block8, smem8, st>>>(gy, x, mean, rstd, gamma, gr, gx, ggamma, gbeta, T, C, Cp);
    else if (nvl == 2) layernorm_bwd8_kernel<2><<<g8, block8, smem8, st>>>(gy, x, mean, rstd, gamma, gr, gx, ggamma, gbeta, T, C, Cp);
    else if (nvl == 3) layernorm_bwd8_kernel<3><<<g8, block8, smem8, st>>>(gy, x, mean, rstd, gamma, gr, gx, ggamma, gbeta, T, C, Cp);
    else layernorm_bwd8_kernel<4><<<g8, block8, smem8, st>>>(gy, x, mean, rstd, gamma, gr, gx, ggamma, gbeta, T, C, Cp);
    return launch_status();
  }
  const int g = static_cast<int>(blocks);
  if (Cp <= 256)
    layernorm_bwd_kernel<1, 2><<<g, block, smem, st>>>(gy, x, mean, rstd, gamma, gr, gx, ggamma, gbeta, T, C, Cp);
  else if (Cp <= 512)
    layernorm_bwd_kernel<2, 2><<<g, block, smem, st>>>(gy, x, mean, rstd, gamma, gr, gx, ggamma, gbeta, T, C, Cp);
  else
    layernorm_bwd_kernel<4, 1><<<g, block, smem, st>>>(gy, x, mean, rstd, gamma, gr, gx, ggamma, gbeta, T, C, Cp);
  return launch_status();
}

extern "C" int srb200_scale_rows(const void* g_bf16, const float* alpha, void* out_bf16, int B,
                                 int64_t elems_per_sample, srb200_stream_t stream) {
  if (!g_bf16 || !alpha || !out_bf16 || B <= 0 || elems_per_sample <= 0 || elems_per_sample % 8 != 0)
    return SRB200_EINVAL;
  const size_t vps = static_cast<size_t>(elems_per_sample) / 8;
  const size_t nvec = vps * B;
  size_t blocks = (nvec + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  scale_rows_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(g_bf16), alpha, static_cast<uint4*>(out_bf16), nvec, vps);
  return launch_status();
}
