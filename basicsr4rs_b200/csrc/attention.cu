// Fused shifted-window attention of SwinIR (swinir_arch.py:144-175, 283-323), forward and backward.
//
// One CTA = one (window, head); 4 warps x 16 query rows.  Reads q,k,v straight from the un-partitioned
// NHWC qkv tensor [B,H,W,3*Cp] (channel = which*Cp + head*32 + d, head_dim padded to 32): the cyclic shift
// (torch.roll, :293-296) and window_partition (:299) are address arithmetic, the relative-position bias
// (:157-160) is a 225-entry smem table indexed analytically, the shifted-window mask (:262-281, 0 / -100)
// is computed from 3x3 region ids -- no [B*nW, nH, 64, 64] attention matrix ever reaches memory.
//   S = scale * Q K^T + bias + mask -> softmax (fp32 registers) -> O = P V   on mma.sync bf16 tensor cores,
// window_reverse + roll back (:309-316) is again the store address.  Algorithmic traffic = read qkv +
// write o = 4*T*Cp*2 bytes (HBM-bound, 32 flop/B).
// Backward recomputes S / P, then dP = dO V^T, dS = P o (dP - rowsum(P o dP)), dQ = scale dS K,
// dK = scale dS^T Q, dV = P^T dO and the bias-table gradient (smem histogram -> global atomics).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace srb {

constexpr int WS_MAX = 8;   // window sizes 2..8: a window's ws*ws tokens occupy the first rows of the 64-row tiles,
constexpr int NTOK = 64;    // the rest is padding (zero q/k/v, keys masked out of the softmax, rows never stored)
constexpr int HD = 32;
constexpr int QROW = 40;   // padded smem row (elements) of the 32-wide tiles: conflict-free ldmatrix
constexpr int PROW = 72;   // padded smem row of the 64-wide P / dS tiles
constexpr int NBIAS = (2 * WS_MAX - 1) * (2 * WS_MAX - 1);

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnGeom {
  int B, H, W, nH, Cp, shift;
  float scale;
  int ws;  // window size (<= WS_MAX)
  int ones;  // forward: head 0 writes 1.0 into its pad channel 31 (SRB200_ATTN_ONES)
};

// global element offset of token n of window (b, wy, wx) in an NHWC tensor with `ld` channels
__device__ __forceinline__ size_t token_offset(const AttnGeom& g, int b, int wy, int wx, int n,
                                               int ld) {
  const int ny = n / g.ws;
  int y = wy * g.ws + ny + g.shift;
  int x = wx * g.ws + (n - ny * g.ws) + g.shift;
  if (y >= g.H) y -= g.H;
  if (x >= g.W) x -= g.W;
  return ((static_cast<size_t>(b) * g.H + y) * g.W + x) * ld;
}

// Each of the 128 threads moves the same two 16-byte pieces (tokens tid/4 and tid/4 + 32, part tid%4) of every
// [64 tokens x 32] head slice: their pixel indices are computed ONCE per CTA (TokenSlots) -- these kernels are
// instruction-bound and the 64-bit offset arithmetic per piece was a fifth of their instructions.
struct TokenSlots {
  long long pix[2];  // pixel index ((b*H + y)*W + x) of the thread's two tokens, -1 = padding token (ws < 8)
};
__device__ __forceinline__ TokenSlots token_slots(const AttnGeom& g, int b, int wy, int wx) {
  TokenSlots t;
  const int ntok = g.ws * g.ws;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = (threadIdx.x >> 2) + 32 * i;
    t.pix[i] = n < ntok ? static_cast<long long>(token_offset(g, b, wy, wx, n, 1)) : -1;
  }
  return t;
}
// copy a [64 tokens x 32] head slice global -> smem (row stride QROW)
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* gsrc, const TokenSlots& t, int ld,
                                          int chan0) {
  const int part = threadIdx.x & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = (threadIdx.x >> 2) + 32 * i;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t.pix[i] >= 0) v = __ldg(reinterpret_cast<const uint4*>(gsrc + t.pix[i] * ld + chan0 + part * 8));
    *reinterpret_cast<uint4*>(s + n * QROW + part * 8) = v;
  }
}
__device__ __forceinline__ void store_tile(const __nv_bfloat16* s, __nv_bfloat16* gdst, const TokenSlots& t, int ld,
                                           int chan0) {
  const int part = threadIdx.x & 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int n = (threadIdx.x >> 2) + 32 * i;
    if (t.pix[i] >= 0)
      *reinterpret_cast<uint4*>(gdst + t.pix[i] * ld + chan0 + part * 8) =
          *reinterpret_cast<const uint4*>(s + n * QROW + part * 8);
  }
}

// C[16 x 64] (this warp's rows) = A[rows 16w.., 32] * Bmat[64 x 32]^T, both tiles stored [row][32]
__device__ __forceinline__ void gemm_nt_16x64x32(float (&acc)[8][4], const __nv_bfloat16* sA,
                                                 const __nv_bfloat16* sB, int warp, int lane) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t a[4];
    ldsm_x4(a, smem_u32(sA + (16 * warp + (lane & 15)) * QROW + kk * 16 + (lane >> 4) * 8));
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t r[4];
      const int idx = lane >> 3, rr = lane & 7;
      ldsm_x4(r, smem_u32(sB + (16 * np + rr + (idx >= 2 ? 8 : 0)) * QROW + kk * 16 + (idx & 1) * 8));
      mma16816(acc[2 * np], a, r[0], r[1]);
      mma16816(acc[2 * np + 1], a, r[2], r[3]);
    }
  }
}

// C[16 x 32] = A[16 x 64] (registers, bf16 fragments per 16-key step) * Bmat[64 x 32] stored [k][32]
__device__ __forceinline__ void gemm_regA_16x32x64(float (&acc)[4][4], const uint32_t (&afrag)[4][4],
                                                   const __nv_bfloat16* sB, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int dp = 0; dp < 2; ++dp) {
      uint32_t r[4];
      const int idx = lane >> 3, rr = lane & 7;
      ldsm_x4_t(r, smem_u32(sB + (16 * kk + rr + (idx & 1) * 8) * QROW + 16 * dp + (idx >= 2 ? 8 : 0)));
      mma16816(acc[2 * dp], afrag[kk], r[0], r[1]);
      mma16816(acc[2 * dp + 1], afrag[kk], r[2], r[3]);
    }
  }
}

// C[16 x 32] (rows = columns 16w.. of sT) = sT^T[16 x 64] * Bmat[64 x 32]; sT stored [k=64][64] (row PROW)
__device__ __forceinline__ void gemm_tA_16x32x64(float (&acc)[4][4], const __nv_bfloat16* sT,
                                                 const __nv_bfloat16* sB, int warp, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    const int idx = lane >> 3, rr = lane & 7;
    ldsm_x4_t(a, smem_u32(sT + (16 * kk + rr + (idx >= 2 ? 8 : 0)) * PROW + 16 * warp + (idx & 1) * 8));
#pragma unroll
    for (int dp = 0; dp < 2; ++dp) {
      uint32_t r[4];
      ldsm_x4_t(r, smem_u32(sB + (16 * kk + rr + (idx & 1) * 8) * QROW + 16 * dp + (idx >= 2 ? 8 : 0)));
      mma16816(acc[2 * dp], a, r[0], r[1]);
      mma16816(acc[2 * dp + 1], a, r[2], r[3]);
    }
  }
}

__device__ __forceinline__ int rel_index(int row, int col, int ws) {
  const int ry = row / ws, cy = col / ws;
  return (ry - cy + ws - 1) * (2 * ws - 1) + ((row - ry * ws) - (col - cy * ws) + ws - 1);
}

// logits -> probabilities in place (unnormalised exp); returns 1/rowsum for the two rows of this thread
__device__ __forceinline__ void softmax_rows(float (&s)[8][4], const float* sBias, const int* sRegion,
                                             bool masked, float scale, int warp, int lane,
                                             float (&inv)[2], int ws) {
  const int g = lane >> 2, t = lane & 3;
  const int ntok = ws * ws;
  float mx[2] = {-1e30f, -1e30f};
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int row = 16 * warp + g + ((e & 2) ? 8 : 0);
      const int col = nt * 8 + 2 * t + (e & 1);
      float v = -1e30f;  // padding keys (ws < 8) drop out of the softmax; padding query rows are never stored
      if (col < ntok) {
        v = 0.0f;
        if (row < ntok) {
          v = s[nt][e] * scale + sBias[rel_index(row, col, ws)];
          if (masked && sRegion[row] != sRegion[col]) v += -100.0f;
        }
      }
      s[nt][e] = v;
      mx[e >> 1] = fmaxf(mx[e >> 1], v);
    }
  float sum[2] = {0.0f, 0.0f};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
    mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
  }
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float p = __expf(s[nt][e] - mx[e >> 1]);
      s[nt][e] = p;
      sum[e >> 1] += p;
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
    sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
    inv[h] = 1.0f / sum[h];
  }
}

__device__ __forceinline__ void setup_window(const AttnGeom& g, int& b, int& wy, int& wx, int& head,
                                             float* sBias, int* sRegion,
                                             const float* __restrict__ table) {
  // grid = (heads * windows per image, batch): gridDim.x holds 2^31 - 1 blocks, so even a 16k x 16k scene fits
  const int nWw = g.W / g.ws;
  const int win = blockIdx.x / g.nH;
  head = blockIdx.x - win * g.nH;
  wy = win / nWw;
  wx = win - wy * nWw;
  b = blockIdx.z;
  for (int i = threadIdx.x; i < (2 * g.ws - 1) * (2 * g.ws - 1); i += blockDim.x)
    sBias[i] = __ldg(table + i * g.nH + head);
  if (threadIdx.x < NTOK) {
    int id = 0;
    if (g.shift > 0 && threadIdx.x < g.ws * g.ws) {
      const int ty = threadIdx.x / g.ws;
      const int ys = wy * g.ws + ty, xs = wx * g.ws + (threadIdx.x - ty * g.ws);
      const int rh = ys < g.H - g.ws ? 0 : (ys < g.H - g.shift ? 1 : 2);
      const int rw = xs < g.W - g.ws ? 0 : (xs < g.W - g.shift ? 1 : 2);
      id = rh * 3 + rw;
    }
    sRegion[threadIdx.x] = id;
  }
}

// WST > 0: window size known at compile time (8 for every classical-SR recipe) -- the token / bias index arithmetic
// folds to shifts; WST = 0 reads it from the geometry (the generic path costs ~2x in these ALU-bound kernels).
template <int WST>
__global__ void __launch_bounds__(128) window_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                              const float* __restrict__ table,
                                                              __nv_bfloat16* __restrict__ out,
                                                              AttnGeom g_in) {
  pdl_trigger();  // a following programmatically serialized GEMM may start its prologue
  AttnGeom g = g_in;
  if (WST > 0) g.ws = WST;
  __shared__ __align__(16) __nv_bfloat16 sQ[NTOK * QROW];
  __shared__ __align__(16) __nv_bfloat16 sK[NTOK * QROW];
  __shared__ __align__(16) __nv_bfloat16 sV[NTOK * QROW];
  __shared__ float sBias[NBIAS];
  __shared__ int sRegion[NTOK];
  int b, wy, wx, head;
  setup_window(g, b, wy, wx, head, sBias, sRegion, table);
  const int ld = 3 * g.Cp;
  const TokenSlots slots = token_slots(g, b, wy, wx);
  load_tile(sQ, qkv, slots, ld, head * HD);
  load_tile(sK, qkv, slots, ld, g.Cp + head * HD);
  load_tile(sV, qkv, slots, ld, 2 * g.Cp + head * HD);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) s[i][e] = 0.0f;
  gemm_nt_16x64x32(s, sQ, sK, warp, lane);
  float inv[2];
  softmax_rows(s, sBias, sRegion, g.shift > 0, g.scale, warp, lane, inv, g.ws);
  uint32_t pf[4][4];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    pf[kk][0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    pf[kk][1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    pf[kk][2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pf[kk][3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
  }
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[i][e] = 0.0f;
  gemm_regA_16x32x64(o, pf, sV, lane);
  // stage O in this warp's own rows of sQ (only this warp ever read them), then 16-byte global stores
  const int gq = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    *reinterpret_cast<uint32_t*>(sQ + (16 * warp + gq) * QROW + nt * 8 + 2 * t) =
        pack_bf16x2(o[nt][0] * inv[0], o[nt][1] * inv[0]);
    *reinterpret_cast<uint32_t*>(sQ + (16 * warp + gq + 8) * QROW + nt * 8 + 2 * t) =
        pack_bf16x2(o[nt][2] * inv[1], o[nt][3] * inv[1]);
  }
  __syncthreads();
  if (g.ones && head == 0 && threadIdx.x < NTOK) sQ[threadIdx.x * QROW + 31] = __float2bfloat16(1.0f);
  __syncthreads();
  store_tile(sQ, out, slots, g.Cp, head * HD);
}

template <int WST>
__global__ void __launch_bounds__(128) window_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                              const __nv_bfloat16* __restrict__ gout,
                                                              const float* __restrict__ table,
                                                              __nv_bfloat16* __restrict__ gqkv,
                                                              float* __restrict__ gtable, AttnGeom g_in) {
  pdl_trigger();  // a following programmatically serialized GEMM may start its prologue
  AttnGeom g = g_in;
  if (WST > 0) g.ws = WST;
  __shared__ __align__(16) __nv_bfloat16 sQ[NTOK * QROW];
  __shared__ __align__(16) __nv_bfloat16 sK[NTOK * QROW];
  __shared__ __align__(16) __nv_bfloat16 sV[NTOK * QROW];
  __shared__ __align__(16) __nv_bfloat16 sdO[NTOK * QROW];
  __shared__ __align__(16) __nv_bfloat16 sP[NTOK * PROW];
  __shared__ __align__(16) __nv_bfloat16 sdS[NTOK * PROW];
  __shared__ float sBias[NBIAS];
  __shared__ float sdBias[NBIAS];
  __shared__ int sRegion[NTOK];
  int b, wy, wx, head;
  setup_window(g, b, wy, wx, head, sBias, sRegion, table);
  const int ld = 3 * g.Cp;
  const TokenSlots slots = token_slots(g, b, wy, wx);
  load_tile(sQ, qkv, slots, ld, head * HD);
  load_tile(sK, qkv, slots, ld, g.Cp + head * HD);
  load_tile(sV, qkv, slots, ld, 2 * g.Cp + head * HD);
  load_tile(sdO, gout, slots, g.Cp, head * HD);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, t = lane & 3;

  float s[8][4], dp[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s[i][e] = 0.0f;
      dp[i][e] = 0.0f;
    }
  gemm_nt_16x64x32(s, sQ, sK, warp, lane);
  float inv[2];
  softmax_rows(s, sBias, sRegion, g.shift > 0, g.scale, warp, lane, inv, g.ws);
  gemm_nt_16x64x32(dp, sdO, sV, warp, lane);  // dP = dO V^T
  float delta[2] = {0.0f, 0.0f};
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s[nt][e] *= inv[e >> 1];  // P
      delta[e >> 1] += s[nt][e] * dp[nt][e];
    }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    delta[h] += __shfl_xor_sync(0xffffffffu, delta[h], 1);
    delta[h] += __shfl_xor_sync(0xffffffffu, delta[h], 2);
  }
  // dS = P o (dP - delta); bias-table histogram; P and dS to smem (bf16) for the transposed products
  uint32_t dsf[4][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    float ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ds[e] = s[nt][e] * (dp[nt][e] - delta[e >> 1]);
    }
    const int r0 = 16 * warp + gq, c0 = nt * 8 + 2 * t;
    const uint32_t p01 = pack_bf16x2(s[nt][0], s[nt][1]), p23 = pack_bf16x2(s[nt][2], s[nt][3]);
    const uint32_t d01 = pack_bf16x2(ds[0], ds[1]), d23 = pack_bf16x2(ds[2], ds[3]);
    *reinterpret_cast<uint32_t*>(sP + r0 * PROW + c0) = p01;
    *reinterpret_cast<uint32_t*>(sP + (r0 + 8) * PROW + c0) = p23;
    *reinterpret_cast<uint32_t*>(sdS + r0 * PROW + c0) = d01;
    *reinterpret_cast<uint32_t*>(sdS + (r0 + 8) * PROW + c0) = d23;
    dsf[nt >> 1][(nt & 1) * 2 + 0] = d01;
    dsf[nt >> 1][(nt & 1) * 2 + 1] = d23;
  }
  __syncthreads();
  float dq[4][4], dk[4][4], dv[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      dq[i][e] = 0.0f;
      dk[i][e] = 0.0f;
      dv[i][e] = 0.0f;
    }
  // Bias-table gradient: bin (dy, dx) collects dS[(ry,rx),(ry-dy,rx-dx)].  One thread per bin sums its <= 64
  // entries of the staged dS tile -- no shared-memory float atomics (4096 contended CAS loops per CTA were the
  // bulk of this kernel's time), then one global atomic per bin as before.
  const int ws = g.ws, nbias = (2 * ws - 1) * (2 * ws - 1);
  for (int bin = threadIdx.x; bin < nbias; bin += blockDim.x) {
    const int dy = bin / (2 * ws - 1) - (ws - 1), dx = bin % (2 * ws - 1) - (ws - 1);
    float acc = 0.0f;
    for (int ry = (dy > 0 ? dy : 0); ry < (dy < 0 ? ws + dy : ws); ++ry)
      for (int rx = (dx > 0 ? dx : 0); rx < (dx < 0 ? ws + dx : ws); ++rx)
        acc += __bfloat162float(sdS[(ry * ws + rx) * PROW + (ry - dy) * ws + (rx - dx)]);
    sdBias[bin] = acc;
  }
  gemm_regA_16x32x64(dq, dsf, sK, lane);          // dQ = dS K      (rows = this warp's queries)
  gemm_tA_16x32x64(dk, sdS, sQ, warp, lane);      // dK = dS^T Q    (rows = keys 16w..)
  gemm_tA_16x32x64(dv, sP, sdO, warp, lane);      // dV = P^T dO
  __syncthreads();                                // every warp is done reading sQ / sK / sV / sdO
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int r0 = 16 * warp + gq, c0 = nt * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(sQ + r0 * QROW + c0) = pack_bf16x2(dq[nt][0] * g.scale, dq[nt][1] * g.scale);
    *reinterpret_cast<uint32_t*>(sQ + (r0 + 8) * QROW + c0) = pack_bf16x2(dq[nt][2] * g.scale, dq[nt][3] * g.scale);
    *reinterpret_cast<uint32_t*>(sK + r0 * QROW + c0) = pack_bf16x2(dk[nt][0] * g.scale, dk[nt][1] * g.scale);
    *reinterpret_cast<uint32_t*>(sK + (r0 + 8) * QROW + c0) = pack_bf16x2(dk[nt][2] * g.scale, dk[nt][3] * g.scale);
    *reinterpret_cast<uint32_t*>(sV + r0 * QROW + c0) = pack_bf16x2(dv[nt][0], dv[nt][1]);
    *reinterpret_cast<uint32_t*>(sV + (r0 + 8) * QROW + c0) = pack_bf16x2(dv[nt][2], dv[nt][3]);
  }
  __syncthreads();
  store_tile(sQ, gqkv, slots, ld, head * HD);
  store_tile(sK, gqkv, slots, ld, g.Cp + head * HD);
  store_tile(sV, gqkv, slots, ld, 2 * g.Cp + head * HD);
  for (int i = threadIdx.x; i < nbias; i += blockDim.x) atomicAdd(gtable + i * g.nH + head, sdBias[i]);
}

static int check_geom(int B, int H, int W, int nH, int Cp, int ws, int shift) {
  if (B <= 0 || H <= 0 || W <= 0 || nH <= 0) return SRB200_EINVAL;
  if (ws < 2 || ws > WS_MAX || H % ws != 0 || W % ws != 0 || shift < 0 || shift >= ws) return SRB200_EINVAL;
  if (Cp != nH * HD) return SRB200_EINVAL;
  return SRB200_OK;
}

}  // namespace srb

using namespace srb;

// attention_tc.cu: the tcgen05 / TMEM / TMA kernels (window 8); SRB200_EINVAL = shape outside their domain
int srb_window_attention_fwd_tc(const void* qkv_bf16, const float* rpb_table, void* out_bf16, float* stats, int B,
                                int H, int W, int num_heads, int Ca, int shift, float scale, int flags,
                                const float* out_alpha, cudaStream_t stream);
int srb_window_attention_bwd_tc(const void* qkv_bf16, const void* gout_bf16, const float* rpb_table,
                                const float* stats, void* gqkv_bf16, float* g_rpb_table, float* workspace, int B,
                                int H, int W, int num_heads, int Ca, int shift, float scale, cudaStream_t stream);

extern "C" int srb200_window_attention_fwd(const void* qkv_bf16, const float* rpb_table,
                                           void* out_bf16, float* stats, int B, int H, int W, int num_heads,
                                           int Cp, int window_size, int shift, float scale, int flags,
                                           const float* out_alpha, srb200_stream_t stream) {
  if (!qkv_bf16 || !rpb_table || !out_bf16) return SRB200_EINVAL;
  const int rc = check_geom(B, H, W, num_heads, Cp, window_size, shift);
  if (rc != SRB200_OK) return rc;
  if (window_size == 8 && SRB_ENV("SRB_ATTN_MMA_SYNC") == nullptr) {
    // window 8 (every classical-SR recipe): the tcgen05 kernel; other windows / odd head counts: mma.sync below
    const int rc_tc = srb_window_attention_fwd_tc(qkv_bf16, rpb_table, out_bf16, stats, B, H, W, num_heads, Cp, shift, scale,
                                                  flags, out_alpha, static_cast<cudaStream_t>(stream));
    if (rc_tc != SRB200_EINVAL) return rc_tc;
  }
  if (out_alpha != nullptr) return SRB200_EINVAL;  // the per-sample output scale exists in the tcgen05 kernel only
  AttnGeom g{B, H, W, num_heads, Cp, shift, scale, window_size, (flags & SRB200_ATTN_ONES) ? 1 : 0};
  const long long wins = static_cast<long long>(H / window_size) * (W / window_size);
  if (wins * num_heads > 0x7fffffffLL || B > 65535) return SRB200_EINVAL;
  const dim3 grid(static_cast<unsigned>(wins * num_heads), 1, B);
  if (window_size == 8)
    window_attn_fwd_kernel<8><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(qkv_bf16), rpb_table, static_cast<__nv_bfloat16*>(out_bf16), g);
  else
    window_attn_fwd_kernel<0><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(qkv_bf16), rpb_table, static_cast<__nv_bfloat16*>(out_bf16), g);
  return launch_status();
}

extern "C" int srb200_window_attention_bwd(const void* qkv_bf16, const void* gout_bf16,
                                           const float* rpb_table, const float* stats, void* gqkv_bf16,
                                           float* g_rpb_table, float* workspace, int B, int H, int W, int num_heads,
                                           int Cp, int window_size, int shift, float scale,
                                           srb200_stream_t stream) {
  if (!qkv_bf16 || !gout_bf16 || !rpb_table || !gqkv_bf16 || !g_rpb_table) return SRB200_EINVAL;
  const int rc = check_geom(B, H, W, num_heads, Cp, window_size, shift);
  if (rc != SRB200_OK) return rc;
  if (window_size == 8 && stats != nullptr && SRB_ENV("SRB_ATTN_MMA_SYNC") == nullptr &&
      SRB_ENV("SRB_ATTN_BWD_MMA_SYNC") == nullptr) {
    const int rc_tc = srb_window_attention_bwd_tc(qkv_bf16, gout_bf16, rpb_table, stats, gqkv_bf16, g_rpb_table, workspace, B, H, W,
                                                  num_heads, Cp, shift, scale, static_cast<cudaStream_t>(stream));
    if (rc_tc != SRB200_EINVAL) return rc_tc;
  }
  AttnGeom g{B, H, W, num_heads, Cp, shift, scale, window_size, 0};
  const long long wins = static_cast<long long>(H / window_size) * (W / window_size);
  if (wins * num_heads > 0x7fffffffLL || B > 65535) return SRB200_EINVAL;
  const dim3 grid(static_cast<unsigned>(wins * num_heads), 1, B);
  if (window_size == 8)
    window_attn_bwd_kernel<8><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(qkv_bf16), static_cast<const __nv_bfloat16*>(gout_bf16), rpb_table,
        static_cast<__nv_bfloat16*>(gqkv_bf16), g_rpb_table, g);
  else
    window_attn_bwd_kernel<0><<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(qkv_bf16), static_cast<const __nv_bfloat16*>(gout_bf16), rpb_table,
        static_cast<__nv_bfloat16*>(gqkv_bf16), g_rpb_table, g);
  return launch_status();
}
