// tap-GEMM: conv3x3 / conv1x1 / Linear as a TMA-fed tcgen05 implicit GEMM (sm_100a).
//
//   out[b,y,x,n] = epi( sum_{tap,k} A[b, y+dy(tap), x+dx(tap), k] * Wp[tap][n][k] )
//
// * A is an NHWC bf16 tensor read through a 4-D TMA tensor map (C, W, H, B): one box is a
//   (64 ch) x (tw x th = 128 pixel) spatial patch; the 3x3 halo is a coordinate offset and the
//   zero padding of nn.Conv2d(padding=1) is TMA's out-of-bounds zero fill -- no im2col buffer,
//   no padded copy, no index arithmetic per element.
// * Wp is [taps][N][K] bf16 (K-major), one 2-D box (64 x BLOCK_N) per k-block.
// * both land in 128B-swizzled smem rows, are consumed by tcgen05.mma (M=128, N=BLOCK_N,
//   K=16) issued by one thread, accumulate in TMEM (fp32, double-buffered across tiles) and
//   are drained by 4 epilogue warps with tcgen05.ld while the next tile's MMAs run.
// * persistent CTAs (one per SM), static round-robin tile schedule.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue: TMEM lane quarter
// = warp_idx % 4, and the two warps of a quarter split every 64-column chunk in halves.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace srb {

struct TapGemmParams {
  CUtensorMap tmap_a[9];    // one per source view (src_r^2 <= 9)
  CUtensorMap tmap_w;
  CUtensorMap tmap_out[9];  // output views for the TMA-store epilogue (out_r^2 views for the fused PixelShuffle)
  CUtensorMap tmap_aux;
  CUtensorMap tmap_res;     // RES kernels: the staged epilogue operand (mask_src or residual), geometry of tmap_out[0]
  int res_kind;             // RES kernels: 1 = mask_src, 2 = residual arrives through shared memory (TMA), 0 = neither
  int tma_store;            // 1: bf16 results leave through smem staging + cp.async.bulk.tensor stores
  unsigned long long* trace;  // debug: per-role clock64 timeline of CTA 0 (NULL = off)
  int B, H, W;
  int tile_w, tile_h, tiles_x, tiles_y;
  int m_tiles, n_tiles;
  // ceil(2^48 / d) for d = n_tiles, tiles_x, tiles_x * tiles_y: the per-tile index arithmetic of all three warp
  // roles without the ~60-cycle integer division sequences (exact for operands < 2^24)
  unsigned long long magic_n, magic_x, magic_xy;
  int kchunks_per_src;  // Cin / 64
  int num_src;          // src_r^2
  int Cout;             // packed N
  int taps, ksize, flip;
  // epilogue
  const float* bias;
  const __nv_bfloat16* mask_src;
  const __nv_bfloat16* residual;
  const float* out_shift;
  void* out;
  __nv_bfloat16* aux_out;
  const float* residual_f32;  // optional fp32 residual (layout of out)
  float* out_f32;             // optional fp32 copy of the result (layout of out)
  const float* alpha_b;       // optional per-sample scale [B] (stochastic depth)
  int alpha_on_bias;          // 1: alpha_b multiplies the bias only (SRB200_EXT_ALPHA_ON_BIAS)
  float* colsum;              // optional fp32 [Cout]: += column sums of the stored result (bias gradient)
  int aux_mode;               // 1: aux_out = act'(pre-activation) instead of the pre-activation (GELU only)
  int pdl;                    // launched with programmatic stream serialization: do the griddepcontrol handoff
  int colsum_per_image;       // 1: colsum is [B][Cout] (per-sample sums: RCAN's global average pool)
  float colsum_scale;         // factor applied to the sums when they are flushed (1/HW for the pool)
  // LayerNorm folded into the GEMMs around it (evaluation): see srb200_tapgemm_ext.ln_*
  const float2* ln_in;        // [pixels] (mean, rstd) of the INPUT rows: v = rstd * (acc - mean * ln_wsum[n]) + bias[n]
  const float* ln_wsum;       // [Cout] row sums of the (gamma-folded, bf16-rounded) weights
  float2* ln_out;             // [pixels] (mean, rstd) of the STORED rows (single N tile, two-team kernels)
  float ln_inv_c, ln_eps;     // 1 / (real channel count), epsilon of the LayerNorm that will consume ln_out
  int act;
  float act_slope, alpha;
  int mask_mode;
  float mask_slope;
  int out_mode, out_r, out_c;
  float out_scale;
};

// TWO = CTA pair (cluster of 2, tcgen05 cta_group::2): one M=256 tile per pair, every CTA stages its own 128
// rows of A but only HALF of B -- 32 KB instead of 48 KB of TMA traffic per 512-cycle k-block (N=256), which is
// what a single SM's TMA path cannot sustain (measured ~80 B/cycle/SM: 616 cycles per k-block with one CTA).
// RES: the bf16 epilogue operand of a Linear (the residual, or the mask / derivative tensor of a data-gradient GEMM)
// is staged through shared memory by TMA, one 64-column chunk ahead of the epilogue team that consumes it (2 slots per
// team), at the price of two operand stages.  Read straight from global memory -- every thread four 16-byte pieces of
// its own row per 32 columns, issued only when the accumulator is already in registers -- that operand made the
// HBM-bound linears latency-bound (ncu: 5.6-7.2 warps stalled on long_scoreboard per issue; fc2's data gradient took
// 53 us for 125 MB).
template <int BLOCK_N, bool TWO = false, bool RES = false>
struct TapCfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int B_ROWS = TWO ? BLOCK_N / 2 : BLOCK_N;  // rows of the weight box this CTA loads
  static constexpr int B_BYTES = B_ROWS * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // epilogue staging: NBUF x [128 rows x 64 ch] bf16 tiles in the TMA 128B-swizzle layout.  A bulk tensor
  // store holds its buffer for ~1.4k cycles, so 4 buffers (where the operand ring leaves room) keep 3 in flight.
  static constexpr int NBUF = (BLOCK_N >= 64 && BLOCK_N <= 192) ? 4 : 2;
  static constexpr int STORE_BYTES = BLOCK_N >= 64 ? NBUF * 128 * 128 : 0;
  static constexpr int RES_BYTES = RES ? 4 * 128 * 128 : 0;
  static constexpr int STAGES_RAW = (232448 - 1024 - 256 - 1024 - STORE_BYTES - RES_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 32)    ? 32
                                   : (2 * BLOCK_N <= 64)  ? 64
                                   : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256
                                                          : 512;
  static constexpr int CHUNK = BLOCK_N >= 32 ? 32 : 16;  // epilogue columns per tcgen05.ld
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + STORE_BYTES + RES_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*bias*/;
};

// erf via Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below bf16 resolution): one ex2 + one rcp + 5 FMA
// instead of libdevice's branchy erff -- the GELU epilogues are ALU-bound otherwise.
// Returns e = exp(-z^2) as well, which d/dx GELU needs anyway.
__device__ __forceinline__ float erf_as(float z_signed, float& e) {
  // MUFU.RCP / MUFU.EX2 directly: the IEEE-rounded __frcp_rn and the range-checked exp2f expand to ~40 more
  // instructions per element, which made the GELU epilogues ALU-bound (18k instead of 9.5k cycles per tile)
  const float z = fabsf(z_signed);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float r = fmaf(-poly * t, e, 1.0f);
  return copysignf(r, z_signed);
}
__device__ __forceinline__ float gelu_erf(float x) {  // nn.GELU() exact (erf) form, swinir_arch.py:46,57
  float e;
  return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752f, e));
}
__device__ __forceinline__ float dgelu_erf(float x) {
  float e;  // e = exp(-x^2/2)
  const float cdf = 0.5f * (1.0f + erf_as(x * 0.70710678118654752f, e));
  return fmaf(x * 0.3989422804014327f, e, cdf);
}

// gelu(x) and gelu'(x) from one erf / exp evaluation (fc1 forward stores the derivative for the backward).
// (Tried and dropped: Phi through ONE MUFU.TANH -- 0.5 (1 + tanh(x (c0 + c1 x^2 + c2 x^4))), fitted to 3e-5 -- is 7 + 1
// instead of 13 + 2 instructions per element and took fc1's epilogue kernel from 35 to 28 us, but the 2^-11 relative
// error of tanh.approx pushed two of the six SwinIR golden fixtures from just under to just over the 1e-2 max-abs bar.)
__device__ __forceinline__ void gelu_and_grad(float x, float& g, float& d) {
  float e;
  const float cdf = 0.5f * (1.0f + erf_as(x * 0.70710678118654752f, e));
  g = x * cdf;
  d = fmaf(x * 0.3989422804014327f, e, cdf);
}

// Column sums over the 32 rows a warp owns: butterfly reduce-scatter (31 shuffles); lane l returns the sum of
// column l.  w is destroyed.
__device__ __forceinline__ float warp_colsum32(float (&w)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? w[i] : w[i + s];
      const float keep = up ? w[i + s] : w[i];
      w[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return w[0];
}

// TEAMS: the 8 epilogue warps work as two independent teams of 4 (one warp per TMEM lane quarter), team a draining
// accumulator buffer a.  For small-K layers (SwinIR's linears, 64-channel convs) a tile's MMAs take ~1.2k cycles but
// its epilogue 4k, so both accumulators are usually full and two tiles' epilogues can run side by side -- each team
// has its own named barrier, staging slots and TMA-store issuer.
// LN: the LayerNorm-folding epilogue stages (srb200_tapgemm_ext.ln_*) are compiled in -- separate instantiations, so
// that the training kernels keep their code size and register budget (as run-time branches in the one kernel they cost
// the SwinIR training step 4 %).
template <int BLOCK_N, bool TWO = false, bool TEAMS = false, bool RES = false, bool LN = false>
__global__ void __launch_bounds__(320, 1) tapgemm_kernel(const __grid_constant__ TapGemmParams p) {
  static_assert(!LN || TEAMS, "LayerNorm folding: two-team kernels only");
  using Cfg = TapCfg<BLOCK_N, TWO, RES>;
  static_assert(!TEAMS || (!TWO && BLOCK_N >= 64 && BLOCK_N <= 192), "teams: single-CTA, TMA-store tile widths");
  static_assert(!RES || TEAMS, "staged epilogue operand: two-team kernels only");
  constexpr int STAGES = Cfg::STAGES;
  constexpr int CHUNK = Cfg::CHUNK;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t store_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t res_base = store_base + Cfg::STORE_BYTES;  // RES: [2 teams][2 slots][128 rows x 64 ch] swizzled
  const uint32_t bar_base = res_base + Cfg::RES_BYTES;
  // barrier layout: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem_ptr, res_full[2 teams][2 slots]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 4);
  auto rfull_bar = [&](int team, int slot) { return bar_base + 8u * (2 * STAGES + 5 + 2 * team + slot); };
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));  // [BLOCK_N]
  auto smem_a = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

  // (the shuffle tells the compiler the warp index is warp-uniform: the role branches below stay convergent and
  // their address / descriptor arithmetic lives in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // work units: one M tile (x N tile) per CTA, or one PAIR of consecutive M tiles per CTA pair
  const int rank = TWO ? static_cast<int>(cluster_ctarank()) : 0;
  const int unit0 = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int unit_stride = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int m_units = TWO ? (p.m_tiles + 1) / 2 : p.m_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.tmap_a[s]);
    tma_prefetch_desc(&p.tmap_w);
    if constexpr (RES) tma_prefetch_desc(&p.tmap_res);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), TWO ? 16 : (TEAMS ? 4 : 8));  // (pair: both CTAs' epilogue warps; teams: one team)
    }
    if constexpr (RES) {
      for (int a = 0; a < 4; ++a) mbar_init(rfull_bar(a >> 1, a & 1), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (TWO) {
      tmem_alloc_2cta(tmem_ptr_smem, Cfg::TMEM_COLS);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / complete_tx
  tc_fence_after();
  if (p.pdl) pdl_handoff();  // everything above touched no global memory
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  const int total_tiles = m_units * p.n_tiles;
  const int kchunks = p.kchunks_per_src * p.num_src;
  // trace layout: [role 0..3][64 slots] (role 3: producer k-blocks of the first tile: before wait / after A issue); role 0 producer (first TMA of tile issued), 1 MMA (tile start / all
  // MMAs issued), 2 epilogue warp 2 (accumulator ready / tile drained)
  unsigned long long* trc = (p.trace != nullptr && blockIdx.x == 0) ? p.trace : nullptr;
  auto stamp = [&](int role, int slot) {
    if (trc != nullptr && slot < 64 && lane == 0) trc[role * 64 + slot] = clock64();
  };
  if (threadIdx.x == 0) stamp(0, 63);
  if (p.trace != nullptr && threadIdx.x == 0) {  // per-CTA wall/cycle start (debug)
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[256 + 4 * blockIdx.x + 0] = gt;
    p.trace[256 + 4 * blockIdx.x + 2] = clock64();
  }
  const int num_kb = p.taps * kchunks;

  if (warp == 0) {
    // ===================================================== TMA producer (convergent warp loop, one elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = unit0; t < total_tiles; t += unit_stride) {
        const int t_div = fast_div(t, p.magic_n);
        const int n_t = t - t_div * p.n_tiles;
        const int m_t = TWO ? 2 * t_div + rank : t_div;
        const int b = fast_div(m_t, p.magic_xy);  // == B for the missing half of an odd last pair: zero fill
        const int m_in = m_t - b * (p.tiles_x * p.tiles_y);
        const int ty = fast_div(m_in, p.magic_x);
        const int tx = m_in - ty * p.tiles_x;
        const int x0 = tx * p.tile_w, y0 = ty * p.tile_h, n0 = n_t * BLOCK_N;
        stamp(0, (t - unit0) / unit_stride);
        int kbi = 0;
        for (int tap = 0; tap < p.taps; ++tap) {
          int dy = 0, dx = 0;
          if (p.ksize == 3) {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
            if (p.flip) {
              dy = -dy;
              dx = -dx;
            }
          }
          int src = 0, c0 = 0;
          for (int kc = 0; kc < kchunks; ++kc, ++kbi) {
            if (t == unit0 && kbi < 16) stamp(3, 2 * kbi);
            mbar_wait(empty_bar(stage), phase ^ 1u);
            if (elect_one()) {
              if constexpr (TWO) {
                // the leader's barrier collects the bytes of BOTH CTAs' loads
                if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
                tma_load_4d_2cta(smem_a(stage), &p.tmap_a[src], full_bar(stage), c0, x0 + dx, y0 + dy, b);
                tma_load_2d_2cta(smem_b(stage), &p.tmap_w, full_bar(stage), kc * 64,
                                 tap * p.Cout + n0 + rank * Cfg::B_ROWS);
              } else {
                mbar_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                tma_load_4d(smem_a(stage), &p.tmap_a[src], full_bar(stage), c0, x0 + dx, y0 + dy, b);
                tma_load_2d(smem_b(stage), &p.tmap_w, full_bar(stage), kc * 64, tap * p.Cout + n0);
              }
            }
            __syncwarp();
            if (t == unit0 && kbi < 16) stamp(3, 2 * kbi + 1);
            c0 += 64;
            if (c0 == p.kchunks_per_src * 64) {
              c0 = 0;
              ++src;
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // The whole warp runs the loop (convergent: barrier waits by all lanes, operands in uniform registers); one
    // elected lane issues the MMAs and their commits.  Written as `if (lane == 0) { loop }` the compiler spends
    // ~20 instructions (R2UR broadcasts inside an election loop, descriptor rebuilds) between two MMAs: measured
    // ~150 cycles per MMA whatever its shape, i.e. above the 128-cycle floor of even the widest one.
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(TWO ? 256 : 128, BLOCK_N, 0, 0);
      constexpr uint32_t d_hi = smem_desc_hi_sw128(1024);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = unit0; t < total_tiles; t += unit_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        stamp(1, 2 * it);
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(smem_a(stage), 16);
          const uint32_t b_lo = smem_desc_lo(smem_b(stage), 16);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // (+32 bytes of K = +2 in the descriptor's address field)
              umma_bf16_lh<TWO>(d_tmem, a_lo + 2 * k, d_hi, b_lo + 2 * k, d_hi, idesc, (kb | k) != 0 ? 1u : 0u);
            if constexpr (TWO) umma_commit_2cta(empty_bar(stage));  // frees the stage in both CTAs
            else umma_commit(empty_bar(stage));
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (elect_one()) {
          if constexpr (TWO) umma_commit_2cta(tfull_bar(acc));
          else umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        stamp(1, 2 * it + 1);
      }
    }
  } else {
    // ===================================================== epilogue (warps 2..9)
    const int quarter = warp & 3;
    const int team = TEAMS ? (warp - 2) >> 2 : 0;   // which accumulator buffer this team drains
    const int half0 = TEAMS ? 0 : (warp - 2) >> 2;  // first 32-column half of every 64-column chunk this warp owns
    constexpr int NHALF = TEAMS ? 2 : 1;            // a team's warp covers both halves
    const int row = quarter * 32 + lane;  // tile row == TMEM lane == smem staging row
    const int ly = row / p.tile_w, lx = row - ly * p.tile_w;
    const bool use_tma = (BLOCK_N >= 64) && p.tma_store != 0;
    const bool issuer = TEAMS ? (((warp - 2) & 3) == 0 && lane == 0) : (warp == 2 && lane == 0);
    const int etid = threadIdx.x - 64;    // 0..255
    auto epi_sync = [&] {
      if constexpr (TEAMS) asm volatile("bar.sync %0, 128;" ::"r"(1 + team) : "memory");
      else asm volatile("bar.sync 1, 256;" ::: "memory");
    };
    constexpr int NSLOT = TEAMS ? Cfg::NBUF / 2 : Cfg::NBUF;  // staging slots per team
    const uint32_t my_store_base = store_base + (TEAMS ? team * NSLOT * 16384 : 0);
    uint32_t store_iter = 0;
    int it = 0;
    int bias_n0 = -1;
    // Which optional epilogue stages are on, in a REGISTER the optimizer cannot see through: tested straight from
    // the kernel parameters every stage costs a dependent constant-bank load + uniform compare + branch (~60
    // cycles each, ~8 of them per 32-column chunk even when every stage is off).
    enum : uint32_t { F_ACT = 3u, F_MASK_SIGN = 4u, F_MASK_DGELU = 8u, F_RES = 16u, F_RES32 = 32u, F_OUT32 = 64u,
                      F_COLSUM = 128u, F_SHUFFLE = 256u, F_MASK_MUL = 512u, F_AUX_GRAD = 1024u, F_LN_IN = 2048u, F_LN_OUT = 4096u,
                      F_MASK = F_MASK_SIGN | F_MASK_DGELU | F_MASK_MUL, F_TAIL = F_MASK | F_RES | F_RES32 | F_OUT32 };
    uint32_t ef = static_cast<uint32_t>(p.act) & F_ACT;
    if (p.mask_mode == SRB200_MASK_SIGN) ef |= F_MASK_SIGN;
    else if (p.mask_mode == SRB200_MASK_MUL) ef |= F_MASK_MUL;
    else if (p.mask_mode != SRB200_MASK_NONE) ef |= F_MASK_DGELU;
    if (p.aux_mode == 1) ef |= F_AUX_GRAD;
    if (p.residual != nullptr) ef |= F_RES;
    if (p.residual_f32 != nullptr) ef |= F_RES32;
    if (p.out_f32 != nullptr) ef |= F_OUT32;
    if (p.colsum != nullptr) ef |= F_COLSUM;
    if (p.out_mode == SRB200_OUT_SHUFFLE) ef |= F_SHUFFLE;
    if constexpr (LN) {
      if (p.ln_in != nullptr) ef |= F_LN_IN;
      if (p.ln_out != nullptr) ef |= F_LN_OUT;
    }
    asm volatile("mov.b32 %0, %0;" : "+r"(ef));
    // bias-gradient column sums of this warp's (32 rows x 32 columns) of every 64-column chunk; kept in
    // registers across the CTA's tiles when there is a single N tile, flushed with one 128-byte red per chunk
    constexpr int NCH = BLOCK_N >= 64 ? BLOCK_N / 64 : 1;
    float csum[NCH * NHALF];
#pragma unroll
    for (int q = 0; q < NCH * NHALF; ++q) csum[q] = 0.0f;
    int csum_n0 = -1, csum_b = 0;
    auto flush_colsum = [&]() {
      if (csum_n0 < 0) return;
      float* dst = p.colsum + (p.colsum_per_image ? static_cast<size_t>(csum_b) * p.Cout : 0) + csum_n0;
#pragma unroll
      for (int q = 0; q < NCH * NHALF; ++q) {
        if (csum_b < p.B)
          atomicAdd(dst + (q / NHALF) * 64 + (half0 + q % NHALF) * 32 + lane, csum[q] * p.colsum_scale);
        csum[q] = 0.0f;
      }
      csum_n0 = -1;
    };
    // RES: this team's staged-operand chunks, numbered g = 0, 1, ... over (tile, 64-column chunk); chunk g lives in slot
    // g & 1 and is requested one chunk ahead by the team's issuer thread
    const uint32_t my_res_base = res_base + team * 2 * 16384;
    uint32_t res_g = 0;
    auto res_issue = [&](int t, int cc, uint32_t g) {  // (issuer thread only)
      const int t_div = fast_div(t, p.magic_n);
      const int n_t = t - t_div * p.n_tiles;
      const int b = fast_div(t_div, p.magic_xy);
      const int m_in = t_div - b * (p.tiles_x * p.tiles_y);
      const int ty = fast_div(m_in, p.magic_x);
      const int tx = m_in - ty * p.tiles_x;
      const uint32_t bar = rfull_bar(team, g & 1u);
      mbar_expect_tx(bar, 16384);
      tma_load_4d(my_res_base + (g & 1u) * 16384, &p.tmap_res, bar, n_t * BLOCK_N + cc * 64, tx * p.tile_w,
                  ty * p.tile_h, b);
    };
    if constexpr (RES) {
      const int t_first = unit0 + team * unit_stride;
      if (issuer && t_first < total_tiles) res_issue(t_first, 0, 0u);
    }
    if constexpr (TEAMS) {
      // the launch grid is a multiple of n_tiles (host-checked): one N tile, one bias vector per CTA, loaded once
      // by all eight warps -- the teams never meet again
      const int n00 = (unit0 - fast_div(unit0, p.magic_n) * p.n_tiles) * BLOCK_N;
      for (int i = etid; i < BLOCK_N; i += 256) s_bias[i] = p.bias != nullptr ? __ldg(p.bias + n00 + i) : 0.0f;
      asm volatile("bar.sync 3, 256;" ::: "memory");
      bias_n0 = n00;
    }
    for (int t = unit0; t < total_tiles; t += unit_stride, ++it) {
      if (TEAMS && (it & 1) != team) continue;
      const int t_div = fast_div(t, p.magic_n);
      const int n_t = t - t_div * p.n_tiles;
      const int m_t = TWO ? 2 * t_div + rank : t_div;
      const int b = fast_div(m_t, p.magic_xy);  // == B for the missing half of an odd last pair: zero fill
      const int m_in = m_t - b * (p.tiles_x * p.tiles_y);
      const int ty = fast_div(m_in, p.magic_x);
      const int tx = m_in - ty * p.tiles_x;
      const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
      const int x = x0 + lx, y = y0 + ly;
      const int n0 = n_t * BLOCK_N;
      const bool valid = (x < p.W) && (y < p.H) && (b < p.B);
      if (p.colsum != nullptr) {
        if (csum_n0 != n0 || (p.colsum_per_image && csum_b != b)) flush_colsum();
        csum_n0 = n0;
        csum_b = p.colsum_per_image ? b : 0;
      }
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      // folded LayerNorm of the input rows: this row's rstd and -rstd * mean (requested before the accumulator wait)
      float ln_r = 1.0f, ln_nrm = 0.0f;
      if constexpr (LN) {
        if ((ef & F_LN_IN) && valid) {
          const float2 st = __ldg(p.ln_in + (static_cast<size_t>(b) * p.H + y) * p.W + x);
          ln_r = st.y;
          ln_nrm = -st.y * st.x;
        }
      }
      float ln_s = 0.0f, ln_ss = 0.0f;  // row sum / sum of squares of the stored (bf16-rounded) values
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (issuer) stamp(2, 2 * it);
      const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(quarter * 32) << 16);
      const size_t pix = (static_cast<size_t>(b) * p.H + y) * p.W + x;
      const float ab = (p.alpha_b != nullptr && b < p.B) ? __ldg(p.alpha_b + b) : 1.0f;
      // (SRB200_EXT_ALPHA_ON_BIAS: the input already carries alpha[b]; only the bias still needs it)
      const float al = p.alpha_on_bias ? p.alpha : p.alpha * ab;
      const float al_bias = p.alpha_on_bias ? ab : 1.0f;
      // bias of this N tile -> smem (every epilogue warp has passed the last barrier of the previous tile); the
      // launch grid is a multiple of n_tiles, so a CTA keeps its N tile for all its tiles and this runs once
      if (bias_n0 != n0) {
        if constexpr (TEAMS) __trap();  // N-tile affinity violated
        if (it > 0) epi_sync();  // every warp is done reading the previous tile's bias
        for (int i = etid; i < BLOCK_N; i += 256) s_bias[i] = p.bias != nullptr ? __ldg(p.bias + n0 + i) : 0.0f;
        epi_sync();
        bias_n0 = n0;
      }

      // element offset of (this thread's pixel, channel nc) in an NHWC tensor shaped like `out`
      auto out_offset = [&](int nc) -> size_t {
        if (ef & F_SHUFFLE) {
          const int r2 = p.out_r * p.out_r;
          const int C = p.Cout / r2;
          const int ij = nc / C, cc = nc - ij * C;
          const int i = ij / p.out_r, j = ij - i * p.out_r;
          return ((static_cast<size_t>(b) * p.H * p.out_r + static_cast<size_t>(y) * p.out_r + i) *
                      (static_cast<size_t>(p.W) * p.out_r) +
                  static_cast<size_t>(x) * p.out_r + j) *
                     C +
                 cc;
        }
        return pix * p.Cout + nc;
      };
      // v: accumulators of CHUNK channels starting at nc -> final values (everything between MMA and store)
      // (staged: shared-memory address of this thread's 64-byte piece of the staged operand, 0 = read global memory)
      auto finish = [&](float (&v)[CHUNK], size_t off, uint32_t staged = 0u, int piece0 = 0) {
        auto ld_staged = [&](int j) {
          uint4 m;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w)
                       : "r"(staged + (((piece0 + j) ^ (row & 7)) << 4)));
          return m;
        };
        const uint32_t act = ef & F_ACT;
        if (act != SRB200_ACT_NONE && !(ef & F_AUX_GRAD)) {  // (with F_AUX_GRAD the activation is already applied)
          if (act == SRB200_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) v[j] = fmaxf(v[j], 0.0f);
          } else if (act == SRB200_ACT_LRELU) {
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) v[j] = v[j] > 0.0f ? v[j] : v[j] * p.act_slope;
          } else {
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) v[j] = gelu_erf(v[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) v[j] *= al;
        if (!valid || (ef & F_TAIL) == 0) return;  // out-of-image rows of a partial tile: never stored / loaded
        if (ef & F_MASK) {
          // (the mode test sits OUTSIDE the element loops: tested per element pair, the three-way branch -- one arm of it
          // the long GELU-derivative code -- made this stage 2.2k cycles per 64-column chunk instead of ~0.6k)
          const uint4* mp = reinterpret_cast<const uint4*>(p.mask_src + off);
          uint4 mm[CHUNK / 8];
#pragma unroll
          for (int j = 0; j < CHUNK / 8; ++j) mm[j] = (RES && p.res_kind == 1) ? ld_staged(j) : __ldg(mp + j);
          if (ef & F_MASK_MUL) {
#pragma unroll
            for (int j = 0; j < CHUNK / 8; ++j) {
              const uint32_t mw[4] = {mm[j].x, mm[j].y, mm[j].z, mm[j].w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[8 * j + 2 * q + 0] *= bf16_lo(mw[q]);
                v[8 * j + 2 * q + 1] *= bf16_hi(mw[q]);
              }
            }
          } else if (ef & F_MASK_SIGN) {
#pragma unroll
            for (int j = 0; j < CHUNK / 8; ++j) {
              const uint32_t mw[4] = {mm[j].x, mm[j].y, mm[j].z, mm[j].w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[8 * j + 2 * q + 0] *= (bf16_lo(mw[q]) > 0.0f) ? 1.0f : p.mask_slope;
                v[8 * j + 2 * q + 1] *= (bf16_hi(mw[q]) > 0.0f) ? 1.0f : p.mask_slope;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < CHUNK / 8; ++j) {
              const uint32_t mw[4] = {mm[j].x, mm[j].y, mm[j].z, mm[j].w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[8 * j + 2 * q + 0] *= dgelu_erf(bf16_lo(mw[q]));
                v[8 * j + 2 * q + 1] *= dgelu_erf(bf16_hi(mw[q]));
              }
            }
          }
        }
        if (ef & F_RES) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + off);
#pragma unroll
          for (int j = 0; j < CHUNK / 8; ++j) {
            const uint4 m = (RES && p.res_kind == 2) ? ld_staged(j) : __ldg(rp + j);
            const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v[8 * j + 2 * q + 0] += bf16_lo(mw[q]);
              v[8 * j + 2 * q + 1] += bf16_hi(mw[q]);
            }
          }
        }
        if (ef & F_RES32) {
          const float4* rp = reinterpret_cast<const float4*>(p.residual_f32 + off);
#pragma unroll
          for (int j = 0; j < CHUNK / 4; ++j) {
            const float4 m = __ldg(rp + j);
            v[4 * j + 0] += m.x;
            v[4 * j + 1] += m.y;
            v[4 * j + 2] += m.z;
            v[4 * j + 3] += m.w;
          }
        }
        if (ef & F_OUT32) {
          float4* fp = reinterpret_cast<float4*>(p.out_f32 + off);
#pragma unroll
          for (int j = 0; j < CHUNK / 4; ++j)
            fp[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      };
      auto load_acc = [&](int col, float (&v)[CHUNK]) {
        uint32_t r[CHUNK];
        if constexpr (CHUNK == 32) {
          tmem_ld32(taddr + col, r);
        } else {
          tmem_ld16(taddr + col, r);
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CHUNK; ++j) v[j] = __uint_as_float(r[j]);
        const float4* bp = reinterpret_cast<const float4*>(s_bias + col);
        if (LN && (ef & F_LN_IN)) {
          // LayerNorm folded into this GEMM: acc = sum_c W'[n,c] x[c] with W' = W diag(gamma) over the RAW rows, so
          // W LN(x) + b = rstd * (acc - mean * sum_c W'[n,c]) + (W beta + b): two FMAs per element, no LN pass
          const float4* wp = reinterpret_cast<const float4*>(p.ln_wsum + n0 + col);
#pragma unroll
          for (int j = 0; j < CHUNK / 4; ++j) {
            const float4 bv = bp[j];
            const float4 wv = __ldg(wp + j);
            v[4 * j + 0] = fmaf(v[4 * j + 0], ln_r, fmaf(ln_nrm, wv.x, bv.x));
            v[4 * j + 1] = fmaf(v[4 * j + 1], ln_r, fmaf(ln_nrm, wv.y, bv.y));
            v[4 * j + 2] = fmaf(v[4 * j + 2], ln_r, fmaf(ln_nrm, wv.z, bv.z));
            v[4 * j + 3] = fmaf(v[4 * j + 3], ln_r, fmaf(ln_nrm, wv.w, bv.w));
          }
          return;
        }
#pragma unroll
        for (int j = 0; j < CHUNK / 4; ++j) {
          const float4 bv = bp[j];
          v[4 * j + 0] = fmaf(al_bias, bv.x, v[4 * j + 0]);  // (al_bias == 1: exactly v + bias)
          v[4 * j + 1] = fmaf(al_bias, bv.y, v[4 * j + 1]);
          v[4 * j + 2] = fmaf(al_bias, bv.z, v[4 * j + 2]);
          v[4 * j + 3] = fmaf(al_bias, bv.w, v[4 * j + 3]);
        }
      };

      if (use_tma) {
        // -------- results leave through swizzled smem staging + TMA tensor stores (full 128-byte lines,
        // partial tiles clipped by the tensor map); the store of chunk c overlaps the math of chunk c+1
        if constexpr (BLOCK_N >= 64) {
          const bool has_aux = p.aux_out != nullptr;
#pragma unroll 1
          for (int cc = 0; cc < BLOCK_N / 64; ++cc) {
            uint32_t buf_out, buf_aux = 0;
            if (has_aux) {
              buf_aux = my_store_base;
              buf_out = my_store_base + 16384;
              if (issuer) tma_store_wait_read<0>();
            } else {
              buf_out = my_store_base + (store_iter % NSLOT) * 16384;
              if (issuer) tma_store_wait_read<NSLOT - 1>();
            }
            const bool tr = issuer && it == 2 && cc < 4;  // debug timeline of team 0's second tile: 7 stamps per chunk
            if (tr) stamp(3, 32 + 7 * cc);
            epi_sync();
            if (tr) stamp(3, 33 + 7 * cc);
            uint32_t staged_row = 0u;
            if constexpr (RES) {
              // request the NEXT chunk of the staged operand (its slot was last read two chunks ago, before the closing
              // barrier of that chunk), then wait for this one
              if (issuer) {
                if (cc + 1 < BLOCK_N / 64) res_issue(t, cc + 1, res_g + 1);
                else if (t + 2 * unit_stride < total_tiles) res_issue(t + 2 * unit_stride, 0, res_g + 1);
              }
              mbar_wait(rfull_bar(team, res_g & 1u), (res_g >> 1) & 1u);
              staged_row = my_res_base + (res_g & 1u) * 16384 + row * 128;
              ++res_g;
            }
            if (tr) stamp(3, 34 + 7 * cc);
#pragma unroll 1
            for (int hh = 0; hh < NHALF; ++hh) {
              const int hf = half0 + hh;
              float v[CHUNK];
              const int col = cc * 64 + hf * 32;
              load_acc(col, v);
              if (has_aux) {
                float a[CHUNK];
                if (ef & F_AUX_GRAD) {
#pragma unroll
                  for (int j = 0; j < CHUNK; ++j) gelu_and_grad(v[j], v[j], a[j]);  // v <- gelu, aux <- gelu'
                } else {
#pragma unroll
                  for (int j = 0; j < CHUNK; ++j) a[j] = v[j];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const uint32_t dst = buf_aux + row * 128 + (((hf * 4 + j) ^ (row & 7)) << 4);
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                               "r"(pack_bf16x2(a[8 * j + 0], a[8 * j + 1])),
                               "r"(pack_bf16x2(a[8 * j + 2], a[8 * j + 3])),
                               "r"(pack_bf16x2(a[8 * j + 4], a[8 * j + 5])),
                               "r"(pack_bf16x2(a[8 * j + 6], a[8 * j + 7]))
                               : "memory");
                }
              }
              finish(v, out_offset(n0 + col), staged_row, hf * 4);
              if constexpr (CHUNK == 32) {
                if (ef & F_COLSUM) {
                  float w[32];
#pragma unroll
                  for (int j = 0; j < 32; ++j) w[j] = valid ? v[j] : 0.0f;
                  const float cs = warp_colsum32(w, lane);
#pragma unroll
                  for (int q = 0; q < NCH * NHALF; ++q)
                    if (q == cc * NHALF + hh) csum[q] += cs;
                }
              }
              if (LN && (ef & F_LN_OUT)) {
                // row statistics of what is stored (the bf16-rounded values, as a LayerNorm kernel would read them)
#pragma unroll
                for (int j = 0; j < CHUNK / 2; ++j) {
                  const uint32_t pk = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                  const float lo = bf16_lo(pk), hi = bf16_hi(pk);
                  ln_s += lo + hi;
                  ln_ss = fmaf(lo, lo, fmaf(hi, hi, ln_ss));
                }
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t dst = buf_out + row * 128 + (((hf * 4 + j) ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                             "r"(pack_bf16x2(v[8 * j + 0], v[8 * j + 1])),
                             "r"(pack_bf16x2(v[8 * j + 2], v[8 * j + 3])),
                             "r"(pack_bf16x2(v[8 * j + 4], v[8 * j + 5])),
                             "r"(pack_bf16x2(v[8 * j + 6], v[8 * j + 7]))
                             : "memory");
              }
            }
            if (tr) stamp(3, 35 + 7 * cc);
            fence_proxy_async_smem();
            epi_sync();
            if (tr) stamp(3, 36 + 7 * cc);
            if (issuer) {
              const int nc = n0 + cc * 64;
              int view = 0, c0 = nc;
              if (p.out_mode == SRB200_OUT_SHUFFLE) {
                const int C = p.Cout / (p.out_r * p.out_r);
                view = nc / C;
                c0 = nc - view * C;
              }
              if (has_aux) tma_store_4d(&p.tmap_aux, buf_aux, c0, x0, y0, b);
              tma_store_4d(&p.tmap_out[view], buf_out, c0, x0, y0, b);
              tma_store_commit();
            }
            ++store_iter;
          }
          if constexpr (TEAMS && LN) {
            // (a two-team warp covers both halves of every chunk: the thread has seen its whole row; host-checked: one N tile)
            if ((ef & F_LN_OUT) && valid) {
              const float mean = ln_s * p.ln_inv_c;
              const float var = fmaxf(fmaf(ln_ss, p.ln_inv_c, -mean * mean), 0.0f);
              p.ln_out[pix] = make_float2(mean, 1.0f / sqrtf(var + p.ln_eps));
            }
          }
        }
      } else {
        // -------- direct stores: final NCHW fp32 image (conv_last), or N < 64
#pragma unroll 1
        for (int c = half0; c < BLOCK_N / CHUNK; c += (TEAMS ? 1 : 2)) {
          float v[CHUNK];
          load_acc(c * CHUNK, v);
          const int nc = n0 + c * CHUNK;
          const size_t off = out_offset(nc);
          if (valid && p.aux_out != nullptr) {
            uint4* ap = reinterpret_cast<uint4*>(p.aux_out + off);
#pragma unroll
            for (int j = 0; j < CHUNK / 8; ++j)
              ap[j] = make_uint4(pack_bf16x2(v[8 * j + 0], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
          finish(v, off);
          if (!valid) continue;
          if (p.out_mode == SRB200_OUT_NCHW_F32) {
            float* op = reinterpret_cast<float*>(p.out);
            const size_t plane = static_cast<size_t>(p.H) * p.W;
#pragma unroll
            for (int j = 0; j < CHUNK; ++j) {
              const int ch = nc + j;
              if (ch < p.out_c) {
                const float sh = p.out_shift != nullptr ? __ldg(p.out_shift + ch) : 0.0f;
                op[(static_cast<size_t>(b) * p.out_c + ch) * plane + static_cast<size_t>(y) * p.W + x] =
                    v[j] * p.out_scale + sh;
              }
            }
          } else {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
#pragma unroll
            for (int j = 0; j < CHUNK / 8; ++j)
              op[j] = make_uint4(pack_bf16x2(v[8 * j + 0], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (TWO) mbar_arrive_cluster(tempty_bar(acc), 0);  // the leader CTA issues the MMAs
        else mbar_arrive(tempty_bar(acc));
      }
      if (issuer) stamp(2, 2 * it + 1);
    }
    if (p.colsum != nullptr) flush_colsum();
    if (use_tma && issuer) tma_store_wait_all<0>();  // smem must outlive the last bulk stores
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (TWO) {
    cluster_sync_all();  // the peer may still be signalling our barriers / the pair's TMEM is released together
    if (warp == 2) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
  } else {
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
  if (p.trace != nullptr && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[256 + 4 * blockIdx.x + 1] = gt;
    p.trace[256 + 4 * blockIdx.x + 3] = clock64();
  }
}

template <int BLOCK_N>
static int launch_tapgemm2(const TapGemmParams& p, cudaStream_t stream) {
  using Cfg = TapCfg<BLOCK_N, true>;
  static PerDeviceOnce configured;
  if (configured.ensure(tapgemm_kernel<BLOCK_N, true>, Cfg::SMEM_BYTES) != SRB200_OK) return SRB200_ELAUNCH;
  const int units = ((p.m_tiles + 1) / 2) * p.n_tiles;
  int pairs = units < num_sms() / 2 ? units : num_sms() / 2;
  if (pairs >= p.n_tiles) pairs -= pairs % p.n_tiles;  // a CTA pair keeps its N tile for all its work units
  return launch_ex(tapgemm_kernel<BLOCK_N, true>, 2 * pairs, 320, Cfg::SMEM_BYTES, stream, 2, p);
}

template <int BLOCK_N, bool TEAMS = false, bool RES = false, bool LN = false>
static int launch_tapgemm(const TapGemmParams& p, cudaStream_t stream) {
  using Cfg = TapCfg<BLOCK_N, false, RES>;
  static PerDeviceOnce configured;
  if (configured.ensure(tapgemm_kernel<BLOCK_N, false, TEAMS, RES, LN>, Cfg::SMEM_BYTES) != SRB200_OK)
    return SRB200_ELAUNCH;
  const int total = p.m_tiles * p.n_tiles;
  int grid = total < num_sms() ? total : num_sms();
  if (grid >= p.n_tiles) grid -= grid % p.n_tiles;  // a CTA keeps its N tile (bias, weight columns) for all its tiles
  if (TEAMS && grid % p.n_tiles != 0) return SRB200_EINVAL;
  return launch_ex(tapgemm_kernel<BLOCK_N, false, TEAMS, RES, LN>, grid, 320, Cfg::SMEM_BYTES, stream, 1, p);
}

static std::atomic<unsigned long long*> g_trace{nullptr};

static int pick_block_n(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout % 192 == 0) return 192;
  if (Cout % 128 == 0) return 128;
  if (Cout % 64 == 0) return 64;
  if (Cout == 16 || Cout == 32 || Cout == 48) return 16;
  return 0;
}

}  // namespace srb

using namespace srb;

extern "C" int srb200_tapgemm(const srb200_tapgemm_desc* d, const void* in_bf16,
                              const void* w_packed, const float* bias, const void* mask_src,
                              const void* residual, const float* out_shift, void* out,
                              void* aux_out, const srb200_tapgemm_ext* ext,
                              srb200_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d || !in_bf16 || !w_packed || !out) return SRB200_EINVAL;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0) return SRB200_EINVAL;
  if (d->Cin <= 0 || d->Cin % 64 != 0) return SRB200_EINVAL;
  if (d->ksize != 1 && d->ksize != 3) return SRB200_EINVAL;
  if (d->src_r < 1 || d->src_r > 3) return SRB200_EINVAL;
  const int bn = pick_block_n(d->Cout);
  if (bn == 0) return SRB200_EINVAL;
  if (d->mask_mode != SRB200_MASK_NONE && !mask_src) return SRB200_EINVAL;
  if (d->out_mode == SRB200_OUT_SHUFFLE) {
    const int r2 = d->out_r * d->out_r;
    if (d->out_r < 2 || d->out_r > 3 || d->Cout % r2 != 0 || (d->Cout / r2) % 32 != 0)
      return SRB200_EINVAL;
  } else if (d->out_mode == SRB200_OUT_NCHW_F32) {
    if (d->out_c <= 0 || d->out_c > d->Cout || mask_src || residual || aux_out)
      return SRB200_EINVAL;
  } else if (d->out_mode != SRB200_OUT_NHWC) {
    return SRB200_EINVAL;
  }
  if ((reinterpret_cast<uintptr_t>(in_bf16) | reinterpret_cast<uintptr_t>(w_packed) |
       reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(mask_src) |
       reinterpret_cast<uintptr_t>(residual) | reinterpret_cast<uintptr_t>(aux_out) |
       reinterpret_cast<uintptr_t>(bias)) &
      15u)
    return SRB200_EINVAL;

  TapGemmParams p;
  p.B = d->B;
  p.H = d->H;
  p.W = d->W;
  pick_tile(d->H, d->W, 128, &p.tile_w, &p.tile_h);
  p.tiles_x = (d->W + p.tile_w - 1) / p.tile_w;
  p.tiles_y = (d->H + p.tile_h - 1) / p.tile_h;
  p.m_tiles = d->B * p.tiles_x * p.tiles_y;
  p.n_tiles = d->Cout / bn;
  auto magic = [](long long dv) { return static_cast<unsigned long long>(((1ULL << 48) + dv - 1) / dv); };
  p.magic_n = magic(p.n_tiles);
  p.magic_x = magic(p.tiles_x);
  p.magic_xy = magic(static_cast<long long>(p.tiles_x) * p.tiles_y);
  if (static_cast<long long>(p.m_tiles) * p.n_tiles >= (1LL << 24)) return SRB200_EINVAL;
  p.kchunks_per_src = d->Cin / 64;
  p.num_src = d->src_r * d->src_r;
  p.Cout = d->Cout;
  p.ksize = d->ksize;
  p.taps = d->ksize * d->ksize;
  p.flip = d->flip;
  p.bias = bias;
  p.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
  p.residual = static_cast<const __nv_bfloat16*>(residual);
  p.out_shift = out_shift;
  p.out = out;
  p.aux_out = static_cast<__nv_bfloat16*>(aux_out);
  p.residual_f32 = ext ? ext->residual_f32 : nullptr;
  p.out_f32 = ext ? ext->out_f32 : nullptr;
  p.alpha_b = ext ? ext->alpha_per_sample : nullptr;
  p.alpha_on_bias = (ext && (ext->flags & SRB200_EXT_ALPHA_ON_BIAS) && p.alpha_b != nullptr) ? 1 : 0;
  p.colsum = ext ? ext->colsum : nullptr;
  p.aux_mode = ext ? ext->aux_mode : 0;
  p.colsum_per_image = ext ? ext->colsum_per_image : 0;
  p.colsum_scale = (ext && ext->colsum_scale != 0.0f) ? ext->colsum_scale : 1.0f;
  p.ln_in = ext ? reinterpret_cast<const float2*>(ext->ln_stats_in) : nullptr;
  p.ln_wsum = ext ? ext->ln_wsum : nullptr;
  p.ln_out = ext ? reinterpret_cast<float2*>(ext->ln_stats_out) : nullptr;
  p.ln_inv_c = (ext && ext->ln_channels > 0) ? 1.0f / static_cast<float>(ext->ln_channels) : 0.0f;
  p.ln_eps = ext ? ext->ln_eps : 0.0f;
  if ((p.ln_in != nullptr) != (p.ln_wsum != nullptr)) return SRB200_EINVAL;
  if (p.ln_in != nullptr && (p.alpha_b != nullptr || bn < 32 || d->out_mode == SRB200_OUT_NCHW_F32)) return SRB200_EINVAL;
  if ((reinterpret_cast<uintptr_t>(p.ln_in) | reinterpret_cast<uintptr_t>(p.ln_out)) & 7u) return SRB200_EINVAL;
  if (reinterpret_cast<uintptr_t>(p.ln_wsum) & 15u) return SRB200_EINVAL;
  if (p.ln_out != nullptr && (ext->ln_channels <= 0 || ext->ln_channels > d->Cout || d->Cout != bn ||
                              d->out_mode != SRB200_OUT_NHWC || aux_out != nullptr))
    return SRB200_EINVAL;
  if (p.aux_mode != 0 && (p.aux_mode != 1 || d->act != SRB200_ACT_GELU || !aux_out || bn < 64 ||
                          d->out_mode != SRB200_OUT_NHWC))
    return SRB200_EINVAL;
  if (p.colsum && (bn < 64 || d->out_mode != SRB200_OUT_NHWC)) return SRB200_EINVAL;
  if ((p.residual_f32 || p.out_f32) && d->out_mode == SRB200_OUT_NCHW_F32) return SRB200_EINVAL;
  if ((reinterpret_cast<uintptr_t>(p.residual_f32) | reinterpret_cast<uintptr_t>(p.out_f32)) & 15u)
    return SRB200_EINVAL;
  p.act = d->act;
  p.act_slope = d->act_slope;
  p.alpha = d->alpha;
  p.mask_mode = d->mask_mode;
  p.mask_slope = d->mask_slope;
  p.out_mode = d->out_mode;
  p.out_r = d->out_r;
  p.out_c = d->out_c;
  p.out_scale = d->out_scale;
  p.trace = g_trace.load(std::memory_order_relaxed);
  p.pdl = pdl_enabled() ? 1 : 0;
  p.res_kind = 0;

  // A views: in[B, H*r, W*r, Cin], view (i,j): element (b,y,x,c) at ((b*H*r + y*r+i)*W*r + x*r+j)*Cin + c
  const int r = d->src_r;
  const uint64_t C = static_cast<uint64_t>(d->Cin);
  const uint64_t Wf = static_cast<uint64_t>(d->W) * r, Hf = static_cast<uint64_t>(d->H) * r;
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < r; ++j) {
      const __nv_bfloat16* base =
          static_cast<const __nv_bfloat16*>(in_bf16) + (static_cast<uint64_t>(i) * Wf + j) * C;
      const uint64_t dims[4] = {C, static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                                static_cast<uint64_t>(d->B)};
      const uint64_t strides[3] = {r * C * 2, r * Wf * C * 2, Hf * Wf * C * 2};
      const uint32_t box[4] = {64, static_cast<uint32_t>(p.tile_w), static_cast<uint32_t>(p.tile_h),
                               1};
      const int rc = make_tmap_bf16(&p.tmap_a[i * r + j], base, 4, dims, strides, box);
      if (rc != SRB200_OK) return rc;
    }
  // CTA pairs (cta_group::2) for the wide layers: each CTA loads half of the weight box
  const bool two_cta = (bn == 256) && p.m_tiles >= 2 && SRB_ENV("SRB_TAPGEMM_1CTA") == nullptr;
  {
    const uint64_t K = static_cast<uint64_t>(p.num_src) * C;
    const uint64_t dims[2] = {K, static_cast<uint64_t>(p.taps) * d->Cout};
    const uint64_t strides[1] = {K * 2};
    const uint32_t box[2] = {64, static_cast<uint32_t>(two_cta ? bn / 2 : bn)};
    const int rc = make_tmap_bf16(&p.tmap_w, w_packed, 2, dims, strides, box);
    if (rc != SRB200_OK) return rc;
  }
  // output views for the TMA-store epilogue: NHWC [B,H,W,Cout], or the out_r^2 phases of [B,H*r,W*r,Cout/r^2]
  p.tma_store = (bn >= 64 && d->out_mode != SRB200_OUT_NCHW_F32) ? 1 : 0;
  if (p.tma_store) {
    const int ro = d->out_mode == SRB200_OUT_SHUFFLE ? d->out_r : 1;
    const uint64_t Co = static_cast<uint64_t>(d->Cout) / (ro * ro);
    const uint64_t Wo = static_cast<uint64_t>(d->W) * ro, Ho = static_cast<uint64_t>(d->H) * ro;
    const uint64_t dims[4] = {Co, static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H),
                              static_cast<uint64_t>(d->B)};
    const uint64_t strides[3] = {ro * Co * 2, ro * Wo * Co * 2, Ho * Wo * Co * 2};
    const uint32_t box[4] = {64, static_cast<uint32_t>(p.tile_w), static_cast<uint32_t>(p.tile_h), 1};
    for (int i = 0; i < ro; ++i)
      for (int j = 0; j < ro; ++j) {
        const size_t eoff = (static_cast<uint64_t>(i) * Wo + j) * Co;
        int rc = make_tmap_bf16(&p.tmap_out[i * ro + j], static_cast<__nv_bfloat16*>(out) + eoff, 4, dims,
                                strides, box);
        if (rc != SRB200_OK) return rc;
        if (aux_out && i == 0 && j == 0) {
          rc = make_tmap_bf16(&p.tmap_aux, static_cast<__nv_bfloat16*>(aux_out), 4, dims, strides, box);
          if (rc != SRB200_OK) return rc;
        }
      }
    if (aux_out && ro != 1) return SRB200_EINVAL;
  }
  if (two_cta) return (p.ln_in != nullptr || p.ln_out != nullptr) ? SRB200_EINVAL : launch_tapgemm2<256>(p, stream);
  // small-K, wide-N layers (SwinIR's linears) are epilogue-bound: two epilogue teams (needs the TMA-store path and
  // N-tile affinity).  Not for 64-wide tiles: their 2-tiles-per-CTA kernels only see a longer last-tile epilogue.
  {
    const int total = p.m_tiles * p.n_tiles;
    int grid = total < num_sms() ? total : num_sms();
    if (grid >= p.n_tiles) grid -= grid % p.n_tiles;
    const bool teams = p.tma_store != 0 && grid % p.n_tiles == 0 && SRB_ENV("SRB_TAPGEMM_NO_TEAMS") == nullptr;
    const bool ln = p.ln_in != nullptr || p.ln_out != nullptr;
    // Linears whose epilogue reads a bf16 tensor shaped like the output (residual, or the mask / derivative of a data
    // gradient): that operand goes through shared memory (TapCfg RES)
    if (teams && bn == 192 && d->ksize == 1 && d->out_mode == SRB200_OUT_NHWC && (mask_src || residual) &&
        SRB_ENV("SRB_TAPGEMM_NO_RES") == nullptr) {
      p.res_kind = mask_src ? 1 : 2;
      const uint64_t Co = static_cast<uint64_t>(d->Cout);
      const uint64_t dims[4] = {Co, static_cast<uint64_t>(d->W), static_cast<uint64_t>(d->H), static_cast<uint64_t>(d->B)};
      const uint64_t strides[3] = {Co * 2, static_cast<uint64_t>(d->W) * Co * 2, static_cast<uint64_t>(d->H) * d->W * Co * 2};
      const uint32_t box[4] = {64, static_cast<uint32_t>(p.tile_w), static_cast<uint32_t>(p.tile_h), 1};
      const int rc = make_tmap_bf16(&p.tmap_res, mask_src ? mask_src : residual, 4, dims, strides, box);
      if (rc != SRB200_OK) return rc;
      if (ln) return launch_tapgemm<192, true, true, true>(p, stream);
      return launch_tapgemm<192, true, true>(p, stream);
    }
    if (teams && bn == 192) return ln ? launch_tapgemm<192, true, false, true>(p, stream) : launch_tapgemm<192, true>(p, stream);
    if (teams && bn == 128) return ln ? launch_tapgemm<128, true, false, true>(p, stream) : launch_tapgemm<128, true>(p, stream);
  }
  if (p.ln_in != nullptr || p.ln_out != nullptr)
    return SRB200_EINVAL;  // LayerNorm folding: two-team kernels (128- / 192-wide N tiles) only
  switch (bn) {
    case 256: return launch_tapgemm<256>(p, stream);
    case 192: return launch_tapgemm<192>(p, stream);
    case 128: return launch_tapgemm<128>(p, stream);
    case 64: return launch_tapgemm<64>(p, stream);
    case 16: return launch_tapgemm<16>(p, stream);
  }
  return SRB200_EINVAL;
}

extern "C" int srb200_set_pdl(int on) {
  return pdl_flag().exchange(on ? 1 : 0);
}

/* debug only: device buffer of 3*64 uint64 receiving CTA 0's per-role clock64 timeline of subsequent launches */
extern "C" int srb200_debug_set_trace(void* dev_buf) {
  g_trace.store(static_cast<unsigned long long*>(dev_buf), std::memory_order_relaxed);
  return SRB200_OK;
}
