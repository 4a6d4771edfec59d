// Backward of the fused shifted-window attention on tcgen05 / TMEM / TMA (sm_100a), window 8.
// (swinir_arch.py:144-175, 293-316 differentiated: dq, dk, dv into the un-partitioned NHWC gradient of qkv, and
// the gradient of relative_position_bias_table.)  Companion of attention_tc.cu -- same stage (two windows x one
// pair of heads), same TMA boxes and token order, same stacking tricks:
//
//   phase A, per head h (windows stacked to M = 128):   S = Q_h K_h^T      dP = dO_h V_h^T          (N = 128, K = 32)
//   threads (TMEM lane = query row, 16 columns each):   P = exp2(S*scale*log2e + bias + mask - lse)  (lse: forward's
//             statistics buffer -- no row max / sum here),  delta = rowsum(P o dP) (one 128-thread named barrier),
//             dS = P o (dP - delta);  bf16 P and dS tiles [(head, query)][key] per window, 128B-swizzled
//   phase B, per window w (heads stacked to M = 128):    dQ = dS K      (A = dS tile K-major,  B = K rows MN-major)
//                                                       dK = dS^T Q    (A = dS tile MN-major, B = Q rows MN-major)
//                                                       dV = P^T dO    (A = P tile MN-major,  B = dO rows MN-major)
//             rows of head h keep columns [32h, 32h+32) of the 64-wide results (N = 64, K = 64).
//   epilogue: dQ | dK | dV -> bf16 -> staging -> TMA stores (window_reverse + roll).
//   bias-table gradient: every thread owns fixed (query, 16 keys) cells, so it accumulates its dS in REGISTERS over all
//             windows of the CTA (a CTA works on ONE pair of heads); at the end the CTA bins them by relative position
//             in shared memory and adds its 2 x 225 sums to the global [225][heads] table.
// TMEM: S | dP of ONE head (256 columns) + dQ | dK | dV of ONE window (192 columns): phase A of the next head / stage
// overlaps phase B and the thread work.  Warp roles (800 threads): warps 0-15 = P / dS (lane quarter = warp & 3, column
// chunk = warp >> 2), 16-19 = epilogue, 21 = MMA issuer, 20 / 22 / 23 / 24 = TMA producers (one tensor each).
// Algorithmic traffic: read q, k, v, dO + write dq, dk, dv = 7 * T * C * 2 bytes.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "attention_tc.cuh"
#include "host_util.h"
#include "ptx.cuh"

namespace srb {

struct AttnBwdParams {
  CUtensorMap tmap_qkv4, tmap_qkv8;  // loads of q, k, v: [B,H,W,3*Ca], boxes (64 ch, 4, 4, 1) / (64 ch, 4, 8, 1)
  CUtensorMap tmap_do4, tmap_do8;    // loads of dO: [B,H,W,Ca]
  CUtensorMap tmap_gq4, tmap_gq8;    // stores of dq, dk, dv: [B,H,W,3*Ca]
  const float* table;                // [225][nH]
  const float* stats;                // [n_windows][nH][64] log2-domain log-sum-exp from the forward
  float* gtable;                     // [225][nH] (zeroed by the caller)
  int B, H, W, nH, Ca, shift;
  int nWx, nWy, n_windows, n_units;
  unsigned long long magic_per, magic_x, magic_units;
  float scale, scale2;               // head_dim^-0.5 and the same times log2(e)
  int pdl;
  unsigned long long* trace;
};

constexpr int kBwdThreads = 800;  // 16 P/dS warps, 4 epilogue warps, warp 21 = MMA issuer, warps 20, 22, 23, 24 = TMA producers
constexpr int kBwdStage = 4 * kSlab;                     // q, k, v, dO
constexpr int kBOffLoad = 0;                             // 2 stages
constexpr int kBOffTile = kBOffLoad + 2 * kBwdStage;     // P tiles (2 windows) | dS tiles (2 windows)
constexpr int kBOffOut = kBOffTile + 4 * kSlab;          // dq | dk | dv of ONE window: 3 x [64 tokens][64 ch] staging
constexpr int kBOffTab = kBOffOut + 3 * 8192;            // [2 heads of the current pair][15][40] fp32 bias * log2e
constexpr int kBOffPart = kBOffTab + 2 * kTabHead * 4;   // [2 buffers][16 warps][32 lanes] partial delta
constexpr int kBOffBar = kBOffPart + 2 * 16 * 32 * 4;
constexpr int kBwdSmem = kBOffBar + 256 + 1024;
// TMEM columns
constexpr uint32_t kTmS = 0, kTmDP = 128, kTmDQ = 256, kTmDK = 320, kTmDV = 384;

__global__ void __launch_bounds__(kBwdThreads, 1) window_attn_bwd_tc_kernel(const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kBOffBar;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };        // TMA -> MMA
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 + s); };  // MMA -> TMA (phase B of the stage retired)
  const uint32_t afull_bar = bar0 + 8u * 4;                     // MMA -> threads (S, dP of one head)
  const uint32_t afree_bar = bar0 + 8u * 5;                     // threads -> MMA (S, dP read out), 16 warps
  const uint32_t tfull_bar = bar0 + 8u * 6;                     // threads -> MMA (P, dS tiles of a stage written), 16 warps
  const uint32_t tfree_bar = bar0 + 8u * 7;                     // MMA -> threads (phase B of a stage retired)
  const uint32_t bfull_bar = bar0 + 8u * 8;                     // MMA -> epilogue (dQ, dK, dV of one window)
  const uint32_t bfree_bar = bar0 + 8u * 9;                     // epilogue -> MMA, 4 warps
  const uint32_t tmem_ptr_smem = bar0 + 8u * 10;
  float* s_tab = reinterpret_cast<float*>(gbase + kBOffTab);
  float* s_part = reinterpret_cast<float*>(gbase + kBOffPart);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  unsigned long long* trc = (p.trace != nullptr && blockIdx.x == 0) ? p.trace : nullptr;
  auto stamp = [&](int role, int it, int slot) {
    if (trc != nullptr && lane == 0 && it < 64) trc[(role * 64 + it) * 8 + slot] = clock64();
  };

  if (warp == 20 && lane == 0) {
    tma_prefetch_desc(&p.tmap_qkv4);
    tma_prefetch_desc(&p.tmap_qkv8);
    tma_prefetch_desc(&p.tmap_do4);
    tma_prefetch_desc(&p.tmap_do8);
    tma_prefetch_desc(&p.tmap_gq4);
    tma_prefetch_desc(&p.tmap_gq8);
  }
  if (warp == 21 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), 4);  // four producer warps, one tensor (q | k | v | dO) each
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(afull_bar, 1);
    mbar_init(afree_bar, 16);  // the 16 P/dS warps
    mbar_init(tfull_bar, 16);
    mbar_init(tfree_bar, 1);
    mbar_init(bfull_bar, 1);
    mbar_init(bfree_bar, 4);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.pdl) pdl_handoff();
  else pdl_trigger();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  auto wall = [&](int slot) {  // debug: per-CTA wall-clock stamps behind the role timeline of CTA 0
    if (p.trace != nullptr && threadIdx.x == 0) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      p.trace[4 * 64 * 8 + 4 * blockIdx.x + slot] = gt;
    }
  };
  wall(0);

  // A CTA works on ONE pair of heads for a strided subset of the window units: its bias-table gradient accumulators
  // then live in registers for the whole kernel and leave once, at the end (an earlier version walked all pairs per CTA
  // and paid ~14k cycles of reductions + table reload at every change of pair).
  const int n_pairs = p.nH >> 1;
  const int my_pair = static_cast<int>(blockIdx.x) % n_pairs;
  const int cta_in_pair = static_cast<int>(blockIdx.x) / n_pairs;
  const int ctas_of_pair = (static_cast<int>(gridDim.x) - my_pair + n_pairs - 1) / n_pairs;
  const int my_units = cta_in_pair < p.n_units ? (p.n_units - cta_in_pair + ctas_of_pair - 1) / ctas_of_pair : 0;
  const int n_iter = my_units;  // stage it = unit cta_in_pair + it * ctas_of_pair, pair my_pair
  auto window_of = [&](int unit, int w, int& b, int& wy, int& wx) -> bool {
    int idx = 2 * unit + w;
    const bool real = idx < p.n_windows;
    if (!real) idx = 2 * unit;
    b = fast_div(idx, p.magic_per);
    const int r = idx - b * (p.nWx * p.nWy);
    wy = fast_div(r, p.magic_x);
    wx = r - wy * p.nWx;
    return real;
  };
  auto is_quad = [&](int wy, int wx) { return p.shift > 0 && (wy == p.nWy - 1 || wx == p.nWx - 1); };
  // pixel origin of TMA box `quad` (quadrant (qx, qy) or x-half) of a window, and which tensor map moves it
  auto box_xy = [&](int wy, int wx, int quad, bool quadrants, int& x, int& y) {
    if (quadrants) {
      y = wy * 8 + (quad & 1) * 4 + p.shift;
      x = wx * 8 + (quad >> 1) * 4 + p.shift;
      if (y >= p.H) y -= p.H;
      if (x >= p.W) x -= p.W;
    } else {
      y = wy * 8 + p.shift;
      x = wx * 8 + quad * 4 + p.shift;
    }
  };

  if (warp == 20 || warp >= 22) {
    // ===================================================== TMA producers: one warp per tensor (q | k | v | dO), lane =
    // (window, box).  A TMA instruction costs its issuing warp ~150 cycles and a warp's lanes issue one after another
    // (16 loads from one warp took 2.4k cycles on the critical path slot-free -> data-landed); four warps issue in parallel.
    const int t = warp == 20 ? 0 : warp - 21;
    const int w = lane >> 2, quad = lane & 3;
    const int pair = my_pair;
    for (int it = 0; it < n_iter; ++it) {
      const int st = it & 1;
      const int unit = cta_in_pair + it * ctas_of_pair;
      if (t == 0) stamp(0, it, 0);
      mbar_wait(empty_bar(st), ((it >> 1) & 1u) ^ 1u);
      if (t == 0) stamp(0, it, 1);
      if (lane == 0) mbar_expect_tx(full_bar(st), kSlab);
      __syncwarp();
      if (lane < 8) {
        int b, wy, wx, x, y;
        window_of(unit, w, b, wy, wx);
        const bool quadrants = is_quad(wy, wx);
        const uint32_t dst = base + kBOffLoad + st * kBwdStage + t * kSlab + w * 8192;
        if (quadrants || quad < 2) {
          box_xy(wy, wx, quad, quadrants, x, y);
          const CUtensorMap* map = t < 3 ? (quadrants ? &p.tmap_qkv4 : &p.tmap_qkv8)
                                         : (quadrants ? &p.tmap_do4 : &p.tmap_do8);
          const int ch = (t < 3 ? t * p.Ca : 0) + pair * 64;
          tma_load_4d(dst + quad * (quadrants ? 2048 : 4096), map, full_bar(st), ch, x, y, b);
        }
      }
      __syncwarp();
      if (t == 0) stamp(0, it, 2);
    }
  } else if (warp == 21) {
    // ===================================================== MMA issuer
    constexpr uint32_t idesc_a = make_idesc_bf16(128, 128, 0, 0);   // S, dP: both operands K-major
    constexpr uint32_t idesc_q = make_idesc_bf16(128, 64, 0, 1);    // dQ: A = dS tile K-major, B = K rows MN-major
    constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, 1, 1);    // dK, dV: A = tile MN-major, B = rows MN-major
    constexpr uint32_t d_hi = smem_desc_hi_sw128(1024);
    auto phase_a = [&](int it, int h) {
      const int st = it & 1;
      const int c = 2 * it + h;
      mbar_wait(afree_bar, (c & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t slab = base + kBOffLoad + st * kBwdStage;
      const uint32_t q_lo = smem_desc_lo(slab, 16) + 4 * h, k_lo = smem_desc_lo(slab + kSlab, 16) + 4 * h;
      const uint32_t v_lo = smem_desc_lo(slab + 2 * kSlab, 16) + 4 * h, o_lo = smem_desc_lo(slab + 3 * kSlab, 16) + 4 * h;
      if (elect_one()) {
        umma_bf16_lh<false>(tmem_base + kTmS, q_lo, d_hi, k_lo, d_hi, idesc_a, 0u);
        umma_bf16_lh<false>(tmem_base + kTmS, q_lo + 2, d_hi, k_lo + 2, d_hi, idesc_a, 1u);
        umma_bf16_lh<false>(tmem_base + kTmDP, o_lo, d_hi, v_lo, d_hi, idesc_a, 0u);
        umma_bf16_lh<false>(tmem_base + kTmDP, o_lo + 2, d_hi, v_lo + 2, d_hi, idesc_a, 1u);
        umma_commit(afull_bar);
      }
      __syncwarp();
    };
    auto phase_b = [&](int j) {
      const int st = j & 1;
      mbar_wait(tfull_bar, j & 1u);
      tc_fence_after();
      stamp(1, j, 3);
      const uint32_t slab = base + kBOffLoad + st * kBwdStage;
#pragma unroll 1
      for (int w = 0; w < 2; ++w) {
        const int cw = 2 * j + w;
        mbar_wait(bfree_bar, (cw & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t pt = base + kBOffTile + w * kSlab, dst = base + kBOffTile + (2 + w) * kSlab;
        const uint32_t ds_k = smem_desc_lo(dst, 16);        // dS tile, K-major: rows (head, query), K = keys
        const uint32_t ds_m = smem_desc_lo(dst, 8192);      // dS tile, MN-major: K rows = queries, M blocks = heads
        const uint32_t p_m = smem_desc_lo(pt, 8192);        // P tile, MN-major
        const uint32_t q_r = smem_desc_lo(slab + w * 8192, 8192);
        const uint32_t k_r = smem_desc_lo(slab + kSlab + w * 8192, 8192);
        const uint32_t o_r = smem_desc_lo(slab + 3 * kSlab + w * 8192, 8192);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dQ = dS K: 16 keys = +32 B of a dS row (+2) / +16 rows of K (+128)
            umma_bf16_lh<false>(tmem_base + kTmDQ, ds_k + 2 * k, d_hi, k_r + 128 * k, d_hi, idesc_q, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dK = dS^T Q: 16 queries = +16 rows of the tile and of Q
            umma_bf16_lh<false>(tmem_base + kTmDK, ds_m + 128 * k, d_hi, q_r + 128 * k, d_hi, idesc_t, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // dV = P^T dO
            umma_bf16_lh<false>(tmem_base + kTmDV, p_m + 128 * k, d_hi, o_r + 128 * k, d_hi, idesc_t, k > 0 ? 1u : 0u);
          umma_commit(bfull_bar);
          if (w == 1) {
            umma_commit(tfree_bar);
            umma_commit(empty_bar(st));  // q, k, v, dO of the stage are dead: the producer may refill the slot
          }
        }
        __syncwarp();
      }
      stamp(1, j, 4);
    };
    // Issue order per stage: A(it, head 0) | B(it-1) | A(it, head 1).  A(it, 0) ahead of B(it-1) lets the threads start on
    // stage it while the tensor core is busy with phase B -- unless the loads of stage it have not landed yet: then
    // B(it-1) goes first (waiting for them would stall the very MMAs whose retirement frees the next load slot).
    for (int it = 0; it <= n_iter; ++it) {
      bool b_done = it == 0;
      if (it < n_iter) {
        stamp(1, it, 0);
        if (!b_done && !mbar_test(full_bar(it & 1), (it >> 1) & 1u)) {
          phase_b(it - 1);
          b_done = true;
        }
        mbar_wait(full_bar(it & 1), (it >> 1) & 1u);
        tc_fence_after();
        stamp(1, it, 1);
        phase_a(it, 0);
        stamp(1, it, 2);
      }
      if (!b_done) phase_b(it - 1);
      if (it < n_iter) {
        phase_a(it, 1);
        stamp(1, it, 5);
      }
    }
  } else if (warp < 16) {
    // ===================================================== P / dS warps: lane quarter q, column chunk c
    const int q = warp & 3, c = warp >> 2;
    const int L = q * 32 + lane;
    const int tok = L & 63, half = L >> 6;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int iy = tok_y(tok), ix = tok_x(tok);
    // key (y, x) of column 16c + jj is (cy + (jj >> 2), cx + (jj & 3))
    const int cy = (4 * c) & 7, cx = (c >> 1) * 4;
    const int tab_ofs = (iy + 7 - cy) * kTabPitch + (ix + 7 - cx);  // bias(i, j) = tab[tab_ofs - ((jj>>2)*40 + (jj&3))]
    const int bar_q = 2 + q;  // named barrier of the four warps that share this lane quarter
    float acc[2][16];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[h][j] = 0.0f;

    auto load_table = [&](int pair) {  // both heads of `pair`, pre-multiplied by log2(e)
      for (int i = threadIdx.x; i < 2 * kNumBias; i += 512) {
        const int h = i / kNumBias, bin = i - h * kNumBias;
        const int r = bin / 15, cc = bin - r * 15;
        s_tab[h * kTabHead + r * kTabPitch + cc] = __ldg(p.table + bin * p.nH + 2 * pair + h) * kLog2e;
      }
    };
    // bias-table gradient of `pair`: this thread's (query, 16 keys) sums of dS over all its windows are binned by
    // relative position into the CTA's (dead) shared bias table -- same (dy + 7, dx + 7) addressing as the lookup --
    // and the CTA adds its 2 x 225 bins to the global table.  (An earlier version merged all CTAs in a global
    // [head][query][key] sheet that the last CTA binned: a 3 us flush on every CTA plus an 11 us single-CTA tail.)
    auto flush = [&](int pair) {
      asm volatile("bar.sync 6, 512;" ::: "memory");  // every P / dS warp is done with the bias table
      for (int i = threadIdx.x; i < 2 * kTabHead; i += 512) s_tab[i] = 0.0f;
      asm volatile("bar.sync 6, 512;" ::: "memory");
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float* tab = s_tab + h * kTabHead + tab_ofs;
#pragma unroll
        for (int j = 0; j < 16; ++j) atomicAdd(tab - ((j >> 2) * kTabPitch + (j & 3)), acc[h][j]);
      }
      asm volatile("bar.sync 6, 512;" ::: "memory");
      for (int i = threadIdx.x; i < 2 * kNumBias; i += 512) {
        const int h = i / kNumBias, bin = i - h * kNumBias;
        const int r = bin / 15, cc = bin - r * 15;
        atomicAdd(p.gtable + bin * p.nH + 2 * pair + h, s_tab[h * kTabHead + r * kTabPitch + cc]);
      }
    };

    const int pair = my_pair;
    load_table(pair);
    asm volatile("bar.sync 6, 512;" ::: "memory");
    for (int it = 0; it < n_iter; ++it) {
      const int unit = cta_in_pair + it * ctas_of_pair;
      int b, wy, wx;
      const bool real = window_of(unit, half, b, wy, wx);
      const int win = real ? 2 * unit + half : 2 * unit;
      const bool masked = is_quad(wy, wx);
      uint32_t allow = 0xFFFFu;  // which of this thread's 16 keys share the query's mask region
      if (masked) {
        uint32_t lo = 0xFFFFFFFFu, hi = 0xFFFFFFFFu;
        if (wy == p.nWy - 1) {
          const uint32_t m = iy >= 4 ? 0xFFFF0000u : 0x0000FFFFu;
          lo &= m;
          hi &= m;
        }
        if (wx == p.nWx - 1) {
          if (ix >= 4) lo = 0u;
          else hi = 0u;
        }
        allow = ((c < 2 ? lo : hi) >> (16 * (c & 1))) & 0xFFFFu;
      }
#pragma unroll  // (fully: acc[h][.] must stay in registers)
      for (int h = 0; h < 2; ++h) {
        const int ch = 2 * it + h;
        // (the duplicate window of an odd last unit gets lse = +inf: P = 0, so it adds nothing to the table gradient)
        const float lse = real ? __ldg(p.stats + (static_cast<size_t>(win) * p.nH + 2 * pair + h) * 64 + tok)
                               : __int_as_float(0x7f800000);
        if (q == 0 && c == 0) stamp(2, it, 3 * h);
        mbar_wait(afull_bar, ch & 1u);
        tc_fence_after();
        if (q == 0 && c == 0) stamp(2, it, 3 * h + 1);
        const uint32_t ta = tmem_base + half * 64 + 16 * c + lane_addr;
        uint32_t pk[8];
        float part = 0.0f;
        {
          uint32_t rs[16], rd[16];
          tmem_ld16(ta + kTmS, rs);
          tmem_ld16(ta + kTmDP, rd);
          tmem_ld_wait();
          const float* tab = s_tab + h * kTabHead + tab_ofs;
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float pv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float v = fmaf(__uint_as_float(rs[j + e]), p.scale2,
                             tab[-(((j + e) >> 2) * kTabPitch + ((j + e) & 3))]);
              if (masked && !((allow >> (j + e)) & 1u)) v += -100.0f * kLog2e;
              pv[e] = ex2_approx(v - lse);
              part = fmaf(pv[e], __uint_as_float(rd[j + e]), part);
            }
            pk[j >> 1] = pack_bf16x2(pv[0], pv[1]);
          }
        }
        // delta = rowsum(P o dP): the four column chunks of a row live in the four warps of this lane quarter
        s_part[((ch & 1) * 16 + warp) * 32 + lane] = part;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_q) : "memory");
        const float* qp = s_part + ((ch & 1) * 16 + q) * 32 + lane;
        const float delta = (qp[0] + qp[4 * 32]) + (qp[8 * 32] + qp[12 * 32]);
        // dP again (a second TMEM read is cheaper than 16 registers held across the barrier next to the accumulators)
        uint32_t rd[16], dk[8];
        tmem_ld16(ta + kTmDP, rd);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(afree_bar);
#pragma unroll
        for (int j = 0; j < 16; j += 2) {  // dS = P o (dP - delta) with the bf16 P the tensor core will see
          const float t0 = bf16_lo(pk[j >> 1]) * (__uint_as_float(rd[j]) - delta);
          const float t1 = bf16_hi(pk[j >> 1]) * (__uint_as_float(rd[j + 1]) - delta);
          acc[h][j] += t0;
          acc[h][j + 1] += t1;
          dk[j >> 1] = pack_bf16x2(t0, t1);  // (dQ and dK pick up head_dim^-0.5 in the epilogue)
        }
        // the tiles were last read by phase B of the previous stage (the head-0 values waited in registers meanwhile)
        if (h == 0) mbar_wait(tfree_bar, (it & 1u) ^ 1u);
        const int row = h * 64 + tok;
        const uint32_t prow = base + kBOffTile + half * kSlab + row * 128;
        const uint32_t drow = prow + 2 * kSlab;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const uint32_t sw = static_cast<uint32_t>(((2 * c + v) ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + sw), "r"(pk[4 * v]), "r"(pk[4 * v + 1]),
                       "r"(pk[4 * v + 2]), "r"(pk[4 * v + 3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(drow + sw), "r"(dk[4 * v]), "r"(dk[4 * v + 1]),
                       "r"(dk[4 * v + 2]), "r"(dk[4 * v + 3])
                       : "memory");
        }
        if (q == 0 && c == 0) stamp(2, it, 3 * h + 2);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(tfull_bar);
    }
    flush(pair);  // (every CTA: the named barriers inside count all 512 threads)
  } else if (warp < 20) {
    // ===================================================== epilogue: warps 16..19 (lane quarter = warp & 3)
    const int L = (warp & 3) * 32 + lane;
    const int tok = L & 63, half = L >> 6;  // half = head of the pair
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const bool store_warp = warp == 16;
    const int st_t = lane >> 2, st_quad = lane & 3;  // store lane -> (tensor dq|dk|dv, box)
    const int pair = my_pair;
    for (int j = 0; j < n_iter; ++j) {
      const int unit = cta_in_pair + j * ctas_of_pair;
#pragma unroll 1
      for (int w = 0; w < 2; ++w) {
        const int cw = 2 * j + w;
        if (warp == 16) stamp(3, j, 3 * w);
        mbar_wait(bfull_bar, cw & 1u);
        tc_fence_after();
        if (warp == 16) stamp(3, j, 3 * w + 1);
        // one window's dq | dk | dv staging: free once the previous window's bulk stores have read it
        if (store_warp && lane < 12) tma_store_wait_read<0>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
        for (int t = 0; t < 3; ++t) {  // dQ, dK (x head_dim^-0.5), dV -- one accumulator at a time (56-register budget)
          uint32_t r[32];
          tmem_ld32(tmem_base + kTmDQ + t * 64 + half * 32 + lane_addr, r);
          tmem_ld_wait();
          const float sc = t < 2 ? p.scale : 1.0f;
          const uint32_t orow = base + kBOffOut + t * 8192 + tok * 128;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int chunk = half * 4 + cc;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + ((chunk ^ (tok & 7)) << 4)),
                         "r"(pack_bf16x2(__uint_as_float(r[8 * cc + 0]) * sc, __uint_as_float(r[8 * cc + 1]) * sc)),
                         "r"(pack_bf16x2(__uint_as_float(r[8 * cc + 2]) * sc, __uint_as_float(r[8 * cc + 3]) * sc)),
                         "r"(pack_bf16x2(__uint_as_float(r[8 * cc + 4]) * sc, __uint_as_float(r[8 * cc + 5]) * sc)),
                         "r"(pack_bf16x2(__uint_as_float(r[8 * cc + 6]) * sc, __uint_as_float(r[8 * cc + 7]) * sc))
                         : "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bfree_bar);
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (store_warp && lane < 12) {
          int b, wy, wx, x, y;
          const bool real = window_of(unit, w, b, wy, wx);
          const bool quadrants = is_quad(wy, wx);
          if (real && (quadrants || st_quad < 2)) {
            box_xy(wy, wx, st_quad, quadrants, x, y);
            tma_store_4d(quadrants ? &p.tmap_gq4 : &p.tmap_gq8,
                         base + kBOffOut + st_t * 8192 + st_quad * (quadrants ? 2048 : 4096),
                         st_t * p.Ca + pair * 64, x, y, b);
          }
          tma_store_commit();
        }
        if (warp == 16) stamp(3, j, 3 * w + 2);
      }
    }
    if (store_warp && lane < 12) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  wall(1);
  if (warp == 0) tmem_dealloc(tmem_base, 512);

  wall(2);
}

}  // namespace srb

using namespace srb;

static std::atomic<unsigned long long*> g_attn_bwd_trace{nullptr};
void srb_attn_bwd_set_trace(unsigned long long* buf) { g_attn_bwd_trace.store(buf, std::memory_order_relaxed); }

int srb_window_attention_bwd_tc(const void* qkv_bf16, const void* gout_bf16, const float* rpb_table,
                                const float* stats, void* gqkv_bf16, float* g_rpb_table, float* workspace, int B,
                                int H, int W, int num_heads, int Ca, int shift, float scale, cudaStream_t stream) {
  if ((shift != 0 && shift != 4) || H % 8 != 0 || W % 8 != 0 || (num_heads & 1) || num_heads > kMaxHeads ||
      Ca != num_heads * 32 || stats == nullptr)
    return SRB200_EINVAL;
  if ((reinterpret_cast<uintptr_t>(qkv_bf16) | reinterpret_cast<uintptr_t>(gout_bf16) |
       reinterpret_cast<uintptr_t>(gqkv_bf16)) & 15u)
    return SRB200_EINVAL;
  AttnBwdParams p;
  const uint32_t box4[4] = {64, 4, 4, 1}, box8[4] = {64, 4, 8, 1};
  {
    const uint64_t C3 = 3ull * Ca;
    const uint64_t dims[4] = {C3, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {C3 * 2, C3 * 2 * W, C3 * 2 * W * H};
    int rc = make_tmap_bf16(&p.tmap_qkv4, qkv_bf16, 4, dims, strides, box4);
    if (rc == SRB200_OK) rc = make_tmap_bf16(&p.tmap_qkv8, qkv_bf16, 4, dims, strides, box8);
    if (rc == SRB200_OK) rc = make_tmap_bf16(&p.tmap_gq4, gqkv_bf16, 4, dims, strides, box4);
    if (rc == SRB200_OK) rc = make_tmap_bf16(&p.tmap_gq8, gqkv_bf16, 4, dims, strides, box8);
    if (rc != SRB200_OK) return rc;
  }
  {
    const uint64_t C = static_cast<uint64_t>(Ca);
    const uint64_t dims[4] = {C, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {C * 2, C * 2 * W, C * 2 * W * H};
    int rc = make_tmap_bf16(&p.tmap_do4, gout_bf16, 4, dims, strides, box4);
    if (rc == SRB200_OK) rc = make_tmap_bf16(&p.tmap_do8, gout_bf16, 4, dims, strides, box8);
    if (rc != SRB200_OK) return rc;
  }
  p.table = rpb_table;
  p.stats = stats;
  p.gtable = g_rpb_table;
  p.B = B;
  p.H = H;
  p.W = W;
  p.nH = num_heads;
  p.Ca = Ca;
  p.shift = shift;
  p.nWx = W / 8;
  p.nWy = H / 8;
  const long long nwin = static_cast<long long>(B) * p.nWx * p.nWy;
  if (nwin >= (1LL << 24)) return SRB200_EINVAL;
  p.n_windows = static_cast<int>(nwin);
  p.n_units = (p.n_windows + 1) / 2;
  p.magic_per = div_magic(static_cast<long long>(p.nWx) * p.nWy);
  p.magic_x = div_magic(p.nWx);
  p.magic_units = 0;
  p.scale = scale;
  p.scale2 = scale * kLog2e;
  p.pdl = pdl_enabled() ? 1 : 0;
  p.trace = g_attn_bwd_trace.load(std::memory_order_relaxed);
  static PerDeviceOnce configured;
  if (configured.ensure(window_attn_bwd_tc_kernel, kBwdSmem) != SRB200_OK) return SRB200_ELAUNCH;
  const long long work = static_cast<long long>(p.n_units) * (num_heads / 2);
  const int grid = work < num_sms() ? static_cast<int>(work) : num_sms();
  return launch_ex(window_attn_bwd_tc_kernel, grid, kBwdThreads, kBwdSmem, stream, 1, p);
}
