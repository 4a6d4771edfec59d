// Fused shifted-window attention on tcgen05 / TMEM / TMA (sm_100a), window 8, forward.
// Replaces roll + window_partition + (q k^T) + bias + mask + softmax + (p v) + window_reverse + roll of
// swinir_arch.py:144-175, 293-316.  No attention matrix, no partitioned copy ever reaches memory.
//
// Work item ("stage") = two windows x one PAIR of heads (64 channels = one 128-byte swizzle row per token):
//   * q, k, v rows arrive by TMA straight from the un-partitioned NHWC qkv tensor [B,H,W,3*Ca]: a window is four
//     4x4-pixel quadrant boxes (64 ch x 4 px x 4 rows), so the cyclic shift (always ws/2 = 4) is just a per-quadrant
//     coordinate ((y + 4) mod H never wraps inside a quadrant) -- window_partition / torch.roll are TMA coordinates.
//     Every window holds its tokens in x-half-major order n = (x >> 2) * 32 + y * 4 + (x & 3) (q, k, v, bias, mask and
//     the output all use it): four 4x4 quadrant boxes produce it, and so do two 4-wide x 8-tall half boxes, which is
//     how windows that cannot wrap are moved (half the TMA instructions).
//   * S = Q K^T: the two windows are stacked to M = 128 rows (A = [Q_w0; Q_w1], B = [K_w0; K_w1], N = 128, K = 32);
//     rows of window w only read columns [64w, 64w+64) -- the off-diagonal blocks are never looked at.
//   * softmax: TMEM lane = query row, so one thread owns a whole row (tcgen05.ld 32x32b): no shuffles.  The
//     relative-position bias is one shared-memory load at an immediate offset (index = base(query) - const(key)),
//     the SW-MSA mask two 32-bit key masks per thread (only windows on the last window row / column have one).
//   * O = P V: the two HEADS of the pair are stacked to M = 128 (A = [P_h0; P_h1] of one window, written to
//     128B-swizzled shared memory as bf16; B = that window's V rows, [key][64 ch] = MN-major, N = 64, K = 64):
//     rows of head h only read columns [32h, 32h+32).
//   * O / rowsum -> bf16 -> swizzled staging -> TMA stores through the same quadrant boxes (window_reverse + roll).
// Warp roles (448 threads): warps 0-7 = softmax in two groups of four (TMEM lane quarter = warp & 3; group g = HEAD g
// of the pair, so every scheduler holds two busy softmax warps -- a single group ran at 0.4 IPC, latency-bound),
// warps 8-11 = epilogue (O / rowsum -> staging -> TMA stores), warp 12 = TMA producer, warp 13 = MMA issuer.  Every
// role runs its own loop over the stages, coupled only by mbarriers: the tensor core runs P V of stage s-1 and Q K^T
// of stage s+1 while the softmax warps work on stage s and the epilogue warps drain stage s-1.  A TMA instruction costs ~90 cycles of the SM's TMA unit whatever its size (measured: 24
// quadrant loads = 2.1k cycles), so windows that cannot wrap are moved as ONE 8x8 box (row-major token order) and only
// the shifted layers' last window row / column use four 4x4 quadrant boxes.
// Algorithmic traffic: read q, k, v + write o = 4 * T * C * 2 bytes (HBM-bound); see DESIGN.md section 3.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "attention_tc.cuh"
#include "host_util.h"
#include "ptx.cuh"

namespace srb {

struct AttnTcParams {
  CUtensorMap tmap_qkv4;  // [B,H,W,3*Ca] bf16, box (64 ch, 4, 4, 1): quadrant of a window that may wrap around
  CUtensorMap tmap_qkv8;  // same tensor, box (64 ch, 4, 8, 1): the left / right half of a window that does not wrap
  CUtensorMap tmap_out4;  // [B,H,W,Ca] bf16, box (64 ch, 4, 4, 1)
  CUtensorMap tmap_out8;  // box (64 ch, 4, 8, 1)
  const float* table;     // [(2*8-1)^2][nH] relative_position_bias_table (swinir_arch.py:117-118)
  float* stats;           // optional [n_windows][nH][64]: log2-domain log-sum-exp of every softmax row (for the backward)
  int B, H, W, nH, Ca, shift;
  int nWx, nWy, n_windows, n_units;  // windows per row / column, B*nWx*nWy, ceil(n_windows / 2)
  unsigned long long magic_per, magic_x;  // srb::div_magic(nWx*nWy), div_magic(nWx)
  float scale2;           // head_dim^-0.5 * log2(e)
  int pdl;
  unsigned skew_ns;       // start-up delay of softmax group 1 (see the kernel)
  int ones;               // SRB200_ATTN_ONES: out channel 31 (pad lane of head 0) = 1.0
  const float* out_alpha; // optional [B]: every output row of sample b is multiplied by out_alpha[b] (DropPath)
  unsigned long long* trace;  // debug: clock64 timeline of CTA 0, [role 0..3][64 stages][8 slots] (NULL = off)
};

constexpr int kAttnThreads = 448;            // 8 softmax warps + 4 epilogue warps + TMA producer + MMA issuer
constexpr int kLoadStage = 3 * kSlab;        // q, k, v slabs of one stage
// shared memory map (offsets from the 1024-aligned base)
constexpr int kOffLoad = 0;                          // 2 stages x (q, k, v)
constexpr int kOffP = kOffLoad + 2 * kLoadStage;     // 2 buffers x 2 windows x [128 (head,query)][64 keys]
constexpr int kOffOut = kOffP + 2 * 2 * kSlab;       // [128 tokens (w0, w1)][64 ch] output staging
constexpr int kOffTab = kOffOut + kSlab;             // [nH <= 16][15][40] fp32 bias * log2e
constexpr int kOffInv = kOffTab + kMaxHeads * kTabHead * 4;  // [4 buffers][2 windows][2 heads][64] 1/rowsum
constexpr int kOffBar = kOffInv + 4 * 2 * 2 * 64 * 4;
constexpr int kAttnSmem = kOffBar + 256 + 1024 /*align*/;

// logits of one query row (64 keys, fp32 accumulators r0 | r1) -> unnormalised probabilities exp2(v - max) as bf16 in
// the 128B-swizzled P tile row `prow`; returns 1 / rowsum.  tab = this head's bias table already offset by the query's
// base index: bias(i, j) = tab[-(jy*40 + jx)], an immediate offset once the loop is unrolled.
template <bool MASKED>
__device__ __forceinline__ float softmax_row(const uint32_t (&r0)[32], const uint32_t (&r1)[32], const float* tab,
                                             float scale2, uint32_t allow_lo, uint32_t allow_hi,
                                             uint32_t prow, int row, float& lse2) {
  float v[64];
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const float s = __uint_as_float(j < 32 ? r0[j] : r1[j - 32]);
    v[j] = fmaf(s, scale2, tab[-(tok_y(j) * kTabPitch + tok_x(j))]);
  }
  if (MASKED) {
#pragma unroll
    for (int j = 0; j < 64; ++j) {
      const uint32_t bits = j < 32 ? allow_lo : allow_hi;
      if (!((bits >> (j & 31)) & 1u)) v[j] += -100.0f * kLog2e;
    }
  }
  // four independent chains each (a single 64-long dependent chain is ~250 cycles of exposed latency per row)
  float m4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
  for (int j = 4; j < 64; ++j) m4[j & 3] = fmaxf(m4[j & 3], v[j]);
  const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
  float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    v[j] = ex2_approx(v[j] - m);
    s4[j & 3] += v[j];
  }
  const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
  lse2 = m + __log2f(sum);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + ((c ^ (row & 7)) << 4)),
                 "r"(pack_bf16x2(v[8 * c + 0], v[8 * c + 1])), "r"(pack_bf16x2(v[8 * c + 2], v[8 * c + 3])),
                 "r"(pack_bf16x2(v[8 * c + 4], v[8 * c + 5])), "r"(pack_bf16x2(v[8 * c + 6], v[8 * c + 7]))
                 : "memory");
  }
  return 1.0f / sum;
}

__global__ void __launch_bounds__(kAttnThreads, 1) window_attn_fwd_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kOffBar;
  // barriers
  auto full_bar = [&](int s) { return bar0 + 8u * s; };           // TMA -> MMA (q, k landed)
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 + s); };    // MMA -> TMA (q, k consumed: S issued and retired)
  auto vfull_bar = [&](int s) { return bar0 + 8u * (16 + s); };   // TMA -> MMA (v landed)
  auto vempty_bar = [&](int s) { return bar0 + 8u * (18 + s); };  // MMA -> TMA (v consumed: P V retired)
  auto sfull_bar = [&](int h) { return bar0 + 8u * (4 + h); };    // MMA -> softmax (S of head h ready)
  auto sfree_bar = [&](int h) { return bar0 + 8u * (6 + h); };    // softmax -> MMA (S of head h read out)
  auto pfull_bar = [&](int b) { return bar0 + 8u * (8 + b); };    // softmax -> MMA (P buffer b written)
  auto ofull_bar = [&](int b) { return bar0 + 8u * (10 + b); };   // MMA -> epilogue (O buffer b ready)
  auto ofree_bar = [&](int b) { return bar0 + 8u * (12 + b); };   // epilogue -> MMA (O buffer b read out)
  const uint32_t tmem_ptr_smem = bar0 + 8u * 14;
  float* s_tab = reinterpret_cast<float*>(gbase + kOffTab);
  float* s_inv = reinterpret_cast<float*>(gbase + kOffInv);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  unsigned long long* trc = (p.trace != nullptr && blockIdx.x == 0) ? p.trace : nullptr;
  auto stamp = [&](int role, int it, int slot) {
    if (trc != nullptr && lane == 0 && it < 64) trc[(role * 64 + it) * 8 + slot] = clock64();
  };

  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&p.tmap_qkv4);
    tma_prefetch_desc(&p.tmap_qkv8);
    tma_prefetch_desc(&p.tmap_out4);
    tma_prefetch_desc(&p.tmap_out8);
  }
  if (warp == 13 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(vfull_bar(s), 1);
      mbar_init(vempty_bar(s), 1);
      mbar_init(sfull_bar(s), 1);
      mbar_init(sfree_bar(s), 4);   // the four warps that read S(head s)
      mbar_init(pfull_bar(s), 8);
      mbar_init(ofull_bar(s), 1);
      mbar_init(ofree_bar(s), 4);
    }
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
  // bias table * log2e -> smem, [head][15][pitch 40]  (global reads only after the PDL handoff)
  for (int i = threadIdx.x; i < kNumBias * p.nH; i += kAttnThreads) {
    const int bin = i / p.nH, h = i - bin * p.nH;
    const int r = bin / 15, c = bin - r * 15;
    s_tab[h * kTabHead + r * kTabPitch + c] = __ldg(p.table + i) * kLog2e;
  }
  __syncthreads();

  const int n_pairs = p.nH >> 1;
  const int my_units = (p.n_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                       static_cast<int>(gridDim.x);
  const int n_iter = my_units * n_pairs;  // stages of this CTA; stage it = (unit it / n_pairs, pair it % n_pairs)
  // window w (0/1) of a unit -> (b, wy, wx); the second window of an odd last unit duplicates the first
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
  // quadrant order <=> the shifted window touches the image's last window row / column (it wraps and is masked)
  auto is_quad = [&](int wy, int wx) { return p.shift > 0 && (wy == p.nWy - 1 || wx == p.nWx - 1); };

  if (warp == 12) {
    // ===================================================== TMA producer: lane = (window, tensor, quadrant)
    // q and k of a stage are dead as soon as its two S MMAs retire, v only after P V a whole softmax later: they
    // have separate barriers, and v(it-1) is issued AFTER q,k(it) -- the order in which their slots free up.  (With one
    // barrier per stage the ~2.5k-cycle load latency sat on the critical path of every stage.)
    const int w = lane / 12, t = (lane % 12) / 4, quad = lane & 3;
    auto issue = [&](int it, int unit, int pair, bool want_v) {
      const int st = it & 1;
      if (lane < 24 && (t == 2) == want_v) {
        int b, wy, wx;
        window_of(unit, w, b, wy, wx);
        const uint32_t dst = base + kOffLoad + st * kLoadStage + t * kSlab + w * 8192;
        const uint32_t bar = want_v ? vfull_bar(st) : full_bar(st);
        if (is_quad(wy, wx)) {  // quadrant (qx, qy) = (quad >> 1, quad & 1): 16 smem rows each
          int y = wy * 8 + (quad & 1) * 4 + p.shift, x = wx * 8 + (quad >> 1) * 4 + p.shift;
          if (y >= p.H) y -= p.H;
          if (x >= p.W) x -= p.W;
          tma_load_4d(dst + quad * 2048, &p.tmap_qkv4, bar, t * p.Ca + pair * 64, x, y, b);
        } else if (quad < 2) {  // x-half `quad`: 4 px wide, 8 rows tall, 32 smem rows
          tma_load_4d(dst + quad * 4096, &p.tmap_qkv8, bar, t * p.Ca + pair * 64, wx * 8 + quad * 4 + p.shift,
                      wy * 8 + p.shift, b);
        }
      }
    };
    int unit = static_cast<int>(blockIdx.x), pair = 0;  // of stage `it`
    int vunit = unit, vpair = 0;                         // of stage `it - 1`
    for (int it = 0; it <= n_iter; ++it) {
      const uint32_t ph = (it >> 1) & 1u;
      if (it < n_iter) {
        stamp(0, it, 0);
        mbar_wait(empty_bar(it & 1), ph ^ 1u);
        stamp(0, it, 1);
        if (lane == 0) mbar_expect_tx(full_bar(it & 1), 2 * kSlab);
        __syncwarp();
        issue(it, unit, pair, false);
        __syncwarp();
        stamp(0, it, 2);
      }
      if (it >= 1) {
        const int j = it - 1;
        mbar_wait(vempty_bar(j & 1), ((j >> 1) & 1u) ^ 1u);
        if (lane == 0) mbar_expect_tx(vfull_bar(j & 1), kSlab);
        __syncwarp();
        issue(j, vunit, vpair, true);
        __syncwarp();
        stamp(0, j, 3);
        if (++vpair == n_pairs) {
          vpair = 0;
          vunit += static_cast<int>(gridDim.x);
        }
      }
      if (++pair == n_pairs) {
        pair = 0;
        unit += static_cast<int>(gridDim.x);
      }
    }
  } else if (warp == 13) {
    // ===================================================== MMA issuer (convergent warp loop, one elected lane issues)
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // S: A, B K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // O: A K-major (P), B MN-major (V rows)
    constexpr uint32_t d_hi = smem_desc_hi_sw128(1024);
    for (int it = 0; it <= n_iter; ++it) {
      if (it < n_iter) {
        const int st = it & 1;
        stamp(1, it, 0);
        mbar_wait(full_bar(st), (it >> 1) & 1u);
        tc_fence_after();
        stamp(1, it, 1);
        const uint32_t q_lo = smem_desc_lo(base + kOffLoad + st * kLoadStage, 16);
        const uint32_t k_lo = smem_desc_lo(base + kOffLoad + st * kLoadStage + kSlab, 16);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(sfree_bar(h), (it & 1u) ^ 1u);
          tc_fence_after();
          if (elect_one()) {
            // head h of the pair = bytes [64h, 64h+64) of every 128-byte row: +4 in the address field, +2 per k16
            umma_bf16_lh<false>(tmem_base + h * 128, q_lo + 4 * h, d_hi, k_lo + 4 * h, d_hi, idesc_s, 0u);
            umma_bf16_lh<false>(tmem_base + h * 128, q_lo + 4 * h + 2, d_hi, k_lo + 4 * h + 2, d_hi, idesc_s, 1u);
            umma_commit(sfull_bar(h));
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(empty_bar(st));  // q, k of this stage are free once the S MMAs retire
        __syncwarp();
        stamp(1, it, 2);
      }
      if (it >= 1) {
        const int j = it - 1, st = j & 1, pb = j & 1;
        const uint32_t ph = (j >> 1) & 1u;
        mbar_wait(pfull_bar(pb), ph);
        stamp(1, j, 3);
        mbar_wait(vfull_bar(st), ph);
        mbar_wait(ofree_bar(pb), ph ^ 1u);
        tc_fence_after();
        stamp(1, j, 4);
        if (elect_one()) {
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const uint32_t p_lo = smem_desc_lo(base + kOffP + (pb * 2 + w) * kSlab, 16);
            const uint32_t v_lo = smem_desc_lo(base + kOffLoad + st * kLoadStage + 2 * kSlab + w * 8192, 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 16 keys: +32 B of a P row (+2), +16 rows of V (+2048 B = +128)
              umma_bf16_lh<false>(tmem_base + 256 + pb * 128 + w * 64, p_lo + 2 * k, d_hi, v_lo + 128 * k, d_hi,
                                  idesc_o, k > 0 ? 1u : 0u);
          }
          umma_commit(ofull_bar(pb));
          umma_commit(vempty_bar(st));
        }
        __syncwarp();
        stamp(1, j, 5);
      }
    }
  } else if (warp < 8) {
    // ===================================================== softmax: two groups of four warps, group g = HEAD g of the
    // pair; TMEM lane quarter = warp & 3, lane L = stacked-window row (window L >> 6, query L & 63)
    const int h = warp >> 2;
    const int L = threadIdx.x & 127;
    const int tok = L & 63;
    const int half = L >> 6;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int iy = tok_y(tok), ix = tok_x(tok);
    // A softmax row runs through pipe-bound phases (LDS + FFMA, FMNMX, MUFU.EX2, F2FP + STS).  The two warps that share
    // a scheduler would hit every phase together; starting group 1 half a stage late keeps them out of step for the
    // whole kernel (each group is paced only by its own S / P barriers), so one warp's MUFU phase overlaps the other's
    // load phase.
    if (h == 1 && p.skew_ns > 0) __nanosleep(p.skew_ns);
    int unit = static_cast<int>(blockIdx.x), pair = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int pb = it & 1;
      int b, wy, wx;
      window_of(unit, half, b, wy, wx);
      const bool masked = is_quad(wy, wx);
      // SW-MSA mask (swinir_arch.py:262-281): only windows of the last window row / column see more than one region
      uint32_t allow_lo = 0xFFFFFFFFu, allow_hi = 0xFFFFFFFFu;
      if (masked) {
        if (wy == p.nWy - 1) {  // keys with y >= 4: bit 4 of the token index
          const uint32_t m = iy >= 4 ? 0xFFFF0000u : 0x0000FFFFu;
          allow_lo &= m;
          allow_hi &= m;
        }
        if (wx == p.nWx - 1) {  // keys with x >= 4 are tokens 32..63
          if (ix >= 4) allow_lo = 0u;
          else allow_hi = 0u;
        }
      }
      if ((warp & 3) == 0) stamp(2, it, 3 * h);
      mbar_wait(sfull_bar(h), it & 1u);
      tc_fence_after();
      if ((warp & 3) == 0) stamp(2, it, 3 * h + 1);
      uint32_t r0[32], r1[32];
      const uint32_t taddr = tmem_base + h * 128 + half * 64 + lane_addr;
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sfree_bar(h));  // the tensor core may overwrite S(h) with the next stage
      // the P buffer was last read by P V of stage it - 2: wait for that MMA to retire (o_full of the same parity slot)
      if (it >= 2) mbar_wait(ofull_bar(pb), ((it - 2) >> 1) & 1u);
      // bias(i, j) = table[(iy - jy + 7) * 15 + (ix - jx + 7)] lives at tab[-(jy * 40 + jx)]
      const float* tab = s_tab + (2 * pair + h) * kTabHead + ((iy + 7) * kTabPitch + ix + 7);
      const int row = h * 64 + tok;  // P tile of this thread's window: row (head, query)
      const uint32_t prow = base + kOffP + (pb * 2 + half) * kSlab + row * 128;
      float inv, lse2;
      if (masked) inv = softmax_row<true>(r0, r1, tab, p.scale2, allow_lo, allow_hi, prow, row, lse2);
      else inv = softmax_row<false>(r0, r1, tab, p.scale2, 0u, 0u, prow, row, lse2);
      if (p.stats != nullptr && 2 * unit + half < p.n_windows)
        p.stats[(static_cast<size_t>(2 * unit + half) * p.nH + 2 * pair + h) * 64 + tok] = lse2;
      // (four buffers: the epilogue can lag the softmax by at most three stages -- P V of stage it needs the epilogue of
      // stage it - 2 to have freed its accumulator)
      s_inv[(((it & 3) * 2 + half) * 2 + h) * 64 + tok] = inv;
      if ((warp & 3) == 0) stamp(2, it, 3 * h + 2);
      fence_proxy_async_smem();  // the P tiles are read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull_bar(pb));
      if (++pair == n_pairs) {
        pair = 0;
        unit += static_cast<int>(gridDim.x);
      }
    }
  } else {
    // ===================================================== epilogue: warps 8..11 (TMEM lane quarter = warp & 3),
    // lane L = stacked-head row (head L >> 6, token L & 63) of BOTH windows' accumulators
    const int L = threadIdx.x & 127;
    const int tok = L & 63;
    const int half = L >> 6;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const bool store_warp = warp == 8;
    const int sw = lane >> 2, squad = lane & 3;  // store lane -> (window, quadrant)
    int unit = static_cast<int>(blockIdx.x), pair = 0;
    for (int j = 0; j < n_iter; ++j) {
      const int pb = j & 1;
      if (warp == 8) stamp(3, j, 0);
      if (store_warp && lane < 8) tma_store_wait_read<0>();  // my previous bulk store has read the staging tile
      // o_full(j) is causally after p_full(j) (the MMA warp waited for it before issuing P V), so the softmax warps'
      // 1/rowsum of stage j is in shared memory by now.  (Waiting on p_full here as well would be wrong: p_full can
      // complete TWICE before a late epilogue looks at it -- stage j+2 only needs o_full(j) -- and a parity wait that
      // is two completions behind blocks forever; o_full(j+2) needs this epilogue's o_free, so it cannot run ahead.)
      mbar_wait(ofull_bar(pb), (j >> 1) & 1u);
      tc_fence_after();
      if (warp == 8) stamp(3, j, 1);
      asm volatile("bar.sync 1, 128;" ::: "memory");         // the staging tile is free (store lanes have waited)
      uint32_t r[2][32];
      tmem_ld32(tmem_base + 256 + pb * 128 + half * 32 + lane_addr, r[0]);
      tmem_ld32(tmem_base + 256 + pb * 128 + 64 + half * 32 + lane_addr, r[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ofree_bar(pb));
      const bool ones_here = p.ones && pair == 0 && half == 0;  // this row ends in out channel 31 (pad lane of head 0)
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        float inv = s_inv[(((j & 3) * 2 + w) * 2 + half) * 64 + tok];
        // SRB200_ATTN_ONES: the pad channel carries 1.0 (x inv below) so that proj's weight-gradient GEMM sums dY there
        if (ones_here) r[w][31] = __float_as_uint(__fdividef(1.0f, inv));
        if (p.out_alpha != nullptr) {  // per-sample DropPath factor folded into the normalisation (the ones lane too)
          int b, wy, wx;
          window_of(unit, w, b, wy, wx);
          inv *= __ldg(p.out_alpha + b);
        }
        const int row = w * 64 + tok;
        const uint32_t orow = base + kOffOut + row * 128;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = half * 4 + c;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + ((chunk ^ (row & 7)) << 4)),
                       "r"(pack_bf16x2(__uint_as_float(r[w][8 * c + 0]) * inv, __uint_as_float(r[w][8 * c + 1]) * inv)),
                       "r"(pack_bf16x2(__uint_as_float(r[w][8 * c + 2]) * inv, __uint_as_float(r[w][8 * c + 3]) * inv)),
                       "r"(pack_bf16x2(__uint_as_float(r[w][8 * c + 4]) * inv, __uint_as_float(r[w][8 * c + 5]) * inv)),
                       "r"(pack_bf16x2(__uint_as_float(r[w][8 * c + 6]) * inv, __uint_as_float(r[w][8 * c + 7]) * inv))
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (store_warp && lane < 8) {
        int b, wy, wx;
        const bool real = window_of(unit, sw, b, wy, wx);  // (the duplicate window of an odd last unit is not stored)
        if (real && is_quad(wy, wx)) {
          int y = wy * 8 + (squad & 1) * 4 + p.shift, x = wx * 8 + (squad >> 1) * 4 + p.shift;
          if (y >= p.H) y -= p.H;
          if (x >= p.W) x -= p.W;
          tma_store_4d(&p.tmap_out4, base + kOffOut + sw * 8192 + squad * 2048, pair * 64, x, y, b);
        } else if (real && squad < 2) {
          tma_store_4d(&p.tmap_out8, base + kOffOut + sw * 8192 + squad * 4096, pair * 64,
                       wx * 8 + squad * 4 + p.shift, wy * 8 + p.shift, b);
        }
        tma_store_commit();
      }
      if (warp == 8) stamp(3, j, 2);
      if (++pair == n_pairs) {
        pair = 0;
        unit += static_cast<int>(gridDim.x);
      }
    }
    if (store_warp && lane < 8) tma_store_wait_all<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace srb

using namespace srb;

void srb_attn_bwd_set_trace(unsigned long long* buf);  // attention_tc_bwd.cu
static std::atomic<unsigned long long*> g_attn_trace{nullptr};
/* debug only: device buffer of 4*64*8 uint64 receiving CTA 0's clock64 timeline of subsequent tcgen05 attention launches */
extern "C" int srb200_debug_set_attn_trace(void* dev_buf) {
  g_attn_trace.store(static_cast<unsigned long long*>(dev_buf), std::memory_order_relaxed);
  srb_attn_bwd_set_trace(static_cast<unsigned long long*>(dev_buf));
  return SRB200_OK;
}

// returns SRB200_OK when the tcgen05 kernel was launched, SRB200_EINVAL when the shape is outside its domain
// (the caller then uses the generic mma.sync kernel of attention.cu)
int srb_window_attention_fwd_tc(const void* qkv_bf16, const float* rpb_table, void* out_bf16, float* stats, int B,
                                int H, int W, int num_heads, int Ca, int shift, float scale, int flags,
                                const float* out_alpha, cudaStream_t stream) {
  if ((shift != 0 && shift != 4) || H % 8 != 0 || W % 8 != 0 || (num_heads & 1) || num_heads > kMaxHeads ||
      Ca != num_heads * 32)
    return SRB200_EINVAL;
  if ((reinterpret_cast<uintptr_t>(qkv_bf16) | reinterpret_cast<uintptr_t>(out_bf16)) & 15u) return SRB200_EINVAL;
  AttnTcParams p;
  {
    const uint64_t C3 = 3ull * Ca;
    const uint64_t dims[4] = {C3, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {C3 * 2, C3 * 2 * W, C3 * 2 * W * H};
    const uint32_t box4[4] = {64, 4, 4, 1}, box8[4] = {64, 4, 8, 1};
    int rc = make_tmap_bf16(&p.tmap_qkv4, qkv_bf16, 4, dims, strides, box4);
    if (rc == SRB200_OK) rc = make_tmap_bf16(&p.tmap_qkv8, qkv_bf16, 4, dims, strides, box8);
    if (rc != SRB200_OK) return rc;
  }
  {
    const uint64_t C = static_cast<uint64_t>(Ca);
    const uint64_t dims[4] = {C, static_cast<uint64_t>(W), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {C * 2, C * 2 * W, C * 2 * W * H};
    const uint32_t box4[4] = {64, 4, 4, 1}, box8[4] = {64, 4, 8, 1};
    int rc = make_tmap_bf16(&p.tmap_out4, out_bf16, 4, dims, strides, box4);
    if (rc == SRB200_OK) rc = make_tmap_bf16(&p.tmap_out8, out_bf16, 4, dims, strides, box8);
    if (rc != SRB200_OK) return rc;
  }
  p.table = rpb_table;
  p.stats = stats;
  p.ones = (flags & SRB200_ATTN_ONES) ? 1 : 0;
  p.out_alpha = out_alpha;
  p.B = B;
  p.H = H;
  p.W = W;
  p.nH = num_heads;
  p.Ca = Ca;
  p.shift = shift;
  p.nWx = W / 8;
  p.nWy = H / 8;
  const long long nwin = static_cast<long long>(B) * p.nWx * p.nWy;
  if (nwin >= (1LL << 24)) return SRB200_EINVAL;  // (fast_div domain; 16M windows = a 32k x 32k scene)
  p.n_windows = static_cast<int>(nwin);
  p.n_units = (p.n_windows + 1) / 2;
  p.scale2 = scale * kLog2e;
  p.magic_per = div_magic(static_cast<long long>(p.nWx) * p.nWy);
  p.magic_x = div_magic(p.nWx);
  p.pdl = pdl_enabled() ? 1 : 0;
  p.trace = g_attn_trace.load(std::memory_order_relaxed);
  p.skew_ns = SRB_ENV("SRB_ATTN_SKEW") ? static_cast<unsigned>(atoi(SRB_ENV("SRB_ATTN_SKEW"))) : 600u;
  static PerDeviceOnce configured;
  if (configured.ensure(window_attn_fwd_tc_kernel, kAttnSmem) != SRB200_OK) return SRB200_ELAUNCH;
  const int grid = p.n_units < num_sms() ? p.n_units : num_sms();
  return launch_ex(window_attn_fwd_tc_kernel, grid, kAttnThreads, kAttnSmem, stream, 1, p);
}
