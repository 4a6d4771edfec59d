// Weight gradient of conv3x3 / conv1x1 / Linear as a tcgen05 GEMM whose reduction runs over
// pixels (sm_100a):
//
//   acc[tap][n][k] += sum_{b,y,x} dY[b,y,x,n] * X[b, y+dy(tap), x+dx(tap), k]
//
// Both operands are read straight from the NHWC activations with the same 4-D TMA tensor maps
// the forward uses: a box is (64 ch) x (64-pixel spatial patch), i.e. 64 smem rows of 128 B with
// the GEMM-K (pixel) index selecting the row -- the "MN-major" UMMA operand layout
// (LBO = distance between 64-channel blocks, SBO = distance between 8-pixel groups).
// A = dY (M = 128 output channels), B = shifted X (N = BLOCK_N input channels), D in TMEM.
// Work unit = (tap, m-tile, n-tile, pixel split); partial sums are merged with vector fp32
// atomics into `acc` (zeroed by the caller), so any split count is legal.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "host_util.h"
#include "ptx.cuh"

namespace srb {

struct WgradParams {
  CUtensorMap tmap_dy[9];
  CUtensorMap tmap_x;
  float* acc;
  int B, H, W;
  int tile_w, tile_h, tiles_x, tiles_y;
  int total_kb;  // B * tiles_x * tiles_y
  int splits;
  int N, K;      // dY channels (rows of acc), X channels (cols of acc)
  int m_tiles, n_tiles;
  int taps, ksize;
  int dy_c;      // channels per dY view (N / dy_r^2)
  unsigned long long* trace;  // debug timeline of CTA 0 (NULL = off)
  int pdl;       // launched with programmatic stream serialization: do the griddepcontrol handoff
};

// KPIX = pixels (GEMM-K) per stage: a cp.async.bulk.tensor costs its issuing lane 400-500 cycles whatever the box
// size (several lanes issuing at once), so small stages leave the loop TMA-issue-bound (~1000 cycles per 64-pixel
// k-block whose MMAs take 128-512); bigger boxes move more bytes per instruction.
// FOLD (3x3 layers with 64 input channels): the three taps of one stencil row share a work unit -- N = 3 x 64: box nb of
// B is the SAME 64 input channels shifted by tap (row, nb), accumulator columns [64 nb, 64 nb + 64) belong to that tap.
// An M=128 MMA costs >= ~118 cycles whatever its N (the A operand is read from shared memory at ~32 B/cycle), so
// N = 64 units run the tensor core at a quarter of its rate; folded, one MMA does the work of three and dY is loaded
// once per three taps.  A_BOXES = 1: the layer has only 64 output channels, the upper half of A is not even staged
// (the MMA reads the first B box in its place; those accumulator rows are never stored).
template <int BLOCK_N, int KPIX = 64, int A_BOXES = 2>
struct WgCfg {
  static constexpr int BOX_BYTES = KPIX * 128;           // one box: [KPIX px][64 ch]
  static constexpr int A_BYTES = A_BOXES * BOX_BYTES;    // 128 (64) out-ch
  static constexpr int B_BYTES = (BLOCK_N / 64) * BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (216 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BLOCK_N, int KPIX = 64, bool FOLD = false, int A_BOXES = 2>
__global__ void __launch_bounds__(192, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = WgCfg<BLOCK_N, KPIX, A_BOXES>;
  static_assert(!FOLD || BLOCK_N == 192, "fold: three taps x 64 input channels");
  static_assert(A_BOXES == 2 || FOLD, "single-box A only in the folded kernel");
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 1);
  auto smem_a = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;

  // work unit
  int u = blockIdx.x;
  const int split = u % p.splits;
  u /= p.splits;
  const int n_t = u % p.n_tiles;
  u /= p.n_tiles;
  const int m_t = u % p.m_tiles;
  const int tap = u / p.m_tiles;
  const int kb_begin = static_cast<int>((static_cast<long long>(p.total_kb) * split) / p.splits);
  const int kb_end = static_cast<int>((static_cast<long long>(p.total_kb) * (split + 1)) / p.splits);
  const int m0 = m_t * 128, n0 = n_t * (FOLD ? 64 : BLOCK_N);
  unsigned long long* trc = (p.trace != nullptr && blockIdx.x == 0) ? p.trace : nullptr;
  if (trc != nullptr && threadIdx.x == 0) trc[0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_dy[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.pdl) pdl_handoff();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    // TMA producer.  One cp.async.bulk.tensor costs its issuing thread ~170 cycles, so the 2 + BLOCK_N/64 boxes
    // of a stage are issued by as many lanes in parallel (measured: 1070 -> MMA-bound cycles per k-block).
    int dy = 0, dx = 0;
    if (FOLD) {
      dy = tap - 1;  // (`tap` is the stencil row; the column comes from the box index)
    } else if (p.ksize == 3) {
      dy = tap / 3 - 1;
      dx = tap % 3 - 1;
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      const int tx = kb % p.tiles_x;
      const int ty = (kb / p.tiles_x) % p.tiles_y;
      const int b = kb / (p.tiles_x * p.tiles_y);
      const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
      mbar_wait(empty_bar(stage), phase ^ 1u);
      // the upper 64 output channels of the tile may not exist (N = 64): their half of A is never loaded --
      // whatever the MMA makes of the stale smem lands in accumulator rows the epilogue never reads
      const bool upper = A_BOXES == 2 && (m0 + 64) < p.N;
      if (lane == 0)
        mbar_expect_tx(full_bar(stage), (upper || A_BOXES == 1) ? Cfg::STAGE_BYTES : Cfg::STAGE_BYTES - Cfg::BOX_BYTES);
      __syncwarp();
      if (lane < 2) {
        if (lane == 0 || upper) {
          const int n = m0 + lane * 64;      // packed dY channel
          const int view = n / p.dy_c;       // pixel-unshuffle view (0 when dy_r == 1)
          const int c = n - view * p.dy_c;
          tma_load_4d(smem_a(stage) + lane * Cfg::BOX_BYTES, &p.tmap_dy[view], full_bar(stage), c, x0, y0, b);
        }
      } else if (lane < 2 + BLOCK_N / 64) {
        const int nb = lane - 2;
        if (FOLD)
          tma_load_4d(smem_b(stage) + nb * Cfg::BOX_BYTES, &p.tmap_x, full_bar(stage), n0, x0 + nb - 1, y0 + dy, b);
        else
          tma_load_4d(smem_b(stage) + nb * Cfg::BOX_BYTES, &p.tmap_x, full_bar(stage), n0 + nb * 64, x0 + dx, y0 + dy,
                      b);
      }
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    {
      // convergent warp loop, one elected lane issues (see tapgemm.cu: a lane-0-only loop costs ~150 cycles per MMA)
      constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, 1, 1);
      constexpr uint32_t d_hi = smem_desc_hi_sw128(1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (trc != nullptr && lane == 0 && kb - kb_begin < 40) trc[8 + kb - kb_begin] = clock64();
        const uint32_t a_lo = smem_desc_lo(smem_a(stage), Cfg::BOX_BYTES);
        const uint32_t b_lo = smem_desc_lo(smem_b(stage), Cfg::BOX_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KPIX / 16; ++k)  // (16 pixel rows = 2048 bytes = +128 in the address field)
            umma_bf16_lh<false>(tmem_base, a_lo + 128 * k, d_hi, b_lo + 128 * k, d_hi, idesc,
                                (kb > kb_begin || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) umma_commit(tfull_bar);
      __syncwarp();
      if (trc != nullptr && lane == 0) trc[1] = clock64();
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    if (kb_end > kb_begin) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      if (trc != nullptr && threadIdx.x == 64) trc[2] = clock64();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      const int n = m0 + row;
      float* dst = p.acc + (static_cast<size_t>(FOLD ? 3 * tap : tap) * p.N + n) * p.K + n0;
      // Split merge: every thread parks its accumulator row in the (dead) operand ring and hands it to the TMA unit as
      // ONE bulk reduction per contiguous run -- whole lines at L2 instead of 48 sixteen-byte vector atomics per thread
      // (measured on 192-wide units: ~8-10k of a CTA's ~20k cycles went into that atomic tail).  The row pitch keeps
      // the 16-byte shared stores of a quarter-warp on distinct banks.
      constexpr uint32_t PITCH = BLOCK_N * 4 + 16;
      static_assert(128 * PITCH <= STAGES * Cfg::STAGE_BYTES, "accumulator staging must fit the operand ring");
      const uint32_t my_row = smem_base + row * PITCH;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + c * 128 + j * 16), "r"(r[4 * j]),
                       "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
      }
      fence_proxy_async_smem();
      if (n < p.N) {
        if constexpr (FOLD) {  // columns [64 nb, 64 nb + 64) are tap 3 * row + nb
#pragma unroll
          for (int nb = 0; nb < 3; ++nb) bulk_reduce_add_f32(dst + static_cast<size_t>(nb) * p.N * p.K, my_row + nb * 256, 256);
        } else {
          bulk_reduce_add_f32(dst, my_row, BLOCK_N * 4);
        }
        tma_store_commit();
        tma_store_wait_all<0>();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (trc != nullptr && threadIdx.x == 64) trc[3] = clock64();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------ 2-CTA variant (cta_group::2)
// One CTA pair per work unit computes a 256 (out-ch) x 256 (in-ch) block: each CTA stages only ITS 128 output
// channels of dY and ITS 128 input channels of X (32 KB / stage instead of 48 KB -> 6 stages instead of 4 and
// 1/3 less L2->smem traffic per SM); the leader issues M=256 UMMAs that write both CTAs' TMEM.  The 1-CTA
// kernel is load-latency bound (~1000 cycles per k-block vs 512 of MMA work); this one is not.
// KPIX = pixels (GEMM-K) per stage.  A cp.async.bulk.tensor costs its issuing lane 400-500 cycles when four
// lanes issue at once, whatever the box size, so with 64-pixel stages (4 x 8 KB boxes) the producer, not the
// tensor pipe, paces the loop (~980 cycles per 512-cycle k-block); 128-pixel stages move twice the bytes per
// instruction.
template <int KPIX>
struct Wg2Cfg {
  static constexpr int BOX_BYTES = KPIX * 128;      // one box: [KPIX px][64 ch]
  static constexpr int A_BYTES = 2 * BOX_BYTES;     // own 128 out-ch
  static constexpr int B_BYTES = 2 * BOX_BYTES;     // own 128 in-ch
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = KPIX == 64 ? 6 : 3;
  static constexpr int TMEM_COLS = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int KPIX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
    wgrad2_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = Wg2Cfg<KPIX>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 1);
  auto smem_a = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + Cfg::A_BYTES; };

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);  // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  int u = blockIdx.x >> 1;  // work unit of the pair
  const int split = u % p.splits;
  u /= p.splits;
  const int n_t = u % p.n_tiles;   // 256-wide in-channel tiles
  u /= p.n_tiles;
  const int m_t = u % p.m_tiles;   // 256-wide out-channel tiles
  const int tap = u / p.m_tiles;
  const int kb_begin = static_cast<int>((static_cast<long long>(p.total_kb) * split) / p.splits);
  const int kb_end = static_cast<int>((static_cast<long long>(p.total_kb) * (split + 1)) / p.splits);
  const int m0 = m_t * 256 + static_cast<int>(rank) * 128;  // this CTA's out channels
  const int n0 = n_t * 256;                                 // the pair's in channels
  const int nb0 = n0 + static_cast<int>(rank) * 128;        // this CTA's in channels (its half of B)
  unsigned long long* trc = (p.trace != nullptr && blockIdx.x == 0) ? p.trace : nullptr;
  if (trc != nullptr && threadIdx.x == 0) trc[0] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmap_x);
    tma_prefetch_desc(&p.tmap_dy[0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote signal can arrive
  tc_fence_after();
  if (p.pdl) pdl_handoff();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));

  if (warp == 0) {
    int dy = 0, dx = 0;
    if (p.ksize == 3) {
      dy = tap / 3 - 1;
      dx = tap % 3 - 1;
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb_begin; kb < kb_end; ++kb) {
      const int tx = kb % p.tiles_x;
      const int ty = (kb / p.tiles_x) % p.tiles_y;
      const int b = kb / (p.tiles_x * p.tiles_y);
      const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
      if (trc != nullptr && lane == 0 && kb - kb_begin < 20) trc[64 + 3 * (kb - kb_begin)] = clock64();
      mbar_wait(empty_bar(stage), phase ^ 1u);
      if (trc != nullptr && lane == 0 && kb - kb_begin < 20) trc[64 + 3 * (kb - kb_begin) + 1] = clock64();
      // the leader's barrier collects the bytes of BOTH CTAs' loads
      if (rank == 0 && lane == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
      __syncwarp();
      if (lane < 2) {
        const int n = m0 + lane * 64;
        const int view = n / p.dy_c;
        const int c = n - view * p.dy_c;
        tma_load_4d_2cta(smem_a(stage) + lane * Cfg::BOX_BYTES, &p.tmap_dy[view], full_bar(stage), c, x0, y0, b);
      } else if (lane < 4) {
        const int nb = lane - 2;
        tma_load_4d_2cta(smem_b(stage) + nb * Cfg::BOX_BYTES, &p.tmap_x, full_bar(stage), nb0 + nb * 64, x0 + dx,
                         y0 + dy, b);
      }
      __syncwarp();
      if (trc != nullptr && lane == 0 && kb - kb_begin < 20) trc[64 + 3 * (kb - kb_begin) + 2] = clock64();
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, 256, 1, 1);
      constexpr uint32_t d_hi = smem_desc_hi_sw128(1024);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (trc != nullptr && lane == 0 && kb - kb_begin < 40) trc[8 + kb - kb_begin] = clock64();
        const uint32_t a_lo = smem_desc_lo(smem_a(stage), Cfg::BOX_BYTES);
        const uint32_t b_lo = smem_desc_lo(smem_b(stage), Cfg::BOX_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < KPIX / 16; ++k)
            umma_bf16_lh<true>(tmem_base, a_lo + 128 * k, d_hi, b_lo + 128 * k, d_hi, idesc,
                               (kb > kb_begin || k > 0) ? 1u : 0u);
          umma_commit_2cta(empty_bar(stage));  // frees the stage in both CTAs
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) umma_commit_2cta(tfull_bar);
      __syncwarp();
      if (trc != nullptr && lane == 0) trc[1] = clock64();
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    if (kb_end > kb_begin) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
      float* dst = p.acc + (static_cast<size_t>(tap) * p.N + (m0 + row)) * p.K + n0;
      // split merge through the TMA unit (see wgrad_kernel): this CTA's operand ring is dead once the pair's last MMA
      // has retired -- the peer's MMAs read the peer's own shared memory... and ours: tfull is committed by the leader
      // after ALL MMAs of the pair, which read both CTAs' rings
      constexpr uint32_t PITCH = 256 * 4 + 16;
      static_assert(128 * PITCH <= STAGES * Cfg::STAGE_BYTES, "accumulator staging must fit the operand ring");
      const uint32_t my_row = smem_base + row * PITCH;
#pragma unroll 1
      for (int c = 0; c < 256 / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + c * 128 + j * 16), "r"(r[4 * j]),
                       "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
      }
      fence_proxy_async_smem();
      bulk_reduce_add_f32(dst, my_row, 256 * 4);
      tma_store_commit();
      tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be signalling our barriers / reading our TMEM allocation state
  if (warp == 2) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
}

template <int KPIX>
static int launch_wgrad2(const WgradParams& p, cudaStream_t stream) {
  static PerDeviceOnce configured;
  if (configured.ensure(wgrad2_kernel<KPIX>, Wg2Cfg<KPIX>::SMEM_BYTES) != SRB200_OK) return SRB200_ELAUNCH;
  const int grid = 2 * p.taps * p.m_tiles * p.n_tiles * p.splits;
  return launch_ex(wgrad2_kernel<KPIX>, grid, 192, Wg2Cfg<KPIX>::SMEM_BYTES, stream, 1, p);  // (__cluster_dims__ 2)
}

template <int BLOCK_N, int KPIX, bool FOLD = false, int A_BOXES = 2>
static int launch_wgrad(const WgradParams& p, cudaStream_t stream) {
  using Cfg = WgCfg<BLOCK_N, KPIX, A_BOXES>;
  static PerDeviceOnce configured;
  if (configured.ensure(wgrad_kernel<BLOCK_N, KPIX, FOLD, A_BOXES>, Cfg::SMEM_BYTES) != SRB200_OK)
    return SRB200_ELAUNCH;
  const int grid = (FOLD ? 3 : p.taps) * p.m_tiles * p.n_tiles * p.splits;
  return launch_ex(wgrad_kernel<BLOCK_N, KPIX, FOLD, A_BOXES>, grid, 192, Cfg::SMEM_BYTES, stream, 1, p);
}

}  // namespace srb

using namespace srb;

static std::atomic<unsigned long long*> g_wgrad_trace{nullptr};
extern "C" int srb200_debug_set_wgrad_trace(void* dev_buf) {
  g_wgrad_trace.store(static_cast<unsigned long long*>(dev_buf), std::memory_order_relaxed);
  return SRB200_OK;
}

extern "C" int srb200_wgrad(const void* dy_bf16, const void* x_bf16, float* acc, int B, int H,
                            int W, int N, int K, int ksize, int dy_r, srb200_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dy_bf16 || !x_bf16 || !acc) return SRB200_EINVAL;
  if (B <= 0 || H <= 0 || W <= 0 || N <= 0 || K <= 0) return SRB200_EINVAL;
  if (N % 64 != 0 || K % 64 != 0) return SRB200_EINVAL;
  if (ksize != 1 && ksize != 3) return SRB200_EINVAL;
  if (dy_r < 1 || dy_r > 3 || N % (dy_r * dy_r) != 0 || (N / (dy_r * dy_r)) % 64 != 0)
    return SRB200_EINVAL;
  int bn;
  if (K % 256 == 0) bn = 256;
  else if (K % 192 == 0) bn = 192;
  else if (K % 128 == 0) bn = 128;
  else bn = 64;

  WgradParams p;
  p.trace = g_wgrad_trace.load(std::memory_order_relaxed);
  p.pdl = pdl_enabled() ? 1 : 0;
  p.acc = acc;
  p.B = B;
  p.H = H;
  p.W = W;
  pick_tile(H, W, 64, &p.tile_w, &p.tile_h);
  p.tiles_x = (W + p.tile_w - 1) / p.tile_w;
  p.tiles_y = (H + p.tile_h - 1) / p.tile_h;
  p.total_kb = B * p.tiles_x * p.tiles_y;
  p.N = N;
  p.K = K;
  p.m_tiles = (N + 127) / 128;
  p.n_tiles = K / bn;
  p.ksize = ksize;
  p.taps = ksize * ksize;
  p.dy_c = N / (dy_r * dy_r);
  const int units = p.taps * p.m_tiles * p.n_tiles;
  int splits = num_sms() / units;
  if (splits < 1) splits = 1;
  if (splits > p.total_kb) splits = p.total_kb;
  // keep each split long enough to amortise the pipeline fill / atomic epilogue
  while (splits > 1 && p.total_kb / splits < 8) --splits;
  if (const char* e = SRB_ENV("SRB_WG_SPLITS")) {  // tuning aid
    splits = atoi(e);
    if (splits < 1) splits = 1;
    if (splits > p.total_kb) splits = p.total_kb;
  }
  p.splits = splits;

  const bool two_cta = (N % 256 == 0) && (K % 256 == 0) && SRB_ENV("SRB_WGRAD_1CTA") == nullptr;
  // 64-wide input-channel tiles of a 3x3 layer: fold the three taps of a stencil row into one N = 192 unit (WgCfg)
  // Measured (tools/sweep_fold.py, B16): 96x96 64->64 32.9 -> 24.7 us, but 48x48 14.5 -> 15.6 us and N = 256 23.1 ->
  // 27.0 us: with few pixels per unit the fp32 atomics of the split-K merge (three taps' worth per CTA) outweigh the
  // shorter mainloop, so it is used for 64-output-channel layers with >= 100k pixels.  SRB_WG_FOLD=1|0 forces.
  const char* fold_env = SRB_ENV("SRB_WG_FOLD");
  const bool fold_shape = ksize == 3 && bn == 64 && !two_cta;
  const bool fold = fold_shape && (fold_env ? atoi(fold_env) != 0
                                            : (N == 64 && static_cast<long long>(B) * H * W >= 100000));
  const bool fold_a1 = fold && N == 64;
  int kpix = 64;
  if (fold) {
    const char* e = SRB_ENV("SRB_WG_KPIX");
    kpix = fold_a1 ? 128 : 64;  // (stage = (A_BOXES + 3) boxes: 64 KB x 3 stages, or 40 KB x 5)
    if (e && fold_a1 && (atoi(e) == 64 || atoi(e) == 128)) kpix = atoi(e);
    pick_tile(H, W, kpix, &p.tile_w, &p.tile_h);
    p.tiles_x = (W + p.tile_w - 1) / p.tile_w;
    p.tiles_y = (H + p.tile_h - 1) / p.tile_h;
    p.total_kb = B * p.tiles_x * p.tiles_y;
    int s3 = num_sms() / (3 * p.m_tiles * p.n_tiles);
    if (s3 < 1) s3 = 1;
    if (s3 > p.total_kb) s3 = p.total_kb;
    while (s3 > 1 && p.total_kb / s3 < 4) --s3;
    if (const char* es = SRB_ENV("SRB_WG_SPLITS")) {
      s3 = atoi(es);
      if (s3 < 1) s3 = 1;
      if (s3 > p.total_kb) s3 = p.total_kb;
    }
    p.splits = s3;
  }
  if (!two_cta && !fold) {
    // largest stage (pixels per k-block) that still leaves every split a few k-blocks
    const char* e = SRB_ENV("SRB_WG_KPIX");
    const int cand[3] = {bn == 64 ? 256 : 128, 128, 64};
    for (int ci = 0; ci < 3; ++ci) {
      const int kp = e ? atoi(e) : cand[ci];
      if (kp != 64 && kp != 128 && !(kp == 256 && bn == 64)) break;
      int tw, th;
      pick_tile(H, W, kp, &tw, &th);
      const int kbs = B * ((W + tw - 1) / tw) * ((H + th - 1) / th);
      if (kp == 64 || e || kbs / splits >= 4 || (splits == 1 && kbs >= 2)) {
        kpix = kp;
        p.tile_w = tw;
        p.tile_h = th;
        p.tiles_x = (W + tw - 1) / tw;
        p.tiles_y = (H + th - 1) / th;
        p.total_kb = kbs;
        if (p.splits > p.total_kb) p.splits = p.total_kb;
        while (p.splits > 1 && p.total_kb / p.splits < 4) --p.splits;
        break;
      }
    }
  }
  if (two_cta) {
    // 128-pixel stages when every split still gets a few of them
    const char* e = SRB_ENV("SRB_WG2_KPIX");
    kpix = e ? atoi(e) : 128;
    if (kpix == 128) {
      int tw, th;
      pick_tile(H, W, 128, &tw, &th);
      const int kb128 = B * ((W + tw - 1) / tw) * ((H + th - 1) / th);
      const int units2 = p.taps * (N / 256) * (K / 256);
      int s2 = (num_sms() / 2) / units2;
      if (s2 < 1) s2 = 1;
      if (kb128 / s2 < 6 && !e) kpix = 64;
      if (kpix == 128) {
        p.tile_w = tw;
        p.tile_h = th;
        p.tiles_x = (W + tw - 1) / tw;
        p.tiles_y = (H + th - 1) / th;
        p.total_kb = kb128;
      }
    }
    p.m_tiles = N / 256;
    p.n_tiles = K / 256;
    const int units2 = p.taps * p.m_tiles * p.n_tiles;
    int s2 = (num_sms() / 2) / units2;
    if (s2 < 1) s2 = 1;
    if (s2 > p.total_kb) s2 = p.total_kb;
    while (s2 > 1 && p.total_kb / s2 < 8) --s2;
    p.splits = s2;
  }
  const uint32_t box[4] = {64, static_cast<uint32_t>(p.tile_w), static_cast<uint32_t>(p.tile_h), 1};
  {
    const uint64_t C = static_cast<uint64_t>(K);
    const uint64_t dims[4] = {C, static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                              static_cast<uint64_t>(B)};
    const uint64_t strides[3] = {C * 2, static_cast<uint64_t>(W) * C * 2,
                                 static_cast<uint64_t>(H) * W * C * 2};
    const int rc = make_tmap_bf16(&p.tmap_x, x_bf16, 4, dims, strides, box);
    if (rc != SRB200_OK) return rc;
  }
  {
    const int r = dy_r;
    const uint64_t C = static_cast<uint64_t>(p.dy_c);
    const uint64_t Wf = static_cast<uint64_t>(W) * r, Hf = static_cast<uint64_t>(H) * r;
    for (int i = 0; i < r; ++i)
      for (int j = 0; j < r; ++j) {
        const __nv_bfloat16* base =
            static_cast<const __nv_bfloat16*>(dy_bf16) + (static_cast<uint64_t>(i) * Wf + j) * C;
        const uint64_t dims[4] = {C, static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                                  static_cast<uint64_t>(B)};
        const uint64_t strides[3] = {r * C * 2, r * Wf * C * 2, Hf * Wf * C * 2};
        const int rc = make_tmap_bf16(&p.tmap_dy[i * r + j], base, 4, dims, strides, box);
        if (rc != SRB200_OK) return rc;
      }
  }
  if (two_cta) return kpix == 128 ? launch_wgrad2<128>(p, stream) : launch_wgrad2<64>(p, stream);
  if (fold) {
    if (!fold_a1) return launch_wgrad<192, 64, true, 2>(p, stream);
    return kpix == 128 ? launch_wgrad<192, 128, true, 1>(p, stream) : launch_wgrad<192, 64, true, 1>(p, stream);
  }
  switch (bn) {
    case 256: return kpix == 128 ? launch_wgrad<256, 128>(p, stream) : launch_wgrad<256, 64>(p, stream);
    case 192: return kpix == 128 ? launch_wgrad<192, 128>(p, stream) : launch_wgrad<192, 64>(p, stream);
    case 128: return kpix == 128 ? launch_wgrad<128, 128>(p, stream) : launch_wgrad<128, 64>(p, stream);
    case 64:
      return kpix == 256   ? launch_wgrad<64, 256>(p, stream)
             : kpix == 128 ? launch_wgrad<64, 128>(p, stream)
                           : launch_wgrad<64, 64>(p, stream);
  }
  return SRB200_EINVAL;
}
