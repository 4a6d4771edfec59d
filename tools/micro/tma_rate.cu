// TMA load throughput per SM for different box shapes / ranks (bf16 NHWC tensor, 128B swizzle), all SMs active,
// source L2-resident.  Each CTA keeps DEPTH boxes in flight from one thread.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../basicsr4rs_b200/csrc/host_util.h"
#include "../../basicsr4rs_b200/csrc/ptx.cuh"
using namespace srb;

struct P { CUtensorMap m; int rank; int bw, bh; int W, H, B; int box_bytes; int iters; };

template <int DEPTH>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ P p, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + DEPTH * 32768;
  if (threadIdx.x == 0) { for (int i = 0; i < DEPTH; ++i) mbar_init(bar + 8 * i, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int tiles_x = p.W / p.bw, tiles_y = p.H / p.bh;
    const int ntile = tiles_x * tiles_y * p.B;
    long long t0 = clock64();
    uint32_t phase = 0; int slot = 0;
    int tile = (blockIdx.x * 37) % ntile;
    for (int i = 0; i < p.iters + DEPTH; ++i) {
      if (i >= DEPTH) { mbar_wait(bar + 8 * slot, phase); }
      if (i < p.iters) {
        mbar_expect_tx(bar + 8 * slot, p.box_bytes);
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
        if (p.rank == 4) tma_load_4d(base + slot * 32768, &p.m, bar + 8 * slot, 0, tx * p.bw, ty * p.bh, b);
        else tma_load_2d(base + slot * 32768, &p.m, bar + 8 * slot, 0, tile * p.bw);
        tile = (tile + 1) % ntile;
      }
      if (++slot == DEPTH) { slot = 0; if (i >= DEPTH) phase ^= 1u; }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

template <int DEPTH>
static void run(const char* name, int rank, int bw, int bh, int C, int W, int H, int B, void* d_src) {
  P p; p.rank = rank; p.bw = bw; p.bh = bh; p.W = W; p.H = H; p.B = B; p.iters = 400;
  if (rank == 4) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
    make_tmap_bf16(&p.m, d_src, 4, dims, str, box);
    p.box_bytes = 128 * bw * bh;
  } else {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)W * H * B};
    uint64_t str[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {64, (uint32_t)bw};
    make_tmap_bf16(&p.m, d_src, 2, dims, str, box);
    p.box_bytes = 128 * bw; p.W = W * H * B; p.H = 1; p.B = 1; p.bh = 1;
  }
  long long* d; cudaMalloc(&d, 148 * 8);
  const int smem = DEPTH * 32768 + 2048;
  cudaFuncSetAttribute(k<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int r = 0; r < 2; ++r) k<DEPTH><<<148, 128, smem>>>(p, d);
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("depth %2d %-30s box %5d B: %7.1f cycles/box  %6.1f B/cycle/SM  [%s]\n", DEPTH, name, p.box_bytes, avg / p.iters,
         p.box_bytes / (avg / p.iters), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  const int C = 256, W = 48, H = 48, B = 16;
  void* src; cudaMalloc(&src, (size_t)C * W * H * B * 2); cudaMemset(src, 0, (size_t)C * W * H * B * 2);
  run<1>("4D [64ch][16px][4rows]", 4, 16, 4, C, W, H, B, src);
  run<2>("4D [64ch][16px][4rows]", 4, 16, 4, C, W, H, B, src);
  run<4>("4D [64ch][16px][4rows]", 4, 16, 4, C, W, H, B, src);
  run<6>("4D [64ch][16px][4rows]", 4, 16, 4, C, W, H, B, src);
  run<1>("4D [64ch][16px][8rows]", 4, 16, 8, C, W, H, B, src);
  run<6>("4D [64ch][16px][8rows]", 4, 16, 8, C, W, H, B, src);
  run<1>("2D [64ch][64 rows]", 2, 64, 1, C, W, H, B, src);
  run<2>("2D [64ch][64 rows]", 2, 64, 1, C, W, H, B, src);
  run<6>("2D [64ch][64 rows]", 2, 64, 1, C, W, H, B, src);
  run<1>("2D [64ch][256 rows]", 2, 256, 1, C, W, H, B, src);
  run<6>("2D [64ch][256 rows]", 2, 256, 1, C, W, H, B, src);
  return 0;
}
