// Cost of ISSUING cp.async.bulk.tensor loads: one thread (or several lanes) fires `n` loads back to back into the
// same smem buffers, all completing on one mbarrier; reports cycles per instruction as seen by the issuer and the
// total time until all bytes landed.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../basicsr4rs_b200/csrc/host_util.h"
#include "../../basicsr4rs_b200/csrc/ptx.cuh"
using namespace srb;
struct P { CUtensorMap m4, m2; };

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ P p, long long* out, int mode, int n, int lanes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 6 * 32768;
  const int box_bytes = mode == 0 ? 8192 : (mode == 1 ? 16384 : (mode == 2 ? 8192 : 32768));
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); tma_prefetch_desc(&p.m4); tma_prefetch_desc(&p.m2); }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    if (lane == 0) mbar_expect_tx(bar, (unsigned)(box_bytes) * n * lanes);
    __syncwarp();
    long long t0 = clock64();
    if (lane < lanes) {
      for (int i = 0; i < n; ++i) {
        const uint32_t dst = base + ((i * lanes + lane) % 6) * 32768;
        const int tile = (blockIdx.x * 7 + i * lanes + lane) % 96;
        if (mode == 0) tma_load_4d(dst, &p.m4, bar, 0, (tile % 3) * 16, (tile / 3 % 12) * 4, tile % 16);
        else if (mode == 1) tma_load_4d(dst, &p.m4, bar, 0, (tile % 3) * 16, (tile / 3 % 6) * 8, tile % 16);  // needs box h=8 map
        else if (mode == 2) tma_load_2d(dst, &p.m2, bar, 0, tile * 64);
        else tma_load_2d(dst, &p.m2, bar, 0, tile * 256);
      }
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(bar, 0);
    long long t2 = clock64();
    if (lane == 0) { out[2 * blockIdx.x] = t1 - t0; out[2 * blockIdx.x + 1] = t2 - t0; }
  }
}

int main() {
  const int C = 256, W = 48, H = 48, B = 16;
  void* src; cudaMalloc(&src, (size_t)C * W * H * B * 2); cudaMemset(src, 0, (size_t)C * W * H * B * 2);
  long long* d; cudaMalloc(&d, 148 * 16);
  const int smem = 6 * 32768 + 2048;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[4] = {"4D [64][16][4] 8KB", "4D [64][16][8] 16KB", "2D [64][64] 8KB", "2D [64][256] 32KB"};
  for (int mode = 0; mode < 4; ++mode) {
    P p;
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, 16, (uint32_t)(mode == 1 ? 8 : 4), 1};
    make_tmap_bf16(&p.m4, src, 4, dims, str, box);
    uint64_t d2[2] = {(uint64_t)C, (uint64_t)W * H * B};
    uint64_t s2[1] = {(uint64_t)C * 2};
    uint32_t b2[2] = {64, (uint32_t)(mode == 3 ? 256 : 64)};
    make_tmap_bf16(&p.m2, src, 2, d2, s2, b2);
    for (int lanes : {1, 4}) {
      const int n = 32 / lanes;
      for (int r = 0; r < 2; ++r) k<<<148, 128, smem>>>(p, d, mode, n, lanes);
      long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double a = 0, b = 0; for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
      printf("%-22s lanes=%d: issue %.0f cycles/instr (per lane), all %d boxes landed after %.0f cycles (%.0f/box)  [%s]\n",
             names[mode], lanes, a / 148 / n, n * lanes, b / 148, b / 148 / (n * lanes), cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
