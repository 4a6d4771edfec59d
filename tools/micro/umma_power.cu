// Whole-GPU sustained UMMA rate: 148 CTAs (or 74 CTA pairs) issue back-to-back bf16 MMAs from resident smem
// operands.  Reports cycles per k-block (clock64), wall time (events) -> effective SM clock and TFLOP/s.
// This is the power-limited ceiling of the tile shape, with no TMA / epilogue traffic.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../basicsr4rs_b200/csrc/ptx.cuh"
using namespace srb;

template <int TWO>
__global__ void __launch_bounds__(128, 1) power_kernel(long long* out, int iters, int nmma_n, int vstages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  constexpr int STAGES = 4;
  const uint32_t bar = base + STAGES * 49152;
  const uint32_t tptr = bar + 8 * (32 + 1);
  for (int i = threadIdx.x; i < STAGES * 49152 / 2; i += blockDim.x)
    reinterpret_cast<unsigned short*>(gbase)[i] = 0x3C00 + ((i * 2654435761u) >> 20 & 0x1FF);
  if (threadIdx.x == 0) { for (int s = 0; s <= 32; ++s) mbar_init(bar + 8 * s, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) {
    if (TWO) { tmem_alloc_2cta(tptr, 256); tmem_relinquish_2cta(); } else { tmem_alloc(tptr, 256); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads();
  if (TWO) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  const bool leader = TWO ? (cluster_ctarank() == 0) : true;
  if (threadIdx.x == 0 && leader) {
    const uint32_t idesc = make_idesc_bf16(TWO ? 256 : 128, nmma_n, 0, 0);
    long long t0 = clock64();
    uint32_t phase = 0; int stage = 0;
    for (int i = 0; i < iters; ++i) {
      if (i >= vstages) mbar_wait(bar + 8 * stage, phase ^ 1u);
      tc_fence_after();
      const uint32_t a0 = base + (stage & 3) * 49152, b0 = a0 + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_smem_desc_sw128(a0 + k * 32, 16, 1024);
        const uint64_t db = make_smem_desc_sw128(b0 + k * 32, 16, 1024);
        if (TWO) umma_bf16_2cta(tmem, da, db, idesc, 1u); else umma_bf16(tmem, da, db, idesc, 1u);
      }
      if (TWO) umma_commit_2cta(bar + 8 * stage); else umma_commit(bar + 8 * stage);
      if (++stage == vstages) { stage = 0; phase ^= 1u; }
    }
    if (TWO) umma_commit_2cta(bar + 8 * 32); else umma_commit(bar + 8 * 32);
    mbar_wait(bar + 8 * 32, 0);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (TWO) cluster_sync_all();
  if (threadIdx.x < 32) { if (TWO) tmem_dealloc_2cta(tmem, 256); else tmem_dealloc(tmem, 256); }
}

template <int TWO>
void run(const char* name, int nmma_n, int iters, int vstages = 4) {
  long long* d; cudaMalloc(&d, 16);
  const int smem = 4 * 49152 + 1024 + 512 + 1024;
  cudaFuncSetAttribute(power_kernel<TWO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(d, 0, 16);
    cudaEventRecord(e0);
    if (TWO) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, power_kernel<TWO>, d, iters, nmma_n, vstages);
    } else {
      power_kernel<TWO><<<148, 128, smem>>>(d, iters, nmma_n, vstages);
    }
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double flop = 2.0 * 128 * nmma_n * 16 * 4 * double(iters) * 148;  // per-SM M=128 rows either way
  printf("%-10s N=%3d stages=%2d iters=%d: %.1f cyc/k-block, %.3f ms, eff clock %.3f GHz, %.0f TFLOP/s  [%s]\n", name, nmma_n, vstages, iters,
         double(h) / iters, ms, double(h) / (ms * 1e6), flop / (ms * 1e9), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  for (int st : {1, 2, 3, 4, 6, 8, 16}) {
    run<0>("1-CTA", 256, 20000, st);
    run<0>("1-CTA", 64, 20000, st);
    run<1>("2-CTA", 256, 20000, st);
  }
  return 0;
}
