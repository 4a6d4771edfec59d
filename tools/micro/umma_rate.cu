// Micro-benchmark: tcgen05.mma issue rate for K-major vs MN-major smem operands (bf16, M=128, N=256, K=16).
// Build: nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../basicsr4rs_b200/csrc/ptx.cuh"
using namespace srb;

template <int A_MN, int B_MN, int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024;
  const uint32_t tptr = bar + 16;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tptr, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, N, A_MN, B_MN);
    const uint32_t a0 = base, b0 = base + 32 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = A_MN ? make_smem_desc_sw128(a0 + k * 2048, 8192, 1024) : make_smem_desc_sw128(a0 + k * 32, 16, 1024);
        const uint64_t db = B_MN ? make_smem_desc_sw128(b0 + k * 2048, 8192, 1024) : make_smem_desc_sw128(b0 + k * 32, 16, 1024);
        umma_bf16(tmem, da, db, idesc, 1u);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

template <int A_MN, int B_MN, int N>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 100 * 1024, iters = 2000;
  cudaFuncSetAttribute(rate_kernel<A_MN, B_MN, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  rate_kernel<A_MN, B_MN, N><<<1, 128, smem>>>(d, iters);
  rate_kernel<A_MN, B_MN, N><<<1, 128, smem>>>(d, iters);
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d: %.1f cycles per MMA (ideal %d)  [%s]\n", name, N, double(h) / (iters * 4.0), N / 2,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<0, 0, 256>("A K-major,  B K-major");
  run<1, 0, 256>("A MN-major, B K-major");
  run<0, 1, 256>("A K-major,  B MN-major");
  run<1, 1, 256>("A MN-major, B MN-major");
  run<0, 0, 128>("A K-major,  B K-major");
  run<1, 1, 128>("A MN-major, B MN-major");
  run<0, 0, 64>("A K-major,  B K-major");
  run<1, 1, 64>("A MN-major, B MN-major");
  return 0;
}
