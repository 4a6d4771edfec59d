// UMMA rate with a realistic issue pattern: rotating smem stages, a tcgen05.commit per 4 MMAs, waiting on the
// commit barrier of stage s before re-using it (as the real mainloop does), random non-zero operands, optional
// concurrent smem traffic from other warps.  bf16, M=128, N=256.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../basicsr4rs_b200/csrc/ptx.cuh"
using namespace srb;

template <int MN, int STAGES>
__global__ void __launch_bounds__(256, 1) rate_kernel(long long* out, int iters, int traffic) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar = base + STAGES * 49152;
  const uint32_t tptr = bar + 8 * (STAGES + 1);
  // random-ish bf16 data (finite values)
  for (int i = threadIdx.x; i < STAGES * 49152 / 2; i += blockDim.x)
    reinterpret_cast<unsigned short*>(gbase)[i] = 0x3C00 + ((i * 2654435761u) >> 20 & 0x1FF);
  if (threadIdx.x == 0) { for (int s = 0; s <= STAGES; ++s) mbar_init(bar + 8 * s, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tptr, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(tptr));
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, MN, MN);
    long long t0 = clock64();
    uint32_t phase = 0; int stage = 0;
    for (int i = 0; i < iters; ++i) {
      if (i >= STAGES) mbar_wait(bar + 8 * stage, phase ^ 1u);  // stage free again (its previous MMAs retired)
      tc_fence_after();
      const uint32_t a0 = base + stage * 49152, b0 = a0 + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = MN ? make_smem_desc_sw128(a0 + k * 2048, 8192, 1024) : make_smem_desc_sw128(a0 + k * 32, 16, 1024);
        const uint64_t db = MN ? make_smem_desc_sw128(b0 + k * 2048, 8192, 1024) : make_smem_desc_sw128(b0 + k * 32, 16, 1024);
        umma_bf16(tmem, da, db, idesc, 1u);
      }
      umma_commit(bar + 8 * stage);
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    umma_commit(bar + 8 * STAGES);
    mbar_wait(bar + 8 * STAGES, 0);
    out[0] = clock64() - t0;
    out[1] = 1;
  } else if (traffic && threadIdx.x >= 128) {
    // background smem writes (stand-in for TMA fills): 128 threads streaming 16-byte stores into a scratch area
    const uint32_t scratch = base + STAGES * 49152 + 1024;
    while (reinterpret_cast<volatile long long*>(out)[1] == 0)
      for (int r = 0; r < 64; ++r)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(scratch + 16 * ((threadIdx.x - 128) + 128 * (r & 7))), "r"(r) : "memory");
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

template <int MN, int STAGES>
void run(const char* name, int traffic) {
  long long* d; cudaMalloc(&d, 16);
  const int smem = STAGES * 49152 + 1024 + 256 + 16384 + 1024, iters = 2000;
  cudaFuncSetAttribute(rate_kernel<MN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) { cudaMemset(d, 0, 16); rate_kernel<MN, STAGES><<<1, 256, smem>>>(d, iters, traffic); }
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-12s stages=%d traffic=%d: %.1f cycles per k-block of 4 MMAs (ideal 512)  [%s]\n", name, STAGES, traffic,
         double(h) / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main() {
  run<0, 4>("K-major", 0);
  run<1, 4>("MN-major", 0);
  run<0, 4>("K-major", 1);
  run<1, 4>("MN-major", 1);
  run<1, 2>("MN-major", 0);
  return 0;
}
