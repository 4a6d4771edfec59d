// Host-side helpers shared by the C-ABI entry points: TMA tensor-map encoding through the
// driver entry point (no link-time libcuda dependency) and launch checks.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include "srb200.h"  // include/srb200.h (-I): the public C ABI

namespace srb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) !=
            cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// bf16 tensor map, 128B swizzle, zero OOB fill.  dims/strides innermost first; strides[i] is the
// byte stride of dim i+1 (dim 0 is contiguous).
inline int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return SRB200_EDRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr,
                  bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? SRB200_OK : SRB200_EDRIVER;
}

inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64) ? dev : 0;
}

// SM count of the CURRENT device (cached per device: one process may drive several GPUs, e.g. tiled inference)
inline int num_sms() {
  static std::atomic<int> cache[64];
  const int dev = current_device();
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (v <= 0) v = 148;
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: remember it per device, thread-safely
// (the entry points are called from the main thread in forward and from autograd worker threads in backward).
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0};
  template <typename Kernel>
  int ensure(Kernel kernel, size_t smem_bytes) {
    const unsigned long long bit = 1ULL << current_device();
    if (mask.load(std::memory_order_acquire) & bit) return SRB200_OK;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)) !=
        cudaSuccess)
      return SRB200_ELAUNCH;
    mask.fetch_or(bit, std::memory_order_release);
    return SRB200_OK;
  }
};

// a tuning / debugging environment variable, read ONCE per process (not on every launch)
#define SRB_ENV(name)                        \
  ([]() -> const char* {                     \
    static const char* const v = getenv(name); \
    return v;                                \
  }())

inline int launch_status() {
  return cudaPeekAtLastError() == cudaSuccess ? SRB200_OK : SRB200_ELAUNCH;
}

// Programmatic dependent launch (SRB_PDL=1): the GEMM kernels call griddepcontrol.launch_dependents right after
// their prologue and griddepcontrol.wait before touching global memory, so the barrier init / TMEM allocation /
// tensor-map fetch of kernel N+1 overlaps the tail of kernel N on SMs that have already drained.
// Off by default: it is a win for chains of GEMM kernels (EDSR, RCAN: +1..5 %) but when a multi-wave kernel such as
// the window attention sits between them, the early-resident GEMM CTAs starve it (SwinIR: -15 %).  The archs switch
// it on around their CUDA-graph capture with srb200_set_pdl(); SRB_PDL=0/1 in the environment overrides.
inline std::atomic<int>& pdl_flag() {
  static std::atomic<int> flag{0};
  return flag;
}
inline bool pdl_enabled() {
  static const int env = [] {
    const char* e = getenv("SRB_PDL");
    return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
  }();
  return env >= 0 ? env == 1 : pdl_flag().load(std::memory_order_relaxed) != 0;
}

template <typename Kernel, typename Params>
inline int launch_ex(Kernel kernel, int grid, int block, size_t smem, cudaStream_t stream, int cluster,
                     const Params& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  if (cudaLaunchKernelEx(&cfg, kernel, p) != cudaSuccess) return SRB200_ELAUNCH;
  return launch_status();
}

// spatial tile (tw x th pixels, tw*th == npix) minimising padded area for an H x W image
inline void pick_tile(int H, int W, int npix, int* tw, int* th) {
  long best = -1;
  int btw = 16, bth = npix / 16;
  for (int w = 4; w <= 64 && w <= npix; w *= 2) {
    int h = npix / w;
    if (h < 1 || w * h != npix) continue;
    long tx = (W + w - 1) / w, ty = (H + h - 1) / h;
    long area = tx * ty;
    // prefer fewer tiles; tie-break towards wider tiles (longer contiguous runs)
    if (best < 0 || area < best || (area == best && w > btw && w <= 16)) {
      best = area;
      btw = w;
      bth = h;
    }
  }
  *tw = btw;
  *th = bth;
}

}  // namespace srb
