// Multi-tensor Adam: the optimizer step of SRModel.optimize_parameters (basicsr/models/sr_model.py:113;
// optimizer built by BaseModel.get_optimizer, base_model.py:107-124, from the YAML's `optim_g: {type: Adam}`) for
// EVERY parameter of a param group in ONE launch over a device table (step overhead).
// torch.optim.Adam(fused=True) packs ~36 tensors per launch into kernel arguments: 61 launches / 1.1 ms per step for
// RCAN's 1660 parameters, 89 / 0.7 ms for SwinIR's, 6 / 0.24 ms for EDSR-L's -- against 60-440 MB of traffic
// (7 passes over the parameters), i.e. 10-70 us at HBM speed.  Same arithmetic as torch's fused kernel:
//   g += wd * p;  m = lerp(m, g, 1 - b1);  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
#include <stdlib.h>

#include "host_util.h"

namespace srb {

__global__ void __launch_bounds__(256) multi_adam_kernel(const srb200_adam_item* __restrict__ items, int n_items,
                                                         long long total_chunks, float step_size, float omb1,
                                                         float beta2, float omb2, float eps, float weight_decay,
                                                         float inv_bc2_sqrt) {
  __shared__ int s_item;
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    if (threadIdx.x == 0) {
      int lo = 0, hi = n_items - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&items[mid].chunk_begin) <= chunk) lo = mid;
        else hi = mid - 1;
      }
      s_item = lo;
    }
    __syncthreads();
    const srb200_adam_item it = items[s_item];
    __syncthreads();
    float* p = static_cast<float*>(it.p);
    const float* g = static_cast<const float*>(it.g);
    float* m = static_cast<float*>(it.m);
    float* v = static_cast<float*>(it.v);
    const long long i0 = (chunk - it.chunk_begin) * 1024 + 4 * threadIdx.x;
    if (i0 >= it.n) continue;
    auto update = [&](float& pv, float gv, float& mv, float& vv) {
      gv = fmaf(weight_decay, pv, gv);
      mv = fmaf(omb1, gv - mv, mv);                // lerp(m, g, 1 - beta1)
      vv = fmaf(beta2, vv, omb2 * gv * gv);
      pv -= step_size * mv / fmaf(sqrtf(vv), inv_bc2_sqrt, eps);
    };
    const bool vec = (i0 + 4 <= it.n) &&
                     (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                        reinterpret_cast<uintptr_t>(v)) & 15u) == 0);
    if (vec) {
      float4 pv = *reinterpret_cast<float4*>(p + i0);
      const float4 gv = __ldg(reinterpret_cast<const float4*>(g + i0));
      float4 mv = *reinterpret_cast<float4*>(m + i0);
      float4 vv = *reinterpret_cast<float4*>(v + i0);
      update(pv.x, gv.x, mv.x, vv.x);
      update(pv.y, gv.y, mv.y, vv.y);
      update(pv.z, gv.z, mv.z, vv.z);
      update(pv.w, gv.w, mv.w, vv.w);
      *reinterpret_cast<float4*>(p + i0) = pv;
      *reinterpret_cast<float4*>(m + i0) = mv;
      *reinterpret_cast<float4*>(v + i0) = vv;
    } else {
      for (long long i = i0; i < i0 + 4 && i < it.n; ++i) {
        float pv = p[i], mv = m[i], vv = v[i];
        update(pv, __ldg(g + i), mv, vv);
        p[i] = pv;
        m[i] = mv;
        v[i] = vv;
      }
    }
  }
}

}  // namespace srb

using namespace srb;

extern "C" int srb200_multi_adam(const srb200_adam_item* items_dev, int n_items, int64_t total_chunks, double lr,
                                 double beta1, double beta2, double eps, double weight_decay, int64_t step,
                                 srb200_stream_t stream) {
  if (!items_dev || n_items <= 0 || total_chunks <= 0 || step <= 0) return SRB200_EINVAL;
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  const float step_size = static_cast<float>(lr / bc1);
  const float inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const int grid = static_cast<int>(total_chunks < cap ? total_chunks : cap);
  // (1 - beta in double, as torch computes it: 1.0f - 0.99f is 2e-6 away from float(0.01))
  const float omb1 = static_cast<float>(1.0 - beta1);
  const float omb2 = static_cast<float>(1.0 - beta2);
  multi_adam_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(items_dev, n_items, total_chunks, step_size, omb1,
                                                                        static_cast<float>(beta2), omb2,
                                                                        static_cast<float>(eps),
                                                                        static_cast<float>(weight_decay), inv_bc2_sqrt);
  return launch_status();
}
