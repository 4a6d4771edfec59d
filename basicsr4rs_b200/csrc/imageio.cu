// Image entry / exit kernels on the callers' side of the hot path, all HBM-bound
// byte shuffles:
//   patch_from_u8     paired_random_crop + augment (hflip / vflip / rot90) + img2tensor (BGR->RGB, HWC->CHW, /255) of
//                     basicsr/data/transforms.py:28-96,166-225 and utils/img_util.py:9-37, batched, from uint8 HWC images
//                     that already live in device memory -- also the uint8 tile entry of the tiled-inference driver
//   tile_blend_add    linear-ramp overlap blending of super-resolved tiles into an fp32 scene accumulator
//   tensor2img_u8     tensor2img of utils/img_util.py:40-96: clamp, normalise, x255, round-half-even, uint8, RGB->BGR, HWC
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "host_util.h"

namespace srb {

struct PatchItem {          // mirror of srb200_patch_item
  const uint8_t* src;       // uint8 HWC image (device)
  long long pitch;          // bytes per image row
  int top, left;            // crop origin
  int flags;                // bit 0 hflip, bit 1 vflip, bit 2 rot90 (transpose), applied in this order (transforms.py:194-200)
  int reserved;
};

// out[b][c][y][x] (fp32, CHW) = crop_b[r'][s'][c'] / div,  (r, s) = rot90 ? (x, y) : (y, x),
// r' = vflip ? ph-1-r : r,  s' = hflip ? pw-1-s : s,  c' = bgr2rgb ? C-1-c : c (3-channel images only)
__global__ void patch_from_u8_kernel(const PatchItem* __restrict__ items, int C, int ph, int pw, int bgr2rgb,
                                     float div, float* __restrict__ out) {
  const PatchItem it = items[blockIdx.z];
  const bool rot = (it.flags & 4) != 0;
  const int oh = rot ? pw : ph, ow = rot ? ph : pw;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  int r = rot ? x : y, s = rot ? y : x;
  if (it.flags & 2) r = ph - 1 - r;
  if (it.flags & 1) s = pw - 1 - s;
  const uint8_t* px = it.src + static_cast<long long>(it.top + r) * it.pitch + static_cast<long long>(it.left + s) * C;
  float* o = out + (static_cast<size_t>(blockIdx.z) * C * oh + y) * ow + x;
  const size_t plane = static_cast<size_t>(oh) * ow;
  for (int c = 0; c < C; ++c) {
    const int cs = (bgr2rgb && C == 3) ? 2 - c : c;
    o[c * plane] = __fdiv_rn(static_cast<float>(px[cs]), div);  // a true division: bit-equal to numpy's img / 255.
  }
}

__device__ __forceinline__ float ramp_later(int p, int ov, int guard) {
  // weight of the LATER tile at position p (0-based) of an overlap ov pixels wide; the earlier tile gets 1 - w.
  // The outer `guard` pixels of every tile (contaminated by its zero padding) carry no weight.
  const float span = static_cast<float>(ov - 2 * guard);
  const float w = (static_cast<float>(p - guard) + 0.5f) / span;
  return fminf(fmaxf(w, 0.0f), 1.0f);
}

struct BlendGeom {
  int C, th, tw;        // tile (SR pixels)
  int H, W;             // accumulator
  int y0, x0;           // tile origin in accumulator coordinates (may be negative: rows above the band are clipped)
  int ov_top, ov_bottom, ov_left, ov_right;  // overlap with the neighbouring tile on each side (0 = image border)
  int guard;
};

__global__ void tile_blend_add_kernel(const float* __restrict__ sr, float* __restrict__ acc, BlendGeom g) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= g.tw || y >= g.th) return;
  const int Y = g.y0 + y, X = g.x0 + x;
  if (Y < 0 || Y >= g.H || X < 0 || X >= g.W) return;
  float wy = 1.0f, wx = 1.0f;
  if (g.ov_top > 0 && y < g.ov_top) wy = ramp_later(y, g.ov_top, g.guard);
  if (g.ov_bottom > 0 && y >= g.th - g.ov_bottom) wy *= 1.0f - ramp_later(y - (g.th - g.ov_bottom), g.ov_bottom, g.guard);
  if (g.ov_left > 0 && x < g.ov_left) wx = ramp_later(x, g.ov_left, g.guard);
  if (g.ov_right > 0 && x >= g.tw - g.ov_right) wx *= 1.0f - ramp_later(x - (g.tw - g.ov_right), g.ov_right, g.guard);
  const float w = wy * wx;
  if (w == 0.0f) return;
  const size_t tp = static_cast<size_t>(g.th) * g.tw, ap = static_cast<size_t>(g.H) * g.W;
  const float* s = sr + static_cast<size_t>(y) * g.tw + x;
  float* a = acc + static_cast<size_t>(Y) * g.W + X;
  for (int c = 0; c < g.C; ++c) a[c * ap] += w * s[c * tp];
}

// dst[y][x][c'] = uint8(rint(255 * (clamp(src[c][y][x], lo, hi) - lo) / (hi - lo)))   (numpy .round() = half to even)
__global__ void tensor2img_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int C, int H, int W,
                                     float lo, float inv_range, int rgb2bgr) {
  const size_t plane = static_cast<size_t>(H) * W;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < plane;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    uint8_t* d = dst + i * C;
    for (int c = 0; c < C; ++c) {
      const int cd = (rgb2bgr && C == 3) ? 2 - c : c;
      float v = (fminf(fmaxf(src[c * plane + i], lo), lo + 1.0f / inv_range) - lo) * inv_range;
      d[cd] = static_cast<uint8_t>(rintf(v * 255.0f));
    }
  }
}

}  // namespace srb

using namespace srb;

extern "C" int srb200_patch_from_u8(const void* items_dev, int n, int C, int ph, int pw, int bgr2rgb, float div,
                                    float* out, srb200_stream_t stream) {
  if (!items_dev || !out || n <= 0 || n > 65535 || C <= 0 || C > 16 || ph <= 0 || pw <= 0 || !(div > 0.0f))
    return SRB200_EINVAL;
  const int m = ph > pw ? ph : pw;
  const dim3 block(32, 8), grid((m + 31) / 32, (m + 7) / 8, n);
  patch_from_u8_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const PatchItem*>(items_dev), C,
                                                                             ph, pw, bgr2rgb, div, out);
  return launch_status();
}

extern "C" int srb200_tile_blend_add(const float* sr_tile, float* acc, int C, int th, int tw, int acc_h, int acc_w,
                                     int y0, int x0, int ov_top, int ov_bottom, int ov_left, int ov_right, int guard,
                                     srb200_stream_t stream) {
  if (!sr_tile || !acc || C <= 0 || th <= 0 || tw <= 0 || acc_h <= 0 || acc_w <= 0 || guard < 0) return SRB200_EINVAL;
  const int ovs[4] = {ov_top, ov_bottom, ov_left, ov_right};
  for (int i = 0; i < 4; ++i)
    if (ovs[i] < 0 || (ovs[i] > 0 && ovs[i] <= 2 * guard)) return SRB200_EINVAL;  // the ramp needs room inside the guard
  if (ov_top + ov_bottom > th || ov_left + ov_right > tw) return SRB200_EINVAL;
  BlendGeom g{C, th, tw, acc_h, acc_w, y0, x0, ov_top, ov_bottom, ov_left, ov_right, guard};
  const dim3 block(32, 8), grid((tw + 31) / 32, (th + 7) / 8);
  tile_blend_add_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(sr_tile, acc, g);
  return launch_status();
}

extern "C" int srb200_tensor2img_u8(const float* src_chw, void* dst_hwc_u8, int C, int H, int W, float lo, float hi,
                                    int rgb2bgr, srb200_stream_t stream) {
  if (!src_chw || !dst_hwc_u8 || C <= 0 || C > 16 || H <= 0 || W <= 0 || !(hi > lo)) return SRB200_EINVAL;
  const size_t plane = static_cast<size_t>(H) * W;
  size_t blocks = (plane + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  tensor2img_u8_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src_chw, static_cast<uint8_t*>(dst_hwc_u8), C, H, W, lo, 1.0f / (hi - lo), rgb2bgr);
  return launch_status();
}

// ------------------------------------------------------------------ fp32 mode: error-compensated bf16 operands
// x (fp32) = hi + lo + O(2^-17 |x|) with hi = bf16(x), lo = bf16(x - hi).  The tap-GEMM then sees the activation
// [hi | lo | hi] (3 * C channels) against the weight [w_hi ; w_hi ; w_lo]: sum_k (x_hi w_hi + x_lo w_hi + x_hi w_lo),
// every product exact in the fp32 TMEM accumulator -- fp32-class results (relative 2^-16 per product, 32x tighter than
// TF32's 2^-11) from the bf16 tensor-core kernels (north_star's "TF32/fp32 mode", BASELINE.md section 4: <= 1e-4).
namespace srb {
__global__ void split3_kernel(const float4* __restrict__ x, uint4* __restrict__ out, size_t nvec8, int c8) {
  // one thread: 8 consecutive channels of one pixel; out row = [hi (C) | lo (C) | hi (C)]
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < nvec8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(x + 2 * i), b = __ldg(x + 2 * i + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * e]), h1 = __float2bfloat16_rn(v[2 * e + 1]);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * e] - __bfloat162float(h0));
      const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * e + 1] - __bfloat162float(h1));
      hi[e] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      lo[e] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
    const size_t pix = i / c8, v8 = i - pix * c8;
    uint4* row = out + pix * (3 * static_cast<size_t>(c8));
    const uint4 h = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    row[v8] = h;
    row[c8 + v8] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    row[2 * c8 + v8] = h;
  }
}
}  // namespace srb

extern "C" int srb200_split3_bf16(const float* x_f32, void* out_bf16, int64_t rows, int C, srb200_stream_t stream) {
  if (!x_f32 || !out_bf16 || rows <= 0 || C <= 0 || C % 8 != 0) return SRB200_EINVAL;
  if ((reinterpret_cast<uintptr_t>(x_f32) | reinterpret_cast<uintptr_t>(out_bf16)) & 15u) return SRB200_EINVAL;
  const size_t nvec8 = static_cast<size_t>(rows) * (C / 8);
  size_t blocks = (nvec8 + 255) / 256;
  const size_t cap = static_cast<size_t>(srb::num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  srb::split3_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x_f32), static_cast<uint4*>(out_bf16), nvec8, C / 8);
  return srb::launch_status();
}
