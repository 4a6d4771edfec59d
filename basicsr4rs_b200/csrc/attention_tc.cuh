// Shared by the tcgen05 window-attention kernels (attention_tc.cu forward, attention_tc_bwd.cu backward).
#pragma once
#include <stdint.h>

namespace srb {

constexpr int kSlab = 16384;     // [128 tokens][64 ch] bf16, 128-byte rows, SWIZZLE_128B
constexpr int kNumBias = 225;    // (2*8-1)^2 relative positions
constexpr int kTabPitch = 40;    // row pitch of the smem bias table: a warp's 32 queries (4 rows x 8 columns) hit 32 banks
constexpr int kTabHead = 15 * kTabPitch;
constexpr int kMaxHeads = 16;
constexpr float kLog2e = 1.4426950408889634f;

// Token order inside a window's 64 smem rows = the order its TMA boxes deliver: x-half-major,
//   n = (x >> 2) * 32 + y * 4 + (x & 3)
// (two 4x8 half boxes, or -- for windows that may wrap -- four 4x4 quadrant boxes in the order (qx, qy)).
__host__ __device__ constexpr int tok_y(int n) { return (n >> 2) & 7; }
__host__ __device__ constexpr int tok_x(int n) { return ((n >> 5) & 1) * 4 + (n & 3); }

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


}  // namespace srb
