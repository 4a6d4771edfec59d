// Bit-exact index remaps and layout conversions (HBM-bound; 128-bit accesses on both sides):
// pixel_shuffle / unshuffle (NCHW and NHWC), window partition / reverse fused with the cyclic
// shift, roll, NCHW fp32 <-> NHWC bf16 entry/exit, weight pack / unpack, column sums.
#include <cuda_bf16.h>

#include "host_util.h"
#include "ptx.cuh"

namespace srb {

static inline int grid_for(size_t work, int block) {
  size_t g = (work + block - 1) / block;
  return static_cast<int>(g > 0x7FFFFFFF ? 0x7FFFFFFF : (g == 0 ? 1 : g));
}

// ------------------------------------------------------------------ pixel shuffle, NCHW
// thread: VEC consecutive x of one (b, c, y); moves r*r*VEC elements with 16-byte accesses.
template <typename T, int R, bool INVERSE>
__global__ void pixel_shuffle_nchw_vec(const T* __restrict__ in, T* __restrict__ out, int B, int C,
                                       int H, int W) {
  constexpr int VEC = 16 / sizeof(T);
  const int Wv = W / VEC;
  const size_t total = static_cast<size_t>(B) * C * H * Wv;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int xv = static_cast<int>(idx % Wv);
    const int y = static_cast<int>((idx / Wv) % H);
    const int c = static_cast<int>((idx / (static_cast<size_t>(Wv) * H)) % C);
    const int b = static_cast<int>(idx / (static_cast<size_t>(Wv) * H * C));
    const int x = xv * VEC;
    const size_t plane = static_cast<size_t>(H) * W;
    // "small" side: [B, C*R*R, H, W]; "big" side: [B, C, H*R, W*R]
    const T* sm_base = (INVERSE ? out : in) +
                       (static_cast<size_t>(b) * C * R * R + static_cast<size_t>(c) * R * R) * plane +
                       static_cast<size_t>(y) * W + x;
    const T* bg_base = (INVERSE ? in : out) +
                       ((static_cast<size_t>(b) * C + c) * H * R + static_cast<size_t>(y) * R) *
                           (static_cast<size_t>(W) * R) +
                       static_cast<size_t>(x) * R;
    T big[R][VEC * R];
    if (!INVERSE) {
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          T v[VEC];
          *reinterpret_cast<uint4*>(v) =
              __ldg(reinterpret_cast<const uint4*>(sm_base + (i * R + j) * plane));
#pragma unroll
          for (int e = 0; e < VEC; ++e) big[i][e * R + j] = v[e];
        }
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int q = 0; q < R; ++q)
          *reinterpret_cast<uint4*>(const_cast<T*>(bg_base) + static_cast<size_t>(i) * W * R +
                                    q * VEC) = *reinterpret_cast<uint4*>(&big[i][q * VEC]);
    } else {
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int q = 0; q < R; ++q)
          *reinterpret_cast<uint4*>(&big[i][q * VEC]) = __ldg(reinterpret_cast<const uint4*>(
              bg_base + static_cast<size_t>(i) * W * R + q * VEC));
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          T v[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e) v[e] = big[i][e * R + j];
          *reinterpret_cast<uint4*>(const_cast<T*>(sm_base) + (i * R + j) * plane) =
              *reinterpret_cast<uint4*>(v);
        }
    }
  }
}

template <typename T, bool INVERSE>
__global__ void pixel_shuffle_nchw_scalar(const T* __restrict__ in, T* __restrict__ out, int B,
                                          int C, int H, int W, int R) {
  const size_t total = static_cast<size_t>(B) * C * H * R * W * R;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % (W * R));
    const int Y = static_cast<int>((idx / (W * R)) % (H * R));
    const int c = static_cast<int>((idx / (static_cast<size_t>(W) * R * H * R)) % C);
    const int b = static_cast<int>(idx / (static_cast<size_t>(W) * R * H * R * C));
    const int y = Y / R, i = Y % R, x = X / R, j = X % R;
    const size_t small =
        ((static_cast<size_t>(b) * C * R * R + static_cast<size_t>(c) * R * R + i * R + j) * H + y) *
            W +
        x;
    if (!INVERSE)
      out[idx] = in[small];
    else
      out[small] = in[idx];
  }
}

// ------------------------------------------------------------------ pixel shuffle, NHWC
// small: [B,H,W,C*R*R] (channel = c*R*R + i*R + j); big: [B,H*R,W*R,C].
// thread: VEC consecutive c of one pixel -> R*R 16-byte loads, R*R 16-byte stores.
template <typename T, int R, bool INVERSE>
__global__ void pixel_shuffle_nhwc_vec(const T* __restrict__ in, T* __restrict__ out, int B, int C,
                                       int H, int W) {
  constexpr int VEC = 16 / sizeof(T);
  constexpr int R2 = R * R;
  const int Cv = C / VEC;
  const size_t total = static_cast<size_t>(B) * H * W * Cv;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % Cv);
    const size_t pix = idx / Cv;
    const int x = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<size_t>(W) * H));
    const T* sm = (INVERSE ? out : in) + pix * (static_cast<size_t>(C) * R2) +
                  static_cast<size_t>(cv) * VEC * R2;
    const T* bg = (INVERSE ? in : out) +
                  ((static_cast<size_t>(b) * H * R + static_cast<size_t>(y) * R) *
                       (static_cast<size_t>(W) * R) +
                   static_cast<size_t>(x) * R) *
                      C +
                  static_cast<size_t>(cv) * VEC;
    T s[VEC * R2];
    if (!INVERSE) {
#pragma unroll
      for (int q = 0; q < R2; ++q)
        *reinterpret_cast<uint4*>(&s[q * VEC]) = __ldg(reinterpret_cast<const uint4*>(sm) + q);
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          T v[VEC];
#pragma unroll
          for (int e = 0; e < VEC; ++e) v[e] = s[e * R2 + i * R + j];
          *reinterpret_cast<uint4*>(const_cast<T*>(bg) +
                                    (static_cast<size_t>(i) * W * R + j) * C) =
              *reinterpret_cast<uint4*>(v);
        }
    } else {
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          T v[VEC];
          *reinterpret_cast<uint4*>(v) = __ldg(
              reinterpret_cast<const uint4*>(bg + (static_cast<size_t>(i) * W * R + j) * C));
#pragma unroll
          for (int e = 0; e < VEC; ++e) s[e * R2 + i * R + j] = v[e];
        }
#pragma unroll
      for (int q = 0; q < R2; ++q)
        *(reinterpret_cast<uint4*>(const_cast<T*>(sm)) + q) = *reinterpret_cast<uint4*>(&s[q * VEC]);
    }
  }
}

template <typename T, bool INVERSE>
__global__ void pixel_shuffle_nhwc_scalar(const T* __restrict__ in, T* __restrict__ out, int B,
                                          int C, int H, int W, int R) {
  const size_t total = static_cast<size_t>(B) * H * R * W * R * C;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const int X = static_cast<int>((idx / C) % (W * R));
    const int Y = static_cast<int>((idx / (static_cast<size_t>(C) * W * R)) % (H * R));
    const int b = static_cast<int>(idx / (static_cast<size_t>(C) * W * R * H * R));
    const int y = Y / R, i = Y % R, x = X / R, j = X % R;
    const size_t small = ((static_cast<size_t>(b) * H + y) * W + x) * (static_cast<size_t>(C) * R * R) +
                         static_cast<size_t>(c) * R * R + i * R + j;
    if (!INVERSE)
      out[idx] = in[small];
    else
      out[small] = in[idx];
  }
}

// ------------------------------------------------------------------ token-row remaps
// Both window partition(+shift) and roll move whole C-element token rows; VT is the access type.
// map_mode 0: window remap, 1: roll.
template <typename VT>
__global__ void token_remap(const VT* __restrict__ in, VT* __restrict__ out, int B, int H, int W,
                            int Cv /* row length in VT units */, int ws, int s0, int s1,
                            int map_mode, int inverse) {
  const size_t total = static_cast<size_t>(B) * H * W * Cv;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(idx % Cv);
    const size_t tok = idx / Cv;  // index on the "destination-ordered" side (see below)
    size_t img_tok, other_tok;
    if (map_mode == 0) {
      // tok enumerates window order: ((b*nWh + wy)*nWw + wx)*ws*ws + iy*ws + ix
      const int nWw = W / ws, nWh = H / ws;
      const int ix = static_cast<int>(tok % ws);
      const int iy = static_cast<int>((tok / ws) % ws);
      const int wx = static_cast<int>((tok / (ws * ws)) % nWw);
      const int wy = static_cast<int>((tok / (static_cast<size_t>(ws) * ws * nWw)) % nWh);
      const int b = static_cast<int>(tok / (static_cast<size_t>(ws) * ws * nWw * nWh));
      int yy = wy * ws + iy + s0;
      int xx = wx * ws + ix + s0;
      if (yy >= H) yy -= H;
      if (xx >= W) xx -= W;
      img_tok = (static_cast<size_t>(b) * H + yy) * W + xx;
      other_tok = tok;
      // forward: win[tok] = img[img_tok]; inverse: img[img_tok] = win[tok]
      if (!inverse)
        out[other_tok * Cv + cv] = __ldg(in + img_tok * Cv + cv);
      else
        out[img_tok * Cv + cv] = __ldg(in + other_tok * Cv + cv);
    } else {
      // roll: out[b, (y+s0)%H, (x+s1)%W] = in[b, y, x]; tok enumerates the output
      const int x = static_cast<int>(tok % W);
      const int y = static_cast<int>((tok / W) % H);
      const int b = static_cast<int>(tok / (static_cast<size_t>(W) * H));
      int ys = y - s0, xs = x - s1;
      ys %= H;
      if (ys < 0) ys += H;
      xs %= W;
      if (xs < 0) xs += W;
      img_tok = (static_cast<size_t>(b) * H + ys) * W + xs;
      out[tok * Cv + cv] = __ldg(in + img_tok * Cv + cv);
    }
  }
}

// ------------------------------------------------------------------ NCHW fp32 <-> NHWC bf16
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                    int B, int C, int H, int W, int Cp,
                                    const float* __restrict__ shift, float scale) {
  const int groups = Cp / 8;
  const size_t npix = static_cast<size_t>(B) * H * W;
  const size_t total = npix * groups;
  const size_t plane = static_cast<size_t>(H) * W;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = idx % npix;  // pixel fastest: coalesced plane reads
    const int g = static_cast<int>(idx / npix);
    const size_t b = pix / plane, yx = pix % plane;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = g * 8 + e;
      v[e] = 0.0f;
      if (c < C) {
        const float sh = shift != nullptr ? __ldg(shift + c) : 0.0f;
        v[e] = (__ldg(in + (b * C + c) * plane + yx) - sh) * scale;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + pix * Cp + g * 8) = o;
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                    int B, int C, int H, int W, int Cp,
                                    const float* __restrict__ shift, float scale) {
  const size_t npix = static_cast<size_t>(B) * H * W;
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = npix * C;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t pix = idx % npix;
    const int c = static_cast<int>(idx / npix);
    const size_t b = pix / plane, yx = pix % plane;
    const float sh = shift != nullptr ? __ldg(shift + c) : 0.0f;
    out[(b * C + c) * plane + yx] = __bfloat162float(in[pix * Cp + c]) * scale + sh;
  }
}

// ------------------------------------------------------------------ nearest-neighbour x2 (SwinIR 'nearest+conv')
// torch.nn.functional.interpolate(x, scale_factor=2, mode='nearest') on NHWC bf16 (swinir_arch.py:911-912):
// out[b, Y, X, :] = in[b, Y/2, X/2, :] -- a pure index remap (bit-exact); backward sums each 2x2 block in fp32.
__global__ void nearest_up2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W,
                                   int vec) {
  const size_t total = static_cast<size_t>(B) * (2 * H) * (2 * W) * vec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vec);
    size_t pix = idx / vec;
    const int X = static_cast<int>(pix % (2 * W));
    pix /= (2 * W);
    const int Y = static_cast<int>(pix % (2 * H));
    const size_t b = pix / (2 * H);
    out[idx] = __ldg(in + ((b * H + (Y >> 1)) * W + (X >> 1)) * vec + v);
  }
}

__global__ void nearest_up2_bwd_kernel(const uint4* __restrict__ g, uint4* __restrict__ out, int B, int H, int W,
                                       int vec) {
  const size_t total = static_cast<size_t>(B) * H * W * vec;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vec);
    size_t pix = idx / vec;
    const int x = static_cast<int>(pix % W);
    pix /= W;
    const int y = static_cast<int>(pix % H);
    const size_t b = pix / H;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint4 m = __ldg(g + ((b * 2 * H + 2 * y + i) * (2 * W) + 2 * x + j) * vec + v);
        acc[0] += bf16_lo(m.x);
        acc[1] += bf16_hi(m.x);
        acc[2] += bf16_lo(m.y);
        acc[3] += bf16_hi(m.y);
        acc[4] += bf16_lo(m.z);
        acc[5] += bf16_hi(m.z);
        acc[6] += bf16_lo(m.w);
        acc[7] += bf16_hi(m.w);
      }
    out[idx] = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                          pack_bf16x2(acc[6], acc[7]));
  }
}

// ------------------------------------------------------------------ few-channel exit conv (conv_last, F -> 3)
// A 3x3 conv with C <= 7 output channels is HBM-bound on its INPUT, yet as a 9-tap implicit GEMM it re-reads every
// input pixel nine times through L2 (EDSR-L: 2.7 GB of L2->SM traffic for a 302 MB tensor, 330 us).  It is linear:
//   out[c](p) = sum_tap (X . Wtap[c])(p + tap)
// so ONE 1x1 GEMM T[p][tap*C + c] = X[p] . W[c][:][tap] reads X once, and this kernel adds the nine shifted taps
// of the tiny T tensor (zero outside the image, exactly the conv's zero padding), the bias and the image affine.
__global__ void tap_stencil_kernel(const float* __restrict__ t, float* __restrict__ out, int B, int C, int H, int W,
                                   int Tp, const float* __restrict__ bias, const float* __restrict__ shift,
                                   float scale) {
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * plane;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = pix / plane;
    const int yx = static_cast<int>(pix - b * plane);
    const int y = yx / W, x = yx - y * W;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.0f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const float* src = t + ((b * H + yy) * W + xx) * Tp + tap * C;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < C) acc[c] += __ldg(src + c);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) {
        const float bv = bias != nullptr ? __ldg(bias + c) : 0.0f;
        const float sh = shift != nullptr ? __ldg(shift + c) : 0.0f;
        out[(b * C + c) * plane + yx] = (acc[c] + bv) * scale + sh;
      }
  }
}

// Backward twin: G[p][tap*C + c] = scale * g[c](p - tap) (zero outside the image; channels >= 9C are zero), bf16.
// With it the data gradient dX = G . Wf and the weight gradient dWf = G^T X are single-tap GEMMs as well.
template <int C>
__global__ void tap_im2col_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ out, int B, int H, int W,
                                  int Gp, float scale) {
  // thread = one pixel: gathers its 3x3xC neighbourhood (coalesced along x across the warp, L1-resident re-use) and
  // writes the pixel's whole Gp-channel row (full 128-byte lines)
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * plane;
  for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = pix / plane;
    const int yx = static_cast<int>(pix - b * plane);
    const int y = yx / W, x = yx - y * W;
    float v[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = 0.0f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y - (tap / 3 - 1), xx = x - (tap % 3 - 1);
      const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (in) v[tap * C + c] = scale * __ldg(g + (b * C + c) * plane + yy * W + xx);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + pix * Gp);
    for (int grp = 0; grp < Gp / 8; ++grp) {
      uint4 o = make_uint4(0u, 0u, 0u, 0u);
      if (grp < 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (q == grp) {
            o.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
            o.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
            o.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
            o.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
          }
      }
      dst[grp] = o;
    }
  }
}

// ------------------------------------------------------------------ weight pack / unpack
__global__ void pack_weight_kernel(const float* __restrict__ w, int Co, int Ci, int taps,
                                   const int32_t* __restrict__ perm_out, int Np,
                                   const int32_t* __restrict__ perm_in, int Kp, int transpose,
                                   __nv_bfloat16* __restrict__ out) {
  // thread = one (n, k) pair, all taps: the `taps` source floats are contiguous (OIHW), and for a fixed
  // tap consecutive threads write consecutive bf16 (k fastest, or n fastest when transposing).
  const size_t pairs = static_cast<size_t>(Np) * Kp;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < pairs;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int n, k;
    if (!transpose) {
      n = static_cast<int>(idx / Kp);
      k = static_cast<int>(idx % Kp);
    } else {
      k = static_cast<int>(idx / Np);
      n = static_cast<int>(idx % Np);
    }
    const int o = perm_out != nullptr ? perm_out[n] : (n < Co ? n : -1);
    const int i = perm_in != nullptr ? perm_in[k] : (k < Ci ? k : -1);
    const bool live = (o >= 0 && i >= 0);
    const float* src = w + (static_cast<size_t>(live ? o : 0) * Ci + (live ? i : 0)) * taps;
    for (int t = 0; t < taps; ++t)
      out[static_cast<size_t>(t) * pairs + idx] = __float2bfloat16_rn(live ? __ldg(src + t) : 0.0f);
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ acc, float* __restrict__ gw, int Co,
                                    int Ci, int taps, const int32_t* __restrict__ perm_out, int Np,
                                    const int32_t* __restrict__ perm_in, int Kp, float alpha) {
  const size_t pairs = static_cast<size_t>(Np) * Kp;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < pairs;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx / Kp);
    const int k = static_cast<int>(idx % Kp);
    const int o = perm_out != nullptr ? perm_out[n] : (n < Co ? n : -1);
    const int i = perm_in != nullptr ? perm_in[k] : (k < Ci ? k : -1);
    if (o < 0 || i < 0) continue;
    float* dst = gw + (static_cast<size_t>(o) * Ci + i) * taps;
    for (int t = 0; t < taps; ++t) dst[t] = alpha * __ldg(acc + static_cast<size_t>(t) * pairs + idx);
  }
}

// Batched variants: ONE launch packs (or unpacks) every weight of a network.  `items` is a device table; item i
// owns the 256-pair chunks [chunk_begin_i, chunk_begin_{i+1}) of the flattened work list, so a block never
// straddles two items and finds its item with a binary search over `chunk_begin`.
__device__ __forceinline__ int find_item(const srb200_pack_item* __restrict__ items, int n_items, long long chunk) {
  int lo = 0, hi = n_items - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&items[mid].chunk_begin) <= chunk) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) pack_weights_kernel(const srb200_pack_item* __restrict__ items, int n_items,
                                                           long long total_chunks) {
  __shared__ int s_item;
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    if (threadIdx.x == 0) s_item = find_item(items, n_items, chunk);
    __syncthreads();
    const srb200_pack_item it = items[s_item];
    __syncthreads();
    const int Np = static_cast<int>(it.Np), Kp = static_cast<int>(it.Kp), Co = static_cast<int>(it.Co),
              Ci = static_cast<int>(it.Ci), taps = static_cast<int>(it.taps);
    const size_t pairs = static_cast<size_t>(Np) * Kp;
    const size_t idx = static_cast<size_t>(chunk - it.chunk_begin) * 256 + threadIdx.x;
    if (idx >= pairs) continue;
    int n, k;
    if (!it.transpose) {
      n = static_cast<int>(idx / Kp);
      k = static_cast<int>(idx % Kp);
    } else {
      k = static_cast<int>(idx / Np);
      n = static_cast<int>(idx % Np);
    }
    const int32_t* perm_out = reinterpret_cast<const int32_t*>(it.perm_out);
    const int32_t* perm_in = reinterpret_cast<const int32_t*>(it.perm_in);
    const int o = perm_out != nullptr ? perm_out[n] : (n < Co ? n : -1);
    const int i = perm_in != nullptr ? perm_in[k] : (k < Ci ? k : -1);
    const bool live = (o >= 0 && i >= 0);
    const float* src = reinterpret_cast<const float*>(it.src) + (static_cast<size_t>(live ? o : 0) * Ci + (live ? i : 0)) * taps;
    if (it.transpose == 2) {  // fp32 copy (padded / permuted bias vectors ride in the same launch)
      float* outf = reinterpret_cast<float*>(it.dst);
      // (reserved = 1 + index of ONE pad element that gets the constant `alpha` instead of zero: e.g. the bias that
      // makes a pad channel of a GELU output a constant one, see srb200_pack_item)
      const float pad = (it.reserved > 0 && idx == static_cast<size_t>(it.reserved - 1)) ? it.alpha : 0.0f;
      for (int t = 0; t < taps; ++t) outf[static_cast<size_t>(t) * pairs + idx] = live ? __ldg(src + t) : pad;
      continue;
    }
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(it.dst);
    for (int t = 0; t < taps; ++t)
      out[static_cast<size_t>(t) * pairs + idx] = __float2bfloat16_rn(live ? __ldg(src + t) : 0.0f);
  }
}

__global__ void __launch_bounds__(256) unpack_wgrads_kernel(const srb200_pack_item* __restrict__ items, int n_items,
                                                            long long total_chunks) {
  __shared__ int s_item;
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    if (threadIdx.x == 0) s_item = find_item(items, n_items, chunk);
    __syncthreads();
    const srb200_pack_item it = items[s_item];
    __syncthreads();
    const int Np = static_cast<int>(it.Np), Kp = static_cast<int>(it.Kp), Co = static_cast<int>(it.Co),
              Ci = static_cast<int>(it.Ci), taps = static_cast<int>(it.taps);
    const size_t pairs = static_cast<size_t>(Np) * Kp;
    const size_t idx = static_cast<size_t>(chunk - it.chunk_begin) * 256 + threadIdx.x;
    if (idx >= pairs) continue;
    const int n = static_cast<int>(idx / Kp);
    const int k = static_cast<int>(idx % Kp);
    const int32_t* perm_out = reinterpret_cast<const int32_t*>(it.perm_out);
    const int32_t* perm_in = reinterpret_cast<const int32_t*>(it.perm_in);
    const int o = perm_out != nullptr ? perm_out[n] : (n < Co ? n : -1);
    const int i = perm_in != nullptr ? perm_in[k] : (k < Ci ? k : -1);
    if (o < 0 || i < 0) continue;
    const float* acc = reinterpret_cast<const float*>(it.src);
    float* dst = reinterpret_cast<float*>(it.dst) + (static_cast<size_t>(o) * Ci + i) * taps;
    for (int t = 0; t < taps; ++t) dst[t] = it.alpha * __ldg(acc + static_cast<size_t>(t) * pairs + idx);
  }
}

// ------------------------------------------------------------------ multi-tensor  dst = a*dst + b*src  (EMA)
// BaseModel.model_ema (basicsr/models/base_model.py:75-82) walks named_parameters() and issues two tiny kernels
// per tensor -- 3260 launches per step for RCAN's 1630 parameters.  One launch over a device table instead.
__global__ void __launch_bounds__(256) multi_axpby_kernel(const srb200_vec_item* __restrict__ items, int n_items,
                                                          long long total_chunks, float a, float b) {
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
    const srb200_vec_item it = items[s_item];
    __syncthreads();
    const long long base = (chunk - it.chunk_begin) * 1024;
    const float* src = static_cast<const float*>(it.src);
    float* dst = static_cast<float*>(it.dst);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = base + u * 256 + threadIdx.x;
      if (i < it.n) dst[i] = a * dst[i] + b * __ldg(src + i);
    }
  }
}

// ------------------------------------------------------------------ column sums (bias gradient)
// dy [rows, C] bf16 -> out[view*C + c] += sum; view = pixel-unshuffle phase of the row (R > 1).
// thread = one 8-channel group x one row lane; R*R register accumulator sets, one smem reduction and
// one global atomic per (block, column).
template <int R>
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ out,
                              long long rows, int C, int Wf, int rows_per_block) {
  pdl_trigger();  // a following programmatically serialized GEMM may start its prologue
  extern __shared__ float s_sum[];  // [R*R*C]
  constexpr int V = R * R;
  const int nsum = V * C;
  for (int i = threadIdx.x; i < nsum; i += blockDim.x) s_sum[i] = 0.0f;
  __syncthreads();
  const int groups = C / 8;
  const int g = threadIdx.x % groups;
  const int rl = threadIdx.x / groups;
  const int lanes = blockDim.x / groups;
  const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  if (rl < lanes) {
    float a[V][8];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) a[v][e] = 0.0f;
    const __nv_bfloat16* base = dy + g * 8;
    auto add = [&](const uint4& m, long long row) {
      int view = 0;
      if (R > 1) {
        const int X = static_cast<int>(row % Wf);
        const int Y = static_cast<int>(row / Wf);
        view = (Y % R) * R + (X % R);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float w = (v == view) ? 1.0f : 0.0f;  // predicated adds keep a[][] in registers
        a[v][0] += w * bf16_lo(m.x);
        a[v][1] += w * bf16_hi(m.x);
        a[v][2] += w * bf16_lo(m.y);
        a[v][3] += w * bf16_hi(m.y);
        a[v][4] += w * bf16_lo(m.z);
        a[v][5] += w * bf16_hi(m.z);
        a[v][6] += w * bf16_lo(m.w);
        a[v][7] += w * bf16_hi(m.w);
      }
    };
    long long row = r0 + rl;
    for (; row + 3LL * lanes < r1; row += 4LL * lanes) {  // 4 independent 16-byte loads in flight
      uint4 m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        m[u] = __ldg(reinterpret_cast<const uint4*>(base + (row + static_cast<long long>(u) * lanes) * C));
#pragma unroll
      for (int u = 0; u < 4; ++u) add(m[u], row + static_cast<long long>(u) * lanes);
    }
    for (; row < r1; row += lanes) add(__ldg(reinterpret_cast<const uint4*>(base + row * C)), row);
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(&s_sum[v * C + g * 8 + e], a[v][e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nsum; i += blockDim.x) atomicAdd(out + i, s_sum[i]);
}

// ------------------------------------------------------------------ activation backward
// out = g * act'(y): y is the (post-activation) forward output, so sign(y) == sign(pre-activation)
// for ReLU / LeakyReLU (arch_util.py:81, swinir_arch.py:836).
__global__ void act_bwd_kernel(const uint4* __restrict__ g, const uint4* __restrict__ y,
                               uint4* __restrict__ out, size_t nvec, float slope) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < nvec;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const uint4 gv = __ldg(g + idx), yv = __ldg(y + idx);
    const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w};
    uint32_t ow[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a = bf16_lo(gw[q]) * (bf16_lo(yw[q]) > 0.0f ? 1.0f : slope);
      const float b = bf16_hi(gw[q]) * (bf16_hi(yw[q]) > 0.0f ? 1.0f : slope);
      ow[q] = pack_bf16x2(a, b);
    }
    out[idx] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

template <typename T>
static int dispatch_ps_nchw(const void* in, void* out, int B, int C, int H, int W, int r,
                            int inverse, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (r == 2 || r == 3) && (W % VEC == 0) &&
                      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  const T* i = static_cast<const T*>(in);
  T* o = static_cast<T*>(out);
  if (vec_ok) {
    const size_t work = static_cast<size_t>(B) * C * H * (W / VEC);
    const int g = grid_for(work, 256);
    if (r == 2) {
      if (!inverse) pixel_shuffle_nchw_vec<T, 2, false><<<g, 256, 0, st>>>(i, o, B, C, H, W);
      else pixel_shuffle_nchw_vec<T, 2, true><<<g, 256, 0, st>>>(i, o, B, C, H, W);
    } else {
      if (!inverse) pixel_shuffle_nchw_vec<T, 3, false><<<g, 256, 0, st>>>(i, o, B, C, H, W);
      else pixel_shuffle_nchw_vec<T, 3, true><<<g, 256, 0, st>>>(i, o, B, C, H, W);
    }
  } else {
    const size_t work = static_cast<size_t>(B) * C * H * r * W * r;
    const int g = grid_for(work, 256);
    if (!inverse) pixel_shuffle_nchw_scalar<T, false><<<g, 256, 0, st>>>(i, o, B, C, H, W, r);
    else pixel_shuffle_nchw_scalar<T, true><<<g, 256, 0, st>>>(i, o, B, C, H, W, r);
  }
  return launch_status();
}

template <typename T>
static int dispatch_ps_nhwc(const void* in, void* out, int B, int C, int H, int W, int r,
                            int inverse, cudaStream_t st) {
  constexpr int VEC = 16 / sizeof(T);
  const bool vec_ok = (r == 2 || r == 3) && (C % VEC == 0) &&
                      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  const T* i = static_cast<const T*>(in);
  T* o = static_cast<T*>(out);
  if (vec_ok) {
    const size_t work = static_cast<size_t>(B) * H * W * (C / VEC);
    const int g = grid_for(work, 256);
    if (r == 2) {
      if (!inverse) pixel_shuffle_nhwc_vec<T, 2, false><<<g, 256, 0, st>>>(i, o, B, C, H, W);
      else pixel_shuffle_nhwc_vec<T, 2, true><<<g, 256, 0, st>>>(i, o, B, C, H, W);
    } else {
      if (!inverse) pixel_shuffle_nhwc_vec<T, 3, false><<<g, 256, 0, st>>>(i, o, B, C, H, W);
      else pixel_shuffle_nhwc_vec<T, 3, true><<<g, 256, 0, st>>>(i, o, B, C, H, W);
    }
  } else {
    const size_t work = static_cast<size_t>(B) * C * H * r * W * r;
    const int g = grid_for(work, 256);
    if (!inverse) pixel_shuffle_nhwc_scalar<T, false><<<g, 256, 0, st>>>(i, o, B, C, H, W, r);
    else pixel_shuffle_nhwc_scalar<T, true><<<g, 256, 0, st>>>(i, o, B, C, H, W, r);
  }
  return launch_status();
}

static int dispatch_token_remap(const void* in, void* out, int B, int H, int W, int C,
                                int elem_bytes, int ws, int s0, int s1, int mode, int inverse,
                                cudaStream_t st) {
  const size_t row_bytes = static_cast<size_t>(C) * elem_bytes;
  const bool a16 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (row_bytes % 16 == 0 && a16) {
    const int Cv = static_cast<int>(row_bytes / 16);
    const int g = grid_for(static_cast<size_t>(B) * H * W * Cv, 256);
    token_remap<uint4><<<g, 256, 0, st>>>(static_cast<const uint4*>(in), static_cast<uint4*>(out), B,
                                          H, W, Cv, ws, s0, s1, mode, inverse);
  } else if (row_bytes % 4 == 0) {
    const int Cv = static_cast<int>(row_bytes / 4);
    const int g = grid_for(static_cast<size_t>(B) * H * W * Cv, 256);
    token_remap<uint32_t><<<g, 256, 0, st>>>(static_cast<const uint32_t*>(in),
                                             static_cast<uint32_t*>(out), B, H, W, Cv, ws, s0, s1,
                                             mode, inverse);
  } else {
    const int Cv = static_cast<int>(row_bytes / 2);
    const int g = grid_for(static_cast<size_t>(B) * H * W * Cv, 256);
    token_remap<uint16_t><<<g, 256, 0, st>>>(static_cast<const uint16_t*>(in),
                                             static_cast<uint16_t*>(out), B, H, W, Cv, ws, s0, s1,
                                             mode, inverse);
  }
  return launch_status();
}

}  // namespace srb

using namespace srb;

extern "C" int srb200_pixel_shuffle_nchw(const void* in, void* out, int B, int C_out, int H, int W,
                                         int r, int elem_bytes, int inverse,
                                         srb200_stream_t stream) {
  if (!in || !out || B <= 0 || C_out <= 0 || H <= 0 || W <= 0 || r < 1) return SRB200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 2) return dispatch_ps_nchw<uint16_t>(in, out, B, C_out, H, W, r, inverse, st);
  if (elem_bytes == 4) return dispatch_ps_nchw<uint32_t>(in, out, B, C_out, H, W, r, inverse, st);
  return SRB200_EINVAL;
}

extern "C" int srb200_pixel_shuffle_nhwc(const void* in, void* out, int B, int C_out, int H, int W,
                                         int r, int elem_bytes, int inverse,
                                         srb200_stream_t stream) {
  if (!in || !out || B <= 0 || C_out <= 0 || H <= 0 || W <= 0 || r < 1) return SRB200_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 2) return dispatch_ps_nhwc<uint16_t>(in, out, B, C_out, H, W, r, inverse, st);
  if (elem_bytes == 4) return dispatch_ps_nhwc<uint32_t>(in, out, B, C_out, H, W, r, inverse, st);
  return SRB200_EINVAL;
}

extern "C" int srb200_window_remap(const void* in, void* out, int B, int H, int W, int C, int ws,
                                   int shift, int elem_bytes, int inverse, srb200_stream_t stream) {
  if (!in || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0 || ws <= 0) return SRB200_EINVAL;
  if (H % ws != 0 || W % ws != 0 || shift < 0 || shift >= ws) return SRB200_EINVAL;
  if (elem_bytes != 2 && elem_bytes != 4) return SRB200_EINVAL;
  return dispatch_token_remap(in, out, B, H, W, C, elem_bytes, ws, shift, shift, 0, inverse,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int srb200_roll_nhwc(const void* in, void* out, int B, int H, int W, int C, int sy,
                                int sx, int elem_bytes, srb200_stream_t stream) {
  if (!in || !out || B <= 0 || H <= 0 || W <= 0 || C <= 0) return SRB200_EINVAL;
  if (elem_bytes != 2 && elem_bytes != 4) return SRB200_EINVAL;
  return dispatch_token_remap(in, out, B, H, W, C, elem_bytes, 1, sy, sx, 1, 0,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int srb200_nchw_to_nhwc(const float* in, void* out_bf16, int B, int C, int H, int W,
                                   int C_pad, const float* shift, float scale,
                                   srb200_stream_t stream) {
  if (!in || !out_bf16 || B <= 0 || C <= 0 || H <= 0 || W <= 0) return SRB200_EINVAL;
  if (C_pad < C || C_pad % 8 != 0 || (reinterpret_cast<uintptr_t>(out_bf16) & 15u)) return SRB200_EINVAL;
  const size_t work = static_cast<size_t>(B) * H * W * (C_pad / 8);
  nchw_to_nhwc_kernel<<<grid_for(work, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      in, static_cast<__nv_bfloat16*>(out_bf16), B, C, H, W, C_pad, shift, scale);
  return launch_status();
}

extern "C" int srb200_nhwc_to_nchw(const void* in_bf16, float* out, int B, int C, int H, int W,
                                   int C_pad, const float* shift, float scale,
                                   srb200_stream_t stream) {
  if (!in_bf16 || !out || B <= 0 || C <= 0 || H <= 0 || W <= 0 || C_pad < C) return SRB200_EINVAL;
  const size_t work = static_cast<size_t>(B) * H * W * C;
  nhwc_to_nchw_kernel<<<grid_for(work, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_bf16), out, B, C, H, W, C_pad, shift, scale);
  return launch_status();
}

extern "C" int srb200_nearest_up2(const void* in_bf16, void* out_bf16, int B, int H, int W, int C, int inverse,
                                  srb200_stream_t stream) {
  if (!in_bf16 || !out_bf16 || B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 != 0) return SRB200_EINVAL;
  if ((reinterpret_cast<uintptr_t>(in_bf16) | reinterpret_cast<uintptr_t>(out_bf16)) & 15u) return SRB200_EINVAL;
  const int vec = C / 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!inverse) {
    const size_t work = static_cast<size_t>(B) * 4 * H * W * vec;
    nearest_up2_kernel<<<grid_for(work, 256), 256, 0, st>>>(static_cast<const uint4*>(in_bf16),
                                                           static_cast<uint4*>(out_bf16), B, H, W, vec);
  } else {
    const size_t work = static_cast<size_t>(B) * H * W * vec;
    nearest_up2_bwd_kernel<<<grid_for(work, 256), 256, 0, st>>>(static_cast<const uint4*>(in_bf16),
                                                               static_cast<uint4*>(out_bf16), B, H, W, vec);
  }
  return launch_status();
}

extern "C" int srb200_tap_stencil(const float* t_f32, float* out, int B, int C, int H, int W, int Tp,
                                  const float* bias, const float* shift, float scale, srb200_stream_t stream) {
  if (!t_f32 || !out || B <= 0 || C <= 0 || C > 7 || H <= 0 || W <= 0 || Tp < 9 * C) return SRB200_EINVAL;
  const size_t work = static_cast<size_t>(B) * H * W;
  tap_stencil_kernel<<<grid_for(work, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(t_f32, out, B, C, H, W, Tp,
                                                                                        bias, shift, scale);
  return launch_status();
}

extern "C" int srb200_tap_im2col(const float* g, void* out_bf16, int B, int C, int H, int W, int Gp, float scale,
                                 srb200_stream_t stream) {
  if (!g || !out_bf16 || B <= 0 || C <= 0 || C > 7 || H <= 0 || W <= 0 || Gp < 9 * C || Gp % 8 != 0)
    return SRB200_EINVAL;
  if (reinterpret_cast<uintptr_t>(out_bf16) & 15u) return SRB200_EINVAL;
  const size_t work = static_cast<size_t>(B) * H * W;
  const int grid = grid_for(work, 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out_bf16);
  switch (C) {
    case 1: tap_im2col_kernel<1><<<grid, 256, 0, st>>>(g, o, B, H, W, Gp, scale); break;
    case 2: tap_im2col_kernel<2><<<grid, 256, 0, st>>>(g, o, B, H, W, Gp, scale); break;
    case 3: tap_im2col_kernel<3><<<grid, 256, 0, st>>>(g, o, B, H, W, Gp, scale); break;
    case 4: tap_im2col_kernel<4><<<grid, 256, 0, st>>>(g, o, B, H, W, Gp, scale); break;
    case 5: tap_im2col_kernel<5><<<grid, 256, 0, st>>>(g, o, B, H, W, Gp, scale); break;
    case 6: tap_im2col_kernel<6><<<grid, 256, 0, st>>>(g, o, B, H, W, Gp, scale); break;
    default: tap_im2col_kernel<7><<<grid, 256, 0, st>>>(g, o, B, H, W, Gp, scale); break;
  }
  return launch_status();
}

extern "C" int srb200_pack_weight(const float* w, int Co, int Ci, int taps,
                                  const int32_t* perm_out, int Np, const int32_t* perm_in, int Kp,
                                  int transpose, void* out_bf16, srb200_stream_t stream) {
  if (!w || !out_bf16 || Co <= 0 || Ci <= 0 || taps <= 0 || Np <= 0 || Kp <= 0) return SRB200_EINVAL;
  const size_t work = static_cast<size_t>(Np) * Kp;
  pack_weight_kernel<<<grid_for(work, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, Co, Ci, taps, perm_out, Np, perm_in, Kp, transpose, static_cast<__nv_bfloat16*>(out_bf16));
  return launch_status();
}

extern "C" int srb200_unpack_wgrad(const float* acc, float* gw, int Co, int Ci, int taps,
                                   const int32_t* perm_out, int Np, const int32_t* perm_in, int Kp,
                                   float alpha, srb200_stream_t stream) {
  if (!acc || !gw || Co <= 0 || Ci <= 0 || taps <= 0 || Np <= 0 || Kp <= 0) return SRB200_EINVAL;
  const size_t work = static_cast<size_t>(Np) * Kp;
  unpack_wgrad_kernel<<<grid_for(work, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      acc, gw, Co, Ci, taps, perm_out, Np, perm_in, Kp, alpha);
  return launch_status();
}

extern "C" int srb200_pack_weights(const srb200_pack_item* items_dev, int n_items, int64_t total_chunks,
                                   srb200_stream_t stream) {
  if (!items_dev || n_items <= 0 || total_chunks <= 0) return SRB200_EINVAL;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const int grid = static_cast<int>(total_chunks < cap ? total_chunks : cap);
  pack_weights_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(items_dev, n_items, total_chunks);
  return launch_status();
}

extern "C" int srb200_unpack_wgrads(const srb200_pack_item* items_dev, int n_items, int64_t total_chunks,
                                    srb200_stream_t stream) {
  if (!items_dev || n_items <= 0 || total_chunks <= 0) return SRB200_EINVAL;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const int grid = static_cast<int>(total_chunks < cap ? total_chunks : cap);
  unpack_wgrads_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(items_dev, n_items, total_chunks);
  return launch_status();
}

// Up to SRB200_MAX_INLINE_ITEMS items passed BY VALUE (kernel parameters): no device table, so the launch can be
// recorded into a CUDA graph and the per-layer backward needs ONE launch for all its weight and bias gradients
// (a bias gradient is the [Np x 1] case: src = fp32 column sums, taps = 1, Kp = Ci = 1).
struct InlineItems {
  srb200_pack_item it[SRB200_MAX_INLINE_ITEMS];
  int n;
};

__global__ void __launch_bounds__(256) unpack_inline_kernel(const __grid_constant__ InlineItems tab, long long total_chunks) {
  pdl_trigger();  // the next layer's first GEMM may start its prologue now
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    srb200_pack_item it = tab.it[0];  // statically indexed, predicated copies: the table stays in param space
#pragma unroll
    for (int j = 1; j < SRB200_MAX_INLINE_ITEMS; ++j)
      if (j < tab.n && tab.it[j].chunk_begin <= chunk) it = tab.it[j];
    const int Kp = static_cast<int>(it.Kp), Co = static_cast<int>(it.Co), Ci = static_cast<int>(it.Ci),
              taps = static_cast<int>(it.taps);
    const size_t pairs = static_cast<size_t>(it.Np) * Kp;
    const size_t idx = static_cast<size_t>(chunk - it.chunk_begin) * 256 + threadIdx.x;
    if (idx >= pairs) continue;
    const int n = static_cast<int>(idx / Kp);
    const int k = static_cast<int>(idx % Kp);
    const int o = it.perm_out != nullptr ? it.perm_out[n] : (n < Co ? n : -1);
    const int ii = it.perm_in != nullptr ? it.perm_in[k] : (k < Ci ? k : -1);
    if (o < 0 || ii < 0) continue;
    const float* acc = reinterpret_cast<const float*>(it.src);
    float* dst = reinterpret_cast<float*>(it.dst) + (static_cast<size_t>(o) * Ci + ii) * taps;
    for (int t = 0; t < taps; ++t) dst[t] = it.alpha * __ldg(acc + static_cast<size_t>(t) * pairs + idx);
  }
}

extern "C" int srb200_unpack_wgrads_inline(const srb200_pack_item* items_host, int n_items,
                                           srb200_stream_t stream) {
  if (!items_host || n_items <= 0 || n_items > SRB200_MAX_INLINE_ITEMS) return SRB200_EINVAL;
  InlineItems tab;
  tab.n = n_items;
  long long chunk = 0;
  for (int i = 0; i < n_items; ++i) {
    tab.it[i] = items_host[i];
    if (!tab.it[i].src || !tab.it[i].dst || tab.it[i].Np <= 0 || tab.it[i].Kp <= 0) return SRB200_EINVAL;
    tab.it[i].chunk_begin = chunk;
    chunk += (tab.it[i].Np * tab.it[i].Kp + 255) / 256;
  }
  for (int i = n_items; i < SRB200_MAX_INLINE_ITEMS; ++i) tab.it[i] = tab.it[0];
  const long long cap = static_cast<long long>(num_sms()) * 8;
  const int grid = static_cast<int>(chunk < cap ? chunk : cap);
  unpack_inline_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(tab, chunk);
  return launch_status();
}

extern "C" int srb200_multi_axpby(const srb200_vec_item* items_dev, int n_items, int64_t total_chunks, float a,
                                  float b, srb200_stream_t stream) {
  if (!items_dev || n_items <= 0 || total_chunks <= 0) return SRB200_EINVAL;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  const int grid = static_cast<int>(total_chunks < cap ? total_chunks : cap);
  multi_axpby_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(items_dev, n_items, total_chunks, a, b);
  return launch_status();
}

extern "C" int srb200_colsum(const void* dy_bf16, float* out, int64_t rows, int C, int r, int Wf,
                             srb200_stream_t stream) {
  if (!dy_bf16 || !out || rows <= 0 || C <= 0 || C % 8 != 0 || r < 1 || r > 3) return SRB200_EINVAL;
  if (C / 8 > 256 || (r > 1 && Wf <= 0)) return SRB200_EINVAL;
  // ~2 blocks per SM, each covering a contiguous slab of rows: every block ends with one global atomic per
  // column onto the SAME C addresses, so the block count (not the row count) bounds the serialised tail
  long long rpb = (rows + static_cast<long long>(num_sms()) * 2 - 1) / (static_cast<long long>(num_sms()) * 2);
  if (rpb < 64) rpb = 64;
  const int rows_per_block = static_cast<int>(rpb);
  const int grid = static_cast<int>((rows + rows_per_block - 1) / rows_per_block);
  const size_t smem = static_cast<size_t>(r) * r * C * sizeof(float);
  const __nv_bfloat16* dy = static_cast<const __nv_bfloat16*>(dy_bf16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (r == 1) colsum_kernel<1><<<grid, 256, smem, st>>>(dy, out, rows, C, Wf, rows_per_block);
  else if (r == 2) colsum_kernel<2><<<grid, 256, smem, st>>>(dy, out, rows, C, Wf, rows_per_block);
  else colsum_kernel<3><<<grid, 256, smem, st>>>(dy, out, rows, C, Wf, rows_per_block);
  return launch_status();
}

extern "C" int srb200_act_bwd(const void* g_bf16, const void* y_bf16, void* out_bf16, int64_t n,
                              float slope, srb200_stream_t stream) {
  if (!g_bf16 || !y_bf16 || !out_bf16 || n <= 0 || n % 8 != 0) return SRB200_EINVAL;
  if ((reinterpret_cast<uintptr_t>(g_bf16) | reinterpret_cast<uintptr_t>(y_bf16) |
       reinterpret_cast<uintptr_t>(out_bf16)) & 15u)
    return SRB200_EINVAL;
  const size_t nvec = static_cast<size_t>(n) / 8;
  act_bwd_kernel<<<grid_for(nvec, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(g_bf16), static_cast<const uint4*>(y_bf16),
      static_cast<uint4*>(out_bf16), nvec, slope);
  return launch_status();
}

extern "C" const char* srb200_version(void) { return "srb200 0.1 (sm_100a)"; }

extern "C" const char* srb200_strerror(int code) {
  switch (code) {
    case SRB200_OK: return "ok";
    case SRB200_EINVAL: return "invalid argument (shape / alignment / flags)";
    case SRB200_EARCH: return "device is not sm_100 (B200)";
    case SRB200_EDRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    case SRB200_ELAUNCH: return "kernel launch failed";
  }
  return "unknown error";
}

extern "C" int srb200_check_device(int dev) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return SRB200_EARCH;
  return major == 10 ? SRB200_OK : SRB200_EARCH;
}
