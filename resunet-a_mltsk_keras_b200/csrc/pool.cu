// pool.cu — PSPPooling pyramid (model2.py:41-79 / model.py:35-64).
//
// The reference materialises MaxPooling2D(k) -> UpSampling2D(k) -> Conv2DN for k = 1,2,4,8 plus a
// concat, i.e. ~10 full-resolution round trips.  Here one pass over x produces all pooled levels
// (each thread owns one BSxBS window of one channel, channels contiguous across the warp so every
// access is a coalesced NHWC row segment); the 1x1 convolutions then run at pooled resolution and
// the up-sampling/concat is folded into the gather of the final 1x1 convolution (igemm segments).
// Backward routes each pooled gradient to the first maximum of its window in row-major scan order
// (the arg-max convention of the oracle) and the adjoint of nearest up-sampling is a window sum.
#include "common.cuh"
#include <cstdlib>

namespace {

constexpr int NT = 256;

template <typename T, int BS>
__global__ void __launch_bounds__(NT) maxpool_pyr_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C,
                                                             T* __restrict__ p2, T* __restrict__ p4,
                                                             T* __restrict__ p8) {
  const int HB = H / BS, WB = W / BS;
  const int64_t total = (int64_t)N * HB * WB * C;
  for (int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * NT) {
    int c = (int)(idx % C);
    int64_t t = idx / C;
    int wb = (int)(t % WB); t /= WB;
    int hb = (int)(t % HB);
    int n = (int)(t / HB);
    float v[BS][BS];
    const T* xp = x + (((int64_t)n * H + hb * BS) * W + wb * BS) * C + c;
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
      for (int j = 0; j < BS; ++j) v[i][j] = ldf<T>(xp + ((int64_t)i * W + j) * C);
    float m2[BS / 2][BS / 2];
#pragma unroll
    for (int i = 0; i < BS / 2; ++i)
#pragma unroll
      for (int j = 0; j < BS / 2; ++j)
        m2[i][j] = fmaxf(fmaxf(v[2 * i][2 * j], v[2 * i][2 * j + 1]), fmaxf(v[2 * i + 1][2 * j], v[2 * i + 1][2 * j + 1]));
    if (p2) {
      const int H2 = H / 2, W2 = W / 2;
#pragma unroll
      for (int i = 0; i < BS / 2; ++i)
#pragma unroll
        for (int j = 0; j < BS / 2; ++j)
          stf<T>(p2 + (((int64_t)n * H2 + hb * (BS / 2) + i) * W2 + wb * (BS / 2) + j) * C + c, m2[i][j]);
    }
    if constexpr (BS >= 4) {
      float m4[BS / 4][BS / 4];
#pragma unroll
      for (int i = 0; i < BS / 4; ++i)
#pragma unroll
        for (int j = 0; j < BS / 4; ++j)
          m4[i][j] = fmaxf(fmaxf(m2[2 * i][2 * j], m2[2 * i][2 * j + 1]), fmaxf(m2[2 * i + 1][2 * j], m2[2 * i + 1][2 * j + 1]));
      if (p4) {
        const int H4 = H / 4, W4 = W / 4;
#pragma unroll
        for (int i = 0; i < BS / 4; ++i)
#pragma unroll
          for (int j = 0; j < BS / 4; ++j)
            stf<T>(p4 + (((int64_t)n * H4 + hb * (BS / 4) + i) * W4 + wb * (BS / 4) + j) * C + c, m4[i][j]);
      }
      if constexpr (BS >= 8) {
        float m8 = fmaxf(fmaxf(m4[0][0], m4[0][1]), fmaxf(m4[1][0], m4[1][1]));
        if (p8) stf<T>(p8 + (((int64_t)n * (H / 8) + hb) * (W / 8) + wb) * C + c, m8);
      }
    }
  }
}

// Route d to the first maximum (row-major scan) of the KxK sub-window at (i0,j0): max value first, then
// the first position that equals it.  (An index-tracking arg-max followed by index compares was
// mis-folded by nvcc 12.9 for K=2 — the "first position equal to the max" form has no such hazard.)
template <int BS, int K>
__device__ __forceinline__ void route_first_max(const float (&v)[BS][BS], float (&g)[BS][BS], int i0, int j0, float d) {
  float m = v[i0][j0];
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) m = fmaxf(m, v[i0 + i][j0 + j]);
  bool taken = false;
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const bool hit = !taken && (v[i0 + i][j0 + j] == m);
      g[i0 + i][j0 + j] += hit ? d : 0.f;
      taken = taken || hit;
    }
}

template <typename T, int BS>
__global__ void __launch_bounds__(NT) maxpool_pyr_bwd_kernel(const T* __restrict__ x, int N, int H, int W, int C,
                                                             const T* __restrict__ dp2, const T* __restrict__ dp4,
                                                             const T* __restrict__ dp8, T* __restrict__ dx,
                                                             int accumulate) {
  const int HB = H / BS, WB = W / BS;
  const int64_t total = (int64_t)N * HB * WB * C;
  for (int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * NT) {
    int c = (int)(idx % C);
    int64_t t = idx / C;
    int wb = (int)(t % WB); t /= WB;
    int hb = (int)(t % HB);
    int n = (int)(t / HB);
    float v[BS][BS], g[BS][BS];
    const int64_t base = (((int64_t)n * H + hb * BS) * W + wb * BS) * C + c;
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
      for (int j = 0; j < BS; ++j) {
        v[i][j] = ldf<T>(x + base + ((int64_t)i * W + j) * C);
        g[i][j] = accumulate ? ldf<T>(dx + base + ((int64_t)i * W + j) * C) : 0.f;
      }
    if (dp2) {
      const int H2 = H / 2, W2 = W / 2;
#pragma unroll
      for (int i = 0; i < BS / 2; ++i)
#pragma unroll
        for (int j = 0; j < BS / 2; ++j) {
          float d = ldf<T>(dp2 + (((int64_t)n * H2 + hb * (BS / 2) + i) * W2 + wb * (BS / 2) + j) * C + c);
          route_first_max<BS, 2>(v, g, 2 * i, 2 * j, d);
        }
    }
    if constexpr (BS >= 4) {
      if (dp4) {
        const int H4 = H / 4, W4 = W / 4;
#pragma unroll
        for (int i = 0; i < BS / 4; ++i)
#pragma unroll
          for (int j = 0; j < BS / 4; ++j) {
            float d = ldf<T>(dp4 + (((int64_t)n * H4 + hb * (BS / 4) + i) * W4 + wb * (BS / 4) + j) * C + c);
            route_first_max<BS, 4>(v, g, 4 * i, 4 * j, d);
          }
      }
    }
    if constexpr (BS >= 8) {
      if (dp8) {
        float d = ldf<T>(dp8 + (((int64_t)n * (H / 8) + hb) * (W / 8) + wb) * C + c);
        route_first_max<BS, 8>(v, g, 0, 0, d);
      }
    }
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
      for (int j = 0; j < BS; ++j) stf<T>(dx + base + ((int64_t)i * W + j) * C, g[i][j]);
  }
}

template <typename T, int BS>
__global__ void __launch_bounds__(NT) sumpool_pyr_kernel(const T* __restrict__ x, int N, int H, int W, int C,
                                                         T* __restrict__ s2, T* __restrict__ s4,
                                                         T* __restrict__ s8) {
  const int HB = H / BS, WB = W / BS;
  const int64_t total = (int64_t)N * HB * WB * C;
  for (int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * NT) {
    int c = (int)(idx % C);
    int64_t t = idx / C;
    int wb = (int)(t % WB); t /= WB;
    int hb = (int)(t % HB);
    int n = (int)(t / HB);
    const T* xp = x + (((int64_t)n * H + hb * BS) * W + wb * BS) * C + c;
    float m2[BS / 2][BS / 2];
#pragma unroll
    for (int i = 0; i < BS / 2; ++i)
#pragma unroll
      for (int j = 0; j < BS / 2; ++j) {
        float a = ldf<T>(xp + ((int64_t)(2 * i) * W + 2 * j) * C);
        float b = ldf<T>(xp + ((int64_t)(2 * i) * W + 2 * j + 1) * C);
        float cc = ldf<T>(xp + ((int64_t)(2 * i + 1) * W + 2 * j) * C);
        float d = ldf<T>(xp + ((int64_t)(2 * i + 1) * W + 2 * j + 1) * C);
        m2[i][j] = (a + b) + (cc + d);
      }
    if (s2) {
      const int H2 = H / 2, W2 = W / 2;
#pragma unroll
      for (int i = 0; i < BS / 2; ++i)
#pragma unroll
        for (int j = 0; j < BS / 2; ++j)
          stf<T>(s2 + (((int64_t)n * H2 + hb * (BS / 2) + i) * W2 + wb * (BS / 2) + j) * C + c, m2[i][j]);
    }
    if constexpr (BS >= 4) {
      float m4[BS / 4][BS / 4];
#pragma unroll
      for (int i = 0; i < BS / 4; ++i)
#pragma unroll
        for (int j = 0; j < BS / 4; ++j)
          m4[i][j] = (m2[2 * i][2 * j] + m2[2 * i][2 * j + 1]) + (m2[2 * i + 1][2 * j] + m2[2 * i + 1][2 * j + 1]);
      if (s4) {
        const int H4 = H / 4, W4 = W / 4;
#pragma unroll
        for (int i = 0; i < BS / 4; ++i)
#pragma unroll
          for (int j = 0; j < BS / 4; ++j)
            stf<T>(s4 + (((int64_t)n * H4 + hb * (BS / 4) + i) * W4 + wb * (BS / 4) + j) * C + c, m4[i][j]);
      }
      if constexpr (BS >= 8) {
        float m8 = (m4[0][0] + m4[0][1]) + (m4[1][0] + m4[1][1]);
        if (s8) stf<T>(s8 + (((int64_t)n * (H / 8) + hb) * (W / 8) + wb) * C + c, m8);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 fast paths for the full 8x8 pyramid (psp_out at 256^2 is the case that matters: 67 MB tensors): one thread owns
// an 8x8 window of a channel PAIR, so every access is a 4-byte bf16x2 word and a warp touches 128 contiguous bytes
// (the scalar kernels above move 64).  Same arithmetic and the same first-maximum convention.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ld_bf2(const bf16* p) {
  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(p);
  return make_float2(__low2float(h), __high2float(h));
}
// same load, opaque to common-subexpression elimination: the second pass of the backward kernel must RE-load x (from
// L1/L2) instead of keeping all 64 values of the first pass in registers
__device__ __forceinline__ float2 ld_bf2_again(const bf16* p) {
  uint32_t w;
  asm volatile("ld.global.b32 %0, [%1];" : "=r"(w) : "l"(p));
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ void st_bf2(bf16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

__global__ void __launch_bounds__(NT) sumpool_pyr8_bf16x2_kernel(const bf16* __restrict__ x, int N, int H, int W, int C,
                                                                 bf16* __restrict__ s2, bf16* __restrict__ s4,
                                                                 bf16* __restrict__ s8) {
  const int HB = H / 8, WB = W / 8, CP = C / 2;
  const int64_t total = (int64_t)N * HB * WB * CP;
  for (int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * NT) {
    const int c = (int)(idx % CP) * 2;
    int64_t t = idx / CP;
    const int wb = (int)(t % WB); t /= WB;
    const int hb = (int)(t % HB);
    const int n = (int)(t / HB);
    const bf16* xp = x + (((int64_t)n * H + hb * 8) * W + wb * 8) * C + c;
    float2 m2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = ld_bf2(xp + ((int64_t)(2 * i) * W + 2 * j) * C), b = ld_bf2(xp + ((int64_t)(2 * i) * W + 2 * j + 1) * C);
        const float2 cc = ld_bf2(xp + ((int64_t)(2 * i + 1) * W + 2 * j) * C), d = ld_bf2(xp + ((int64_t)(2 * i + 1) * W + 2 * j + 1) * C);
        m2[i][j] = make_float2((a.x + b.x) + (cc.x + d.x), (a.y + b.y) + (cc.y + d.y));
      }
    if (s2) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_bf2(s2 + (((int64_t)n * (H / 2) + hb * 4 + i) * (W / 2) + wb * 4 + j) * C + c, m2[i][j].x, m2[i][j].y);
    }
    float2 m4[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        m4[i][j] = make_float2((m2[2 * i][2 * j].x + m2[2 * i][2 * j + 1].x) + (m2[2 * i + 1][2 * j].x + m2[2 * i + 1][2 * j + 1].x),
                               (m2[2 * i][2 * j].y + m2[2 * i][2 * j + 1].y) + (m2[2 * i + 1][2 * j].y + m2[2 * i + 1][2 * j + 1].y));
    if (s4) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
          st_bf2(s4 + (((int64_t)n * (H / 4) + hb * 2 + i) * (W / 4) + wb * 2 + j) * C + c, m4[i][j].x, m4[i][j].y);
    }
    if (s8)
      st_bf2(s8 + (((int64_t)n * (H / 8) + hb) * (W / 8) + wb) * C + c, (m4[0][0].x + m4[0][1].x) + (m4[1][0].x + m4[1][1].x),
             (m4[0][0].y + m4[0][1].y) + (m4[1][0].y + m4[1][1].y));
  }
}

// Backward of the 8-pixel pyramid with a WARP per (8 x 8 window, 32 channels), lane = channel: every pixel access of the warp
// is one contiguous 64-byte row and all 64 loads of a lane are in flight at once.  Pass 1 reduces the 64 values of a lane to
// three small words - for every 2x2 / 4x4 / 8x8 window the position of its FIRST maximum in row-major order (the first
// maximum of a 4x4 window is the smallest row-major key among the first maxima of those of its 2x2 windows that attain the
// 4x4 maximum, and likewise one level up) - after which x is no longer needed; pass 2 walks the window two rows at a time
// and routes the pooled gradients by comparing positions.  The thread-per-(window, channel pair) kernel below needs 128
// registers with spills for its two passes over x and ran at 0.23 of the HBM rate (122 us for 201 MB at 16 x 256 x 256 x 32).
// CT = compile-time channel count (column offsets become immediates: 8 row pointers address all 64 loads), 0 = runtime C.
template <int CT>
__global__ void __launch_bounds__(NT, 2) maxpool_pyr8_bwd_warp_kernel(const bf16* __restrict__ x, int N, int H, int W, int Crt,
                                                                      const bf16* __restrict__ dp2, const bf16* __restrict__ dp4,
                                                                      const bf16* __restrict__ dp8, bf16* __restrict__ dx,
                                                                      int accumulate) {
  const int C = CT ? CT : Crt;
  const int HB = H / 8, WB = W / 8, CB = C / 32;
  const int lane = threadIdx.x & 31;
  const int64_t total = (int64_t)N * HB * WB * CB;
  const int64_t wstride = (int64_t)gridDim.x * (NT / 32);
  const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
  for (int64_t wid = (int64_t)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); wid < total; wid += wstride) {
    const int c = (int)(wid % CB) * 32 + lane;
    int64_t t = wid / CB;
    const int wb = (int)(t % WB); t /= WB;
    const int hb = (int)(t % HB);
    const int n = (int)(t / HB);
    const int64_t base = (((int64_t)n * H + hb * 8) * W + wb * 8) * C + c;
    uint32_t xr[8][4];                 // two pixels per register: (i, 2k) in the low half, (i, 2k+1) in the high half
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t lo = __ldg(xs + base + ((int64_t)i * W + 2 * k) * C), hi = __ldg(xs + base + ((int64_t)i * W + 2 * k + 1) * C);
        xr[i][k] = lo | (hi << 16);
      }
    // ---- pass 1: first-maximum positions.  key of a pixel = 8 * row + column inside the 8 x 8 window; the 2 x 2 keys are kept
    // as 2-bit local positions (one word), the 4 x 4 keys as four 6-bit fields
    uint32_t f2 = 0, f4 = 0, k8 = 64;
    float m8 = 0.f;
#pragma unroll
    for (int qi = 0; qi < 2; ++qi)
#pragma unroll
      for (int qj = 0; qj < 2; ++qj) {
        float m4 = 0.f;
        uint32_t k4 = 64;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const int i = 2 * qi + u, j = 2 * qj + v;                  // 2 x 2 window (i, j): rows 2i, 2i+1, columns 2j, 2j+1
            const float a = __uint_as_float(xr[2 * i][j] << 16), b = __uint_as_float(xr[2 * i][j] & 0xffff0000u);
            const float cc = __uint_as_float(xr[2 * i + 1][j] << 16), d = __uint_as_float(xr[2 * i + 1][j] & 0xffff0000u);
            const float m = fmaxf(fmaxf(a, b), fmaxf(cc, d));
            const uint32_t loc = a == m ? 0u : (b == m ? 1u : (cc == m ? 2u : 3u));
            f2 |= loc << (2 * (4 * i + j));
            const uint32_t key = 16 * i + 2 * j + (loc & 1) + 8 * (loc >> 1);
            // row-major order of the 4 x 4 window: a later 2 x 2 window wins only with a larger value or an earlier key
            if ((u == 0 && v == 0) || m > m4 || (m == m4 && key < k4)) { m4 = m; k4 = key; }
          }
        f4 |= k4 << (6 * (2 * qi + qj));
        if ((qi == 0 && qj == 0) || m4 > m8 || (m4 == m8 && k4 < k8)) { m8 = m4; k8 = k4; }
      }
    const float g8 = dp8 ? __bfloat162float(dp8[(((int64_t)n * (H / 8) + hb) * (W / 8) + wb) * C + c]) : 0.f;
    // ---- pass 2: two image rows (four 2 x 2 windows) at a time
#pragma unroll
    for (int part = 0; part < 4; ++part) {
      float prev[2][8], g2v[4];
      if (accumulate) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) prev[i][j] = __bfloat162float(dx[base + ((int64_t)(2 * part + i) * W + j) * C]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        g2v[j] = dp2 ? __bfloat162float(dp2[(((int64_t)n * (H / 2) + hb * 4 + part) * (W / 2) + wb * 4 + j) * C + c]) : 0.f;
      float g4v[2];
#pragma unroll
      for (int j = 0; j < 2; ++j)
        g4v[j] = dp4 ? __bfloat162float(dp4[(((int64_t)n * (H / 4) + hb * 2 + (part >> 1)) * (W / 4) + wb * 2 + j) * C + c]) : 0.f;
#pragma unroll
      for (int ii = 0; ii < 2; ++ii)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = 2 * part + ii;
          const uint32_t key = 8 * i + j;
          float g = accumulate ? prev[ii][j] : 0.f;
          const uint32_t loc = (f2 >> (2 * (4 * part + (j >> 1)))) & 3;       // first maximum of this pixel's 2 x 2 window
          g += loc == (uint32_t)(2 * ii + (j & 1)) ? g2v[j >> 1] : 0.f;
          const uint32_t key4 = (f4 >> (6 * (2 * (part >> 1) + (j >> 2)))) & 63;
          g += key4 == key ? g4v[j >> 2] : 0.f;
          g += k8 == key ? g8 : 0.f;
          dx[base + ((int64_t)i * W + j) * C] = __float2bfloat16_rn(g);
        }
    }
  }
}

// 2 x 2 window sums only (the adjoint of one nearest up-sampling by two: decoder combine, model2.py:84,91), bf16: a thread
// owns a window and 8 channels - four 16-byte loads, one 16-byte store; the scalar generic kernel ran this at 0.2 of the
// HBM rate (54 us for 84 MB at 16 x 256 x 256 x 32)
__global__ void __launch_bounds__(NT) sumpool2_bf16x8_kernel(const bf16* __restrict__ x, int N, int H, int W, int C,
                                                             bf16* __restrict__ s2) {
  const int HB = H / 2, WB = W / 2, CV = C / 8;
  const int64_t total = (int64_t)N * HB * WB * CV;
  for (int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * NT) {
    const int c = (int)(idx % CV) * 8;
    int64_t t = idx / CV;
    const int wb = (int)(t % WB); t /= WB;
    const int hb = (int)(t % HB);
    const int n = (int)(t / HB);
    const bf16* xp = x + (((int64_t)n * H + 2 * hb) * W + 2 * wb) * C + c;
    float a[8], b[8], cc[8], d[8], o[8];
    ldv<bf16>(xp, a);
    ldv<bf16>(xp + C, b);
    ldv<bf16>(xp + (int64_t)W * C, cc);
    ldv<bf16>(xp + (int64_t)W * C + C, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (a[i] + b[i]) + (cc[i] + d[i]);
    stv<bf16>(s2 + (((int64_t)n * HB + hb) * WB + wb) * C + c, o);
  }
}

// backward in two passes over the window: (1) window maxima of every level, (2) each element receives the pooled gradients
// of the windows whose FIRST maximum (row-major scan) it is — one "taken" bit per (window, channel)
__global__ void __launch_bounds__(NT, 2) maxpool_pyr8_bwd_bf16x2_kernel(const bf16* __restrict__ x, int N, int H, int W, int C,
                                                                     const bf16* __restrict__ dp2, const bf16* __restrict__ dp4,
                                                                     const bf16* __restrict__ dp8, bf16* __restrict__ dx,
                                                                     int accumulate) {
  const int HB = H / 8, WB = W / 8, CP = C / 2;
  const int64_t total = (int64_t)N * HB * WB * CP;
  for (int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * NT) {
    const int c = (int)(idx % CP) * 2;
    int64_t t = idx / CP;
    const int wb = (int)(t % WB); t /= WB;
    const int hb = (int)(t % HB);
    const int n = (int)(t / HB);
    const int64_t base = (((int64_t)n * H + hb * 8) * W + wb * 8) * C + c;
    // pass 1 in the packed bf16x2 domain (max is exact): one register per loaded word
    __nv_bfloat162 p2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(x + base + ((int64_t)(2 * i) * W + 2 * j) * C);
        const __nv_bfloat162 bq = *reinterpret_cast<const __nv_bfloat162*>(x + base + ((int64_t)(2 * i) * W + 2 * j + 1) * C);
        const __nv_bfloat162 cc = *reinterpret_cast<const __nv_bfloat162*>(x + base + ((int64_t)(2 * i + 1) * W + 2 * j) * C);
        const __nv_bfloat162 d = *reinterpret_cast<const __nv_bfloat162*>(x + base + ((int64_t)(2 * i + 1) * W + 2 * j + 1) * C);
        p2[i][j] = __hmax2(__hmax2(a, bq), __hmax2(cc, d));
      }
      asm volatile("" ::: "memory");
    }
    // window maxima and pooled gradients stay PACKED (one register per channel pair) and are unpacked at use
    __nv_bfloat162 p4[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        p4[i][j] = __hmax2(__hmax2(p2[2 * i][2 * j], p2[2 * i][2 * j + 1]), __hmax2(p2[2 * i + 1][2 * j], p2[2 * i + 1][2 * j + 1]));
    const __nv_bfloat162 p8 = __hmax2(__hmax2(p4[0][0], p4[0][1]), __hmax2(p4[1][0], p4[1][1]));
    const float2 m8 = make_float2(__low2float(p8), __high2float(p8));
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
    __nv_bfloat162 q2[4][4], q4[2][2], q8 = zero2;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        q2[i][j] = dp2 ? *reinterpret_cast<const __nv_bfloat162*>(dp2 + (((int64_t)n * (H / 2) + hb * 4 + i) * (W / 2) + wb * 4 + j) * C + c) : zero2;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        q4[i][j] = dp4 ? *reinterpret_cast<const __nv_bfloat162*>(dp4 + (((int64_t)n * (H / 4) + hb * 2 + i) * (W / 4) + wb * 2 + j) * C + c) : zero2;
    if (dp8) q8 = *reinterpret_cast<const __nv_bfloat162*>(dp8 + (((int64_t)n * (H / 8) + hb) * (W / 8) + wb) * C + c);
    const float2 g8 = make_float2(__low2float(q8), __high2float(q8));
    uint32_t t2x = 0, t2y = 0, t4x = 0, t4y = 0, t8x = 0, t8y = 0;      // taken bits
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int64_t o = base + ((int64_t)i * W + j) * C;
        const float2 v = ld_bf2_again(x + o);
        float2 g = accumulate ? ld_bf2_again(dx + o) : make_float2(0.f, 0.f);
        const int w2 = (i >> 1) * 4 + (j >> 1), w4 = (i >> 2) * 2 + (j >> 2);
        // level 2
        const __nv_bfloat162 pm2 = p2[i >> 1][j >> 1], pg2 = q2[i >> 1][j >> 1], pm4 = p4[i >> 2][j >> 2], pg4 = q4[i >> 2][j >> 2];
        bool hx = !((t2x >> w2) & 1) && v.x == __low2float(pm2), hy = !((t2y >> w2) & 1) && v.y == __high2float(pm2);
        g.x += hx ? __low2float(pg2) : 0.f; g.y += hy ? __high2float(pg2) : 0.f;
        t2x |= (uint32_t)hx << w2; t2y |= (uint32_t)hy << w2;
        // level 4
        hx = !((t4x >> w4) & 1) && v.x == __low2float(pm4); hy = !((t4y >> w4) & 1) && v.y == __high2float(pm4);
        g.x += hx ? __low2float(pg4) : 0.f; g.y += hy ? __high2float(pg4) : 0.f;
        t4x |= (uint32_t)hx << w4; t4y |= (uint32_t)hy << w4;
        // level 8
        hx = !t8x && v.x == m8.x; hy = !t8y && v.y == m8.y;
        g.x += hx ? g8.x : 0.f; g.y += hy ? g8.y : 0.f;
        t8x |= (uint32_t)hx; t8y |= (uint32_t)hy;
        st_bf2(dx + o, g.x, g.y);
        if (j == 7 && (i & 1)) __threadfence_block();   // two rows of loads in flight, not all 64 (ptxas would hoist them all)
      }
  }
}

inline int pyr_bs(const void* a2, const void* a4, const void* a8) { return a8 ? 8 : (a4 ? 4 : 2); }
inline int pyr_grid(int64_t total) {
  int64_t b = ceil_div64(total, NT);
  int64_t cap = (int64_t)rsa_num_sms() * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

#define PYR_DISPATCH(KERNEL, T, ...)                                                       \
  do {                                                                                     \
    if (bs == 8) KERNEL<T, 8><<<grid, NT, 0, st>>>(__VA_ARGS__);                            \
    else if (bs == 4) KERNEL<T, 4><<<grid, NT, 0, st>>>(__VA_ARGS__);                       \
    else KERNEL<T, 2><<<grid, NT, 0, st>>>(__VA_ARGS__);                                    \
  } while (0)

extern "C" int rsa_maxpool_pyr_fwd(const void* x, int dtype, int N, int H, int W, int C, void* p2, void* p4,
                                   void* p8, void* stream) {
  int bs = pyr_bs(p2, p4, p8);
  RSA_REQUIRE(x && N > 0 && C > 0 && H % bs == 0 && W % bs == 0, RSA_ERR_SHAPE,
              "maxpool_pyr_fwd: H=%d W=%d must be multiples of %d", H, W, bs);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = pyr_grid((int64_t)N * (H / bs) * (W / bs) * C);
  if (dtype == RSA_F32) PYR_DISPATCH(maxpool_pyr_fwd_kernel, float, (const float*)x, N, H, W, C, (float*)p2, (float*)p4, (float*)p8);
  else if (dtype == RSA_BF16) PYR_DISPATCH(maxpool_pyr_fwd_kernel, bf16, (const bf16*)x, N, H, W, C, (bf16*)p2, (bf16*)p4, (bf16*)p8);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "maxpool_pyr_fwd: bad dtype");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_maxpool_pyr_bwd(const void* x, int dtype, int N, int H, int W, int C, const void* dp2,
                                   const void* dp4, const void* dp8, void* dx, int accumulate, void* stream) {
  int bs = pyr_bs(dp2, dp4, dp8);
  RSA_REQUIRE(x && dx && N > 0 && C > 0 && H % bs == 0 && W % bs == 0, RSA_ERR_SHAPE,
              "maxpool_pyr_bwd: H=%d W=%d must be multiples of %d", H, W, bs);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = pyr_grid((int64_t)N * (H / bs) * (W / bs) * C);
  static const int warp_env = getenv("RSA_MAXPOOL_WARP") ? atoi(getenv("RSA_MAXPOOL_WARP")) : 1;
  if (dtype == RSA_BF16 && bs == 8 && C == 32 && warp_env) {      // the full-resolution PSP of the last decoder level
    const int gw = pyr_grid((int64_t)N * (H / 8) * (W / 8) * 32);
    maxpool_pyr8_bwd_warp_kernel<32><<<gw, NT, 0, st>>>((const bf16*)x, N, H, W, C, (const bf16*)dp2, (const bf16*)dp4,
                                                        (const bf16*)dp8, (bf16*)dx, accumulate);
  } else if (dtype == RSA_BF16 && bs == 8 && C % 2 == 0) {
    const int g2 = pyr_grid((int64_t)N * (H / 8) * (W / 8) * (C / 2));
    maxpool_pyr8_bwd_bf16x2_kernel<<<g2, NT, 0, st>>>((const bf16*)x, N, H, W, C, (const bf16*)dp2, (const bf16*)dp4,
                                                      (const bf16*)dp8, (bf16*)dx, accumulate);
  } else if (dtype == RSA_F32) PYR_DISPATCH(maxpool_pyr_bwd_kernel, float, (const float*)x, N, H, W, C, (const float*)dp2, (const float*)dp4, (const float*)dp8, (float*)dx, accumulate);
  else if (dtype == RSA_BF16) PYR_DISPATCH(maxpool_pyr_bwd_kernel, bf16, (const bf16*)x, N, H, W, C, (const bf16*)dp2, (const bf16*)dp4, (const bf16*)dp8, (bf16*)dx, accumulate);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "maxpool_pyr_bwd: bad dtype");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_sumpool_pyr(const void* x, int dtype, int N, int H, int W, int C, void* s2, void* s4, void* s8,
                               void* stream) {
  int bs = pyr_bs(s2, s4, s8);
  RSA_REQUIRE(x && N > 0 && C > 0 && H % bs == 0 && W % bs == 0, RSA_ERR_SHAPE,
              "sumpool_pyr: H=%d W=%d must be multiples of %d", H, W, bs);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = pyr_grid((int64_t)N * (H / bs) * (W / bs) * C);
  if (dtype == RSA_BF16 && bs == 8 && C % 2 == 0) {
    const int g2 = pyr_grid((int64_t)N * (H / 8) * (W / 8) * (C / 2));
    sumpool_pyr8_bf16x2_kernel<<<g2, NT, 0, st>>>((const bf16*)x, N, H, W, C, (bf16*)s2, (bf16*)s4, (bf16*)s8);
  } else if (dtype == RSA_BF16 && bs == 2 && C % 8 == 0 && s2 && (((uintptr_t)x | (uintptr_t)s2) & 15) == 0) {
    sumpool2_bf16x8_kernel<<<pyr_grid((int64_t)N * (H / 2) * (W / 2) * (C / 8)), NT, 0, st>>>((const bf16*)x, N, H, W, C, (bf16*)s2);
  } else if (dtype == RSA_F32) PYR_DISPATCH(sumpool_pyr_kernel, float, (const float*)x, N, H, W, C, (float*)s2, (float*)s4, (float*)s8);
  else if (dtype == RSA_BF16) PYR_DISPATCH(sumpool_pyr_kernel, bf16, (const bf16*)x, N, H, W, C, (bf16*)s2, (bf16*)s4, (bf16*)s8);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "sumpool_pyr: bad dtype");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
