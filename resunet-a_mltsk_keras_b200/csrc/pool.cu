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
  if (dtype == RSA_F32) PYR_DISPATCH(maxpool_pyr_bwd_kernel, float, (const float*)x, N, H, W, C, (const float*)dp2, (const float*)dp4, (const float*)dp8, (float*)dx, accumulate);
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
  if (dtype == RSA_F32) PYR_DISPATCH(sumpool_pyr_kernel, float, (const float*)x, N, H, W, C, (float*)s2, (float*)s4, (float*)s8);
  else if (dtype == RSA_BF16) PYR_DISPATCH(sumpool_pyr_kernel, bf16, (const bf16*)x, N, H, W, C, (bf16*)s2, (bf16*)s4, (bf16*)s8);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "sumpool_pyr: bad dtype");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
