// thin.cu — 1x1 convolutions with one thin side (<= 8 channels) and one 32-channel side.
//
// Two places of the graph have that shape and are pure HBM streaming: the stem Conv2D(32,(1,1)) on the 3- or
// 14-band input (model2.py:101) and the final 1x1 convolutions of the heads (32 -> num_classes / 3,
// model2.py:159,168,180,186) whose gradients arrive as fp32 d(logits).  Tensor cores cannot help (K or N below
// the UMMA minimum) and the generic implicit GEMM wastes a 64x64 tile on them.  Here one lane owns one of the
// 32 wide channels: every pixel row is a single coalesced 64-byte (bf16) access per warp, the thin side is read
// as warp-uniform broadcasts, per-channel reductions (BatchNorm statistics, weight / bias gradients) accumulate
// in registers across the pixels a warp walks and leave the SM as one atomic per lane.
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAXN = 16;

inline int thin_grid() { return rsa_num_sms() * 8; }

// out[p][lane] = sum_j x[p][j] * w[j*32 + lane] + b[lane];  stats += {sum, sumsq} of the stored values
template <typename TX, typename TO>
__global__ void __launch_bounds__(NT) stem_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ b, TO* __restrict__ out, int64_t M, int n,
                                                      double* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * NT) >> 5;
  float wr[MAXN];
#pragma unroll
  for (int j = 0; j < MAXN; ++j) wr[j] = j < n ? w[j * 32 + lane] : 0.f;
  const float bias = b ? b[lane] : 0.f;
  float s = 0.f, sq = 0.f;
  for (int64_t p0 = warp; p0 < M; p0 += 4 * nwarps) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t p = p0 + u * nwarps;
      if (p < M) {
        const TX* xp = x + p * n;
        float acc = bias;
#pragma unroll
        for (int j = 0; j < MAXN; ++j)
          if (j < n) acc = fmaf(ldf<TX>(xp + j), wr[j], acc);
        stf<TO>(out + p * 32 + lane, acc);
        s += acc; sq += acc * acc;
      }
    }
  }
  if (stats) {
    __shared__ float sh[2][NT];
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = sq;
    __syncthreads();
    if (threadIdx.x < 32) {
      double a = 0, d = 0;
      for (int wv = 0; wv < NT / 32; ++wv) { a += sh[0][wv * 32 + lane]; d += sh[1][wv * 32 + lane]; }
      atomicAdd(stats + lane, a);
      atomicAdd(stats + 32 + lane, d);
    }
  }
}

// dw[j*32 + lane] += sum_p x[p][j] * dy[p][lane];  db[lane] += sum_p dy[p][lane]
template <typename TX, typename TG>
__global__ void __launch_bounds__(NT) stem_wgrad_kernel(const TX* __restrict__ x, const TG* __restrict__ dy, int64_t M, int n,
                                                        float* __restrict__ dw, float* __restrict__ db) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * NT) >> 5;
  float acc[MAXN], bs = 0.f;
#pragma unroll
  for (int j = 0; j < MAXN; ++j) acc[j] = 0.f;
  constexpr int U = 4;                       // pixels in flight per warp
  for (int64_t p0 = warp; p0 < M; p0 += U * nwarps) {
    float g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + u * nwarps;
      g[u] = p < M ? ldf<TG>(dy + p * 32 + lane) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + u * nwarps;
      if (p < M) {
        const TX* xp = x + p * n;
        bs += g[u];
#pragma unroll
        for (int j = 0; j < MAXN; ++j)
          if (j < n) acc[j] = fmaf(ldf<TX>(xp + j), g[u], acc[j]);
      }
    }
  }
  __shared__ float sh[NT / 32][32];
  for (int j = 0; j <= n; ++j) {           // j == n: bias gradient
    float v = bs;
#pragma unroll
    for (int t = 0; t < MAXN; ++t) if (t == j && j < n) v = acc[t];
    __syncthreads();
    sh[threadIdx.x >> 5][lane] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      float a = 0.f;
      for (int wv = 0; wv < NT / 32; ++wv) a += sh[wv][lane];
      if (j < n) atomicAdd(dw + j * 32 + lane, a);
      else if (db) atomicAdd(db + lane, a);
    }
  }
}

// heads backward: h [M,32] (bf16/f32), dz [M,n] fp32, w [32][n] fp32
//   dh[p][c] (=|+=) (mask: h>0) * sum_j dz[p][j] * w[c*n + j]
//   dw[c*n + j] += sum_p h[p][c] * dz[p][j];   db[j] += sum_p dz[p][j]
// Tiles of 192 pixels are staged in shared memory with coalesced 16-byte loads; phase A gives every thread one
// (pixel, 8-channel group) of dh (16-byte store), phase B gives every thread one entry of dw/db whose partial sum
// lives in a register across all tiles of the block.
template <typename TH>
__global__ void __launch_bounds__(NT) head_bwd_kernel(const TH* __restrict__ h, const float* __restrict__ dz,
                                                      const float* __restrict__ w, int64_t M, int n, TH* __restrict__ dh,
                                                      int accumulate, int relu_mask, float* __restrict__ dw,
                                                      float* __restrict__ db) {
  constexpr int TP = 192;                 // 192 x (33 + 17) floats + weights = 40.5 KB of static smem
  __shared__ float hs[TP][33];
  __shared__ float zs[TP][MAXN + 1];
  __shared__ float wt[MAXN][32];          // wt[j][c] = w[c*n + j]
  const int tid = threadIdx.x;
  for (int i = tid; i < 32 * n; i += NT) wt[i % n][i / n] = w[i];
  // phase-B ownership: outputs o = j*32 + c for j < n (dw) and o = 32n + j (db)
  const int nout = 32 * n + n;
  float accB[3] = {0.f, 0.f, 0.f};        // up to 3 outputs per thread (nout <= 528)
  const int64_t ntile = (M + TP - 1) / TP;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t p0 = tile * TP;
    const int np = (int)(M - p0 < TP ? M - p0 : TP);
    __syncthreads();
    for (int e = tid; e < TP * 32; e += NT) {
      const int pp = e >> 5, c = e & 31;
      hs[pp][c] = pp < np ? ldf<TH>(h + (p0 + pp) * 32 + c) : 0.f;
    }
    for (int e = tid; e < TP * n; e += NT) {
      const int pp = e / n, j = e % n;
      zs[pp][j] = pp < np ? dz[(p0 + pp) * n + j] : 0.f;
    }
    __syncthreads();
    if (dh) {
      // phase A: 4 threads per pixel, 8 channels each
      for (int it = 0; it < TP / 64; ++it) {
        const int pp = it * 64 + (tid >> 2), g = (tid & 3) * 8;
        if (pp < np) {
          float d[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) d[i] = 0.f;
          for (int j = 0; j < n; ++j) {
            const float z = zs[pp][j];
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] = fmaf(z, wt[j][g + i], d[i]);
          }
          TH* dst = dh + (p0 + pp) * 32 + g;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (relu_mask && !(hs[pp][g + i] > 0.f)) d[i] = 0.f;
            if (accumulate) d[i] += ldf<TH>(dst + i);
            stf<TH>(dst + i, d[i]);
          }
        }
      }
    }
    // phase B
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int o = tid + k * NT;
      if (o < nout) {
        float a = 0.f;
        if (o < 32 * n) {
          const int j = o >> 5, c = o & 31;
          for (int pp = 0; pp < TP; ++pp) a = fmaf(hs[pp][c], zs[pp][j], a);
        } else {
          const int j = o - 32 * n;
          for (int pp = 0; pp < TP; ++pp) a += zs[pp][j];
        }
        accB[k] += a;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int o = tid + k * NT;
    if (o < 32 * n) atomicAdd(dw + (o & 31) * n + (o >> 5), accB[k]);
    else if (o < nout && db) atomicAdd(db + (o - 32 * n), accB[k]);
  }
}

}  // namespace

/* Stem: out[m, 0:32] = x[m, 0:n] . w[n][32] + b, optional BatchNorm statistics of the output (double[64]).
 * Replaces Conv2D(32,(1,1)) on the raw input, model2.py:101. */
extern "C" int rsa_stem_fwd(const void* x, int x_dtype, const float* w, const float* b, void* out, int out_dtype, int64_t M,
                            int n, double* stats, void* stream) {
  RSA_REQUIRE(x && w && out && M > 0 && n >= 1 && n <= MAXN, RSA_ERR_SHAPE, "stem_fwd: bad args (n=%d)", n);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = thin_grid();
  if (x_dtype == RSA_BF16 && out_dtype == RSA_BF16) stem_fwd_kernel<bf16, bf16><<<grid, NT, 0, st>>>((const bf16*)x, w, b, (bf16*)out, M, n, stats);
  else if (x_dtype == RSA_F32 && out_dtype == RSA_F32) stem_fwd_kernel<float, float><<<grid, NT, 0, st>>>((const float*)x, w, b, (float*)out, M, n, stats);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "stem_fwd: dtype combination");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* dw[n][32] += x^T dy, db[32] += column sums of dy (fp32, zeroed by the caller). */
extern "C" int rsa_stem_wgrad(const void* x, const void* dy, int dtype, int64_t M, int n, float* dw, float* db, void* stream) {
  RSA_REQUIRE(x && dy && dw && M > 0 && n >= 1 && n <= MAXN, RSA_ERR_SHAPE, "stem_wgrad: bad args (n=%d)", n);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = thin_grid();
  if (dtype == RSA_BF16) stem_wgrad_kernel<bf16, bf16><<<grid, NT, 0, st>>>((const bf16*)x, (const bf16*)dy, M, n, dw, db);
  else if (dtype == RSA_F32) stem_wgrad_kernel<float, float><<<grid, NT, 0, st>>>((const float*)x, (const float*)dy, M, n, dw, db);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "stem_wgrad: bad dtype");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* Backward of a head's final 1x1 convolution (32 -> n <= 16, fp32 logits gradient dz):
 * dh (=|+=) mask(h>0) * dz . w^T ; dw[32][n] += h^T dz ; db[n] += column sums of dz.  dh may be NULL.
 * Replaces the Conv2D 1x1 backward kernels behind model2.py:159,168,180,186. */
extern "C" int rsa_head_bwd(const void* h, int h_dtype, const float* dz, const float* w, int64_t M, int n, void* dh,
                            int accumulate, int relu_mask, float* dw, float* db, void* stream) {
  RSA_REQUIRE(h && dz && w && dw && M > 0 && n >= 1 && n <= MAXN, RSA_ERR_SHAPE, "head_bwd: bad args (n=%d)", n);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = thin_grid();
  if (h_dtype == RSA_BF16) head_bwd_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)h, dz, w, M, n, (bf16*)dh, accumulate, relu_mask, dw, db);
  else if (h_dtype == RSA_F32) head_bwd_kernel<float><<<grid, NT, 0, st>>>((const float*)h, dz, w, M, n, (float*)dh, accumulate, relu_mask, dw, db);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "head_bwd: bad dtype");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
