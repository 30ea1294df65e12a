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


// ---------------------------------------------------------------------------------------------------------------
// bf16 fast paths: one thread owns (pixel, 8-channel group) so that every access to the 32-channel tensor is one
// 16-byte vector (a warp covers 8 consecutive pixels = 512 contiguous bytes); per-channel reductions live in registers
// for the whole grid-stride walk and leave the block through one shuffle/shared-memory reduction.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack_bf8(const uint4& q, float* v) {
  v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
  v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
  v[4] = __uint_as_float(q.z << 16); v[5] = __uint_as_float(q.z & 0xffff0000u);
  v[6] = __uint_as_float(q.w << 16); v[7] = __uint_as_float(q.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack_bf8(const float* v) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return t;
}
// sum over the 8 lanes of a warp that share (lane & 3), result valid in lanes 0..3
__device__ __forceinline__ float quad_col_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

// stem, bf16: out[p][8g..8g+8) = x[p][0..NI) . w[:, 8g..] + b;  statistics of the stored values
template <int NI>
__global__ void __launch_bounds__(NT) stem_fwd_bf16_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ b, bf16* __restrict__ out, int64_t M,
                                                           double* __restrict__ stats) {
  const int g = threadIdx.x & 3;
  float wr[NI][8], br[8], s[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    br[i] = b ? b[g * 8 + i] : 0.f; s[i] = 0.f; sq[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NI; ++j) wr[j][i] = w[j * 32 + g * 8 + i];
  }
  const int64_t stride = (int64_t)gridDim.x * (NT / 4);
  for (int64_t p = (int64_t)blockIdx.x * (NT / 4) + (threadIdx.x >> 2); p < M; p += stride) {
    float xv[NI];
#pragma unroll
    for (int j = 0; j < NI; ++j) xv[j] = __bfloat162float(x[p * NI + j]);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = br[i];
#pragma unroll
      for (int j = 0; j < NI; ++j) a = fmaf(xv[j], wr[j][i], a);
      o[i] = a;
    }
    const uint4 pk = pack_bf8(o);
    *reinterpret_cast<uint4*>(out + p * 32 + g * 8) = pk;
    if (stats) {
      float r[8];
      unpack_bf8(pk, r);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += r[i]; sq[i] = fmaf(r[i], r[i], sq[i]); }
    }
  }
  if (stats) {
    __shared__ float sh[NT / 32][2][32];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = quad_col_sum(s[i]), d = quad_col_sum(sq[i]);
      if (lane < 4) { sh[wv][0][lane * 8 + i] = a; sh[wv][1][lane * 8 + i] = d; }
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int which = threadIdx.x >> 5, c = threadIdx.x & 31;
      double a = 0;
      for (int k = 0; k < NT / 32; ++k) a += sh[k][which][c];
      atomicAdd(stats + which * 32 + c, a);
    }
  }
}

// stem weight gradient, bf16: dw[j][c] += sum_p x[p][j] dy[p][c];  db[c] += sum_p dy[p][c]
template <int NI>
__global__ void __launch_bounds__(NT) stem_wgrad_bf16_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, int64_t M,
                                                             float* __restrict__ dw, float* __restrict__ db) {
  const int g = threadIdx.x & 3;
  float acc[NI + 1][8];
#pragma unroll
  for (int j = 0; j <= NI; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
  const int64_t stride = (int64_t)gridDim.x * (NT / 4);
  for (int64_t p = (int64_t)blockIdx.x * (NT / 4) + (threadIdx.x >> 2); p < M; p += 2 * stride) {
    uint4 q[2]; float xv[2][NI];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t pp = p + u * stride;
      if (pp < M) {
        q[u] = __ldg(reinterpret_cast<const uint4*>(dy + pp * 32 + g * 8));
#pragma unroll
        for (int j = 0; j < NI; ++j) xv[u][j] = __bfloat162float(x[pp * NI + j]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (p + u * stride < M) {
        float gv[8];
        unpack_bf8(q[u], gv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[NI][i] += gv[i];
#pragma unroll
          for (int j = 0; j < NI; ++j) acc[j][i] = fmaf(xv[u][j], gv[i], acc[j][i]);
        }
      }
    }
  }
  __shared__ float sh[NT / 32][NI + 1][32];
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j <= NI; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = quad_col_sum(acc[j][i]);
      if (lane < 4) sh[wv][j][lane * 8 + i] = a;
    }
  __syncthreads();
  for (int o = threadIdx.x; o < (NI + 1) * 32; o += NT) {
    const int j = o >> 5, c = o & 31;
    float a = 0.f;
    for (int k = 0; k < NT / 32; ++k) a += sh[k][j][c];
    if (j < NI) atomicAdd(dw + j * 32 + c, a);
    else if (db) atomicAdd(db + c, a);
  }
}

// head forward, bf16 features -> fp32 logits: z[p][j] = sum_c h[p][c] w[c*NO + j] + b[j]; one thread per pixel
template <int NO>
__global__ void __launch_bounds__(NT) head_fwd_bf16_kernel(const bf16* __restrict__ h, const float* __restrict__ w,
                                                           const float* __restrict__ b, float* __restrict__ z, int64_t M) {
  __shared__ float ws[32 * NO];
  for (int i = threadIdx.x; i < 32 * NO; i += NT) ws[i] = w[i];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * NT;
  for (int64_t p = (int64_t)blockIdx.x * NT + threadIdx.x; p < M; p += stride) {
    uint4 q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] = __ldg(reinterpret_cast<const uint4*>(h + p * 32) + k);
    float a[NO];
#pragma unroll
    for (int j = 0; j < NO; ++j) a[j] = b ? b[j] : 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float hv[8];
      unpack_bf8(q[k], hv);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NO; ++j) a[j] = fmaf(hv[i], ws[(k * 8 + i) * NO + j], a[j]);
    }
    float* zp = z + p * NO;
#pragma unroll
    for (int j = 0; j < NO; ++j) zp[j] = a[j];
  }
}

// head backward, bf16 features: thread = (pixel, 8-channel group g)
//   dh[p][8g+i] (=|+=) mask(h>0) * sum_j dz[p][j] w[(8g+i)*NO + j];  dw[c*NO+j] += sum_p h[p][c] dz[p][j];  db[j] += sum_p dz[p][j]
// The weights are read from shared memory (the eight lanes that share g read the same words: broadcast), so that the
// 8 x NO weight-gradient accumulators are the only per-thread state and two blocks fit an SM.  Registers have no room for a
// second pixel in flight (128 with the accumulators), and a warp that loads, waits and computes in turn ran this kernel at a
// third of the HBM rate (74 us for 226 MB); so the operands arrive through a per-warp shared-memory FIFO filled HB_DEPTH
// groups of eight pixels ahead with cp.async (pw_stream.cu's recipe: every lane copies the 16-byte pieces it consumes itself,
// only the NO logit gradients of a pixel are shared by its quad).
constexpr int HB_DEPTH = 4;                               // 40 KB of static shared memory with eight warps
constexpr int HB_STAGE = 32 * 16 + 32 * 16 + 8 * 32;      // h pieces, previous dh pieces, dz rows (<= 8 floats) of 8 pixels

__device__ __forceinline__ void hb_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void hb_cp4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

template <int NO>
__global__ void __launch_bounds__(NT, 2) head_bwd_bf16_kernel(const bf16* __restrict__ h, const float* __restrict__ dz,
                                                              const float* __restrict__ w, int64_t M, bf16* __restrict__ dh,
                                                              int accumulate, int relu_mask, float* __restrict__ dw,
                                                              float* __restrict__ db) {
  static_assert(NO <= 8, "a pixel's logit gradients fill at most one 32-byte FIFO row");
  __shared__ __align__(16) float ws[32 * NO];
  __shared__ __align__(16) uint8_t fifo_s[(NT / 32) * HB_DEPTH * HB_STAGE];
  float (*sh)[33][NO] = reinterpret_cast<float (*)[33][NO]>(fifo_s);      // block reduction, after the FIFO has drained
  static_assert(sizeof(float) * (NT / 32) * 33 * NO <= sizeof(fifo_s), "reduction buffer aliases the FIFO");
  for (int i = threadIdx.x; i < 32 * NO; i += NT) ws[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int g = lane & 3, q = lane >> 2;                   // channel group, pixel of the group
  const float* wg = ws + g * 8 * NO;
  float acc[8][NO], accb[NO];
#pragma unroll
  for (int j = 0; j < NO; ++j) {
    accb[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][j] = 0.f;
  }
  const uint32_t fifo = (uint32_t)__cvta_generic_to_shared(fifo_s) + (uint32_t)(wv * HB_DEPTH * HB_STAGE);
  const int64_t ngroups = (M + 7) >> 3, gstride = (int64_t)gridDim.x * (NT / 32);
  const bool with_prev = dh && accumulate;
  auto issue = [&](int64_t grp, int stage) {
    const int64_t p = grp * 8 + q;
    if (grp < ngroups && p < M) {
      const uint32_t base = fifo + (uint32_t)(stage * HB_STAGE);
      hb_cp16(base + lane * 16, h + p * 32 + g * 8);
      if (with_prev) hb_cp16(base + 512 + lane * 16, dh + p * 32 + g * 8);
#pragma unroll
      for (int j = 0; j < NO; j += 4)
        if (j + g < NO) hb_cp4(base + 1024 + q * 32 + (j + g) * 4, dz + p * NO + j + g);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int64_t grp0 = (int64_t)blockIdx.x * (NT / 32) + wv;
  for (int s = 0; s < HB_DEPTH - 1; ++s) issue(grp0 + s * gstride, s);
  int stage = 0;
  for (int64_t grp = grp0; grp < ngroups; grp += gstride) {
    int nstage = stage + HB_DEPTH - 1;
    if (nstage >= HB_DEPTH) nstage -= HB_DEPTH;
    __syncwarp();                      // the quad has finished reading the dz row of the stage refilled now
    issue(grp + (HB_DEPTH - 1) * gstride, nstage);
    asm volatile("cp.async.wait_group %0;" ::"n"(HB_DEPTH - 1) : "memory");
    __syncwarp();
    const uint8_t* base = fifo_s + (wv * HB_DEPTH + stage) * HB_STAGE;
    if (++stage == HB_DEPTH) stage = 0;
    const int64_t p = grp * 8 + q;
    if (p >= M) continue;
    float hv[8], zv[8];
    unpack_bf8(*reinterpret_cast<const uint4*>(base + lane * 16), hv);
    {
      const float4 z0 = *reinterpret_cast<const float4*>(base + 1024 + q * 32);
      const float4 z1 = *reinterpret_cast<const float4*>(base + 1024 + q * 32 + 16);
      zv[0] = z0.x; zv[1] = z0.y; zv[2] = z0.z; zv[3] = z0.w; zv[4] = z1.x; zv[5] = z1.y; zv[6] = z1.z; zv[7] = z1.w;
    }
    if (dh) {
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = 0.f;
#pragma unroll
        for (int j = 0; j < NO; ++j) a = fmaf(zv[j], wg[i * NO + j], a);
        d[i] = (relu_mask && !(hv[i] > 0.f)) ? 0.f : a;
      }
      if (accumulate) {
        float pv[8];
        unpack_bf8(*reinterpret_cast<const uint4*>(base + 512 + lane * 16), pv);
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] += pv[i];
      }
      *reinterpret_cast<uint4*>(dh + p * 32 + g * 8) = pack_bf8(d);
    }
#pragma unroll
    for (int j = 0; j < NO; ++j) {
      accb[j] += zv[j];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(hv[i], zv[j], acc[i][j]);
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();                     // every warp is done with its FIFO: the buffer becomes the reduction scratch
  // lanes that share a channel group differ in bits 2..4
#pragma unroll
  for (int j = 0; j < NO; ++j) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = quad_col_sum(acc[i][j]);
      if (lane < 4) sh[wv][lane * 8 + i][j] = a;
    }
    const float bsum = quad_col_sum(accb[j]);      // every quad lane saw the same dz: lane 0's copy is the pixel sum
    if (lane == 0) sh[wv][32][j] = bsum;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 33 * NO; o += NT) {
    const int c = o / NO, j = o % NO;
    float a = 0.f;
    for (int k = 0; k < NT / 32; ++k) a += sh[k][c][j];
    if (c < 32) atomicAdd(dw + c * NO + j, a);
    else if (db) atomicAdd(db + j, a);
  }
}

template <int NO>
void launch_head_bwd_bf16(const void* h, const float* dz, const float* w, int64_t M, void* dh, int accumulate, int relu_mask,
                          float* dw, float* db, cudaStream_t st) {
  head_bwd_bf16_kernel<NO><<<rsa_num_sms() * 4, NT, 0, st>>>((const bf16*)h, dz, w, M, (bf16*)dh, accumulate, relu_mask, dw, db);
}
template <int NO>
void launch_head_fwd_bf16(const void* h, const float* w, const float* b, float* z, int64_t M, cudaStream_t st) {
  head_fwd_bf16_kernel<NO><<<rsa_num_sms() * 8, NT, 0, st>>>((const bf16*)h, w, b, z, M);
}

}  // namespace

/* Stem: out[m, 0:32] = x[m, 0:n] . w[n][32] + b, optional BatchNorm statistics of the output (double[64]).
 * Replaces Conv2D(32,(1,1)) on the raw input, model2.py:101. */
extern "C" int rsa_stem_fwd(const void* x, int x_dtype, const float* w, const float* b, void* out, int out_dtype, int64_t M,
                            int n, double* stats, void* stream) {
  RSA_REQUIRE(x && w && out && M > 0 && n >= 1 && n <= MAXN, RSA_ERR_SHAPE, "stem_fwd: bad args (n=%d)", n);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = thin_grid();
  if (x_dtype == RSA_BF16 && out_dtype == RSA_BF16 && n == 3) stem_fwd_bf16_kernel<3><<<grid, NT, 0, st>>>((const bf16*)x, w, b, (bf16*)out, M, stats);
  else if (x_dtype == RSA_BF16 && out_dtype == RSA_BF16 && n == 4) stem_fwd_bf16_kernel<4><<<grid, NT, 0, st>>>((const bf16*)x, w, b, (bf16*)out, M, stats);
  else if (x_dtype == RSA_BF16 && out_dtype == RSA_BF16) stem_fwd_kernel<bf16, bf16><<<grid, NT, 0, st>>>((const bf16*)x, w, b, (bf16*)out, M, n, stats);
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
  if (dtype == RSA_BF16 && n == 3) stem_wgrad_bf16_kernel<3><<<rsa_num_sms() * 4, NT, 0, st>>>((const bf16*)x, (const bf16*)dy, M, dw, db);
  else if (dtype == RSA_BF16 && n == 4) stem_wgrad_bf16_kernel<4><<<rsa_num_sms() * 4, NT, 0, st>>>((const bf16*)x, (const bf16*)dy, M, dw, db);
  else if (dtype == RSA_BF16) stem_wgrad_kernel<bf16, bf16><<<grid, NT, 0, st>>>((const bf16*)x, (const bf16*)dy, M, n, dw, db);
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
  if (h_dtype == RSA_BF16 && n <= 8) {
    switch (n) {
#define HB_CASE(NN) case NN: launch_head_bwd_bf16<NN>(h, dz, w, M, dh, accumulate, relu_mask, dw, db, st); break;
      HB_CASE(1) HB_CASE(2) HB_CASE(3) HB_CASE(4) HB_CASE(5) HB_CASE(6) HB_CASE(7) HB_CASE(8)
#undef HB_CASE
    }
  } else if (h_dtype == RSA_BF16) head_bwd_kernel<bf16><<<grid, NT, 0, st>>>((const bf16*)h, dz, w, M, n, (bf16*)dh, accumulate, relu_mask, dw, db);
  else if (h_dtype == RSA_F32) head_bwd_kernel<float><<<grid, NT, 0, st>>>((const float*)h, dz, w, M, n, (float*)dh, accumulate, relu_mask, dw, db);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "head_bwd: bad dtype");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* Head forward: z[m, 0:n] (fp32 logits) = h[m, 0:32] (bf16) . w[32][n] + b, n <= 8.  Replaces the final Conv2D 1x1 of the
 * heads (model2.py:159,168,180,186) in bf16 mode: pure streaming, 64 B read + 4n B written per pixel. */
extern "C" int rsa_head_fwd(const void* h, const float* w, const float* b, float* z, int64_t M, int n, void* stream) {
  RSA_REQUIRE(h && w && z && M > 0 && n >= 1 && n <= 8, RSA_ERR_SHAPE, "head_fwd: bad args (n=%d)", n);
  cudaStream_t st = (cudaStream_t)stream;
  switch (n) {
#define HF_CASE(NN) case NN: launch_head_fwd_bf16<NN>(h, w, b, z, M, st); break;
    HF_CASE(1) HF_CASE(2) HF_CASE(3) HF_CASE(4) HF_CASE(5) HF_CASE(6) HF_CASE(7) HF_CASE(8)
#undef HF_CASE
  }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
