// bn.cu — BatchNormalization forward / backward on NHWC tensors viewed as [M, C].
//
// Keras semantics (model2.py:17,21,38,86,93): training mode normalises with the
// batch mean and *biased* variance, eps = 1e-3; the moving averages take the Bessel-corrected
// variance with momentum 0.99.  Statistics are exchanged as double {sum, sumsq} so that the
// producers (conv epilogues) can accumulate them with atomics and fp32 validation mode keeps
// 1e-4 at a million pixels per channel.  All kernels are HBM-bound: 16-byte vector access,
// channel-contiguous, coefficient tables staged once per block in shared memory.
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAX_OUT = 4;

template <typename T>
__global__ void __launch_bounds__(NT) bn_stats_kernel(const T* __restrict__ x, int64_t M, int C,
                                                      double* __restrict__ stats, int rows_per_block) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sm[];           // [2][NT*V]
  const int tpr = C / V;                  // threads per row
  const int rpi = NT / tpr;               // rows per iteration
  const int tid = threadIdx.x;
  const int cg = tid % tpr, r0 = tid / tpr;
  float s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { s[i] = 0.f; q[i] = 0.f; }
  int64_t rbeg = (int64_t)blockIdx.x * rows_per_block;
  int64_t rend = min(M, rbeg + rows_per_block);
  if (r0 < rpi) {
    for (int64_t r = rbeg + r0; r < rend; r += rpi) {
      float v[V];
      ldv<T>(x + r * C + cg * V, v);
#pragma unroll
      for (int i = 0; i < V; ++i) { s[i] += v[i]; q[i] += v[i] * v[i]; }
    }
  }
  float* ss = sm;
  float* sq = sm + NT * V;
#pragma unroll
  for (int i = 0; i < V; ++i) { ss[tid * V + i] = s[i]; sq[tid * V + i] = q[i]; }
  __syncthreads();
  for (int c = tid; c < C; c += NT) {
    int g = c / V, i = c % V;
    double a = 0, b = 0;
    for (int r = 0; r < rpi; ++r) {
      a += ss[(r * tpr + g) * V + i];
      b += sq[(r * tpr + g) * V + i];
    }
    atomicAdd(stats + c, a);
    atomicAdd(stats + C + c, b);
  }
}

struct ApplyParams {
  void* out[MAX_OUT];
  const float* gamma[MAX_OUT];
  const float* beta[MAX_OUT];
  const float* mmean[MAX_OUT];
  const float* mvar[MAX_OUT];
  int nout;
  float* meaninv;     // optional [2][C] {mean, invstd} table written by block 0 (read by the fused BatchNorm-backward
                      // epilogue of rsa_conv_tc2_fwd)
};

template <typename T, int NOUT>
__global__ void __launch_bounds__(NT) bn_apply_kernel(const T* __restrict__ x, int64_t M, int C, const ApplyParams ap,
                                                      const double* __restrict__ stats, double count, float eps,
                                                      int relu, int rows_per_block) {
  // each thread owns one fixed group of V channels: scale/shift are computed once per block (fp64 -> smem),
  // copied to registers, and the rows are walked with 16-byte loads/stores, two rows in flight per thread
  // (one read of x feeds all NOUT branch outputs)
  constexpr int V = Vec16<T>::N;
  extern __shared__ float tab[];          // [NOUT][2][C]
  for (int i = threadIdx.x; i < NOUT * C; i += NT) {
    const int k = i / C, c = i % C;
    float mean, invstd;
    bn_mean_invstd(stats, count, C, c, eps, ap.mmean[k], ap.mvar[k], mean, invstd);
    const float scv = ap.gamma[k][c] * invstd;
    tab[(2 * k) * C + c] = scv;
    tab[(2 * k + 1) * C + c] = ap.beta[k][c] - mean * scv;
    if (ap.meaninv && blockIdx.x == 0 && k == 0) { ap.meaninv[c] = mean; ap.meaninv[C + c] = invstd; }
  }
  __syncthreads();
  const int tpr = C / V, rpi = NT / tpr, tid = threadIdx.x;
  const int cg = tid % tpr, r0 = tid / tpr;
  float sc[NOUT][V], sh[NOUT][V];
#pragma unroll
  for (int k = 0; k < NOUT; ++k)
#pragma unroll
    for (int j = 0; j < V; ++j) { sc[k][j] = tab[(2 * k) * C + cg * V + j]; sh[k][j] = tab[(2 * k + 1) * C + cg * V + j]; }
  const int64_t rbeg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t rend = min(M, rbeg + rows_per_block);
  for (int64_t r = rbeg + r0; r < rend; r += 2 * rpi) {
    const int64_t o0 = r * C + cg * V, o1 = (r + rpi) * C + cg * V;
    const bool two = r + rpi < rend;
    float v0[V], v1[V];
    ldv<T>(x + o0, v0);
    if (two) ldv<T>(x + o1, v1);
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      float out[V];
#pragma unroll
      for (int j = 0; j < V; ++j) { const float t = fmaf(v0[j], sc[k][j], sh[k][j]); out[j] = relu ? fmaxf(t, 0.f) : t; }
      stv<T>(reinterpret_cast<T*>(ap.out[k]) + o0, out);
      if (two) {
#pragma unroll
        for (int j = 0; j < V; ++j) { const float t = fmaf(v1[j], sc[k][j], sh[k][j]); out[j] = relu ? fmaxf(t, 0.f) : t; }
        stv<T>(reinterpret_cast<T*>(ap.out[k]) + o1, out);
      }
    }
  }
}

// red[c] += sum g ; red[C+c] += sum g*xhat
template <typename T>
__global__ void __launch_bounds__(NT) bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                           const T* __restrict__ act, int64_t M, int C,
                                                           const double* __restrict__ stats, double count,
                                                           float eps, double* __restrict__ red,
                                                           int rows_per_block) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sm[];   // [2][NT*V]
  const int tpr = C / V, rpi = NT / tpr, tid = threadIdx.x;
  const int cg = tid % tpr, r0 = tid / tpr;
  float mean[V], inv[V], s[V], q[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    bn_mean_invstd(stats, count, C, cg * V + i, eps, nullptr, nullptr, mean[i], inv[i]);
    s[i] = 0.f; q[i] = 0.f;
  }
  int64_t rbeg = (int64_t)blockIdx.x * rows_per_block;
  int64_t rend = min(M, rbeg + rows_per_block);
  if (r0 < rpi) {
    for (int64_t r = rbeg + r0; r < rend; r += rpi) {
      float g[V], xv[V], a[V];
      int64_t o = r * C + cg * V;
      ldv<T>(dy + o, g);
      ldv<T>(x + o, xv);
      if (act) {
        ldv<T>(act + o, a);
#pragma unroll
        for (int i = 0; i < V; ++i) g[i] = a[i] > 0.f ? g[i] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < V; ++i) { s[i] += g[i]; q[i] += g[i] * (xv[i] - mean[i]) * inv[i]; }
    }
  }
  float* ss = sm;
  float* sq = sm + NT * V;
#pragma unroll
  for (int i = 0; i < V; ++i) { ss[tid * V + i] = s[i]; sq[tid * V + i] = q[i]; }
  __syncthreads();
  for (int c = tid; c < C; c += NT) {
    int g = c / V, i = c % V;
    double a = 0, b = 0;
    for (int r = 0; r < rpi; ++r) {
      a += ss[(r * tpr + g) * V + i];
      b += sq[(r * tpr + g) * V + i];
    }
    atomicAdd(red + c, a);
    atomicAdd(red + C + c, b);
  }
}

template <typename T>
__global__ void __launch_bounds__(NT) bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                          const T* __restrict__ act, int64_t nvec, int C,
                                                          const double* __restrict__ stats, double count,
                                                          float eps, const float* __restrict__ gamma,
                                                          const double* __restrict__ red, T* __restrict__ dx,
                                                          int accumulate, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sm[];   // [4][C]: mean, invstd, k0 = gamma*invstd, (c1 = sum g / n, c2 = sum g xhat / n) packed
  float* s_mean = sm;
  float* s_inv = sm + C;
  float* s_c1 = sm + 2 * C;
  float* s_c2 = sm + 3 * C;
  float* s_g = sm + 4 * C;
  for (int c = threadIdx.x; c < C; c += NT) {
    float mean, inv;
    bn_mean_invstd(stats, count, C, c, eps, nullptr, nullptr, mean, inv);
    s_mean[c] = mean;
    s_inv[c] = inv;
    s_c1[c] = (float)(red[c] / count);
    s_c2[c] = (float)(red[C + c] / count);
    s_g[c] = gamma[c] * inv;
    if (blockIdx.x == 0) {
      if (dgamma) dgamma[c] = (float)red[C + c];
      if (dbeta) dbeta[c] = (float)red[c];
    }
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * NT) {
    float g[V], xv[V], o[V];
    ldv<T>(dy + i * V, g);
    ldv<T>(x + i * V, xv);
    if (act) {
      float a[V];
      ldv<T>(act + i * V, a);
#pragma unroll
      for (int j = 0; j < V; ++j) g[j] = a[j] > 0.f ? g[j] : 0.f;
    }
    if (accumulate) ldv<T>(dx + i * V, o);
    int c0 = (int)((i * V) % C);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      int c = c0 + j;
      float xh = (xv[j] - s_mean[c]) * s_inv[c];
      float d = s_g[c] * (g[j] - s_c1[c] - xh * s_c2[c]);
      o[j] = accumulate ? o[j] + d : d;
    }
    stv<T>(dx + i * V, o);
  }
}

__global__ void bn_meaninv_kernel(const double* __restrict__ stats, double count, float eps, float* __restrict__ out, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, inv;
  bn_mean_invstd(stats, count, C, c, eps, nullptr, nullptr, mean, inv);
  out[c] = mean;
  out[C + c] = inv;
}

__global__ void bn_derive_stats_kernel(const double* __restrict__ src, double count, const float* gamma,
                                       const float* beta, float eps, double* __restrict__ dst, double dcount,
                                       int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mu = src[c] / count;
  double var = src[C + c] / count - mu * mu;
  if (var < 0) var = 0;
  double g = gamma[c], b = beta[c];
  double vy = g * g * var / (var + (double)eps);   // biased variance of gamma*xhat+beta
  dst[c] = dcount * b;
  dst[C + c] = dcount * (vy + b * b);
}

__global__ void bn_update_moving_kernel(const double* __restrict__ stats_base, float* __restrict__ param_base,
                                        const int64_t* __restrict__ table, const double* __restrict__ counts,
                                        int nlayers, float momentum) {
  int layer = blockIdx.x;
  if (layer >= nlayers) return;
  const int64_t soff = table[layer * 4 + 0];
  const int C = (int)table[layer * 4 + 1];
  float* mm = param_base + table[layer * 4 + 2];
  float* mv = param_base + table[layer * 4 + 3];
  const double n = counts[layer * 2 + 0], nfull = counts[layer * 2 + 1];
  const double* st = stats_base + soff;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double mu = st[c] / n;
    double var = st[C + c] / n - mu * mu;
    if (var < 0) var = 0;
    double var_u = nfull > 1.0 ? var * (nfull / (nfull - 1.0)) : var;
    mm[c] = (float)((double)mm[c] * momentum + mu * (1.0 - (double)momentum));
    mv[c] = (float)((double)mv[c] * momentum + var_u * (1.0 - (double)momentum));
  }
}

// ---- multi-branch backward with recomputed ReLU masks ---------------------------------------------------
// The k branches of a ResBlock-a normalise the SAME input x (model2.py:17), so their BatchNorm backward
// shares every read of x and produces ONE dx:  dx (=|+=) sum_k gamma_k*inv*(g_k - mean(g_k) - xhat*mean(g_k xhat)).
// The ReLU mask is recomputed from x with the forward's own scale/shift (fmaf(x, sc, sh) > 0) instead of
// re-reading the activated tensor.  Each thread owns 4 fixed channels (coefficients live in registers) and
// walks rows; HBM traffic per element: reduce (1+k) reads, apply (1+k) reads + 1 write (+1 read to accumulate).
constexpr int MAXK = 4;

struct BwdMultiParams {
  const void* dy[MAXK];
  const float* gamma[MAXK];
  const float* beta[MAXK];
  double* red[MAXK];
  float* dgamma[MAXK];
  float* dbeta[MAXK];
  int k;
};

// VM channels per thread: 8 (one 16-byte bf16 vector) for a single branch, 4 for several (register budget)
template <typename T, int K> struct MultiV { static constexpr int V = (K == 1 && sizeof(T) == 2) ? 8 : 4; };

template <typename T, int V> __device__ __forceinline__ void ldm(const T* p, float* v) {
  if constexpr (V == 8) ldv<T>(p, v);
  else { float t[4]; ld4<T>(p, t); v[0] = t[0]; v[1] = t[1]; v[2] = t[2]; v[3] = t[3]; }
}
template <typename T, int V> __device__ __forceinline__ void stm(T* p, const float* v) {
  if constexpr (V == 8) stv<T>(p, v);
  else { float t[4] = {v[0], v[1], v[2], v[3]}; st4<T>(p, t); }
}

// raw (still packed) vectors: loads stay in flight as 8/16-byte registers and are unpacked at use, so that several
// rows per thread can be outstanding without the fp32 copies eating the register file
template <typename T, int V> struct RawV;
template <> struct RawV<bf16, 8> { typedef uint4 type; };
template <> struct RawV<bf16, 4> { typedef uint2 type; };
template <> struct RawV<float, 4> { typedef float4 type; };
template <typename T, int V> __device__ __forceinline__ typename RawV<T, V>::type ld_raw(const T* p) {
  return __ldg(reinterpret_cast<const typename RawV<T, V>::type*>(p));
}
__device__ __forceinline__ void unpack_raw(const uint4& q, float* v) {
  v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
  v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
  v[4] = __uint_as_float(q.z << 16); v[5] = __uint_as_float(q.z & 0xffff0000u);
  v[6] = __uint_as_float(q.w << 16); v[7] = __uint_as_float(q.w & 0xffff0000u);
}
__device__ __forceinline__ void unpack_raw(const uint2& q, float* v) {
  v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
  v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
}
__device__ __forceinline__ void unpack_raw(const float4& q, float* v) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }

// smem table [2 + 2K][C]: mean, invstd, then per branch scale = gamma*invstd and shift = beta - mean*scale
template <int K>
__device__ __forceinline__ void stage_coeffs(float* tab, int C, const BwdMultiParams& bp, const double* stats, double count,
                                             float eps) {
  for (int c = threadIdx.x; c < C; c += NT) {
    float mean, inv;
    bn_mean_invstd(stats, count, C, c, eps, nullptr, nullptr, mean, inv);
    tab[c] = mean;
    tab[C + c] = inv;
#pragma unroll
    for (int b = 0; b < K; ++b) {
      const float scv = bp.gamma[b][c] * inv;
      tab[(2 + 2 * b) * C + c] = scv;
      tab[(3 + 2 * b) * C + c] = bp.beta[b][c] - mean * scv;
    }
  }
  __syncthreads();
}

template <typename T, int K>
__global__ void __launch_bounds__(NT) bn_bwd_reduce_multi_kernel(const T* __restrict__ x, int64_t M, int C,
                                                                 const BwdMultiParams bp, const double* __restrict__ stats,
                                                                 double count, float eps, int relu, int rows_per_block) {
  constexpr int V = MultiV<T, K>::V;
  extern __shared__ float tab[];    // coefficient table, reused for the block reduction
  stage_coeffs<K>(tab, C, bp, stats, count, eps);
  const int tpr = C / V, rpi = NT / tpr, tid = threadIdx.x;
  const int cg = tid % tpr, r0 = tid / tpr;
  float mean[V], inv[V], sc[K][V], sh[K][V], s[K][V], q[K][V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    mean[i] = tab[cg * V + i];
    inv[i] = tab[C + cg * V + i];
#pragma unroll
    for (int b = 0; b < K; ++b) {
      sc[b][i] = tab[(2 + 2 * b) * C + cg * V + i];
      sh[b][i] = tab[(3 + 2 * b) * C + cg * V + i];
      s[b][i] = 0.f; q[b][i] = 0.f;
    }
  }
  const int64_t rbeg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t rend = min(M, rbeg + rows_per_block);
  constexpr int U = K == 1 ? 4 : 2;            // rows in flight per thread (loads stay packed)
  typedef typename RawV<T, V>::type Raw;
  for (int64_t r = rbeg + r0; r < rend; r += U * rpi) {
    Raw xr[U], gr[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * rpi;
      if (rr < rend) {
        const int64_t o = rr * C + cg * V;
        xr[u] = ld_raw<T, V>(x + o);
#pragma unroll
        for (int b = 0; b < K; ++b) gr[u][b] = ld_raw<T, V>(reinterpret_cast<const T*>(bp.dy[b]) + o);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (r + u * rpi < rend) {
        float xv[V], xh[V];
        unpack_raw(xr[u], xv);
#pragma unroll
        for (int i = 0; i < V; ++i) xh[i] = (xv[i] - mean[i]) * inv[i];
#pragma unroll
        for (int b = 0; b < K; ++b) {
          float g[V];
          unpack_raw(gr[u][b], g);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float gg = (!relu || fmaf(xv[i], sc[b][i], sh[b][i]) > 0.f) ? g[i] : 0.f;
            s[b][i] += gg;
            q[b][i] = fmaf(gg, xh[i], q[b][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < K; ++b) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < V; ++i) { tab[tid * 2 * V + i] = s[b][i]; tab[tid * 2 * V + V + i] = q[b][i]; }
    __syncthreads();
    for (int c = tid; c < C; c += NT) {
      const int g = c / V, i = c % V;
      double a = 0, d = 0;
      for (int r = 0; r < rpi; ++r) {
        a += tab[(r * tpr + g) * 2 * V + i];
        d += tab[(r * tpr + g) * 2 * V + V + i];
      }
      atomicAdd(bp.red[b] + c, a);
      atomicAdd(bp.red[b] + C + c, d);
    }
  }
}

template <typename T, int K>
__global__ void __launch_bounds__(NT) bn_bwd_apply_multi_kernel(const T* __restrict__ x, int64_t M, int C,
                                                                const BwdMultiParams bp, const double* __restrict__ stats,
                                                                double count, float eps, int relu, T* __restrict__ dx,
                                                                int accumulate, int rows_per_block) {
  constexpr int V = MultiV<T, K>::V;
  extern __shared__ float tab[];
  stage_coeffs<K>(tab, C, bp, stats, count, eps);
  const int tpr = C / V, rpi = NT / tpr, tid = threadIdx.x;
  const int cg = tid % tpr, r0 = tid / tpr;
  float mean[V], inv[V], Bsum[V], Dsum[V], sc[K][V], sh[K][V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = cg * V + i;
    mean[i] = tab[c];
    inv[i] = tab[C + c];
    Bsum[i] = 0.f; Dsum[i] = 0.f;
#pragma unroll
    for (int b = 0; b < K; ++b) {
      sc[b][i] = tab[(2 + 2 * b) * C + c];
      sh[b][i] = tab[(3 + 2 * b) * C + c];
      Bsum[i] += sc[b][i] * (float)(bp.red[b][c] / count);
      Dsum[i] += sc[b][i] * (float)(bp.red[b][C + c] / count);
      if (blockIdx.x == 0 && r0 == 0) {
        if (bp.dgamma[b]) bp.dgamma[b][c] = (float)bp.red[b][C + c];
        if (bp.dbeta[b]) bp.dbeta[b][c] = (float)bp.red[b][c];
      }
    }
  }
  const int64_t rbeg = (int64_t)blockIdx.x * rows_per_block;
  const int64_t rend = min(M, rbeg + rows_per_block);
  constexpr int U = K == 1 ? 4 : 2;
  typedef typename RawV<T, V>::type Raw;
  for (int64_t r = rbeg + r0; r < rend; r += U * rpi) {
    Raw xr[U], ar[U], gr[U][K];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * rpi;
      if (rr < rend) {
        const int64_t o = rr * C + cg * V;
        xr[u] = ld_raw<T, V>(x + o);
#pragma unroll
        for (int b = 0; b < K; ++b) gr[u][b] = ld_raw<T, V>(reinterpret_cast<const T*>(bp.dy[b]) + o);
        if (accumulate) ar[u] = *reinterpret_cast<const Raw*>(dx + o);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + u * rpi;
      if (rr < rend) {
        float xv[V], acc[V];
        unpack_raw(xr[u], xv);
        if (accumulate) unpack_raw(ar[u], acc);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float base = -Bsum[i] - (xv[i] - mean[i]) * inv[i] * Dsum[i];
          acc[i] = accumulate ? acc[i] + base : base;
        }
#pragma unroll
        for (int b = 0; b < K; ++b) {
          float g[V];
          unpack_raw(gr[u][b], g);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float gg = (!relu || fmaf(xv[i], sc[b][i], sh[b][i]) > 0.f) ? g[i] : 0.f;
            acc[i] = fmaf(sc[b][i], gg, acc[i]);
          }
        }
        stm<T, V>(dx + rr * C + cg * V, acc);
      }
    }
  }
}

inline bool multi_shape_ok(int C) { return C >= 8 && (C & (C - 1)) == 0 && C <= 1024; }

template <typename T> bool bn_shape_ok(int C) {
  constexpr int V = Vec16<T>::N;
  if (C < V || C % V) return false;
  int tpr = C / V;
  return tpr <= NT && (NT % tpr) == 0;
}

inline int grid_for(int64_t nvec) {
  int64_t b = ceil_div64(nvec, NT);
  int64_t cap = (int64_t)rsa_num_sms() * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" int rsa_bn_stats(const void* x, int dtype, int64_t M, int C, double* stats, void* stream) {
  RSA_REQUIRE(x && stats && M > 0, RSA_ERR_SHAPE, "bn_stats: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int rows = (int)ceil_div64(M, (int64_t)rsa_num_sms() * 4);
  if (rows < 64) rows = 64;
  int grid = (int)ceil_div64(M, rows);
  if (dtype == RSA_F32) {
    RSA_REQUIRE(bn_shape_ok<float>(C), RSA_ERR_SHAPE, "bn_stats: C=%d unsupported", C);
    bn_stats_kernel<float><<<grid, NT, 2 * NT * 4 * sizeof(float), st>>>((const float*)x, M, C, stats, rows);
  } else if (dtype == RSA_BF16) {
    RSA_REQUIRE(bn_shape_ok<bf16>(C), RSA_ERR_SHAPE, "bn_stats: C=%d unsupported", C);
    bn_stats_kernel<bf16><<<grid, NT, 2 * NT * 8 * sizeof(float), st>>>((const bf16*)x, M, C, stats, rows);
  } else {
    RSA_REQUIRE(false, RSA_ERR_DTYPE, "bn_stats: bad dtype");
  }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_bn_apply(const void* x, int dtype, int64_t M, int C, int nout, void* const* outs,
                            const float* const* gammas, const float* const* betas, const double* stats,
                            double count, const float* const* moving_means, const float* const* moving_vars,
                            float eps, int relu, float* meaninv_out, void* stream) {
  RSA_REQUIRE(x && outs && gammas && betas && M > 0 && nout >= 1 && nout <= MAX_OUT, RSA_ERR_SHAPE,
              "bn_apply: bad args");
  RSA_REQUIRE(stats || (moving_means && moving_vars), RSA_ERR_SHAPE, "bn_apply: no statistics given");
  ApplyParams ap;
  ap.nout = nout;
  ap.meaninv = meaninv_out;
  for (int k = 0; k < MAX_OUT; ++k) {
    int kk = k < nout ? k : 0;
    ap.out[k] = outs[kk]; ap.gamma[k] = gammas[kk]; ap.beta[k] = betas[kk];
    ap.mmean[k] = moving_means ? moving_means[kk] : nullptr;
    ap.mvar[k] = moving_vars ? moving_vars[kk] : nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSA_F32) {
    RSA_REQUIRE(bn_shape_ok<float>(C), RSA_ERR_SHAPE, "bn_apply: C=%d unsupported", C);
    const int rpi = NT / (C / 4);
    int rows = (int)ceil_div64(M, (int64_t)rsa_num_sms() * 8);
    rows = (rows + 2 * rpi - 1) / (2 * rpi) * (2 * rpi);
    if (rows < 4 * rpi) rows = 4 * rpi;
    const size_t sm = (size_t)nout * 2 * C * sizeof(float);
    const int gr = (int)ceil_div64(M, rows);
    switch (nout) {
      case 1: bn_apply_kernel<float, 1><<<gr, NT, sm, st>>>((const float*)x, M, C, ap, stats, count, eps, relu, rows); break;
      case 2: bn_apply_kernel<float, 2><<<gr, NT, sm, st>>>((const float*)x, M, C, ap, stats, count, eps, relu, rows); break;
      case 3: bn_apply_kernel<float, 3><<<gr, NT, sm, st>>>((const float*)x, M, C, ap, stats, count, eps, relu, rows); break;
      default: bn_apply_kernel<float, 4><<<gr, NT, sm, st>>>((const float*)x, M, C, ap, stats, count, eps, relu, rows); break;
    }
  } else if (dtype == RSA_BF16) {
    RSA_REQUIRE(bn_shape_ok<bf16>(C), RSA_ERR_SHAPE, "bn_apply: C=%d unsupported", C);
    const int rpi = NT / (C / 8);
    int rows = (int)ceil_div64(M, (int64_t)rsa_num_sms() * 8);
    rows = (rows + 2 * rpi - 1) / (2 * rpi) * (2 * rpi);
    if (rows < 4 * rpi) rows = 4 * rpi;
    const size_t sm = (size_t)nout * 2 * C * sizeof(float);
    const int gr = (int)ceil_div64(M, rows);
    switch (nout) {
      case 1: bn_apply_kernel<bf16, 1><<<gr, NT, sm, st>>>((const bf16*)x, M, C, ap, stats, count, eps, relu, rows); break;
      case 2: bn_apply_kernel<bf16, 2><<<gr, NT, sm, st>>>((const bf16*)x, M, C, ap, stats, count, eps, relu, rows); break;
      case 3: bn_apply_kernel<bf16, 3><<<gr, NT, sm, st>>>((const bf16*)x, M, C, ap, stats, count, eps, relu, rows); break;
      default: bn_apply_kernel<bf16, 4><<<gr, NT, sm, st>>>((const bf16*)x, M, C, ap, stats, count, eps, relu, rows); break;
    }
  } else {
    RSA_REQUIRE(false, RSA_ERR_DTYPE, "bn_apply: bad dtype");
  }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_bn_bwd_reduce(const void* dy, const void* x, const void* act, int dtype, int64_t M, int C,
                                 const double* stats, double count, float eps, double* red, void* stream) {
  RSA_REQUIRE(dy && x && stats && red && M > 0, RSA_ERR_SHAPE, "bn_bwd_reduce: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int rows = (int)ceil_div64(M, (int64_t)rsa_num_sms() * 4);
  if (rows < 64) rows = 64;
  int grid = (int)ceil_div64(M, rows);
  if (dtype == RSA_F32) {
    RSA_REQUIRE(bn_shape_ok<float>(C), RSA_ERR_SHAPE, "bn_bwd_reduce: C=%d unsupported", C);
    bn_bwd_reduce_kernel<float><<<grid, NT, 2 * NT * 4 * sizeof(float), st>>>(
        (const float*)dy, (const float*)x, (const float*)act, M, C, stats, count, eps, red, rows);
  } else if (dtype == RSA_BF16) {
    RSA_REQUIRE(bn_shape_ok<bf16>(C), RSA_ERR_SHAPE, "bn_bwd_reduce: C=%d unsupported", C);
    bn_bwd_reduce_kernel<bf16><<<grid, NT, 2 * NT * 8 * sizeof(float), st>>>(
        (const bf16*)dy, (const bf16*)x, (const bf16*)act, M, C, stats, count, eps, red, rows);
  } else {
    RSA_REQUIRE(false, RSA_ERR_DTYPE, "bn_bwd_reduce: bad dtype");
  }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_bn_bwd_apply(const void* dy, const void* x, const void* act, int dtype, int64_t M, int C,
                                const double* stats, double count, float eps, const float* gamma,
                                const double* red, void* dx, int accumulate, float* dgamma, float* dbeta,
                                void* stream) {
  RSA_REQUIRE(dy && x && stats && red && gamma && dx && M > 0, RSA_ERR_SHAPE, "bn_bwd_apply: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  size_t smem = (size_t)5 * C * sizeof(float);
  if (dtype == RSA_F32) {
    RSA_REQUIRE(C % 4 == 0, RSA_ERR_SHAPE, "bn_bwd_apply: C=%d must be a multiple of 4", C);
    int64_t nvec = M * C / 4;
    bn_bwd_apply_kernel<float><<<grid_for(nvec), NT, smem, st>>>((const float*)dy, (const float*)x,
        (const float*)act, nvec, C, stats, count, eps, gamma, red, (float*)dx, accumulate, dgamma, dbeta);
  } else if (dtype == RSA_BF16) {
    RSA_REQUIRE(C % 8 == 0, RSA_ERR_SHAPE, "bn_bwd_apply: C=%d must be a multiple of 8", C);
    int64_t nvec = M * C / 8;
    bn_bwd_apply_kernel<bf16><<<grid_for(nvec), NT, smem, st>>>((const bf16*)dy, (const bf16*)x,
        (const bf16*)act, nvec, C, stats, count, eps, gamma, red, (bf16*)dx, accumulate, dgamma, dbeta);
  } else {
    RSA_REQUIRE(false, RSA_ERR_DTYPE, "bn_bwd_apply: bad dtype");
  }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_bn_derive_stats(const double* src_stats, double count, const float* gamma, const float* beta,
                                   float eps, double* dst_stats, double dst_count, int C, void* stream) {
  RSA_REQUIRE(src_stats && gamma && beta && dst_stats && C > 0, RSA_ERR_SHAPE, "bn_derive_stats: bad args");
  bn_derive_stats_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(src_stats, count, gamma, beta, eps,
                                                                            dst_stats, dst_count, C);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_bn_update_moving(const double* stats_base, float* param_base, const int64_t* table,
                                    const double* counts, int nlayers, float momentum, void* stream) {
  RSA_REQUIRE(stats_base && param_base && table && counts && nlayers > 0, RSA_ERR_SHAPE,
              "bn_update_moving: bad args");
  bn_update_moving_kernel<<<nlayers, 128, 0, (cudaStream_t)stream>>>(stats_base, param_base, table, counts,
                                                                     nlayers, momentum);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}


namespace {
int fill_multi(BwdMultiParams& bp, int k, const void* const* dys, const float* const* gammas, const float* const* betas,
               double* const* reds, float* const* dgammas, float* const* dbetas) {
  RSA_REQUIRE(k >= 1 && k <= MAXK && dys && gammas && betas && reds, RSA_ERR_SHAPE, "bn_bwd_multi: bad branch arguments");
  bp.k = k;
  for (int b = 0; b < MAXK; ++b) {
    const int bb = b < k ? b : 0;
    bp.dy[b] = dys[bb]; bp.gamma[b] = gammas[bb]; bp.beta[b] = betas[bb]; bp.red[b] = reds[bb];
    bp.dgamma[b] = dgammas ? dgammas[bb] : nullptr;
    bp.dbeta[b] = dbetas ? dbetas[bb] : nullptr;
    RSA_REQUIRE(bp.dy[b] && bp.gamma[b] && bp.beta[b] && bp.red[b], RSA_ERR_SHAPE, "bn_bwd_multi: null branch pointer");
  }
  return RSA_OK;
}
inline void multi_grid(int64_t M, int C, int V, int& rows, int& grid) {
  // one resident wave (2 blocks of 256 threads per SM at ~120 registers): the per-block prologue (coefficient table)
  // and epilogue (block reduction + atomics) are paid once per SM slot instead of once per ~900 rows
  const int rpi = NT / (C / V);
  rows = (int)ceil_div64(M, (int64_t)rsa_num_sms() * 2);
  rows = (rows + 4 * rpi - 1) / (4 * rpi) * (4 * rpi);
  grid = (int)ceil_div64(M, rows);
}
inline size_t multi_smem(int C, int k, int V) {
  size_t a = (size_t)(2 + 2 * k) * C * sizeof(float), b = (size_t)NT * 2 * V * sizeof(float);
  return a > b ? a : b;
}

template <typename T, int K>
void launch_reduce(const void* x, int64_t M, int C, const BwdMultiParams& bp, const double* stats, double count, float eps,
                   int relu, cudaStream_t st) {
  constexpr int V = MultiV<T, K>::V;
  int rows, grid;
  multi_grid(M, C, V, rows, grid);
  bn_bwd_reduce_multi_kernel<T, K><<<grid, NT, multi_smem(C, K, V), st>>>((const T*)x, M, C, bp, stats, count, eps, relu, rows);
}
template <typename T, int K>
void launch_apply(const void* x, int64_t M, int C, const BwdMultiParams& bp, const double* stats, double count, float eps,
                  int relu, void* dx, int accumulate, cudaStream_t st) {
  constexpr int V = MultiV<T, K>::V;
  int rows, grid;
  multi_grid(M, C, V, rows, grid);
  bn_bwd_apply_multi_kernel<T, K><<<grid, NT, multi_smem(C, K, V), st>>>((const T*)x, M, C, bp, stats, count, eps, relu,
                                                                       (T*)dx, accumulate, rows);
}
}  // namespace

/* Backward of k <= 4 BatchNormalization(+ReLU) layers that share their input x (the ResBlock-a branches,
 * model2.py:17-18): reds[b][2C] (double, zeroed) += { sum g_b, sum g_b*xhat }, g_b = dy_b * (relu ? bn_b(x) > 0 : 1). */
extern "C" int rsa_bn_bwd_reduce_multi(const void* const* dys, const void* x, int dtype, int64_t M, int C, int k,
                                       const double* stats, double count, float eps, const float* const* gammas,
                                       const float* const* betas, int relu, double* const* reds, void* stream) {
  RSA_REQUIRE(x && stats && M > 0 && multi_shape_ok(C), RSA_ERR_SHAPE, "bn_bwd_reduce_multi: bad args (C=%d)", C);
  BwdMultiParams bp;
  int rc = fill_multi(bp, k, dys, gammas, betas, reds, nullptr, nullptr);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  RSA_REQUIRE(dtype == RSA_F32 || dtype == RSA_BF16, RSA_ERR_DTYPE, "bn_bwd_reduce_multi: bad dtype");
#define RED_CASE(KK) case KK: if (dtype == RSA_F32) launch_reduce<float, KK>(x, M, C, bp, stats, count, eps, relu, st); \
                              else launch_reduce<bf16, KK>(x, M, C, bp, stats, count, eps, relu, st); break;
  switch (k) { RED_CASE(1) RED_CASE(2) RED_CASE(3) RED_CASE(4) }
#undef RED_CASE
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* dx (=|+=) sum_b gamma_b*invstd*(g_b - red_b[0]/count - xhat*red_b[1]/count); dgamma_b = red_b[1], dbeta_b = red_b[0]. */
extern "C" int rsa_bn_bwd_apply_multi(const void* const* dys, const void* x, int dtype, int64_t M, int C, int k,
                                      const double* stats, double count, float eps, const float* const* gammas,
                                      const float* const* betas, int relu, double* const* reds, void* dx, int accumulate,
                                      float* const* dgammas, float* const* dbetas, void* stream) {
  RSA_REQUIRE(x && dx && stats && M > 0 && multi_shape_ok(C), RSA_ERR_SHAPE, "bn_bwd_apply_multi: bad args (C=%d)", C);
  BwdMultiParams bp;
  int rc = fill_multi(bp, k, dys, gammas, betas, reds, dgammas, dbetas);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  RSA_REQUIRE(dtype == RSA_F32 || dtype == RSA_BF16, RSA_ERR_DTYPE, "bn_bwd_apply_multi: bad dtype");
#define APP_CASE(KK) case KK: if (dtype == RSA_F32) launch_apply<float, KK>(x, M, C, bp, stats, count, eps, relu, dx, accumulate, st); \
                              else launch_apply<bf16, KK>(x, M, C, bp, stats, count, eps, relu, dx, accumulate, st); break;
  switch (k) { APP_CASE(1) APP_CASE(2) APP_CASE(3) APP_CASE(4) }
#undef APP_CASE
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* out[2C] = {mean, invstd} (fp32) of a BatchNormalization from its {sum, sumsq} statistics — the coefficient table
 * the fused BatchNorm-backward epilogue of rsa_conv_tc2_fwd reads. */
extern "C" int rsa_bn_meaninv(const double* stats, double count, float eps, float* out, int C, void* stream) {
  RSA_REQUIRE(stats && out && C > 0, RSA_ERR_SHAPE, "bn_meaninv: bad args");
  bn_meaninv_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, count, eps, out, C);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
