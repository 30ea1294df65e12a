// heads_loss.cu — head activations, the Tanimoto-with-complement dual loss and the element-wise
// losses, forward and backward, as single-pass reductions over fp32 [M, C] probability maps.
//
// Reference semantics: softmax / sigmoid heads model2.py:162,171,182,186; Tanimoto_loss and
// Tanimoto_dual_loss multitasking_utils.py:38-85 (class weights 1/V^2 with V taken from the FIRST
// argument — the prediction in the first term because of the argument swap at :79 — and
// differentiated through); weighted_categorical_crossentropy utils.py:466-491; keras
// BinaryCrossentropy / MeanSquaredError train_ISPRS.py:426-428; seg metrics :446-449.
// All kernels are HBM-bound: each probability/label element is read exactly once per pass.
#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int MAXC = 32;

inline int grid1d(int64_t n, int per_sm = 8) {
  int64_t b = ceil_div64(n, NT);
  int64_t cap = (int64_t)rsa_num_sms() * per_sm;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

__global__ void __launch_bounds__(NT) softmax_fwd_kernel(const float* __restrict__ z, float* __restrict__ p,
                                                         int64_t M, int C) {
  for (int64_t m = (int64_t)blockIdx.x * NT + threadIdx.x; m < M; m += (int64_t)gridDim.x * NT) {
    const float* zp = z + m * C;
    float v[MAXC];
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) { v[c] = zp[c]; mx = fmaxf(mx, v[c]); }
    float s = 0.f;
    for (int c = 0; c < C; ++c) { v[c] = expf(v[c] - mx); s += v[c]; }
    float inv = 1.f / s;
    for (int c = 0; c < C; ++c) p[m * C + c] = v[c] * inv;
  }
}

__global__ void __launch_bounds__(NT) softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                         float* __restrict__ dz, int64_t M, int C) {
  for (int64_t m = (int64_t)blockIdx.x * NT + threadIdx.x; m < M; m += (int64_t)gridDim.x * NT) {
    float pv[MAXC], gv[MAXC];
    float dot = 0.f;
    for (int c = 0; c < C; ++c) { pv[c] = p[m * C + c]; gv[c] = dp[m * C + c]; dot += pv[c] * gv[c]; }
    for (int c = 0; c < C; ++c) dz[m * C + c] = pv[c] * (gv[c] - dot);
  }
}

__global__ void __launch_bounds__(NT) sigmoid_fwd_kernel(const float* __restrict__ z, float* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT)
    p[i] = 1.f / (1.f + expf(-z[i]));
}

__global__ void __launch_bounds__(NT) sigmoid_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                         float* __restrict__ dz, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
    float pv = p[i];
    dz[i] = dp[i] * pv * (1.f - pv);
  }
}

// ---- Tanimoto -------------------------------------------------------------------------------------
// grid: (blocks per sample, B).  Each thread walks pixels of one sample; per-thread fp32 partials over
// a bounded run, block reduction in double, one double atomic per (b, c, k) per block.
template <int C>
__global__ void __launch_bounds__(NT) tanimoto_sums_kernel(const float* __restrict__ pred,
                                                           const float* __restrict__ label, int64_t HW,
                                                           double* __restrict__ sums) {
  const int b = blockIdx.y;
  const float* pp = pred + (int64_t)b * HW * C;
  const float* lp = label + (int64_t)b * HW * C;
  float acc[C][5];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 5; ++k) acc[c][k] = 0.f;
  for (int64_t x = (int64_t)blockIdx.x * NT + threadIdx.x; x < HW; x += (int64_t)gridDim.x * NT) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float p = pp[x * C + c], l = lp[x * C + c];
      acc[c][0] += p; acc[c][1] += p * p; acc[c][2] += l; acc[c][3] += l * l; acc[c][4] += p * l;
    }
  }
  __shared__ double sh[NT / 32][C * 5];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      double v = warp_sum_d((double)acc[c][k]);
      if (lane == 0) sh[wid][c * 5 + k] = v;
    }
  __syncthreads();
  if (threadIdx.x < C * 5) {
    double v = 0;
    for (int w = 0; w < NT / 32; ++w) v += sh[w][threadIdx.x];
    atomicAdd(sums + (int64_t)b * C * 5 + threadIdx.x, v);
  }
}

// generic C (<= MAXC): one thread per (pixel, channel) element, shared-memory double atomics
__global__ void __launch_bounds__(NT) tanimoto_sums_generic_kernel(const float* __restrict__ pred,
                                                                   const float* __restrict__ label, int64_t HW,
                                                                   int C, double* __restrict__ sums) {
  __shared__ double sh[MAXC * 5];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < C * 5; i += NT) sh[i] = 0;
  __syncthreads();
  const float* pp = pred + (int64_t)b * HW * C;
  const float* lp = label + (int64_t)b * HW * C;
  const int64_t n = HW * C;
  for (int64_t e = (int64_t)blockIdx.x * NT + threadIdx.x; e < n; e += (int64_t)gridDim.x * NT) {
    int c = (int)(e % C);
    double p = pp[e], l = lp[e];
    atomicAdd(&sh[c * 5 + 0], p); atomicAdd(&sh[c * 5 + 1], p * p); atomicAdd(&sh[c * 5 + 2], l);
    atomicAdd(&sh[c * 5 + 3], l * l); atomicAdd(&sh[c * 5 + 4], p * l);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 5; i += NT) atomicAdd(sums + (int64_t)b * C * 5 + i, sh[i]);
}

// one block; dynamic smem: doubles w1[C], w2[C], V1[C], G[C], A1[B], R1[B], A2[B], R2[B], lossb[B]
__global__ void tanimoto_finalize_kernel(const double* __restrict__ sums, int B, double HW, int C, float scale,
                                         float* __restrict__ loss_b, float* __restrict__ loss_mean,
                                         float* __restrict__ coef) {
  extern __shared__ double sd[];
  double* w1 = sd;
  double* w2 = w1 + C;
  double* V1 = w2 + C;
  double* G = V1 + C;
  double* A1 = G + C;
  double* R1 = A1 + B;
  double* A2 = R1 + B;
  double* R2 = A2 + B;
  double* lb = R2 + B;
  const double smooth = 1e-5;
  const int tid = threadIdx.x, nt = blockDim.x;
#define S(b, c, k) sums[((int64_t)(b) * C + (c)) * 5 + (k)]
  for (int c = tid; c < C; c += nt) {
    double v1 = 0, v2 = 0;
    for (int b = 0; b < B; ++b) { v1 += S(b, c, 0); v2 += HW - S(b, c, 2); }
    v1 /= B; v2 /= B;
    V1[c] = v1;
    w1[c] = 1.0 / (v1 * v1);   // +inf when the class has no mass (multitasking_utils.py:47)
    w2[c] = 1.0 / (v2 * v2);
  }
  __syncthreads();
  if (tid == 0) {   // inf -> max finite weight (multitasking_utils.py:52-53)
    double m1 = 0, m2 = 0;
    for (int c = 0; c < C; ++c) {
      if (!isinf(w1[c])) m1 = fmax(m1, w1[c]);
      if (!isinf(w2[c])) m2 = fmax(m2, w2[c]);
    }
    for (int c = 0; c < C; ++c) {
      if (isinf(w1[c])) { w1[c] = m1; V1[c] = 0; }
      if (isinf(w2[c])) w2[c] = m2;
    }
  }
  __syncthreads();
  for (int b = tid; b < B; b += nt) {
    double n1 = 0, d1 = 0, n2 = 0, d2 = 0;
    for (int c = 0; c < C; ++c) {
      double sp = S(b, c, 0), sp2 = S(b, c, 1), sl = S(b, c, 2), sl2 = S(b, c, 3), spl = S(b, c, 4);
      n1 += w1[c] * spl;
      d1 += w1[c] * (sp2 + sl2 - spl);
      double prod = HW - sl - sp + spl;
      double sq = (HW - 2 * sp + sp2) + (HW - 2 * sl + sl2);
      n2 += w2[c] * prod;
      d2 += w2[c] * (sq - prod);
    }
    double t1 = (n1 + smooth) / (d1 + smooth), t2 = (n2 + smooth) / (d2 + smooth);
    A1[b] = 1.0 / (d1 + smooth); R1[b] = (n1 + smooth) / ((d1 + smooth) * (d1 + smooth));
    A2[b] = 1.0 / (d2 + smooth); R2[b] = (n2 + smooth) / ((d2 + smooth) * (d2 + smooth));
    lb[b] = 1.0 - 0.5 * (t1 + t2);
    if (loss_b) loss_b[b] = (float)lb[b];
  }
  __syncthreads();
  for (int c = tid; c < C; c += nt) {
    // gradient through the prediction-derived weights of the first term
    double acc = 0;
    for (int b = 0; b < B; ++b) {
      double d = S(b, c, 1) + S(b, c, 3) - S(b, c, 4);
      acc += A1[b] * S(b, c, 4) - R1[b] * d;
    }
    G[c] = V1[c] > 0 ? (-2.0 / ((double)B * V1[c] * V1[c] * V1[c])) * acc : 0.0;
  }
  if (tid == 0 && loss_mean) {
    double m = 0;
    for (int b = 0; b < B; ++b) m += lb[b];
    loss_mean[0] = (float)(m / B);
  }
  __syncthreads();
  if (coef) {
    const double k = -(double)scale / (2.0 * B);
    for (int i = tid; i < B * C; i += nt) {
      int b = i / C, c = i % C;
      coef[i * 3 + 0] = (float)(k * (G[c] + w2[c] * (R2[b] - A2[b])));
      coef[i * 3 + 1] = (float)(k * (-2.0 * R1[b] * w1[c] - 2.0 * R2[b] * w2[c]));
      coef[i * 3 + 2] = (float)(k * (w1[c] * (A1[b] + R1[b]) + w2[c] * (A2[b] + R2[b])));
    }
  }
#undef S
}

__global__ void __launch_bounds__(NT) tanimoto_bwd_kernel(const float* __restrict__ pred,
                                                          const float* __restrict__ label,
                                                          const float* __restrict__ coef, int64_t HWC, int C,
                                                          float* __restrict__ dpred) {
  extern __shared__ float sc[];   // [C][3] for this sample
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < C * 3; i += NT) sc[i] = coef[(int64_t)b * C * 3 + i];
  __syncthreads();
  const int64_t base = (int64_t)b * HWC;
  for (int64_t e = (int64_t)blockIdx.x * NT + threadIdx.x; e < HWC; e += (int64_t)gridDim.x * NT) {
    int c = (int)(e % C);
    dpred[base + e] = sc[c * 3] + sc[c * 3 + 1] * pred[base + e] + sc[c * 3 + 2] * label[base + e];
  }
}

// ---- element-wise losses -----------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_to_double_atomic(double v, double* out) {
  __shared__ double sh[NT / 32];
  v = warp_sum_d(v);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < NT / 32; ++w) t += sh[w];
    atomicAdd(out, t);
  }
  return 0.f;
}

constexpr float KEPS = 1e-7f;   // keras.backend.epsilon()

// loss of one pixel: kind 0 weighted categorical cross-entropy (utils.py:479-490), 1 binary cross-entropy, 2 squared error
// (keras BinaryCrossentropy / MeanSquaredError: mean over the last axis)
__device__ __forceinline__ float pixel_loss_one(int kind, const float* __restrict__ p, const float* __restrict__ y,
                                                const float* __restrict__ weights, int C) {
  float l = 0.f;
  if (kind == 0) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += p[c];
    for (int c = 0; c < C; ++c) {
      float q = fminf(fmaxf(p[c] / s, KEPS), 1.f - KEPS);
      float w = weights ? weights[c] : 1.f;
      l -= y[c] * logf(q) * w;
    }
  } else if (kind == 1) {
    for (int c = 0; c < C; ++c) {
      float q = fminf(fmaxf(p[c], KEPS), 1.f - KEPS);
      l -= y[c] * logf(q) + (1.f - y[c]) * logf(1.f - q);
    }
    l /= C;
  } else {
    for (int c = 0; c < C; ++c) { float d = p[c] - y[c]; l += d * d; }
    l /= C;
  }
  return l;
}

__global__ void __launch_bounds__(NT) pixel_loss_fwd_kernel(int kind, const float* __restrict__ pred,
                                                            const float* __restrict__ label,
                                                            const float* __restrict__ weights, int64_t M, int C,
                                                            double* __restrict__ loss_sum) {
  double acc = 0;
  for (int64_t m = (int64_t)blockIdx.x * NT + threadIdx.x; m < M; m += (int64_t)gridDim.x * NT)
    acc += (double)pixel_loss_one(kind, pred + m * C, label + m * C, weights, C);
  block_sum_to_double_atomic(acc, loss_sum);
}

// the un-reduced loss map a keras loss callable returns ([B,H,W]): standalone use of the loss functions
__global__ void __launch_bounds__(NT) pixel_loss_elem_kernel(int kind, const float* __restrict__ pred,
                                                             const float* __restrict__ label,
                                                             const float* __restrict__ weights, int64_t M, int C,
                                                             float* __restrict__ out) {
  for (int64_t m = (int64_t)blockIdx.x * NT + threadIdx.x; m < M; m += (int64_t)gridDim.x * NT)
    out[m] = pixel_loss_one(kind, pred + m * C, label + m * C, weights, C);
}

__global__ void __launch_bounds__(NT) pixel_loss_bwd_kernel(int kind, const float* __restrict__ pred,
                                                            const float* __restrict__ label,
                                                            const float* __restrict__ weights, int64_t M, int C,
                                                            float scale, float* __restrict__ dpred) {
  for (int64_t m = (int64_t)blockIdx.x * NT + threadIdx.x; m < M; m += (int64_t)gridDim.x * NT) {
    const float* p = pred + m * C;
    const float* y = label + m * C;
    float* d = dpred + m * C;
    if (kind == 0) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s += p[c];
      float g[MAXC];
      float dot = 0.f;
      for (int c = 0; c < C; ++c) {
        float q = p[c] / s;
        bool in = q > KEPS && q < 1.f - KEPS;   // clip has zero gradient outside
        float w = weights ? weights[c] : 1.f;
        g[c] = in ? -y[c] * w / q : 0.f;
        dot += g[c] * q;
      }
      for (int c = 0; c < C; ++c) d[c] = scale * (g[c] - dot) / s;
    } else if (kind == 1) {
      for (int c = 0; c < C; ++c) {
        float q = p[c];
        bool in = q > KEPS && q < 1.f - KEPS;
        d[c] = in ? scale * (-y[c] / q + (1.f - y[c]) / (1.f - q)) / C : 0.f;
      }
    } else {
      for (int c = 0; c < C; ++c) d[c] = scale * 2.f * (p[c] - y[c]) / C;
    }
  }
}

__global__ void __launch_bounds__(NT) seg_metrics_kernel(const float* __restrict__ pred,
                                                         const float* __restrict__ label, int64_t M, int C,
                                                         unsigned long long* __restrict__ out) {
  unsigned int cnt[5] = {0, 0, 0, 0, 0};
  for (int64_t m = (int64_t)blockIdx.x * NT + threadIdx.x; m < M; m += (int64_t)gridDim.x * NT) {
    const float* p = pred + m * C;
    const float* y = label + m * C;
    int ap = 0, ay = 0;
    float bp = p[0], by = y[0];
    for (int c = 0; c < C; ++c) {
      float pv = p[c], yv = y[c];
      if (pv > bp) { bp = pv; ap = c; }
      if (yv > by) { by = yv; ay = c; }
      bool t = yv > 0.5f, q = pv > 0.5f;
      cnt[1] += (t && q); cnt[2] += (!t && q); cnt[3] += (!t && !q); cnt[4] += (t && !q);
    }
    cnt[0] += (ap == ay);
  }
  __shared__ unsigned long long sh[5];
  if (threadIdx.x < 5) sh[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    unsigned int v = cnt[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sh[k], (unsigned long long)v);
  }
  __syncthreads();
  if (threadIdx.x < 5) atomicAdd(out + threadIdx.x, sh[threadIdx.x]);
}

// argmax (first maximum) + optional confusion matrix, K*K <= 1024 bins staged in shared memory
__global__ void __launch_bounds__(NT) argmax_confusion_kernel(const float* __restrict__ prob, int64_t M, int C,
                                                              int32_t* __restrict__ pred_label,
                                                              const int32_t* __restrict__ true_label, int K,
                                                              unsigned long long* __restrict__ cm) {
  __shared__ unsigned int bins[1024];
  const bool do_cm = cm != nullptr && true_label != nullptr;
  if (do_cm) {
    for (int i = threadIdx.x; i < K * K; i += NT) bins[i] = 0;
    __syncthreads();
  }
  for (int64_t m = (int64_t)blockIdx.x * NT + threadIdx.x; m < M; m += (int64_t)gridDim.x * NT) {
    const float* p = prob + m * C;
    int a = 0;
    float best = p[0];
    for (int c = 1; c < C; ++c) {
      float v = p[c];
      if (v > best) { best = v; a = c; }
    }
    if (pred_label) pred_label[m] = a;
    if (do_cm) {
      int t = true_label[m];
      if (t >= 0 && t < K && a < K) atomicAdd(&bins[t * K + a], 1u);
    }
  }
  if (do_cm) {
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += NT)
      if (bins[i]) atomicAdd(cm + i, (unsigned long long)bins[i]);
  }
}

}  // namespace

extern "C" int rsa_softmax_fwd(const float* z, float* p, int64_t M, int C, void* stream) {
  RSA_REQUIRE(z && p && M > 0 && C >= 1 && C <= MAXC, RSA_ERR_SHAPE, "softmax_fwd: C=%d out of [1,%d]", C, MAXC);
  softmax_fwd_kernel<<<grid1d(M), NT, 0, (cudaStream_t)stream>>>(z, p, M, C);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
extern "C" int rsa_softmax_bwd(const float* p, const float* dp, float* dz, int64_t M, int C, void* stream) {
  RSA_REQUIRE(p && dp && dz && M > 0 && C >= 1 && C <= MAXC, RSA_ERR_SHAPE, "softmax_bwd: bad args");
  softmax_bwd_kernel<<<grid1d(M), NT, 0, (cudaStream_t)stream>>>(p, dp, dz, M, C);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
extern "C" int rsa_sigmoid_fwd(const float* z, float* p, int64_t n, void* stream) {
  RSA_REQUIRE(z && p && n > 0, RSA_ERR_SHAPE, "sigmoid_fwd: bad args");
  sigmoid_fwd_kernel<<<grid1d(n), NT, 0, (cudaStream_t)stream>>>(z, p, n);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
extern "C" int rsa_sigmoid_bwd(const float* p, const float* dp, float* dz, int64_t n, void* stream) {
  RSA_REQUIRE(p && dp && dz && n > 0, RSA_ERR_SHAPE, "sigmoid_bwd: bad args");
  sigmoid_bwd_kernel<<<grid1d(n), NT, 0, (cudaStream_t)stream>>>(p, dp, dz, n);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_tanimoto_sums(const float* pred, const float* label, int B, int64_t HW, int C, double* sums,
                                 void* stream) {
  RSA_REQUIRE(pred && label && sums && B > 0 && HW > 0 && C >= 1 && C <= MAXC, RSA_ERR_SHAPE,
              "tanimoto_sums: bad args (C=%d)", C);
  cudaStream_t st = (cudaStream_t)stream;
  int bx = (int)ceil_div64(HW, (int64_t)NT * 8);
  int cap = (rsa_num_sms() * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, B);
  switch (C) {
    case 1: tanimoto_sums_kernel<1><<<grid, NT, 0, st>>>(pred, label, HW, sums); break;
    case 2: tanimoto_sums_kernel<2><<<grid, NT, 0, st>>>(pred, label, HW, sums); break;
    case 3: tanimoto_sums_kernel<3><<<grid, NT, 0, st>>>(pred, label, HW, sums); break;
    case 4: tanimoto_sums_kernel<4><<<grid, NT, 0, st>>>(pred, label, HW, sums); break;
    case 5: tanimoto_sums_kernel<5><<<grid, NT, 0, st>>>(pred, label, HW, sums); break;
    case 6: tanimoto_sums_kernel<6><<<grid, NT, 0, st>>>(pred, label, HW, sums); break;
    case 7: tanimoto_sums_kernel<7><<<grid, NT, 0, st>>>(pred, label, HW, sums); break;
    case 8: tanimoto_sums_kernel<8><<<grid, NT, 0, st>>>(pred, label, HW, sums); break;
    default: tanimoto_sums_generic_kernel<<<grid, NT, 0, st>>>(pred, label, HW, C, sums); break;
  }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_tanimoto_finalize(const double* sums, int B, int64_t HW, int C, float scale, float* loss_b,
                                     float* loss_mean, float* coef, void* stream) {
  RSA_REQUIRE(sums && B > 0 && HW > 0 && C >= 1 && C <= MAXC && B <= 4096, RSA_ERR_SHAPE,
              "tanimoto_finalize: bad args");
  size_t smem = (size_t)(4 * C + 5 * B) * sizeof(double);
  tanimoto_finalize_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(sums, B, (double)HW, C, scale, loss_b,
                                                                   loss_mean, coef);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_tanimoto_bwd(const float* pred, const float* label, const float* coef, int B, int64_t HW,
                                int C, float* dpred, void* stream) {
  RSA_REQUIRE(pred && label && coef && dpred && B > 0 && HW > 0 && C >= 1 && C <= MAXC, RSA_ERR_SHAPE,
              "tanimoto_bwd: bad args");
  int64_t hwc = HW * C;
  int bx = (int)ceil_div64(hwc, (int64_t)NT * 4);
  int cap = (rsa_num_sms() * 8 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  tanimoto_bwd_kernel<<<dim3(bx, B), NT, C * 3 * sizeof(float), (cudaStream_t)stream>>>(pred, label, coef, hwc, C,
                                                                                        dpred);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_pixel_loss_fwd(int kind, const float* pred, const float* label, const float* weights,
                                  int64_t M, int C, double* loss_sum, void* stream) {
  RSA_REQUIRE(pred && label && loss_sum && M > 0 && C >= 1 && C <= MAXC && kind >= 0 && kind <= 2, RSA_ERR_SHAPE,
              "pixel_loss_fwd: bad args");
  pixel_loss_fwd_kernel<<<grid1d(M, 4), NT, 0, (cudaStream_t)stream>>>(kind, pred, label, weights, M, C, loss_sum);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* out[m] = loss of pixel m (no reduction): what weighted_categorical_crossentropy(w)(y, p) (utils.py:466-491) and the
 * keras BinaryCrossentropy / MeanSquaredError functions return before keras averages them. */
extern "C" int rsa_pixel_loss_elem(int kind, const float* pred, const float* label, const float* weights,
                                   int64_t M, int C, float* out, void* stream) {
  RSA_REQUIRE(pred && label && out && M > 0 && C >= 1 && C <= MAXC && kind >= 0 && kind <= 2, RSA_ERR_SHAPE,
              "pixel_loss_elem: bad args");
  pixel_loss_elem_kernel<<<grid1d(M), NT, 0, (cudaStream_t)stream>>>(kind, pred, label, weights, M, C, out);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_pixel_loss_bwd(int kind, const float* pred, const float* label, const float* weights,
                                  int64_t M, int C, float scale, float* dpred, void* stream) {
  RSA_REQUIRE(pred && label && dpred && M > 0 && C >= 1 && C <= MAXC && kind >= 0 && kind <= 2, RSA_ERR_SHAPE,
              "pixel_loss_bwd: bad args");
  pixel_loss_bwd_kernel<<<grid1d(M), NT, 0, (cudaStream_t)stream>>>(kind, pred, label, weights, M, C, scale, dpred);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_seg_metrics(const float* pred, const float* label, int64_t M, int C, int64_t* out,
                               void* stream) {
  RSA_REQUIRE(pred && label && out && M > 0 && C >= 1, RSA_ERR_SHAPE, "seg_metrics: bad args");
  seg_metrics_kernel<<<grid1d(M, 4), NT, 0, (cudaStream_t)stream>>>(pred, label, M, C,
                                                                    reinterpret_cast<unsigned long long*>(out));
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_argmax_confusion(const float* prob, int64_t M, int C, int32_t* pred_label,
                                    const int32_t* true_label, int K, int64_t* cm, void* stream) {
  RSA_REQUIRE(prob && M > 0 && C >= 1, RSA_ERR_SHAPE, "argmax_confusion: bad args");
  RSA_REQUIRE(!cm || (K >= 1 && K <= 32 && C <= K), RSA_ERR_SHAPE, "argmax_confusion: K=%d out of [C,32]", K);
  argmax_confusion_kernel<<<grid1d(M, 4), NT, 0, (cudaStream_t)stream>>>(
      prob, M, C, pred_label, true_label, K, reinterpret_cast<unsigned long long*>(cm));
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
