// optim.cu — fused optimizer updates over the flat parameter buffer and dtype casts.
// Keras forms (train_ISPRS.py:404-407):
//   Adam: m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr_t * m / (sqrt(v) + eps),
//         lr_t = lr * sqrt(1-b2^t) / (1-b1^t) computed on the host;
//   SGD : vel = momentum*vel - lr*g ; p += vel.
// One launch covers all 42.7 M trainable parameters (the reference issues one resource-apply kernel
// per variable, ~350 launches).  HBM-bound: 16 B/param read + 12 B/param written for Adam.
#include "common.cuh"

namespace {
constexpr int NT = 256;

__global__ void __launch_bounds__(NT) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                  float lr_t, const float* __restrict__ lr_dev, float b1,
                                                  float b2, float eps, float gs) {
  if (lr_dev) lr_t = lr_dev[0];
  const int64_t nv = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < nv; i += (int64_t)gridDim.x * NT) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ga[4] = {gv.x, gv.y, gv.z, gv.w};
    float ma[4] = {mv.x, mv.y, mv.z, mv.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = ga[k] * gs;
      ma[k] = b1 * ma[k] + (1.f - b1) * gg;
      va[k] = b2 * va[k] + (1.f - b2) * gg * gg;
      pa[k] -= lr_t * ma[k] / (sqrtf(va[k]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
  }
  for (int64_t i = (nv << 2) + (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
    float gg = g[i] * gs;
    float mm = b1 * m[i] + (1.f - b1) * gg;
    float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm; v[i] = vv;
    p[i] -= lr_t * mm / (sqrtf(vv) + eps);
  }
}

__global__ void __launch_bounds__(NT) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                 float* __restrict__ vel, int64_t n, float lr,
                                                 const float* __restrict__ lr_dev, float mom, float gs) {
  if (lr_dev) lr = lr_dev[0];
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
    float v = mom * vel[i] - lr * g[i] * gs;
    vel[i] = v;
    p[i] += v;
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(NT) cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT)
    stf<TD>(d + i, ldf<TS>(s + i));
}

inline int grid1d(int64_t n) {
  int64_t b = ceil_div64(n, NT);
  int64_t cap = (int64_t)rsa_num_sms() * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}
}  // namespace

extern "C" int rsa_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr_t,
                             const float* lr_dev, float beta1, float beta2, float eps, float grad_scale,
                             void* stream) {
  RSA_REQUIRE(param && grad && m && v && n > 0, RSA_ERR_SHAPE, "adam_step: bad args");
  RSA_REQUIRE(((uintptr_t)param % 16 == 0) && ((uintptr_t)grad % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                  ((uintptr_t)v % 16 == 0), RSA_ERR_ALIGN, "adam_step: buffers must be 16-byte aligned");
  adam_kernel<<<grid1d(n / 4 + 1), NT, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, lr_t, lr_dev, beta1, beta2,
                                                                  eps, grad_scale);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_sgd_step(float* param, const float* grad, float* vel, int64_t n, float lr, const float* lr_dev,
                            float momentum, float grad_scale, void* stream) {
  RSA_REQUIRE(param && grad && vel && n > 0, RSA_ERR_SHAPE, "sgd_step: bad args");
  sgd_kernel<<<grid1d(n), NT, 0, (cudaStream_t)stream>>>(param, grad, vel, n, lr, lr_dev, momentum, grad_scale);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
  RSA_REQUIRE(src && dst && n > 0, RSA_ERR_SHAPE, "cast: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid1d(n);
  if (src_dtype == RSA_F32 && dst_dtype == RSA_BF16) cast_kernel<float, bf16><<<grid, NT, 0, st>>>((const float*)src, (bf16*)dst, n);
  else if (src_dtype == RSA_BF16 && dst_dtype == RSA_F32) cast_kernel<bf16, float><<<grid, NT, 0, st>>>((const bf16*)src, (float*)dst, n);
  else if (src_dtype == RSA_F32 && dst_dtype == RSA_F32) cast_kernel<float, float><<<grid, NT, 0, st>>>((const float*)src, (float*)dst, n);
  else if (src_dtype == RSA_BF16 && dst_dtype == RSA_BF16) cast_kernel<bf16, bf16><<<grid, NT, 0, st>>>((const bf16*)src, (bf16*)dst, n);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "cast: bad dtype");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

namespace {
template <typename T>
__global__ void __launch_bounds__(NT) axpy_kernel(T* __restrict__ dst, const T* __restrict__ src, int64_t n,
                                                  int accumulate) {
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * NT) {
    float v = ldf<T>(src + i);
    if (accumulate) v += ldf<T>(dst + i);
    stf<T>(dst + i, v);
  }
}
}  // namespace

// dst (=|+=) src — identity branch of the ResBlock-a backward (Add, model2.py:31)
extern "C" int rsa_axpy(void* dst, const void* src, int dtype, int64_t n, int accumulate, void* stream) {
  RSA_REQUIRE(dst && src && n > 0, RSA_ERR_SHAPE, "axpy: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSA_F32) axpy_kernel<float><<<grid1d(n), NT, 0, st>>>((float*)dst, (const float*)src, n, accumulate);
  else if (dtype == RSA_BF16) axpy_kernel<bf16><<<grid1d(n), NT, 0, st>>>((bf16*)dst, (const bf16*)src, n, accumulate);
  else RSA_REQUIRE(false, RSA_ERR_DTYPE, "axpy: bad dtype");
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
