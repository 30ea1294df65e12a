// common.cuh — shared device/host helpers for libresuneta (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/resuneta.h"

typedef __nv_bfloat16 bf16;

void rsa_set_error(const char* fmt, ...);

#define RSA_CHECK_LAUNCH()                                                         \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      rsa_set_error("%s:%d CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return RSA_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

#define RSA_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      rsa_set_error(__VA_ARGS__);      \
      return code;                     \
    }                                  \
  } while (0)

static inline int rsa_num_sms() { return 148; }  // B200

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// A training step is ~560 short kernels replayed from a CUDA graph.  Launched with the programmatic-stream-serialization
// attribute, a kernel's CTAs may be scheduled while the previous kernel drains: they run their private prologue (mbarrier
// init, TMEM allocation, tensor-map prefetch), then block in pdl_wait() until every prerequisite grid has completed and
// flushed, and only then touch global memory.  Kernels without pdl_wait() must NOT be launched through launch_pdl().
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifdef __CUDACC__
#include <cstdlib>
#include <utility>
static inline bool rsa_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("RSA_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = rsa_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
#endif
// streaming path of the thin 1x1 convolutions (pw_stream.cu); -100 = "not mine", otherwise the launch status
int rsa_pw_stream_dispatch(const void* x0, int C0, const void* x1, int C1, const void* wt, const float* bias, void* out,
                           const void* residual, const void* mask, double* stats, int N, int H, int W, int Cout,
                           int in_stride, int nup, const void* const* up_ptrs, const int* up_shifts, int k_base, int k_total,
                           int out_stride, int accumulate, int relu, cudaStream_t st);
int rsa_pw_wgrad_stream_dispatch(const void* x, const void* dz, float* dw, int ldw, int N, int H, int W, int Cin, int Cout,
                                 int in_stride, cudaStream_t st);
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- scalar / vector element access, always computing in fp32 ------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4 consecutive elements (16 B for fp32, 8 B for bf16); pointer must be aligned accordingly
template <typename T> __device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<bf16>(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <typename T> __device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<bf16>(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// 16-byte vectors: 4 fp32 or 8 bf16
template <typename T> struct Vec16;
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<bf16> { static constexpr int N = 8; };

template <typename T> __device__ __forceinline__ void ldv(const T* p, float* v);
template <> __device__ __forceinline__ void ldv<float>(const float* p, float* v) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ldv<bf16>(const bf16* p, float* v) {
  uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
template <typename T> __device__ __forceinline__ void stv(T* p, const float* v);
template <> __device__ __forceinline__ void stv<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void stv<bf16>(bf16* p, const float* v) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// per-channel BN coefficients from {sum, sumsq} (double) or moving statistics
__device__ __forceinline__ void bn_mean_invstd(const double* stats, double count, int C, int c, float eps,
                                               const float* mmean, const float* mvar, float& mean,
                                               float& invstd) {
  if (stats) {
    double mu = stats[c] / count;
    double var = stats[C + c] / count - mu * mu;
    if (var < 0) var = 0;
    mean = (float)mu;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
  } else {
    mean = mmean[c];
    invstd = rsqrtf(mvar[c] + eps);
  }
}
