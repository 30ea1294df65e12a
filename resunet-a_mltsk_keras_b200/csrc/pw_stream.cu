// pw_stream.cu — streaming 1x1 convolution for the thin, high-resolution layers (K <= 64 input channels in total,
// Cout <= 64): combine / PSP / UpSampling / head 1x1 convolutions at 256x256 and 128x128 and their data gradients
// (keras Conv2D(f,(1,1)) at model2.py:37,55-68,84,92,103,159-187).
//
// Why a second kernel beside conv_tc2: these launches move 4..16 FLOP per byte, i.e. they are bound by HBM, and the
// persistent 128-pixel-tile kernel is not built for that (profiles/r2_step_launch_trace.txt: 88-136 us per launch at
// 16 x 256 x 256 x 32 where the bytes take 21-42 us).  Its epilogue - one thread per pixel, side inputs (mask, running sum,
// residual, up-sampled addends) fetched with dependent 16-byte loads after the accumulator arrives, two tiles in flight per
// CTA - runs at the latency of those loads.  Here nothing is staged and nothing waits for anything else:
//
//   * a warp owns groups of 16 consecutive pixels; every thread issues ALL loads of its group (operand rows and side
//     inputs) up front, so a resident warp keeps up to 6 KB in flight and 24 warps per SM cover the HBM latency;
//   * the channel order inside a K step and inside the N dimension of a GEMM is free, so it is chosen such that the
//     m16n8k16 fragment a thread owns is exactly a contiguous 16-byte piece of its pixel's NHWC row: the quad of a row
//     reads / writes one contiguous 64-byte row, a warp instruction 512 contiguous bytes, no shared memory, no shuffles.
//     K step s of a 32-channel block: thread t of the quad holds channels 8t+4s .. 8t+4s+3; output n-tile j: channels
//     2 NT t + 2j, 2 NT t + 2j + 1 (NT = Cout / 8), so a thread stores 2 NT contiguous channels per row;
//   * the weights (at most 64 x 32) live in registers as B fragments for the whole kernel;
//   * BatchNorm statistics of the stored values: per-thread partial sums over all its pixels, three shuffles over the rows
//     of the fragment, per-warp slots summed in a fixed order, one double atomic per channel per CTA (reproducible).
//
// The arithmetic is legacy mma.sync (HMMA) on purpose: 2.1 GFLOP per launch at K = N = 32 need ~100 TFLOP/s to hide under
// the 21 us the bytes take; tcgen05 with its TMEM round trip buys nothing on a kernel whose roofline is HBM.
// Entered through rsa_conv_tc2_fwd (same contract, same results up to summation order), never directly.
#include "common.cuh"

namespace {

constexpr int PWS_THREADS = 256;
constexpr int PWS_WARPS = PWS_THREADS / 32;
constexpr int PWS_OCC = 2;             // resident CTAs per SM the register budget is compiled for

struct PwsParams {
  const bf16* x0; const bf16* x1; const bf16* wt; const float* bias;
  bf16* out; const bf16* residual; const bf16* mask; double* stats;
  int M, lw, lh;            // output pixels, log2 W, log2 H
  int in_stride, out_stride;
  int kt, k_base;           // weight row length (bf16 elements), first column used
  int accumulate, relu, nup;
  const bf16* upq[4]; int upshift[4];
};

template <int C> struct PwsSrc { static constexpr int KS = C >= 32 ? C / 16 : (C > 0 ? 1 : 0); };

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// first channel of the low half (logical columns 2t, 2t+1) of K step s; the high half (2t+8, 2t+9) follows 2 channels on
template <int C> __device__ __forceinline__ int pws_kch(int s, int t) {
  if (C >= 32) return 32 * (s >> 1) + 8 * t + 4 * (s & 1);
  if (C == 16) return 4 * t;
  return 2 * t;
}

// A fragments of one source for rows r0 (fragment row g) and r1 (row g + 8); the pointers already include the quad offset
template <int C, int BASE, int KTOT>
__device__ __forceinline__ void pws_load_a(const bf16* r0, const bf16* r1, int t, uint32_t (&a)[KTOT][4]) {
  if constexpr (C >= 32) {
#pragma unroll
    for (int h = 0; h < C / 32; ++h) {
      const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(r0 + 32 * h + 8 * t));
      const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(r1 + 32 * h + 8 * t));
      a[BASE + 2 * h][0] = v0.x; a[BASE + 2 * h][1] = v1.x; a[BASE + 2 * h][2] = v0.y; a[BASE + 2 * h][3] = v1.y;
      a[BASE + 2 * h + 1][0] = v0.z; a[BASE + 2 * h + 1][1] = v1.z; a[BASE + 2 * h + 1][2] = v0.w; a[BASE + 2 * h + 1][3] = v1.w;
    }
  } else if constexpr (C == 16) {
    const uint2 v0 = __ldg(reinterpret_cast<const uint2*>(r0 + 4 * t));
    const uint2 v1 = __ldg(reinterpret_cast<const uint2*>(r1 + 4 * t));
    a[BASE][0] = v0.x; a[BASE][1] = v1.x; a[BASE][2] = v0.y; a[BASE][3] = v1.y;
  } else if constexpr (C == 8) {
    a[BASE][0] = __ldg(reinterpret_cast<const uint32_t*>(r0 + 2 * t));
    a[BASE][1] = __ldg(reinterpret_cast<const uint32_t*>(r1 + 2 * t));
    a[BASE][2] = 0u; a[BASE][3] = 0u;
  }
}

// 2 NT contiguous channels of one pixel row (this thread's share), bf16 <-> fp32
template <int NT, bool NC> __device__ __forceinline__ void pws_ld_raw(const bf16* p, uint32_t (&w)[NT]) {
  if constexpr (NT >= 4) {
#pragma unroll
    for (int i = 0; i < NT / 4; ++i) {
      const uint4 q = NC ? __ldg(reinterpret_cast<const uint4*>(p) + i) : *(reinterpret_cast<const uint4*>(p) + i);
      w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
    }
  } else if constexpr (NT == 2) {
    const uint2 q = NC ? __ldg(reinterpret_cast<const uint2*>(p)) : *reinterpret_cast<const uint2*>(p);
    w[0] = q.x; w[1] = q.y;
  } else {
    w[0] = NC ? __ldg(reinterpret_cast<const uint32_t*>(p)) : *reinterpret_cast<const uint32_t*>(p);
  }
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
template <int NT, bool NC> __device__ __forceinline__ void pws_ld_row(const bf16* p, float (&v)[2 * NT]) {
  uint32_t w[NT];
  pws_ld_raw<NT, NC>(p, w);
#pragma unroll
  for (int i = 0; i < NT; ++i) { v[2 * i] = bf_lo(w[i]); v[2 * i + 1] = bf_hi(w[i]); }
}
template <int NT> __device__ __forceinline__ void pws_st_row(bf16* p, const float (&v)[2 * NT], uint32_t (&w)[NT]) {
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  if constexpr (NT >= 4) {
#pragma unroll
    for (int i = 0; i < NT / 4; ++i) *(reinterpret_cast<uint4*>(p) + i) = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  } else if constexpr (NT == 2) {
    *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
  } else {
    *reinterpret_cast<uint32_t*>(p) = w[0];
  }
}

template <int C0, int C1, int COUT, bool STATS>
__global__ void __launch_bounds__(PWS_THREADS, PWS_OCC) pw_stream_kernel(const PwsParams p) {
  constexpr int KS0 = PwsSrc<C0>::KS, KS1 = PwsSrc<C1>::KS, NT = COUT / 8, NCH = 2 * NT;
  static_assert((KS0 + KS1) * NT <= 16, "B fragments must fit in registers");
  __shared__ float wsum[STATS ? PWS_WARPS : 1][COUT], wsq[STATS ? PWS_WARPS : 1][COUT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();

  // B fragments: n-tile j, fragment column g  <->  output channel 2 NT (g >> 1) + 2 j + (g & 1)
  uint32_t b0[KS0 + KS1][NT], b1[KS0 + KS1][NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int co = NCH * (g >> 1) + 2 * j + (g & 1);
    const bf16* wrow = p.wt + (size_t)co * p.kt + p.k_base;
#pragma unroll
    for (int s = 0; s < KS0; ++s) {
      const int ch = pws_kch<C0>(s, t);
      b0[s][j] = __ldg(reinterpret_cast<const uint32_t*>(wrow + ch));
      b1[s][j] = C0 == 8 ? 0u : __ldg(reinterpret_cast<const uint32_t*>(wrow + ch + 2));
    }
#pragma unroll
    for (int s = 0; s < KS1; ++s) {
      const int ch = C0 + pws_kch<C1 ? C1 : 8>(s, t);
      b0[KS0 + s][j] = __ldg(reinterpret_cast<const uint32_t*>(wrow + ch));
      b1[KS0 + s][j] = C1 == 8 ? 0u : __ldg(reinterpret_cast<const uint32_t*>(wrow + ch + 2));
    }
  }
  float bias_r[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) bias_r[i] = p.bias ? __ldg(p.bias + NCH * t + i) : 0.f;
  float ps[STATS ? NCH : 1], pq[STATS ? NCH : 1];
#pragma unroll
  for (int i = 0; i < (STATS ? NCH : 1); ++i) { ps[i] = 0.f; pq[i] = 0.f; }

  const int wmask = (1 << p.lw) - 1, hmask = (1 << p.lh) - 1;
  const int ngroups = p.M >> 4;
  for (int grp = blockIdx.x * PWS_WARPS + warp; grp < ngroups; grp += gridDim.x * PWS_WARPS) {
    // the two rows of this thread: pixels m0 + g and m0 + g + 8
    size_t src[2], dst[2];
    int pn[2], ph[2], pw[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int m = (grp << 4) + g + 8 * r;
      pw[r] = m & wmask; ph[r] = (m >> p.lw) & hmask; pn[r] = m >> (p.lw + p.lh);
      src[r] = p.in_stride == 1 ? (size_t)m
             : ((((size_t)pn[r] << (p.lh + 1)) + 2 * ph[r]) << (p.lw + 1)) + 2 * pw[r];
      dst[r] = p.out_stride == 1 ? (size_t)m
             : ((((size_t)pn[r] << (p.lh + 1)) + 2 * ph[r]) << (p.lw + 1)) + 2 * pw[r];
    }
    // ---- every load of the group is issued before anything is consumed
    uint32_t a[KS0 + KS1][4];
    pws_load_a<C0, 0>(p.x0 + src[0] * C0, p.x0 + src[1] * C0, t, a);
    if constexpr (C1 > 0) pws_load_a<C1, KS0>(p.x1 + src[0] * C1, p.x1 + src[1] * C1, t, a);
    float add[2][NCH];
    uint32_t mk[2][NT];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) add[r][i] = bias_r[i];
      float tv[NCH];
      for (int u = 0; u < p.nup; ++u) {
        const int sh = p.upshift[u];
        const size_t qi = ((((size_t)pn[r] << (p.lh - sh)) + (ph[r] >> sh)) << (p.lw - sh)) + (pw[r] >> sh);
        pws_ld_row<NT, true>(p.upq[u] + qi * COUT + NCH * t, tv);
#pragma unroll
        for (int i = 0; i < NCH; ++i) add[r][i] += tv[i];
      }
      if (p.residual) {
        pws_ld_row<NT, true>(p.residual + dst[r] * COUT + NCH * t, tv);
#pragma unroll
        for (int i = 0; i < NCH; ++i) add[r][i] += tv[i];
      }
      if (p.accumulate) {
        pws_ld_row<NT, false>(p.out + dst[r] * COUT + NCH * t, tv);
#pragma unroll
        for (int i = 0; i < NCH; ++i) add[r][i] += tv[i];
      }
      if (p.mask) pws_ld_raw<NT, true>(p.mask + dst[r] * COUT + NCH * t, mk[r]);
    }
    // ---- K steps
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
#pragma unroll
    for (int s = 0; s < KS0 + KS1; ++s)
#pragma unroll
      for (int j = 0; j < NT; ++j) mma_bf16_16816(acc[j], a[s], b0[s][j], b1[s][j]);
    // ---- epilogue: + bias + addends, ReLU, mask, store, statistics (order as conv_tc2's epilogue)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float f[NCH];
#pragma unroll
      for (int j = 0; j < NT; ++j) { f[2 * j] = acc[j][2 * r] + add[r][2 * j]; f[2 * j + 1] = acc[j][2 * r + 1] + add[r][2 * j + 1]; }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      if (p.mask) {
#pragma unroll
        for (int j = 0; j < NT; ++j) { f[2 * j] = bf_lo(mk[r][j]) > 0.f ? f[2 * j] : 0.f; f[2 * j + 1] = bf_hi(mk[r][j]) > 0.f ? f[2 * j + 1] : 0.f; }
      }
      uint32_t pk[NT];
      pws_st_row<NT>(p.out + dst[r] * COUT + NCH * t, f, pk);
      if constexpr (STATS) {        // statistics of the stored (bf16-rounded) values
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const float lo = bf_lo(pk[j]), hi = bf_hi(pk[j]);
          ps[2 * j] += lo; pq[2 * j] = fmaf(lo, lo, pq[2 * j]);
          ps[2 * j + 1] += hi; pq[2 * j + 1] = fmaf(hi, hi, pq[2 * j + 1]);
        }
      }
    }
  }
  if constexpr (STATS) {
    // rows of the fragment live in lanes that differ in bits 2..4
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
#pragma unroll
      for (int off = 4; off <= 16; off <<= 1) {
        ps[i] += __shfl_xor_sync(0xffffffffu, ps[i], off);
        pq[i] += __shfl_xor_sync(0xffffffffu, pq[i], off);
      }
    }
    if (g == 0) {
#pragma unroll
      for (int i = 0; i < NCH; ++i) { wsum[warp][NCH * t + i] = ps[i]; wsq[warp][NCH * t + i] = pq[i]; }
    }
    __syncthreads();
    if (threadIdx.x < COUT) {
      double s = 0.0, q = 0.0;
#pragma unroll
      for (int w = 0; w < PWS_WARPS; ++w) { s += (double)wsum[w][threadIdx.x]; q += (double)wsq[w][threadIdx.x]; }
      atomicAdd(p.stats + threadIdx.x, s);
      atomicAdd(p.stats + COUT + threadIdx.x, q);
    }
  }
}

template <int C0, int C1, int COUT>
int pws_launch(const PwsParams& p, cudaStream_t st) {
  const int ngroups = p.M >> 4;
  int grid = (ngroups + PWS_WARPS - 1) / PWS_WARPS;
  const int cap = PWS_OCC * rsa_num_sms();           // one resident wave, grid-stride inside
  if (grid > cap) grid = cap;
  cudaError_t le = p.stats ? launch_pdl(pw_stream_kernel<C0, C1, COUT, true>, dim3(grid), dim3(PWS_THREADS), (size_t)0, st, p)
                           : launch_pdl(pw_stream_kernel<C0, C1, COUT, false>, dim3(grid), dim3(PWS_THREADS), (size_t)0, st, p);
  if (le != cudaSuccess) { rsa_set_error("pw_stream: launch: %s", cudaGetErrorString(le)); return RSA_ERR_CUDA; }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

template <int C0, int C1>
int pws_by_cout(int Cout, const PwsParams& p, cudaStream_t st) {
  constexpr int KS = PwsSrc<C0>::KS + PwsSrc<C1>::KS;
  switch (Cout) {
    case 8: return pws_launch<C0, C1, 8>(p, st);
    case 16: return pws_launch<C0, C1, 16>(p, st);
    case 32: return pws_launch<C0, C1, 32>(p, st);
    case 64: if constexpr (KS * 8 <= 16) return pws_launch<C0, C1, 64>(p, st);
  }
  return -100;
}

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

}  // namespace

/* Streaming path of rsa_conv_tc2_fwd for taps = 1 (pw_stream.cu header).  Returns -100 when the launch is not one it takes
 * (the caller then runs the persistent tcgen05 kernel), otherwise the launch status.  RSA_PW_STREAM=0 switches it off. */
int rsa_pw_stream_dispatch(const void* x0, int C0, const void* x1, int C1, const void* wt, const float* bias, void* out,
                           const void* residual, const void* mask, double* stats, int N, int H, int W, int Cout,
                           int in_stride, int nup, const void* const* up_ptrs, const int* up_shifts, int k_base, int k_total,
                           int out_stride, int accumulate, int relu, cudaStream_t st) {
  static const int enabled = getenv("RSA_PW_STREAM") ? atoi(getenv("RSA_PW_STREAM")) : 1;
  if (!enabled) return -100;
  if (!x1) C1 = 0;
  const long long M = (long long)N * H * W;
  if (M % 16 || M > 0x7fffffffLL || (H & (H - 1)) || (W & (W - 1))) return -100;
  if (Cout != 8 && Cout != 16 && Cout != 32 && Cout != 64) return -100;
  if (!(C0 == 8 || C0 == 16 || C0 == 32 || C0 == 64) || !(C1 == 0 || (C0 == 32 && C1 == 32))) return -100;
  const int kt = k_total > 0 ? k_total : C0 + C1;
  if ((kt & 1) || (k_base & 1)) return -100;
  // the pieces a thread loads must be aligned: 16 B operand rows, 4 NT bytes of an output row
  if (((uintptr_t)x0 | (uintptr_t)x1 | (uintptr_t)out | (uintptr_t)residual | (uintptr_t)mask) & 15) return -100;
  if ((uintptr_t)wt & 3) return -100;
  PwsParams p;
  p.x0 = (const bf16*)x0; p.x1 = (const bf16*)x1; p.wt = (const bf16*)wt; p.bias = bias;
  p.out = (bf16*)out; p.residual = (const bf16*)residual; p.mask = (const bf16*)mask; p.stats = stats;
  p.M = (int)M; p.lw = ilog2(W); p.lh = ilog2(H);
  p.in_stride = in_stride; p.out_stride = out_stride; p.kt = kt; p.k_base = k_base;
  p.accumulate = accumulate; p.relu = relu; p.nup = nup;
  for (int u = 0; u < 4; ++u) {
    p.upq[u] = u < nup ? (const bf16*)up_ptrs[u] : nullptr;
    p.upshift[u] = u < nup ? up_shifts[u] : 0;
    if (u < nup && (((uintptr_t)up_ptrs[u] & 15) || up_shifts[u] > p.lw || up_shifts[u] > p.lh)) return -100;
  }
  if (C1 == 32) return pws_by_cout<32, 32>(Cout, p, st);
  switch (C0) {
    case 8: return pws_by_cout<8, 0>(Cout, p, st);
    case 16: return pws_by_cout<16, 0>(Cout, p, st);
    case 32: return pws_by_cout<32, 0>(Cout, p, st);
    case 64: return pws_by_cout<64, 0>(Cout, p, st);
  }
  return -100;
}
