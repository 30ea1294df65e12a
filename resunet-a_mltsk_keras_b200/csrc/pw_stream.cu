// pw_stream.cu — streaming 1x1 convolution for the thin, high-resolution layers (K <= 64 input channels in total,
// Cout <= 64): combine / PSP / UpSampling / head 1x1 convolutions at 256x256 and 128x128 and their data gradients
// (keras Conv2D(f,(1,1)) at model2.py:37,55-68,84,92,103,159-187).
//
// Why a second kernel beside conv_tc2: these launches move 4..16 FLOP per byte, i.e. they are bound by HBM, and the
// persistent 128-pixel-tile kernel is not built for that (profiles/r2_step_launch_trace.txt: 88-136 us per launch at
// 16 x 256 x 256 x 32 where the bytes take 21-42 us).  Its epilogue - one thread per pixel, side inputs (mask, running sum,
// residual, up-sampled addends) fetched with dependent 16-byte loads after the accumulator arrives, two tiles in flight per
// CTA - runs at the latency of those loads.  Here:
//
//   * a warp owns groups of 16 consecutive pixels.  The channel order inside a K step and inside the N dimension of a GEMM
//     is free, so it is chosen such that the m16n8k16 fragment a thread owns is exactly a contiguous 16-byte piece of its
//     pixel's NHWC row: the quad of a row reads / writes one contiguous 64-byte row, a warp instruction 512 contiguous
//     bytes, no shuffles.  K step s of a 32-channel block: thread t of the quad holds channels 8t+4s .. 8t+4s+3; output
//     n-tile j: channels 2 NT t + 2j, 2 NT t + 2j + 1 (NT = Cout / 8), so a thread stores 2 NT contiguous channels per row;
//   * HBM latency x bandwidth wants ~45 KB in flight per SM, more than registers can hold beside the fragments (the first
//     version, every load of a group issued up front into registers, reached 0.3-0.65 of the copy peak,
//     gpurun_out/r2s_bench_pw.txt).  So every lane copies ITS pieces of the next groups - operand rows, mask, running sum,
//     residual, up-sampled addends - with cp.async into a private shared-memory FIFO (slot = lane, conflict-free, no
//     barrier: a lane only ever reads what it copied itself), DEPTH groups ahead, and consumes the oldest group with
//     ld.shared; the FIFO depth is sized per launch from the bytes a group moves;
//   * the weights live in registers as B fragments for the whole kernel (64 x 64: in shared memory, one word per lane);
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
// resident CTAs per SM the register budget is compiled for: 3 (80 registers), 2 where the fragments alone need more
// (64 output channels, or 64 input channels into 32: B fragments 32 registers + accumulators 16-32)
template <int C0, int C1, int COUT> struct PwsOcc {
  static constexpr int V = (COUT == 64 || (C0 + C1 == 64 && COUT == 32)) ? 2 : 3;
};
constexpr int PWS_MAXSIDE = 7;         // up-sampled addends (4), residual, running sum, mask

struct PwsParams {
  const bf16* x0; const bf16* x1; const bf16* wt; const float* bias;
  bf16* out; const bf16* mask; double* stats;
  int M, lw, lh;            // output pixels, log2 W, log2 H
  int in_stride, out_stride;
  int kt, k_base;           // weight row length (bf16 elements), first column used
  int relu;
  int nadd;                 // addends: side[0 .. nadd), shift > 0 = nearest up-sampling by 2^shift; the mask is side[nadd]
  const bf16* side[PWS_MAXSIDE]; int shift[PWS_MAXSIDE];
  int depth, stage_bytes;   // FIFO: groups in flight per warp, bytes of one group's pieces (all 32 lanes)
};

template <int C> struct PwsSrc {
  static constexpr int KS = C >= 32 ? C / 16 : (C > 0 ? 1 : 0);      // K steps
  static constexpr int PC = C >= 32 ? C / 32 : (C > 0 ? 1 : 0);      // pieces per pixel row and thread
  static constexpr int PB = C >= 32 ? 16 : C / 2;                    // bytes of a piece
};

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// global -> shared copy of one piece (4 / 8 / 16 bytes); 16-byte pieces of the streamed tensors bypass L1
template <int BYTES, bool STREAM> __device__ __forceinline__ void cp_piece(uint32_t dst, const void* src) {
  if constexpr (BYTES == 16 && STREAM) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait(int pending) {       // all but the newest `pending` groups have landed
  switch (pending) {
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    case 7: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
  }
}
__device__ __forceinline__ uint4 lds16(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}

// first channel of the low half (logical columns 2t, 2t+1) of K step s; the high half (2t+8, 2t+9) follows 2 channels on
template <int C> __device__ __forceinline__ int pws_kch(int s, int t) {
  if (C >= 32) return 32 * (s >> 1) + 8 * t + 4 * (s & 1);
  if (C == 16) return 4 * t;
  return 2 * t;
}

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// copies of one source's pieces for this thread's two rows; slots advance by 512 bytes (32 lanes x 16)
template <int C> __device__ __forceinline__ uint32_t pws_cp_src(uint32_t slot, const bf16* r0, const bf16* r1, int t) {
  constexpr int PC = PwsSrc<C>::PC, PB = PwsSrc<C>::PB;
#pragma unroll
  for (int h = 0; h < PC; ++h) {
    cp_piece<PB, true>(slot, r0 + 32 * h + (PB / 2) * t);
    cp_piece<PB, true>(slot + 512, r1 + 32 * h + (PB / 2) * t);
    slot += 1024;
  }
  return slot;
}
// A fragments of one source from the FIFO (piece order as pws_cp_src)
template <int C, int BASE, int KTOT> __device__ __forceinline__ uint32_t pws_frag_a(uint32_t slot, uint32_t (&a)[KTOT][4]) {
  if constexpr (C >= 32) {
#pragma unroll
    for (int h = 0; h < C / 32; ++h) {
      const uint4 v0 = lds16(slot), v1 = lds16(slot + 512);
      a[BASE + 2 * h][0] = v0.x; a[BASE + 2 * h][1] = v1.x; a[BASE + 2 * h][2] = v0.y; a[BASE + 2 * h][3] = v1.y;
      a[BASE + 2 * h + 1][0] = v0.z; a[BASE + 2 * h + 1][1] = v1.z; a[BASE + 2 * h + 1][2] = v0.w; a[BASE + 2 * h + 1][3] = v1.w;
      slot += 1024;
    }
  } else if constexpr (C > 0) {
    const uint4 v0 = lds16(slot), v1 = lds16(slot + 512);      // only the first PB bytes were copied
    a[BASE][0] = v0.x; a[BASE][1] = v1.x;
    a[BASE][2] = C == 16 ? v0.y : 0u; a[BASE][3] = C == 16 ? v1.y : 0u;
    slot += 1024;
  }
  return slot;
}

template <int C0, int C1, int COUT, bool STATS>
__global__ void __launch_bounds__(PWS_THREADS, PwsOcc<C0, C1, COUT>::V) pw_stream_kernel(const PwsParams p) {
  constexpr int KS0 = PwsSrc<C0>::KS, KS1 = PwsSrc<C1>::KS, NT = COUT / 8, NCH = 2 * NT;
  constexpr int SP = NT >= 4 ? NT / 4 : 1;             // pieces per row of an output-shaped side tensor
  constexpr int SB = NT >= 4 ? 16 : 4 * NT;            // their bytes
  // B fragments in registers while they fit (<= 32 words); 64 -> 64 channels keeps them in shared memory, one conflict-free
  // word per lane and fragment (two ld.shared per MMA on a kernel that moves 4 KB per 32 MMAs)
  constexpr bool BSM = (KS0 + KS1) * NT > 16;
  constexpr int NBF = (KS0 + KS1) * NT;
  __shared__ uint32_t wfr[BSM ? NBF : 1][2][32];
  extern __shared__ uint4 pws_fifo[];
  __shared__ float wsum[STATS ? PWS_WARPS : 1][COUT], wsq[STATS ? PWS_WARPS : 1][COUT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();

  // B fragments: n-tile j, fragment column g  <->  output channel 2 NT (g >> 1) + 2 j + (g & 1)
  uint32_t b0[BSM ? 1 : KS0 + KS1][BSM ? 1 : NT], b1[BSM ? 1 : KS0 + KS1][BSM ? 1 : NT];
  auto bfrag = [&](int s, int j, uint32_t& lo, uint32_t& hi) {          // K step s (both sources), n-tile j, this lane
    const int co = NCH * (g >> 1) + 2 * j + (g & 1);
    const bf16* wrow = p.wt + (size_t)co * p.kt + p.k_base;
    const int ch = s < KS0 ? pws_kch<C0>(s, t) : C0 + pws_kch<C1 ? C1 : 8>(s - KS0, t);
    const bool half = s < KS0 ? C0 == 8 : C1 == 8;                      // 8-channel source: the high half of the K step is empty
    lo = __ldg(reinterpret_cast<const uint32_t*>(wrow + ch));
    hi = half ? 0u : __ldg(reinterpret_cast<const uint32_t*>(wrow + ch + 2));
  };
  if constexpr (BSM) {
    for (int e = warp; e < NBF; e += PWS_WARPS) bfrag(e / NT, e % NT, wfr[e][0][lane], wfr[e][1][lane]);
    __syncthreads();
  } else {
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int s = 0; s < KS0 + KS1; ++s) bfrag(s, j, b0[s][j], b1[s][j]);
  }
  float bias_r[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) bias_r[i] = p.bias ? __ldg(p.bias + NCH * t + i) : 0.f;
  float ps[STATS ? NCH : 1], pq[STATS ? NCH : 1];
#pragma unroll
  for (int i = 0; i < (STATS ? NCH : 1); ++i) { ps[i] = 0.f; pq[i] = 0.f; }

  const int wmask = (1 << p.lw) - 1, hmask = (1 << p.lh) - 1;
  const int ngroups = p.M >> 4;
  const int nside = p.nadd + (p.mask ? 1 : 0);
  const int gstride = gridDim.x * PWS_WARPS;
  const uint32_t fifo = (uint32_t)__cvta_generic_to_shared(pws_fifo) + (uint32_t)(warp * p.depth * p.stage_bytes) + (uint32_t)lane * 16u;
  auto pix = [&](int m, size_t& src, size_t& dst, int& n, int& h, int& w) {
    w = m & wmask; h = (m >> p.lw) & hmask; n = m >> (p.lw + p.lh);
    const size_t strided = ((((size_t)n << (p.lh + 1)) + 2 * h) << (p.lw + 1)) + 2 * w;
    src = p.in_stride == 1 ? (size_t)m : strided;
    dst = p.out_stride == 1 ? (size_t)m : strided;
  };
  // every piece this thread will consume of group `grp`, into FIFO stage `stage`
  auto issue = [&](int grp, int stage) {
    if (grp < ngroups) {
      size_t src[2], dst[2];
      int pn[2], ph[2], pw[2];
      pix((grp << 4) + g, src[0], dst[0], pn[0], ph[0], pw[0]);
      pix((grp << 4) + g + 8, src[1], dst[1], pn[1], ph[1], pw[1]);
      uint32_t slot = fifo + (uint32_t)(stage * p.stage_bytes);
      slot = pws_cp_src<C0>(slot, p.x0 + src[0] * C0, p.x0 + src[1] * C0, t);
      if constexpr (C1 > 0) slot = pws_cp_src<C1>(slot, p.x1 + src[0] * C1, p.x1 + src[1] * C1, t);
      for (int k = 0; k < nside; ++k) {
        const bf16* base = k < p.nadd ? p.side[k] : p.mask;
        const int sh = k < p.nadd ? p.shift[k] : 0;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const size_t q = sh ? ((((size_t)pn[r] << (p.lh - sh)) + (ph[r] >> sh)) << (p.lw - sh)) + (pw[r] >> sh) : dst[r];
          const bf16* row = base + q * COUT + NCH * t;
#pragma unroll
          for (int i = 0; i < SP; ++i) {
            if (sh) cp_piece<SB, false>(slot, row + 8 * i); else cp_piece<SB, true>(slot, row + 8 * i);
            slot += 512;
          }
        }
      }
    }
    cp_commit();           // an empty group past the end keeps the wait count uniform
  };

  const int grp0 = blockIdx.x * PWS_WARPS + warp;
  for (int s = 0; s < p.depth - 1; ++s) issue(grp0 + s * gstride, s);
  int stage = 0;
  for (int grp = grp0; grp < ngroups; grp += gstride) {
    int nstage = stage + p.depth - 1;
    if (nstage >= p.depth) nstage -= p.depth;
    issue(grp + (p.depth - 1) * gstride, nstage);
    cp_wait(p.depth - 1);
    uint32_t slot = fifo + (uint32_t)(stage * p.stage_bytes);
    if (++stage == p.depth) stage = 0;
    uint32_t a[KS0 + KS1][4];
    slot = pws_frag_a<C0, 0>(slot, a);
    slot = pws_frag_a<C1, KS0>(slot, a);
    // ---- K steps
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      acc[j][0] = bias_r[2 * j]; acc[j][1] = bias_r[2 * j + 1]; acc[j][2] = bias_r[2 * j]; acc[j][3] = bias_r[2 * j + 1];
    }
#pragma unroll
    for (int s = 0; s < KS0 + KS1; ++s)
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        if constexpr (BSM) mma_bf16_16816(acc[j], a[s], wfr[s * NT + j][0][lane], wfr[s * NT + j][1][lane]);
        else mma_bf16_16816(acc[j], a[s], b0[s][j], b1[s][j]);
      }
    // ---- epilogue: + addends, ReLU, mask, store, statistics (order as conv_tc2's epilogue)
    for (int k = 0; k < p.nadd; ++k) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < SP; ++i) {
          const uint4 v = lds16(slot);
          slot += 512;
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < (NT < 4 ? NT : 4); ++j) {
            acc[4 * i + j][2 * r] += bf_lo(w[j]);
            acc[4 * i + j][2 * r + 1] += bf_hi(w[j]);
          }
        }
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = fmaxf(acc[j][i], 0.f);
      }
    }
    if (p.mask) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < SP; ++i) {
          const uint4 v = lds16(slot);
          slot += 512;
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < (NT < 4 ? NT : 4); ++j) {
            acc[4 * i + j][2 * r] = bf_lo(w[j]) > 0.f ? acc[4 * i + j][2 * r] : 0.f;
            acc[4 * i + j][2 * r + 1] = bf_hi(w[j]) > 0.f ? acc[4 * i + j][2 * r + 1] : 0.f;
          }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int m = (grp << 4) + g + 8 * r;
      size_t dst = (size_t)m;
      if (p.out_stride != 1) {
        const int w = m & wmask, h = (m >> p.lw) & hmask, n = m >> (p.lw + p.lh);
        dst = ((((size_t)n << (p.lh + 1)) + 2 * h) << (p.lw + 1)) + 2 * w;
      }
      uint32_t pk[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        __nv_bfloat162 hh = __floats2bfloat162_rn(acc[j][2 * r], acc[j][2 * r + 1]);
        pk[j] = *reinterpret_cast<uint32_t*>(&hh);
      }
      bf16* orow = p.out + dst * COUT + NCH * t;
      if constexpr (NT >= 4) {
#pragma unroll
        for (int i = 0; i < NT / 4; ++i)
          *(reinterpret_cast<uint4*>(orow) + i) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      } else if constexpr (NT == 2) {
        *reinterpret_cast<uint2*>(orow) = make_uint2(pk[0], pk[1]);
      } else {
        *reinterpret_cast<uint32_t*>(orow) = pk[0];
      }
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
  cp_wait(0);
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

template <int C0, int C1, int COUT, bool STATS>
int pws_launch_k(PwsParams& p, cudaStream_t st) {
  constexpr int NT = COUT / 8, SP = NT >= 4 ? NT / 4 : 1, PWS_OCC = PwsOcc<C0, C1, COUT>::V;
  const int nside = p.nadd + (p.mask ? 1 : 0);
  p.stage_bytes = (2 * (PwsSrc<C0>::PC + PwsSrc<C1>::PC) + 2 * SP * nside) * 512;
  // FIFO depth: ~6 KB in flight per warp (24 warps per SM: ~140 KB, three times what latency x bandwidth asks for - the
  // queueing under load is what the margin is for), at least 3 groups, bounded by shared memory for PWS_OCC CTAs per SM
  int depth = 1 + (6144 + p.stage_bytes - 1) / p.stage_bytes;
  if (depth < 3) depth = 3;
  if (depth > 8) depth = 8;
  const int budget = (227 * 1024 - PWS_OCC * 2048) / PWS_OCC;          // per CTA, static shared memory and reserve deducted
  while (depth > 2 && PWS_WARPS * depth * p.stage_bytes > budget) --depth;
  p.depth = depth;
  const int smem = PWS_WARPS * depth * p.stage_bytes;
  if (smem > 200 * 1024) return -100;                // more side tensors than the FIFO holds: the tcgen05 kernel takes it
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(pw_stream_kernel<C0, C1, COUT, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { rsa_set_error("pw_stream: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RSA_ERR_CUDA; }
    configured = true;
  }
  const int ngroups = p.M >> 4;
  int grid = (ngroups + PWS_WARPS - 1) / PWS_WARPS;
  int per_sm = (227 * 1024) / (smem + 2048);
  if (per_sm > PWS_OCC) per_sm = PWS_OCC;
  if (per_sm < 1) per_sm = 1;
  const int cap = per_sm * rsa_num_sms();            // one resident wave, grid-stride inside
  if (grid > cap) grid = cap;
  cudaError_t le = launch_pdl(pw_stream_kernel<C0, C1, COUT, STATS>, dim3(grid), dim3(PWS_THREADS), (size_t)smem, st, p);
  if (le != cudaSuccess) { rsa_set_error("pw_stream: launch: %s", cudaGetErrorString(le)); return RSA_ERR_CUDA; }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
template <int C0, int C1, int COUT>
int pws_launch(PwsParams& p, cudaStream_t st) {
  return p.stats ? pws_launch_k<C0, C1, COUT, true>(p, st) : pws_launch_k<C0, C1, COUT, false>(p, st);
}

template <int C0, int C1>
int pws_by_cout(int Cout, PwsParams& p, cudaStream_t st) {
  switch (Cout) {
    case 8: return pws_launch<C0, C1, 8>(p, st);
    case 16: return pws_launch<C0, C1, 16>(p, st);
    case 32: return pws_launch<C0, C1, 32>(p, st);
    case 64: return pws_launch<C0, C1, 64>(p, st);
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
  p.out = (bf16*)out; p.mask = (const bf16*)mask; p.stats = stats;
  p.M = (int)M; p.lw = ilog2(W); p.lh = ilog2(H);
  p.in_stride = in_stride; p.out_stride = out_stride; p.kt = kt; p.k_base = k_base;
  p.relu = relu;
  p.nadd = 0;
  for (int u = 0; u < PWS_MAXSIDE; ++u) { p.side[u] = nullptr; p.shift[u] = 0; }
  for (int u = 0; u < nup; ++u) {        // addend order of conv_tc2's epilogue: up-sampled, residual, running sum
    if (((uintptr_t)up_ptrs[u] & 15) || up_shifts[u] > p.lw || up_shifts[u] > p.lh) return -100;
    p.side[p.nadd] = (const bf16*)up_ptrs[u]; p.shift[p.nadd++] = up_shifts[u];
  }
  if (residual) p.side[p.nadd++] = (const bf16*)residual;
  if (accumulate) p.side[p.nadd++] = (const bf16*)out;
  if (C1 == 32) return pws_by_cout<32, 32>(Cout, p, st);
  switch (C0) {
    case 8: return pws_by_cout<8, 0>(Cout, p, st);
    case 16: return pws_by_cout<16, 0>(Cout, p, st);
    case 32: return pws_by_cout<32, 0>(Cout, p, st);
    case 64: return pws_by_cout<64, 0>(Cout, p, st);
  }
  return -100;
}

// =====================================================================================================
// Weight gradient of the same thin 1x1 convolutions:  dw[ci][co] += sum_pix x[pix, ci] * dz[pix, co]
// (Conv2D 1x1 backward-filter, model2.py:37,84,92,103).  Two tensors are read once and 4-16 KB come out: HBM-bound.  The
// tcgen05 kernel behind rsa_pw_wgrad_tc takes 36-74 us per launch at 16 x 256 x 256 (8-channel operands ride on half-empty
// 16-channel TMA boxes) where the bytes take 13-21 us.  Same recipe as the forward kernel above: a warp owns groups of 16
// pixels = one K step of mma.sync.m16n8k16 with M = ci, N = co; both operands are K-major in memory ([pixel][channel]), so
// the fragments come out of a per-warp shared-memory FIFO through ldmatrix.trans (rows padded to an odd number of 16-byte
// chunks: conflict-free), the FIFO is filled DEPTH groups ahead with cp.async, and the Cin x Cout accumulator lives in
// registers for the whole kernel.  Warps are summed in a fixed order inside the CTA; CTAs add with fp32 atomics like every
// other weight-gradient kernel here.
// =====================================================================================================
namespace {

constexpr int PWG_THREADS = 256, PWG_WARPS = 8;
// FIFO depth: four groups ahead, three where a group is large (64 channels on one side)
__host__ __device__ constexpr int pwg_depth(int stage_bytes) { return stage_bytes > 3000 ? 3 : 4; }

struct PwgParams {
  const bf16* x; const bf16* dz; float* dw;
  int ldw, M, lw, lh, in_stride;
};

// row pitch (bytes) of a [16 pixel][C channel] tile: an odd number of 16-byte chunks
__host__ __device__ constexpr int pwg_pitch(int C) { return C == 8 ? 16 : C * 2 + 16; }

__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(PWG_THREADS, 2) pw_wgrad_stream_kernel(const PwgParams p) {
  constexpr int MT = CIN >= 16 ? CIN / 16 : 1, NT = COUT / 8;
  constexpr int XP = pwg_pitch(CIN), ZP = pwg_pitch(COUT);
  constexpr int XB = 16 * XP, STAGE = (16 * (XP + ZP) + 127) & ~127;
  constexpr int XCH = CIN / 8, ZCH = COUT / 8;                 // 16-byte chunks per pixel row
  constexpr int PWG_DEPTH = pwg_depth(STAGE);
  static_assert(MT * NT <= 16, "accumulator must fit in registers");
  extern __shared__ uint4 pwg_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();
  const uint32_t fifo = (uint32_t)__cvta_generic_to_shared(pwg_smem) + (uint32_t)(warp * PWG_DEPTH * STAGE);
  const int wmask = (1 << p.lw) - 1, hmask = (1 << p.lh) - 1;
  const int ngroups = p.M >> 4, gstride = gridDim.x * PWG_WARPS;
  auto issue = [&](int grp, int stage) {
    if (grp < ngroups) {
      const uint32_t base = fifo + (uint32_t)(stage * STAGE);
      for (int c = lane; c < 16 * XCH; c += 32) {
        const int row = c / XCH, ch = c % XCH, m = (grp << 4) + row;
        size_t src = (size_t)m;
        if (p.in_stride != 1) {
          const int w = m & wmask, h = (m >> p.lw) & hmask, n = m >> (p.lw + p.lh);
          src = ((((size_t)n << (p.lh + 1)) + 2 * h) << (p.lw + 1)) + 2 * w;
        }
        cp_piece<16, true>(base + (uint32_t)(row * XP + ch * 16), p.x + src * CIN + ch * 8);
      }
      for (int c = lane; c < 16 * ZCH; c += 32) {
        const int row = c / ZCH, ch = c % ZCH;
        cp_piece<16, true>(base + (uint32_t)(XB + row * ZP + ch * 16), p.dz + ((size_t)(grp << 4) + row) * COUT + ch * 8);
      }
    }
    cp_commit();
  };
  float acc[MT][NT][4];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j) { acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f; }
  // ldmatrix row addresses of this lane: matrices {pixels 0-7 | 8-15} x {channel block, next channel block}
  const int lrow = (lane & 7) + ((lane >> 4) << 3), lblk = (lane >> 3) & 1;
  const int grp0 = blockIdx.x * PWG_WARPS + warp;
  for (int s = 0; s < PWG_DEPTH - 1; ++s) issue(grp0 + s * gstride, s);
  int stage = 0;
  for (int grp = grp0; grp < ngroups; grp += gstride) {
    int nstage = stage + PWG_DEPTH - 1;
    if (nstage >= PWG_DEPTH) nstage -= PWG_DEPTH;
    __syncwarp();                    // every lane has finished reading the stage that is refilled now
    issue(grp + (PWG_DEPTH - 1) * gstride, nstage);
    cp_wait(PWG_DEPTH - 1);
    __syncwarp();                    // the copies of all lanes are visible
    const uint32_t base = fifo + (uint32_t)(stage * STAGE);
    if (++stage == PWG_DEPTH) stage = 0;
    uint32_t a[MT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      // x4.trans: r0 = (px 0-7, ci 16i..+7), r1 = (px 0-7, ci 16i+8..), r2 = (px 8-15, ci 16i..), r3 = (px 8-15, ci 16i+8..)
      // fragment order a0 a1 a2 a3 = (ci lo, px lo), (ci hi, px lo), (ci lo, px hi), (ci hi, px hi): the same
      const int blk = CIN >= 16 ? 2 * i + lblk : 0;      // 8 channels: the upper rows of the m-tile repeat the lower ones
      ldsm_x4_t(base + (uint32_t)(lrow * XP + blk * 16), a[i]);
    }
#pragma unroll
    for (int j = 0; j < NT; j += 2) {
      // r0 = (px 0-7, co 8j..), r1 = (px 0-7, co 8j+8..), r2 = (px 8-15, co 8j..), r3 = (px 8-15, co 8j+8..)
      uint32_t b[4];
      const int blk = NT >= 2 ? j + lblk : 0;
      ldsm_x4_t(base + (uint32_t)(XB + lrow * ZP + blk * 16), b);
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        mma_bf16_16816(acc[i][j], a[i], b[0], b[2]);
        if constexpr (NT >= 2) mma_bf16_16816(acc[i][j + 1], a[i], b[1], b[3]);
      }
    }
  }
  cp_wait(0);
  __syncthreads();                   // the FIFO is dead: reuse it for the cross-warp sum
  float* red = reinterpret_cast<float*>(pwg_smem);             // [warp][CIN][COUT]
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int ci = 16 * i + g + 8 * r, co = 8 * j + 2 * t;
        if (ci < CIN) {
          red[(warp * CIN + ci) * COUT + co] = acc[i][j][2 * r];
          red[(warp * CIN + ci) * COUT + co + 1] = acc[i][j][2 * r + 1];
        }
      }
  __syncthreads();
  for (int e = threadIdx.x; e < CIN * COUT; e += PWG_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < PWG_WARPS; ++w) s += red[w * CIN * COUT + e];
    atomicAdd(p.dw + (size_t)(e / COUT) * p.ldw + (e % COUT), s);
  }
}

template <int CIN, int COUT>
int pwg_launch(const PwgParams& p, cudaStream_t st) {
  constexpr int STAGE = (16 * (pwg_pitch(CIN) + pwg_pitch(COUT)) + 127) & ~127;
  constexpr int FIFO = PWG_WARPS * pwg_depth(STAGE) * STAGE, RED = PWG_WARPS * CIN * COUT * 4;
  constexpr int SMEM = FIFO > RED ? FIFO : RED;
  static_assert(SMEM <= 100 * 1024, "two CTAs per SM");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(pw_wgrad_stream_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) { rsa_set_error("pw_wgrad_stream: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RSA_ERR_CUDA; }
    configured = true;
  }
  const int ngroups = p.M >> 4;
  int grid = (ngroups + PWG_WARPS - 1) / PWG_WARPS;
  if (grid > 2 * rsa_num_sms()) grid = 2 * rsa_num_sms();
  cudaError_t le = launch_pdl(pw_wgrad_stream_kernel<CIN, COUT>, dim3(grid), dim3(PWG_THREADS), (size_t)SMEM, st, p);
  if (le != cudaSuccess) { rsa_set_error("pw_wgrad_stream: launch: %s", cudaGetErrorString(le)); return RSA_ERR_CUDA; }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
template <int CIN>
int pwg_by_cout(int Cout, const PwgParams& p, cudaStream_t st) {
  constexpr int MT = CIN >= 16 ? CIN / 16 : 1;
  switch (Cout) {
    case 8: return pwg_launch<CIN, 8>(p, st);
    case 16: return pwg_launch<CIN, 16>(p, st);
    case 32: return pwg_launch<CIN, 32>(p, st);
    case 64: if constexpr (MT * 8 <= 16) return pwg_launch<CIN, 64>(p, st);
  }
  return -100;
}

}  // namespace

/* Streaming path of rsa_pw_wgrad_tc (same contract): -100 = "not mine".  RSA_PW_STREAM=0 switches it off. */
int rsa_pw_wgrad_stream_dispatch(const void* x, const void* dz, float* dw, int ldw, int N, int H, int W, int Cin, int Cout,
                                 int in_stride, cudaStream_t st) {
  static const int enabled = getenv("RSA_PW_STREAM") ? atoi(getenv("RSA_PW_STREAM")) : 1;
  if (!enabled) return -100;
  const long long M = (long long)N * H * W;
  if (M % 16 || M > 0x7fffffffLL || (H & (H - 1)) || (W & (W - 1))) return -100;
  if (((uintptr_t)x | (uintptr_t)dz) & 15) return -100;
  PwgParams p;
  p.x = (const bf16*)x; p.dz = (const bf16*)dz; p.dw = dw; p.ldw = ldw; p.M = (int)M; p.lw = ilog2(W); p.lh = ilog2(H);
  p.in_stride = in_stride;
  switch (Cin) {
    case 8: return pwg_by_cout<8>(Cout, p, st);
    case 16: return pwg_by_cout<16>(Cout, p, st);
    case 32: return pwg_by_cout<32>(Cout, p, st);
    case 64: return pwg_by_cout<64>(Cout, p, st);
  }
  return -100;
}
