// conv_tc.cu — tensor-core helpers of the bf16 path that are not convolutions forward:
//   * conv_tc_wgrad_kernel: weight gradient of the dilated 3x3 convolutions with C >= 128 (and the large dilations at
//     C = 64), MN-major operands, K = pixels split across CTAs (Conv2D backward-filter, model2.py:19-24);
//   * pw_wgrad_kernel: weight gradient of the 1x1 convolutions (model2.py:37,84,92,103-111);
//   * bias_grad_kernel, pack_weights_kernel (fp32 HWIO master weights -> bf16 [tap][Cout][Cin] / [tap][Cin][Cout] copies);
//   * rsa_conv_tc_supported: which 3x3 layers the tensor-core path takes.
// The forward / data-gradient kernels live in conv_tc2.cu (C >= 128 and every 1x1) and conv_tc3.cu (C = 32 / 64); the
// first-generation forward kernel that used to be here was retired in round 2 (nothing issued it any more).
#include "tc_common.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int NTHREADS = 192;

struct PackEntry { long long src_off, fwd_off, bwd_off; int taps, Cin, Cout, pad; };

constexpr int PACK_MAXL = 512;
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ params, bf16* __restrict__ shadow,
                                                           const PackEntry* __restrict__ table, int nlayers) {
  // 32x32 tiles through shared memory: reads (HWIO, co contiguous) and both writes ([ci][co] copy and the
  // transposed [co][ci] copy) are coalesced.  The tiles of all layers form one flat list (prefix sums rebuilt per
  // block from the table) that a single wave of blocks walks with a grid stride - layer sizes differ by 1000x, so a
  // (tiles, layer) grid would launch mostly empty blocks.
  __shared__ float tile[32][33];
  __shared__ int prefix[PACK_MAXL + 1];
  for (int l = threadIdx.x; l < nlayers; l += 256) {
    const PackEntry e = table[l];
    prefix[l + 1] = e.taps * ((e.Cin + 31) / 32) * ((e.Cout + 31) / 32);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    prefix[0] = 0;
    for (int l = 0; l < nlayers; ++l) prefix[l + 1] += prefix[l];
  }
  __syncthreads();
  const int total = prefix[nlayers];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  int layer = 0;
  for (int ft = blockIdx.x; ft < total; ft += gridDim.x) {
    while (prefix[layer + 1] <= ft) ++layer;                   // flat ids only grow
    const PackEntry e = table[layer];
    const int coutp = e.pad > 0 ? e.pad : e.Cout;
    const int tci = (e.Cin + 31) / 32, tco = (e.Cout + 31) / 32;
    const float* src = params + e.src_off;
    const int t = ft - prefix[layer];
    const int tap = t / (tci * tco);
    const int r = t % (tci * tco);
    const int ci0 = (r / tco) * 32, co0 = (r % tco) * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ci = ci0 + ty + 8 * i, co = co0 + tx;
      float v = 0.f;
      if (ci < e.Cin && co < e.Cout) {
        const long long idx = ((long long)tap * e.Cin + ci) * e.Cout + co;
        v = src[idx];
        shadow[e.bwd_off + idx] = __float2bfloat16_rn(v);
      }
      tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = co0 + ty + 8 * i, ci = ci0 + tx;
      if (ci < e.Cin && co < e.Cout)
        shadow[e.fwd_off + ((long long)tap * coutp + co) * e.Cin + ci] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
    __syncthreads();
  }
}

}  // namespace

/* 1 if the tcgen05 path handles this shape: channels multiples of 32 (64 above 32), spatial side a
 * power of two >= 4 with a 128-pixel tile. */
extern "C" int rsa_conv_tc_supported(int N, int H, int W, int Cin, int Cout) {
  auto okc = [](int c) { return c == 32 || (c >= 64 && c % 64 == 0); };
  if (!okc(Cin) || !okc(Cout)) return 0;
  if (H != W || (W & (W - 1)) || W < 4) return 0;
  return 1;
}


namespace {
// =====================================================================================================
// Weight gradient on the tensor cores.
//   dW[tap][ci][co] = sum_pix x[pix + off(tap), ci] * dy[pix, co]           (model2.py:19-24 backward-filter)
// GEMM view per tap: D[M=ci][N=co] += X_tap^T[ci x pix] * dY[pix x co], K = pixels.  Both operands are
// "MN-major" in UMMA terms (channels contiguous, the pixel/K index strides by one NHWC row), which is
// exactly what the forward's TMA boxes already deliver: a {KA channels, TW, TH, TN} box of 64 pixels is a
// [64 K-rows][KA*2 bytes] swizzled tile.  One M=128 MMA spans 128/KA consecutive sub-buffers (LBO = one
// sub-buffer), so for thin layers several taps share one instruction: C=32 -> 4 taps per MMA, C=64 -> 2,
// C>=128 -> one tap x 128 input channels.  The ninth tap rides in an overlapping group ((5..8) / (7,8)),
// whose duplicate rows are ignored by the epilogue.  The pixel dimension is split across CTAs; partial
// sums are added to the fp32 gradient with atomics (the buffer is zeroed once per step).
// =====================================================================================================
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct WgradTcParams {
  int N, H, W, Cin, Cout, dil;
  int TW, TH, TN, tiles_w, tiles_h, ntiles;   // 64-pixel tiles
  int tiles_per_cta;
  float* dw;
};

// MN-major smem descriptor: LBO = byte distance between consecutive M/N atoms (sub-buffers),
// SBO = 8 K-rows of one swizzle span.
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, int swizzle_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(((8 * swizzle_bytes) >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(swizzle_bytes == 128 ? 2 : (swizzle_bytes == 64 ? 4 : 6)) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {
  return make_idesc(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
}

// CC = channel class of Cin (32, 64 or 128 = ">=128, processed in blocks of 128"), NB = N tile (co)
template <int CC, int NB>
struct WgradCfg {
  static constexpr int KA = CC >= 64 ? 64 : 32;                    // channels per A sub-buffer row
  static constexpr int KB = NB >= 64 ? 64 : 32;
  static constexpr int TP = 64;                                    // pixels per K tile
  static constexpr int A_SUB = TP * KA * 2;                        // bytes
  static constexpr int B_SUB = TP * KB * 2;
  static constexpr int NSUB_A = CC == 128 ? 6 : 9;                 // 3 taps x 2 halves | 9 taps
  static constexpr int NSUB_B = NB / KB;
  static constexpr int NSLOT = CC == 128 ? 3 : (CC == 64 ? 5 : 3);
  static constexpr int STAGE_BYTES = NSUB_A * A_SUB + NSUB_B * B_SUB;
  static constexpr int STAGES = CC == 128 ? 3 : (CC == 64 ? 2 : 4);
  static constexpr int RING = STAGES * STAGE_BYTES;
  static constexpr int TMEM_COLS = NSLOT * NB <= 128 ? 128 : 512;
  static constexpr int TOTAL = RING + (2 * STAGES + 1) * 8 + 16 + 1024;
  // first A sub-buffer of MMA slot s
  __host__ __device__ static constexpr int slot_sub(int s) {
    return CC == 128 ? 2 * s : (CC == 64 ? (s < 4 ? 2 * s : 7) : (s < 2 ? 4 * s : 5));
  }
};

constexpr int WG_THREADS = 384;   // warp 0 TMA, warps 1..NSLOT MMA issuers (one accumulator slot each), warps 8..11 epilogue

template <int CC, int NB>
__global__ void __launch_bounds__(WG_THREADS) conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                 const __grid_constant__ CUtensorMap tmDY,
                                                                 const WgradTcParams p) {
  using Cfg = WgradCfg<CC, NB>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::RING);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // blockIdx.y -> (tap group, ci block, co block)
  const int ncob = p.Cout / NB;
  const int ncib = CC == 128 ? p.Cin / 128 : 1;
  int by = blockIdx.y;
  const int cob = by % ncob; by /= ncob;
  const int cib = by % ncib; by /= ncib;
  const int tapg = by;                               // 0..2 for CC==128, 0 otherwise
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(p.ntiles, t_begin + p.tiles_per_cta);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], Cfg::NSLOT); }
    mbar_init(tmem_full, Cfg::NSLOT);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int tt = t;
        const int tw = tt % p.tiles_w; tt /= p.tiles_w;
        const int th = tt % p.tiles_h; tt /= p.tiles_h;
        const int n0 = tt * p.TN, h0 = th * p.TH, w0 = tw * p.TW;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
        for (int sb = 0; sb < Cfg::NSUB_A; ++sb) {
          const int tap = CC == 128 ? tapg * 3 + sb / 2 : sb;
          const int c0 = CC == 128 ? cib * 128 + (sb & 1) * 64 : 0;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma_load_4d(sa + sb * Cfg::A_SUB, &tmX, &full_bar[stage], c0, w0 + dx * p.dil, h0 + dy * p.dil, n0);
        }
#pragma unroll
        for (int sb = 0; sb < Cfg::NSUB_B; ++sb)
          tma_load_4d(sa + Cfg::NSUB_A * Cfg::A_SUB + sb * Cfg::B_SUB, &tmDY, &full_bar[stage], cob * NB + sb * Cfg::KB,
                      w0, h0, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp <= Cfg::NSLOT) {
    // one issuing thread per accumulator slot: the single-thread issue rate (~100 cycles per MMA) was the limit
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_mn(128, NB);
      const int s = warp - 1;
      int stage = 0, phase = 0;
      uint32_t accum = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::NSUB_A * Cfg::A_SUB;
#pragma unroll
        for (int k = 0; k < Cfg::TP / 16; ++k) {
          const uint64_t bdesc = make_mnmajor_desc(sb + k * 16 * Cfg::KB * 2, Cfg::KB * 2, Cfg::B_SUB);
          const uint64_t adesc = make_mnmajor_desc(sa + Cfg::slot_sub(s) * Cfg::A_SUB + k * 16 * Cfg::KA * 2,
                                                   Cfg::KA * 2, Cfg::A_SUB);
          umma_bf16(tmem_base + (uint32_t)(s * NB), adesc, bdesc, idesc, accum);
          accum = 1;
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);
    }
  } else if (warp < 8) {
    // idle warps (role slots up to 7 MMA issuers)
  } else {
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (t_begin < t_end) {
      const int m = q * 32 + lane;
#pragma unroll 1
      for (int s = 0; s < Cfg::NSLOT; ++s) {
        int tap, ci;
        bool valid = true;
        if (CC == 128) { tap = tapg * 3 + s; ci = cib * 128 + m; }
        else if (CC == 64) { tap = (s < 4 ? 2 * s : 7) + m / 64; ci = m % 64; valid = s < 4 || m >= 64; }
        else { tap = (s < 2 ? 4 * s : 5) + m / 32; ci = m % 32; valid = s < 2 || m >= 96; }
        float* dst = p.dw + ((size_t)tap * p.Cin + ci) * p.Cout + cob * NB;
#pragma unroll 1
        for (int c0 = 0; c0 < NB; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * NB + c0), v);
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(dst + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                         __uint_as_float(v[j + 3]));
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS));
  }
}

template <int CC, int NB>
int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmDY, const WgradTcParams& p, dim3 grid, cudaStream_t st) {
  using Cfg = WgradCfg<CC, NB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_wgrad_kernel<CC, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::TOTAL);
    if (e != cudaSuccess) { rsa_set_error("conv_tc_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RSA_ERR_CUDA; }
    configured = true;
  }
  cudaError_t le = launch_pdl(conv_tc_wgrad_kernel<CC, NB>, grid, dim3(WG_THREADS), (size_t)Cfg::TOTAL, st, tmX, tmDY, p);
  if (le != cudaSuccess) { rsa_set_error("conv_tc_wgrad: launch: %s", cudaGetErrorString(le)); return RSA_ERR_CUDA; }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

}  // namespace

/* dw[tap][ci][co] (fp32, HWIO, zeroed by the caller once per step) += sum_pix x[pix+off(tap), ci] * dy[pix, co]
 * for the 3x3 taps at dilation dil; x, dy bf16 NHWC.  Replaces cuDNN's Conv2D backward-filter behind
 * model2.py:19-24,153-178.  Supported when rsa_conv_tc_supported() and Cin == Cout. */
extern "C" int rsa_conv_tc_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                                 int dil, void* stream) {
  RSA_REQUIRE(x && dy && dw, RSA_ERR_SHAPE, "conv_tc_wgrad: null pointer");
  RSA_REQUIRE(rsa_conv_tc_supported(N, H, W, Cin, Cout) && Cin == Cout, RSA_ERR_SHAPE,
              "conv_tc_wgrad: unsupported shape N=%d H=%d W=%d Cin=%d Cout=%d", N, H, W, Cin, Cout);
  EncodeTiledFn enc = get_encode();
  RSA_REQUIRE(enc, RSA_ERR_CUDA, "conv_tc_wgrad: cuTensorMapEncodeTiled not available from the driver");
  WgradTcParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.dil = dil; p.dw = dw;
  p.TW = W < 16 ? W : 16;
  p.TH = H < 64 / p.TW ? H : 64 / p.TW;
  p.TN = 64 / (p.TW * p.TH);
  p.tiles_w = W / p.TW; p.tiles_h = H / p.TH;
  p.ntiles = p.tiles_w * p.tiles_h * ((N + p.TN - 1) / p.TN);
  const int CC = Cin >= 128 ? 128 : Cin;
  const int NB = Cout >= 128 ? 128 : Cout;
  const int KA = CC >= 64 ? 64 : 32, KB = NB >= 64 ? 64 : 32;
  const int ygroups = (CC == 128 ? 3 * (Cin / 128) : 1) * (Cout / NB);
  // split the pixel reduction into ONE resident wave (a CTA owns an SM: ~190 KB of shared memory), keep >= 4 tiles per CTA
  int want = ygroups >= 96 ? 1 : rsa_num_sms() / ygroups;   // deep levels: enough (ci,co,tap) groups already
  int maxsplit = (p.ntiles + 3) / 4;
  int split = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
  if (split < 1) split = 1;
  p.tiles_per_cta = (p.ntiles + split - 1) / split;
  split = (p.ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  CUtensorMap tmX, tmDY;
  auto encode = [&](CUtensorMap* tm, const void* base, int C, int KC) -> CUresult {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = encode(&tmX, x, Cin, KA);
  RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc_wgrad: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
  r = encode(&tmDY, dy, Cout, KB);
  RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc_wgrad: cuTensorMapEncodeTiled(dy) failed (%d)", (int)r);
  dim3 grid((unsigned)split, (unsigned)ygroups);
  cudaStream_t st = (cudaStream_t)stream;
  if (CC == 32) return launch_wgrad<32, 32>(tmX, tmDY, p, grid, st);
  if (CC == 64) return launch_wgrad<64, 64>(tmX, tmDY, p, grid, st);
  return launch_wgrad<128, 128>(tmX, tmDY, p, grid, st);
}

namespace {
// =====================================================================================================
// Weight gradient of the 1x1 convolutions (Conv2D 1x1 bwd-filter; model2.py:37,84,92,101-111):
//   dW[k][co] += sum_pix x[pix*stride, k] * dz[pix, co]
// Same MN-major formulation as the 3x3 kernel above with one tap.  M = 128 rows of the accumulator hold
// min(Cin,128) input channels; thinner sources (16/32/64 channels) fill the remaining M atoms with replicas of
// the same TMA box (ignored by the epilogue) so that every MMA is the M=128 shape.
// =====================================================================================================
struct PwWgradParams {
  int N, H, W, Cin, Cout, in_stride, ldw;
  int TW, TH, TN, tiles_w, tiles_h, ntiles, tiles_per_cta;
  int ncob, ncib;
  float* dw;
};

template <int KA, int KB, int NB>
struct PwCfg {
  static constexpr int TP = 64;
  static constexpr int NA = 128 / KA;                 // A atoms per MMA
  static constexpr int NBA = NB / KB;                 // B atoms
  static constexpr int A_SUB = TP * KA * 2;
  static constexpr int B_SUB = TP * KB * 2;
  static constexpr int STAGE_BYTES = NA * A_SUB + NBA * B_SUB;
  static constexpr int STAGES = 5;
  static constexpr int RING = STAGES * STAGE_BYTES;
  static constexpr int TMEM_COLS = NB < 32 ? 32 : NB;
  static constexpr int TOTAL = RING + (2 * STAGES + 1) * 8 + 16 + 1024;
};

template <int KA, int KB, int NB>
__global__ void __launch_bounds__(NTHREADS) pw_wgrad_kernel(const __grid_constant__ CUtensorMap tmX,
                                                            const __grid_constant__ CUtensorMap tmDY,
                                                            const PwWgradParams p) {
  using Cfg = PwCfg<KA, KB, NB>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::RING);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cob = blockIdx.y % p.ncob, cib = blockIdx.y / p.ncob;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  const int t_end = min(p.ntiles, t_begin + p.tiles_per_cta);
  const int real_atoms = p.Cin >= 128 ? 2 : 1;                   // Cin >= 128 -> two 64-channel atoms, else one

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmDY);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        int tt = t;
        const int tw = tt % p.tiles_w; tt /= p.tiles_w;
        const int th = tt % p.tiles_h; tt /= p.tiles_h;
        const int n0 = tt * p.TN, h0 = th * p.TH, w0 = tw * p.TW;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
        for (int a = 0; a < Cfg::NA; ++a) {
          const int c0 = cib * 128 + (a % real_atoms) * KA;       // replicas re-load atom (a % real_atoms)
          tma_load_4d(sa + a * Cfg::A_SUB, &tmX, &full_bar[stage], c0, w0 * p.in_stride, h0 * p.in_stride, n0);
        }
#pragma unroll
        for (int b = 0; b < Cfg::NBA; ++b)
          tma_load_4d(sa + Cfg::NA * Cfg::A_SUB + b * Cfg::B_SUB, &tmDY, &full_bar[stage], cob * NB + b * KB, w0, h0, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_mn(128, NB);
      int stage = 0, phase = 0;
      uint32_t accum = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::NA * Cfg::A_SUB;
#pragma unroll
        for (int k = 0; k < Cfg::TP / 16; ++k) {
          const uint64_t adesc = make_mnmajor_desc(sa + k * 16 * KA * 2, KA * 2, Cfg::A_SUB);
          const uint64_t bdesc = make_mnmajor_desc(sb + k * 16 * KB * 2, KB * 2, Cfg::B_SUB);
          umma_bf16(tmem_base, adesc, bdesc, idesc, accum);
          accum = 1;
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full);
    }
  } else {
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    if (t_begin < t_end) {
      const int m = q * 32 + lane;
      const int cinb = p.Cin >= 128 ? 128 : p.Cin;
      const bool valid = m < cinb;
      float* dst = p.dw + (size_t)(cib * 128 + m) * p.ldw + cob * NB;
      const int ncol = p.Cout - cob * NB < NB ? p.Cout - cob * NB : NB;
#pragma unroll 1
      for (int c0 = 0; c0 < NB; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (c0 + j < ncol)      // ncol is a multiple of 16
              red_add_v4(dst + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                         __uint_as_float(v[j + 3]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS));
  }
}

// SW32 variant of the MN-major descriptor is selected by swizzle_bytes == 32 inside make_mnmajor_desc2
template <int KA, int KB, int NB>
int launch_pw(const CUtensorMap& tmX, const CUtensorMap& tmDY, const PwWgradParams& p, dim3 grid, cudaStream_t st) {
  using Cfg = PwCfg<KA, KB, NB>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(pw_wgrad_kernel<KA, KB, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::TOTAL);
    if (e != cudaSuccess) { rsa_set_error("pw_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RSA_ERR_CUDA; }
    configured = true;
  }
  cudaError_t le = launch_pdl(pw_wgrad_kernel<KA, KB, NB>, grid, dim3(NTHREADS), (size_t)Cfg::TOTAL, st, tmX, tmDY, p);
  if (le != cudaSuccess) { rsa_set_error("pw_wgrad: launch: %s", cudaGetErrorString(le)); return RSA_ERR_CUDA; }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
}  // namespace

/* dw[k*ldw + co] (fp32, zeroed by the caller once per step) += sum_pix x[n, h*s, w*s, k] * dz[n,h,w,co] for a 1x1
 * convolution source with Cin channels (power of two >= 16) and Cout in {16,32,64,128k}; x bf16 [N,H*s,W*s,Cin],
 * dz bf16 [N,H,W,Cout].  Replaces the Conv2D 1x1 backward-filter behind model2.py:37,84,92,101-111. */
extern "C" int rsa_pw_wgrad_tc(const void* x, const void* dz, float* dw, int ldw, int N, int H, int W, int Cin,
                               int Cout, int in_stride, void* stream) {
  RSA_REQUIRE(x && dz && dw, RSA_ERR_SHAPE, "pw_wgrad_tc: null pointer");
  auto p2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
  RSA_REQUIRE(p2(H) && p2(W) && W >= 4 && H >= 4 && p2(Cin) && Cin >= 8 && Cin <= 1024 && p2(Cout) && Cout >= 8 && Cout <= 1024 &&
                  (in_stride == 1 || in_stride == 2), RSA_ERR_SHAPE,
              "pw_wgrad_tc: unsupported shape N=%d H=%d W=%d Cin=%d Cout=%d stride=%d", N, H, W, Cin, Cout, in_stride);
  if (Cin <= 64 && Cout <= 64) {      // thin layers are HBM-bound: streaming kernel (pw_stream.cu), same contract
    const int rc = rsa_pw_wgrad_stream_dispatch(x, dz, dw, ldw, N, H, W, Cin, Cout, in_stride, (cudaStream_t)stream);
    if (rc != -100) return rc;
  }
  EncodeTiledFn enc = get_encode();
  RSA_REQUIRE(enc, RSA_ERR_CUDA, "pw_wgrad_tc: cuTensorMapEncodeTiled not available from the driver");
  PwWgradParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.in_stride = in_stride; p.ldw = ldw; p.dw = dw;
  p.TW = W < 16 ? W : 16;
  p.TH = H < 64 / p.TW ? H : 64 / p.TW;
  p.TN = 64 / (p.TW * p.TH);
  p.tiles_w = W / p.TW; p.tiles_h = H / p.TH;
  p.ntiles = p.tiles_w * p.tiles_h * ((N + p.TN - 1) / p.TN);
  // 8-channel tensors: a 16-channel TMA box whose upper half is out of bounds (zero filled)
  const int KA = Cin >= 64 ? 64 : (Cin < 16 ? 16 : Cin);
  const int NB = Cout >= 128 ? 128 : (Cout < 16 ? 16 : Cout);
  const int KB = NB >= 64 ? 64 : NB;
  p.ncob = (Cout + NB - 1) / NB; p.ncib = Cin >= 128 ? Cin / 128 : 1;
  const int ygroups = p.ncob * p.ncib;
  int want = (2 * rsa_num_sms() + ygroups - 1) / ygroups;
  int maxsplit = (p.ntiles + 3) / 4;
  int split = want > maxsplit ? maxsplit : want;
  if (split < 1) split = 1;
  p.tiles_per_cta = (p.ntiles + split - 1) / split;
  split = (p.ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  auto swz = [](int kc) { return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B); };
  auto encode = [&](CUtensorMap* tm, const void* base, int C, int KC, int s) -> CUresult {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W * s, (cuuint64_t)H * s, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * s * C * 2, (cuuint64_t)H * s * W * s * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)KC, (cuuint32_t)(p.TW * s), (cuuint32_t)(p.TH * s), (cuuint32_t)p.TN};
    cuuint32_t es[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, swz(KC), CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap tmX, tmDY;
  CUresult r = encode(&tmX, x, Cin, KA, in_stride);
  RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "pw_wgrad_tc: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
  r = encode(&tmDY, dz, Cout, KB, 1);
  RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "pw_wgrad_tc: cuTensorMapEncodeTiled(dz) failed (%d)", (int)r);
  dim3 grid((unsigned)split, (unsigned)ygroups);
  cudaStream_t st = (cudaStream_t)stream;
#define PW_CASE(ka, kb, nb) if (KA == ka && KB == kb && NB == nb) return launch_pw<ka, kb, nb>(tmX, tmDY, p, grid, st);
  PW_CASE(16, 16, 16) PW_CASE(16, 32, 32) PW_CASE(16, 64, 64) PW_CASE(16, 64, 128)
  PW_CASE(32, 16, 16) PW_CASE(32, 32, 32) PW_CASE(32, 64, 64) PW_CASE(32, 64, 128)
  PW_CASE(64, 16, 16) PW_CASE(64, 32, 32) PW_CASE(64, 64, 64) PW_CASE(64, 64, 128)
#undef PW_CASE
  RSA_REQUIRE(false, RSA_ERR_SHAPE, "pw_wgrad_tc: no kernel for KA=%d KB=%d NB=%d", KA, KB, NB);
}

namespace {
// per-channel column sums of dy -> up to 4 fp32 bias gradients (all branches of a ResBlock-a share d(out))
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(const T* __restrict__ dy, long long M, int C, float* o0,
                                                        float* o1, float* o2, float* o3, int rows_per_block) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sm[];
  const int tpr = C / V, rpi = 256 / tpr, tid = threadIdx.x;
  const int cg = tid % tpr, r0 = tid / tpr;
  float s[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s[i] = 0.f;
  long long rbeg = (long long)blockIdx.x * rows_per_block;
  long long rend = rbeg + rows_per_block < M ? rbeg + rows_per_block : M;
  for (long long r = rbeg + r0; r < rend; r += rpi) {
    float v[V];
    ldv<T>(dy + r * C + cg * V, v);
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] += v[i];
  }
#pragma unroll
  for (int i = 0; i < V; ++i) sm[tid * V + i] = s[i];
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    int g = c / V, i = c % V;
    float a = 0.f;
    for (int r = 0; r < rpi; ++r) a += sm[(r * tpr + g) * V + i];
    atomicAdd(o0 + c, a);
    if (o1) atomicAdd(o1 + c, a);
    if (o2) atomicAdd(o2 + c, a);
    if (o3) atomicAdd(o3 + c, a);
  }
}
}  // namespace

/* db_k[c] += sum_m dy[m,c] for up to four bias gradients (NULL to skip): the conv biases of all branches of a
 * ResBlock-a receive the same gradient (Add, model2.py:27-31). */
extern "C" int rsa_bias_grad(const void* dy, int dtype, long long M, int C, float* db0, float* db1, float* db2,
                             float* db3, void* stream) {
  RSA_REQUIRE(dy && db0 && M > 0, RSA_ERR_SHAPE, "bias_grad: bad args");
  int rows = (int)ceil_div64(M, (int64_t)rsa_num_sms() * 4);
  if (rows < 64) rows = 64;
  int grid = (int)ceil_div64(M, rows);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RSA_BF16) {
    RSA_REQUIRE(C % 8 == 0 && (256 % (C / 8)) == 0 && C / 8 <= 256, RSA_ERR_SHAPE, "bias_grad: C=%d unsupported", C);
    bias_grad_kernel<bf16><<<grid, 256, 256 * 8 * sizeof(float), st>>>((const bf16*)dy, M, C, db0, db1, db2, db3, rows);
  } else if (dtype == RSA_F32) {
    RSA_REQUIRE(C % 4 == 0 && (256 % (C / 4)) == 0 && C / 4 <= 256, RSA_ERR_SHAPE, "bias_grad: C=%d unsupported", C);
    bias_grad_kernel<float><<<grid, 256, 256 * 4 * sizeof(float), st>>>((const float*)dy, M, C, db0, db1, db2, db3, rows);
  } else {
    RSA_REQUIRE(false, RSA_ERR_DTYPE, "bias_grad: bad dtype");
  }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* bf16 weight copies for the tensor-core path, all layers in one launch.  table (device): n entries of
 * {int64 src_off (into params, fp32 HWIO), int64 fwd_off, int64 bwd_off (into shadow, bf16), int32 taps, Cin, Cout, pad}. */
extern "C" int rsa_pack_weights_tc(const float* params, void* shadow, const void* table, int nlayers,
                                   long long max_elems, void* stream) {
  RSA_REQUIRE(params && shadow && table && nlayers > 0, RSA_ERR_SHAPE, "pack_weights_tc: bad args");
  RSA_REQUIRE(nlayers <= PACK_MAXL, RSA_ERR_SHAPE, "pack_weights_tc: more than %d layers", PACK_MAXL);
  int gx = (int)ceil_div64(max_elems, 1024);          // never more blocks than the largest layer has tiles
  if (gx < 1) gx = 1;
  if (gx > rsa_num_sms() * 8) gx = rsa_num_sms() * 8; // one resident wave
  pack_weights_kernel<<<gx, 256, 0, (cudaStream_t)stream>>>(params, (bf16*)shadow, (const PackEntry*)table, nlayers);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
