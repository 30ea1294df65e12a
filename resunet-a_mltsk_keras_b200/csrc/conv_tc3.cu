// conv_tc3.cu — thin-layer (C = 32 / 64) dilated 3x3 'same' convolution on tcgen05.
//
// History of what bounds these layers (profiles/r1b_*, scripts/exp_mma_rate.cu, scripts/iso_tc3.py):
//   * conv_tc2 on them sat at 7 % of the tensor pipe: nine TMA round trips and a ~5.7k-cycle register epilogue on 4 warps per
//     128-pixel tile.  Hence resident weights, halo tiles whose nine taps are shifted UMMA descriptors into ONE box (measured:
//     with base_offset = 0 any row shift and any stride-byte-offset address a TMA-swizzled tile correctly), KT MMA-issuing
//     warps (a single thread sustains one N=32 MMA per ~95 cycles, four reach the shared-memory operand limit of 40);
//   * round 2 switched parts of the kernel off one at a time (RSA_TC3_DEBUG, scripts/iso_tc3.py; C = 32, d = 1, batch 16):
//     whole kernel 50 us, MMA + loads without the epilogue 35 us, epilogue + stores without any MMA 44 us, and the bare
//     barrier protocol with no data moved at all 23 us.  The epilogue PIPELINE - eight warps meeting at a staging buffer per
//     128-pixel sub-tile, a ninth thread storing it and handing the buffer back, a side ring fed by that same thread - was
//     the bottleneck, not the tensor core and not memory.
// So the epilogue is now eight independent pipelines: every epilogue warp owns 32 pixels x 32 channels slices (one
// tcgen05.ld.32x32b.x32, the TMEM lane quarter it may read), its own two staging buffers, its own TMA stores (box 8 x 4
// pixels x 32 channels) and its own TMA side-input loads with a private mbarrier - no barrier is shared between epilogue
// warps, the only hand-offs left are accumulator full / empty with the MMA warps.
//
//   * weights of all branches stay resident in shared memory for the life of the (persistent) CTA;
//   * an item is 16 x (8*KT) pixels = KT accumulators; |d| <= 3: one halo box per item, larger dilations one box per tap.
//     Box mode is bound by L2 -> SM bandwidth (nine 32 KB boxes per 512-pixel item = 9.3 TB/s over 148 SMs at the measured
//     64 us): releasing its stages in pairs (half the tcgen05.commit count) made it 15 % SLOWER because the producer then
//     refills at twice the granularity - the ring depth, not the barrier traffic, is what it lives on;
//   * up to four branches (ResBlock-a: dilations 1/3/15/31, model2.py:23-31) accumulate into the same TMEM tiles,
//     so the branch sum and the identity add happen once, in the epilogue;
//   * BatchNorm statistics of the stored values: per-thread partial sums over all the slices of a warp, one 32-wide shuffle
//     butterfly per warp at the end, per-warp shared-memory slots summed in a fixed order (reproducible), one double atomic
//     per channel per CTA.
//
// Replaces cuDNN's Conv2D forward / backward-data behind keras Conv2D(C, 3, dilation_rate=d, padding='same') at
// model2.py:19-24,153-178 for the C = 32 / 64 layers (enc1/2, dec1/2, heads).
#include "tc_common.cuh"

namespace {

constexpr int T3_MAXBR = 4;
constexpr int T3_EW = 8;              // epilogue warps
constexpr int T3_SLICE = 32 * 64;     // bytes of one epilogue slice: 32 pixels x 32 bf16 channels

// compile-time geometry of one channel class (C = 32: SWIZZLE_64B rows, C = 64: SWIZZLE_128B rows)
template <int C>
struct T3 {
  static constexpr int PITCH = C * 2;               // bytes per pixel = swizzle span
  static constexpr int BOXB = 128 * PITCH;          // one 16x8-pixel sub-tile
  static constexpr int WBYTES = 9 * C * PITCH;      // one branch's weights
  static constexpr uint32_t LAYOUT = C == 32 ? 4u : 2u;   // UMMA layout code: SWIZZLE_64B / SWIZZLE_128B
};
// warp roles: 0..7 epilogue (the TMEM lane quarter a warp may read is warp % 4), 8 TMA producer + TMEM allocation,
// 9..10 MMA issuers.  scripts/exp_mma_align.cu (profiles/r2_exp_mma_align.txt): ONE thread issues at most one tcgen05.mma per
// ~100 cycles, however many independent accumulator chains it interleaves (1 thread x 4 chains 96-100 cycles per MMA,
// 2 threads x 2 chains 54-60, 4 threads 40 = the shared-memory operand limit of an N=32 MMA), but the limit is per THREAD,
// not per warp: two warps with two issuing lanes each also reach 40.  So every sub-tile (accumulator chain) of an item gets
// its own issuing lane, KT / 2 lanes in each of the two MMA warps, and the CTA stays at 11 warps = 3 per scheduler = 168
// registers per thread, which the epilogue (32 accumulator values + 64 statistics partial sums) needs; a fourth warp per
// scheduler would cap every thread at 128 registers and spill the epilogue.
template <int KT> struct T3Warps {
  static constexpr int EPI0 = 0;
  static constexpr int PROD = T3_EW;
  static constexpr int MMA0 = T3_EW + 1;
  static constexpr int NMW = 2;                                  // MMA warps
  static constexpr int LPW = KT >= 2 ? KT / 2 : 1;               // issuing lanes per MMA warp (KT = 1: warp MMA0 lane 0 only)
  static constexpr int THREADS = (T3_EW + 1 + NMW) * 32;
};

struct Tc3Params {
  int N, H, W;
  int nbr;
  int dil[T3_MAXBR];      // signed: negative = data gradient (taps mirrored)
  int halo[T3_MAXBR];     // 1: one halo box per item, 0: one box per tap
  int band;               // 1: every branch has a large dilation and W is a multiple of 128: an item is KT image rows of 128
                          // pixels (sub-tile = one row), each tap ROW is one band box KT x (128 + 2|d|) pixels and its three
                          // column shifts are start addresses into it - 3 boxes of L2 traffic per item instead of 9
  int IH, IWP;            // item height / width in pixels: 16 x 8 KT, or KT x 128 (band)
  int items, tiles_w, tiles_h;
  int nstages, slot_bytes;
  int alt;                      // 1: the two MMA warps take alternate items (every branch in halo mode, nstages even)
  int sdepth;                   // side-input buffers per epilogue warp: sdepth - 1 slices are in flight ahead of the one in use
  int has_add, has_mask;        // addend present (residual or previous out), ReLU mask present
  int has_bnx, bnr_relu;        // fused BatchNorm backward: BN input slices; the BN was followed by ReLU
  const double* bnr_stats;      // {sum, sumsq} of the BN input (forward statistics)
  double bnr_count;
  float bnr_eps;
  const float* bnr_gamma;
  const float* bnr_beta;
  const float* bias[T3_MAXBR];
  double* stats;
  uint8_t* out;           // output tensor (direct-store mode)
  int direct;             // 1: the epilogue writes its slices with 16-byte global stores instead of staging + TMA store
  int relu;
  int debug;              // diagnostic (RSA_TC3_DEBUG, results are wrong): 1 no MMAs, 2 no epilogue data movement, 4 no TMA
                          // operand loads, 8 no TMA stores, 16 no tcgen05.ld, 32 no proxy fence, 64 no staging stores -
                          // isolates which pipeline bounds the kernel (scripts/iso_tc3.py)
};

__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld32(taddr, v); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& q, float* t) {
  t[0] = __uint_as_float(q.x << 16); t[1] = __uint_as_float(q.x & 0xffff0000u);
  t[2] = __uint_as_float(q.y << 16); t[3] = __uint_as_float(q.y & 0xffff0000u);
  t[4] = __uint_as_float(q.z << 16); t[5] = __uint_as_float(q.z & 0xffff0000u);
  t[6] = __uint_as_float(q.w << 16); t[7] = __uint_as_float(q.w & 0xffff0000u);
}
// K-major swizzled descriptor split in halves: hi carries the stride byte offset (distance between 8-row groups)
template <int C> __device__ __forceinline__ uint32_t t3_desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (T3<C>::LAYOUT << 29);
}
__device__ __forceinline__ uint64_t t3_desc(uint32_t hi, uint32_t saddr) {
  return ((uint64_t)hi << 32) | (uint64_t)(((saddr >> 4) & 0x3FFF) | (1u << 16));
}

// a: operand boxes per branch; w: weights per branch; out / add / mask / bnx: 8 x 4 pixel x 32 channel slices (SWIZZLE_64B)
struct Tc3Maps { CUtensorMap a[T3_MAXBR]; CUtensorMap w[T3_MAXBR]; CUtensorMap out; CUtensorMap add; CUtensorMap mask; CUtensorMap bnx; };

// shared-memory carve-up (offsets from the 1024-aligned base)
struct Tc3Smem {
  int w_off, st_off, side_off, ring_off, misc_off, bar_off, total;
  __host__ __device__ Tc3Smem(int C, int nbr, int nside, int sdepth, int nstages, int slot_bytes, int direct) {
    w_off = 0;
    st_off = (nbr * 9 * C * C * 2 + 1023) & ~1023;
    side_off = st_off + (direct ? 0 : T3_EW * 2 * T3_SLICE);  // [warp][2] staging slices (none with direct global stores)
    ring_off = side_off + T3_EW * sdepth * nside * T3_SLICE;  // [warp][sdepth][nside] side slices
    misc_off = ring_off + nstages * slot_bytes;               // bias[C], csum[8][32], csq[8][32], BN coefficients [4][C]
    bar_off = misc_off + (5 * C + 2 * T3_EW * 32) * 4;
    total = bar_off + (2 * nstages + 8 + 4 * T3_EW + 1) * 8 + 16 + 1024;
  }
};

// STATS = the launch accumulates BatchNorm statistics (64 partial sums per epilogue thread: 168 registers).  Launches without
// them are compiled to 112 registers: 352 threads x 168 registers take 90 % of the register file and keep every other kernel
// off the SM, 352 x 112 leave room for two CTAs of a bandwidth-bound kernel (BatchNorm apply / backward, 256 threads x 40-64
// registers) to run BESIDE the persistent convolution CTA - the lanes of the step executor put them on concurrent streams.
template <int C, int KT, bool STATS>
__global__ void __maxnreg__(STATS ? 168 : 112) conv_tc3_kernel(const __grid_constant__ Tc3Maps maps, const Tc3Params p) {
  constexpr int PITCH = T3<C>::PITCH, BOXB = T3<C>::BOXB, WBYTES = T3<C>::WBYTES;
  constexpr int EPI0 = T3Warps<KT>::EPI0, PROD = T3Warps<KT>::PROD, MMA0 = T3Warps<KT>::MMA0;
  constexpr int LPW = T3Warps<KT>::LPW;
  constexpr int ALT_L = KT >= 2 ? 2 : 1, ALT_CH = KT / ALT_L;     // alternating mode: issuing lanes per warp, chains per lane
  constexpr int NACC = 4;                                  // accumulator stages (power of two)
  constexpr int SPI = C == 32 ? KT / 2 : KT;        // slices an epilogue warp handles per item
  static_assert(C == 64 || KT % 2 == 0, "C = 32: the two epilogue warp groups alternate over the sub-tiles of an item");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nside = p.has_add + p.has_mask + p.has_bnx;
  const Tc3Smem L(C, p.nbr, nside, p.sdepth, p.nstages, p.slot_bytes, p.direct);
  uint8_t* wsm = smem + L.w_off;
  uint8_t* ring = smem + L.ring_off;
  float* bias_s = reinterpret_cast<float*>(smem + L.misc_off);
  float* bnc = bias_s + C;                   // [4][C]: invstd, -mean*invstd, gamma, beta of the fused BatchNorm backward
  float* csum = bnc + 4 * C;                 // [8 warps][32]: one slot per epilogue warp, summed in a fixed order
  float* csq = csum + T3_EW * 32;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* empty_bar = full_bar + p.nstages;
  uint64_t* tfull = empty_bar + p.nstages;   // [NACC]
  uint64_t* tempty = tfull + NACC;           // [NACC]
  uint64_t* sbar = tempty + NACC;            // [warp][4] side slices landed (private to one epilogue warp)
  uint64_t* wbar = sbar + 4 * T3_EW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool has_stats = STATS;
  // NACC accumulator stages of KT sub-tiles each.  The MMA warps may run NACC items ahead of the epilogue; with two stages
  // the loop MMA(it) -> commit -> epilogue sees it (~1.7k cycles later, scripts/iso_tc3.py) -> epilogue has read the item ->
  // MMA(it + 2) bounded an item at (commit latency + epilogue latency + MMA time) / 2 = ~4.5k cycles against 2.9k of MMAs.
  constexpr uint32_t TMEM_COLS = NACC * KT * C <= 128 ? 128 : (NACC * KT * C <= 256 ? 256 : 512);
  static_assert(NACC * KT * C <= 512, "accumulator ring exceeds TMEM");

  if (threadIdx.x == 0) {
    for (int b = 0; b < p.nbr; ++b) { prefetch_tmap(&maps.a[b]); prefetch_tmap(&maps.w[b]); }
    prefetch_tmap(&maps.out);
    if (p.has_add) prefetch_tmap(&maps.add);
    if (p.has_mask) prefetch_tmap(&maps.mask);
    if (p.has_bnx) prefetch_tmap(&maps.bnx);
    const int nissue = p.alt ? ALT_L : KT;          // threads that commit a stage / an accumulator
    for (int s = 0; s < p.nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], nissue); }
    for (int s = 0; s < NACC; ++s) { mbar_init(&tfull[s], nissue); mbar_init(&tempty[s], T3_EW); }
    for (int s = 0; s < 4 * T3_EW; ++s) mbar_init(&sbar[s], 1);
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == PROD) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  pdl_wait();                        // everything above is private to the CTA; global memory is touched only below
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (warp >= EPI0 && warp < EPI0 + C / 32) {
    const int c = (warp - EPI0) * 32 + lane;
    float b = 0.f;
    for (int k = 0; k < p.nbr; ++k) if (p.bias[k]) b += __ldg(p.bias[k] + c);
    bias_s[c] = b;
    if (p.has_bnx) {
      float mean, inv;
      bn_mean_invstd(p.bnr_stats, p.bnr_count, C, c, p.bnr_eps, nullptr, nullptr, mean, inv);
      // ReLU mask: gamma * xhat + beta > 0  <=>  x * A + B > 0;  sum g * xhat = inv * (sum g * x) - mean * inv * (sum g)
      const float ga = p.bnr_gamma[c];
      bnc[c] = ga * inv; bnc[C + c] = p.bnr_beta[c] - ga * mean * inv; bnc[2 * C + c] = inv; bnc[3 * C + c] = -mean * inv;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == PROD) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(wbar, p.nbr * WBYTES);
      for (int b = 0; b < p.nbr; ++b)
        for (int t = 0; t < 9; ++t) tma_load_3d(wsm + b * WBYTES + t * C * PITCH, &maps.w[b], wbar, 0, 0, t);
      int stage = 0, phase = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int r = item;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int n = r, h0 = th * p.IH, w0 = tw * p.IWP;
        for (int b = 0; b < p.nbr; ++b) {
          const int d = p.dil[b], ad = d < 0 ? -d : d;
          if (p.band) {
            for (int dyi = 0; dyi < 3; ++dyi) {
              const int ch = h0 + (dyi - 1) * d;
              if (ch + KT <= 0 || ch >= p.H) continue;
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (p.debug & 4) mbar_expect_tx(&full_bar[stage], 0);
              else {
                mbar_expect_tx(&full_bar[stage], KT * (128 + 2 * ad) * PITCH);
                tma_load_4d(ring + stage * p.slot_bytes, &maps.a[b], &full_bar[stage], 0, w0 - ad, ch, n);
              }
              if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
          } else if (p.halo[b]) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (p.debug & 4) mbar_expect_tx(&full_bar[stage], 0);
            else {
              mbar_expect_tx(&full_bar[stage], (16 + 2 * ad) * (8 * KT + 2 * ad) * PITCH);
              tma_load_4d(ring + stage * p.slot_bytes, &maps.a[b], &full_bar[stage], 0, w0 - ad, h0 - ad, n);
            }
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          } else {
            for (int tap = 0; tap < 9; ++tap) {
              const int ch = h0 + (tap / 3 - 1) * d, cw = w0 + (tap % 3 - 1) * d;
              if (ch + 16 <= 0 || ch >= p.H || cw + 8 * KT <= 0 || cw >= p.W) continue;
              mbar_wait(&empty_bar[stage], phase ^ 1);
              if (p.debug & 4) mbar_expect_tx(&full_bar[stage], 0);
              else {
                mbar_expect_tx(&full_bar[stage], KT * BOXB);
                tma_load_4d(ring + stage * p.slot_bytes, &maps.a[b], &full_bar[stage], 0, cw, ch, n);
              }
              if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp >= MMA0) {
    // ===== MMA issuers =====
    // A tcgen05.commit holds its thread until the MMAs it issued have drained (~1.1k cycles between two items during which
    // that thread issues nothing; with all issuing threads on the same item the tensor pipe idles with them, scripts/iso_tc3.py:
    // 4.0k cycles per 512-pixel item against 2.9k of MMAs).  Halo-only launches therefore give the two MMA warps ALTERNATE
    // items: while one warp commits and waits, the other one's MMAs keep the pipe busy.  Each warp has ALT_L issuing lanes
    // (one thread issues at most one MMA per ~100 cycles) that interleave ALT_CH accumulator chains each.  The operand ring is
    // walked by position (item * branches + branch); an even number of stages keeps every stage with one warp, so a warp
    // never waits for a barrier phase whose predecessor it has not seen complete.
    if (p.alt) {
      if (lane < ALT_L) {
        constexpr uint32_t idesc = make_idesc(128, C);
        const int gsel = warp - MMA0, s0 = lane * ALT_CH;
        mbar_wait(wbar, 0);
        tc_fence_after();
        const uint32_t wbase = smem_u32(wsm);
        const uint32_t bhi = t3_desc_hi<C>(8 * PITCH);
        for (int it = gsel; (long long)blockIdx.x + (long long)it * gridDim.x < p.items; it += 2) {
          const int as = it & (NACC - 1);
          mbar_wait(&tempty[as], ((it / NACC) & 1) ^ 1);
          tc_fence_after();
          const uint32_t acc = tmem_base + (uint32_t)((as * KT + s0) * C);
          uint32_t started = 0;
          for (int b = 0; b < p.nbr; ++b) {
            const int d = p.dil[b], ad = d < 0 ? -d : d;
            const uint32_t wb = wbase + b * WBYTES;
            const int Wh = 8 * KT + 2 * ad;
            const uint32_t ahi = t3_desc_hi<C>(Wh * PITCH);
            const int pos = it * p.nbr + b, stage = pos % p.nstages, phase = (pos / p.nstages) & 1;
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + stage * p.slot_bytes) + (uint32_t)((ad * Wh + ad + 8 * s0) * PITCH);
            const int rowb = d * Wh * PITCH, colb = d * PITCH;
            if (!(p.debug & 1)) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint32_t at = sa + (tap / 3 - 1) * rowb + (tap % 3 - 1) * colb;
                const uint64_t bdesc = t3_desc(bhi, wb + tap * C * PITCH);
#pragma unroll
                for (int k = 0; k < C / 16; ++k)
#pragma unroll
                  for (int u = 0; u < ALT_CH; ++u)
                    umma_bf16(acc + (uint32_t)(u * C), t3_desc(ahi, at + 8 * u * PITCH) + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                              idesc, started | (tap > 0) | (k > 0));
              }
            }
            started = 1;
            umma_commit(&empty_bar[stage]);
          }
          umma_commit(&tfull[as]);
        }
      }
      // fall through to the common tail
    } else
    // same-item mode (a branch with one box per tap): lane l of warp MMA0+m owns sub-tile s = m * LPW + l of every item
    if (const int s = (warp - MMA0) * LPW + lane; lane < LPW && s < KT) {
      constexpr uint32_t idesc = make_idesc(128, C);
      mbar_wait(wbar, 0);
      tc_fence_after();
      const uint32_t wbase = smem_u32(wsm);
      const uint32_t bhi = t3_desc_hi<C>(8 * PITCH);
      int stage = 0, phase = 0, it = 0;
      int tw = (int)blockIdx.x % p.tiles_w, th = ((int)blockIdx.x / p.tiles_w) % p.tiles_h;
      const int dtw = (int)gridDim.x % p.tiles_w, dth = ((int)gridDim.x / p.tiles_w) % p.tiles_h;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int h0 = th * p.IH, w0 = tw * p.IWP;
        tw += dtw; th += dth;                       // next item's tile coordinates without a division
        if (tw >= p.tiles_w) { tw -= p.tiles_w; ++th; }
        if (th >= p.tiles_h) th -= p.tiles_h;
        const int as = it & (NACC - 1);
        mbar_wait(&tempty[as], ((it / NACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)((as * KT + s) * C);
        uint32_t started = 0;
        for (int b = 0; b < p.nbr; ++b) {
          const int d = p.dil[b], ad = d < 0 ? -d : d;
          const uint32_t wb = wbase + b * WBYTES;
          if (p.band) {
            // sub-tile s = image row h0 + s: 16 groups of 8 consecutive pixels (SBO = 8 pixels); tap (dy, dx) = band dy, window
            // shifted by dx * d pixels inside its row of 128 + 2|d|
            const int Wb = 128 + 2 * ad;
            for (int dyi = 0; dyi < 3; ++dyi) {
              const int ch = h0 + (dyi - 1) * d;
              if (ch + KT <= 0 || ch >= p.H) continue;
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t sa = smem_u32(ring + stage * p.slot_bytes) + (uint32_t)((s * Wb + ad) * PITCH);
              if (!(p.debug & 1)) {
#pragma unroll
                for (int dxi = 0; dxi < 3; ++dxi) {
                  const uint64_t adesc = t3_desc(bhi, sa + (dxi - 1) * d * PITCH);
                  const uint64_t bdesc = t3_desc(bhi, wb + (dyi * 3 + dxi) * C * PITCH);
#pragma unroll
                  for (int k = 0; k < C / 16; ++k)
                    umma_bf16(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, started | (dxi > 0) | (k > 0));
                  }
              }
              started = 1;
              umma_commit(&empty_bar[stage]);
              if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
          } else if (p.halo[b]) {
            const int Wh = 8 * KT + 2 * ad;
            const uint32_t ahi = t3_desc_hi<C>(Wh * PITCH);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + stage * p.slot_bytes) + (uint32_t)((ad * Wh + ad + 8 * s) * PITCH);
            const int rowb = d * Wh * PITCH, colb = d * PITCH;
            if (!(p.debug & 1)) {
#pragma unroll
              for (int tap = 0; tap < 9; ++tap) {
                const uint64_t adesc = t3_desc(ahi, sa + (tap / 3 - 1) * rowb + (tap % 3 - 1) * colb);
                const uint64_t bdesc = t3_desc(bhi, wb + tap * C * PITCH);
#pragma unroll
                for (int k = 0; k < C / 16; ++k)
                  umma_bf16(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, started | (tap > 0) | (k > 0));
              }
            }
            started = 1;
            umma_commit(&empty_bar[stage]);
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          } else {
            const uint32_t ahi = t3_desc_hi<C>(8 * KT * PITCH);
            for (int tap = 0; tap < 9; ++tap) {
              const int ch = h0 + (tap / 3 - 1) * d, cw = w0 + (tap % 3 - 1) * d;
              if (ch + 16 <= 0 || ch >= p.H || cw + 8 * KT <= 0 || cw >= p.W) continue;
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint64_t adesc = t3_desc(ahi, smem_u32(ring + stage * p.slot_bytes) + 8 * s * PITCH);
              const uint64_t bdesc = t3_desc(bhi, wb + tap * C * PITCH);
              if (!(p.debug & 1)) {
#pragma unroll
                for (int k = 0; k < C / 16; ++k)
                  umma_bf16(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, started | (k > 0));
              }
              started = 1;
              umma_commit(&empty_bar[stage]);
              if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
          }
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    // ===== epilogue: eight independent warp pipelines.  Warp e = warp - EPI0 reads TMEM lane quarter q = e % 4 (pixels
    // 32q .. 32q+31 of a sub-tile = image rows 4q .. 4q+3, 8 pixels each) and channel block c0 .. c0+31.  C = 32: the two
    // groups of four warps alternate over the sub-tiles of an item; C = 64: group g takes channel half g of every sub-tile.
    const int e = warp - EPI0, q = e & 3, g = e >> 2;
    const int c0 = C == 32 ? 0 : 32 * g;
    uint8_t* ybuf = smem + L.st_off + e * 2 * T3_SLICE;
    uint8_t* sbuf = smem + L.side_off + e * p.sdepth * nside * T3_SLICE;
    uint64_t* mybar = sbar + 4 * e;
    const uint32_t srow = (uint32_t)lane * 64u;                 // this thread's pixel row inside a slice
    const uint32_t sw = (uint32_t)((lane >> 1) & 3);            // SWIZZLE_64B phase of that row
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = srow + (((uint32_t)k ^ sw) << 4);
    float acc_s[STATS ? 32 : 1], acc_q[STATS ? 32 : 1];         // BatchNorm statistics of this thread's pixels (all its slices)
#pragma unroll
    for (int j = 0; j < (STATS ? 32 : 1); ++j) { acc_s[j] = 0.f; acc_q[j] = 0.f; }
    // coordinates of the slices of this warp are walked incrementally (no divisions in the loop): the "current" cursor for
    // the slice being processed, a second cursor `pf` that runs sdepth - 1 slices ahead for the side-input prefetch
    struct Cur { int item, j, n, tw, th; };
    const int dtw = (int)gridDim.x % p.tiles_w, dth = ((int)gridDim.x / p.tiles_w) % p.tiles_h,
              dn = (int)gridDim.x / (p.tiles_w * p.tiles_h);
    auto cur_init = [&](Cur& c) {
      c.item = blockIdx.x; c.j = 0;
      int r = c.item;
      c.tw = r % p.tiles_w; r /= p.tiles_w;
      c.th = r % p.tiles_h; r /= p.tiles_h;
      c.n = r;
    };
    auto cur_next = [&](Cur& c) {          // next slice of this warp; item += gridDim.x by carry arithmetic, no division
      if (++c.j < SPI) return;
      c.j = 0; c.item += gridDim.x;
      c.tw += dtw; c.th += dth; c.n += dn;
      if (c.tw >= p.tiles_w) { c.tw -= p.tiles_w; ++c.th; }
      if (c.th >= p.tiles_h) { c.th -= p.tiles_h; ++c.n; }
    };
    auto sub_of = [&](int j) { return C == 32 ? 2 * j + g : j; };
    auto issue_side = [&](const Cur& c, int idx) {       // lane 0: TMA loads of the side slices of slice idx into buffer idx % sdepth
      if (c.item >= p.items) return;
      const int b = idx % p.sdepth;
      const int hh = c.th * p.IH + (p.band ? sub_of(c.j) : 4 * q), ww = c.tw * p.IWP + (p.band ? 32 * q : 8 * sub_of(c.j));
      uint8_t* dst = sbuf + b * nside * T3_SLICE;
      mbar_expect_tx(&mybar[b], nside * T3_SLICE);
      if (p.has_add) tma_load_4d(dst, &maps.add, &mybar[b], c0, ww, hh, c.n);
      if (p.has_mask) tma_load_4d(dst + p.has_add * T3_SLICE, &maps.mask, &mybar[b], c0, ww, hh, c.n);
      if (p.has_bnx) tma_load_4d(dst + (p.has_add + p.has_mask) * T3_SLICE, &maps.bnx, &mybar[b], c0, ww, hh, c.n);
    };
    Cur cur, pf;
    cur_init(cur);
    cur_init(pf);
    int pf_idx = 0;
    if (nside && lane == 0)
      for (int k = 0; k < (p.sdepth > 1 ? p.sdepth - 1 : 1); ++k) { issue_side(pf, pf_idx); cur_next(pf); ++pf_idx; }
    int it = 0, idx = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int as = it & (NACC - 1);
      mbar_wait(&tfull[as], (it / NACC) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < SPI; ++j, ++idx) {
        const int s = sub_of(j);
        const int n = cur.n, hh = cur.th * p.IH + (p.band ? s : 4 * q), ww = cur.tw * p.IWP + (p.band ? 32 * q : 8 * s);
        cur_next(cur);
        const int sb = nside ? idx % p.sdepth : 0;
        const uint8_t* ib = sbuf + sb * nside * T3_SLICE;
        if (nside) {
          // keep sdepth - 1 slices of side data in flight: the buffer refilled here was read during the previous slice
          if (p.sdepth > 1 && lane == 0) { issue_side(pf, pf_idx); cur_next(pf); ++pf_idx; }
          mbar_wait(&mybar[sb], (idx / p.sdepth) & 1);
        }
        uint32_t v[32];
        if (!(p.debug & (2 | 16))) tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * KT + s) * C + c0), v);
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = 0;
        }
        if (j == SPI - 1) {             // this warp's last read of the accumulator stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[as]);
        }
        if (p.debug & 2) {
          if (nside && p.sdepth == 1) { __syncwarp(); if (lane == 0) { issue_side(pf, pf_idx); cur_next(pf); ++pf_idx; } }
          continue;
        }
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + c0 + i);
          f[i] = __uint_as_float(v[i]) + bv.x; f[i + 1] = __uint_as_float(v[i + 1]) + bv.y;
          f[i + 2] = __uint_as_float(v[i + 2]) + bv.z; f[i + 3] = __uint_as_float(v[i + 3]) + bv.w;
        }
        if (p.has_add) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float t[8];
            unpack8(*reinterpret_cast<const uint4*>(ib + o[k]), t);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[8 * k + i] += t[i];
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (p.has_mask) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float t[8];
            unpack8(*reinterpret_cast<const uint4*>(ib + p.has_add * T3_SLICE + o[k]), t);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[8 * k + i] = t[i] > 0.f ? f[8 * k + i] : 0.f;
          }
        }
        if constexpr (STATS) if (p.has_bnx) {
          // fused BatchNorm(+ReLU) backward reductions: f is d(relu(bn(x))); recompute the ReLU mask from x like the
          // forward did, keep g = f * mask as the stored value and accumulate {sum g, sum g * x}; the affine step from
          // sum g * x to sum g * xhat is applied once per channel at the end
          const float* cf = bnc + c0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float xv[8];
            unpack8(*reinterpret_cast<const uint4*>(ib + (p.has_add + p.has_mask) * T3_SLICE + o[k]), xv);
            if (p.bnr_relu) {
#pragma unroll
              for (int i = 0; i < 8; i += 4) {
                const int c = 8 * k + i;
                const float4 ca = *reinterpret_cast<const float4*>(cf + c), cb = *reinterpret_cast<const float4*>(cf + C + c);
                f[c] = fmaf(xv[i], ca.x, cb.x) > 0.f ? f[c] : 0.f; f[c + 1] = fmaf(xv[i + 1], ca.y, cb.y) > 0.f ? f[c + 1] : 0.f;
                f[c + 2] = fmaf(xv[i + 2], ca.z, cb.z) > 0.f ? f[c + 2] : 0.f; f[c + 3] = fmaf(xv[i + 3], ca.w, cb.w) > 0.f ? f[c + 3] : 0.f;
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc_s[8 * k + i] += f[8 * k + i]; acc_q[8 * k + i] = fmaf(f[8 * k + i], xv[i], acc_q[8 * k + i]); }
          }
        }
        if (nside && p.sdepth == 1) {       // single side buffer: refill it as soon as every lane has read it
          __syncwarp();
          if (lane == 0) { issue_side(pf, pf_idx); cur_next(pf); ++pf_idx; }
        }
        uint4 pk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          pk[k].x = pack_bf16x2(f[8 * k], f[8 * k + 1]); pk[k].y = pack_bf16x2(f[8 * k + 2], f[8 * k + 3]);
          pk[k].z = pack_bf16x2(f[8 * k + 4], f[8 * k + 5]); pk[k].w = pack_bf16x2(f[8 * k + 6], f[8 * k + 7]);
        }
        if constexpr (STATS) if (!p.has_bnx) {
          // statistics of the stored (bf16-rounded) values
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float t[8];
            unpack8(pk[k], t);
#pragma unroll
            for (int i = 0; i < 8; ++i) { acc_s[8 * k + i] += t[i]; acc_q[8 * k + i] = fmaf(t[i], t[i], acc_q[8 * k + i]); }
          }
        }
        if (p.direct) {
          // each thread holds one pixel's 32 channels = 64 contiguous bytes of the NHWC tensor: four 16-byte stores
          const int lr = p.band ? 0 : lane >> 3, lc = p.band ? lane : lane & 7;      // this lane's pixel inside the slice
          uint4* gp = reinterpret_cast<uint4*>(p.out + ((((size_t)n * p.H + hh + lr) * p.W + ww + lc) * C + c0) * 2);
#pragma unroll
          for (int k = 0; k < 4; ++k) gp[k] = pk[k];
        } else {
          // staging buffer idx & 1 was last read by this warp's store of slice idx - 2: at most one younger store may be open
          uint8_t* yb = ybuf + (idx & 1) * T3_SLICE;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          if (!(p.debug & 64)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(yb + o[k]) = pk[k];
          }
          if (!(p.debug & 32)) fence_proxy_async();
          __syncwarp();
          if (lane == 0 && !(p.debug & 8)) tma_store_4d(&maps.out, yb, c0, ww, hh, n);
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if constexpr (STATS) {
      // 32-wide butterfly reduce-scatter over the warp's 32 pixel rows: lane l ends with channel c0 + l
#pragma unroll
      for (int off = 16, nn = 16; nn >= 1; off >>= 1, nn >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < nn; ++i) {
          const float send_s = upper ? acc_s[i] : acc_s[i + nn], keep_s = upper ? acc_s[i + nn] : acc_s[i];
          const float send_q = upper ? acc_q[i] : acc_q[i + nn], keep_q = upper ? acc_q[i + nn] : acc_q[i];
          acc_s[i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
          acc_q[i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
        }
      }
      // after the halving exchanges lane l holds channel bitrev-free index: upper halves were taken when the lane bit was set
      int ch = 0;
#pragma unroll
      for (int off = 16, nn = 16; nn >= 1; off >>= 1, nn >>= 1) if (lane & off) ch += nn;
      if (p.has_bnx) acc_q[0] = fmaf(bnc[2 * C + c0 + ch], acc_q[0], bnc[3 * C + c0 + ch] * acc_s[0]);   // sum g*x -> sum g*xhat
      csum[e * 32 + ch] = acc_s[0];             // exactly one lane of one warp owns a slot: no atomics, fixed order below
      csq[e * 32 + ch] = acc_q[0];
      asm volatile("bar.sync 1, %0;" ::"n"(T3_EW * 32) : "memory");
      if (it > 0 && e < C / 32) {
        // channel c = 32 e' + lane is held by the warps with c0 == 32 e': C = 32 all eight, C = 64 the group g == e'
        const int c = e * 32 + lane;
        double s8 = 0.0, q8 = 0.0;
#pragma unroll
        for (int w = 0; w < T3_EW; ++w) {
          if (C == 64 && (w >> 2) != e) continue;
          s8 += (double)csum[w * 32 + lane];
          q8 += (double)csq[w * 32 + lane];
        }
        atomicAdd(p.stats + c, s8);
        atomicAdd(p.stats + C + c, q8);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == PROD) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

template <int C, int KT, bool STATS>
int launch3s(const Tc3Maps& maps, const Tc3Params& p, int smem_bytes, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc3_kernel<C, KT, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { rsa_set_error("conv_tc3: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RSA_ERR_CUDA; }
    configured = true;
  }
  const int grid = p.items < rsa_num_sms() ? p.items : rsa_num_sms();
  cudaError_t le = launch_pdl(conv_tc3_kernel<C, KT, STATS>, dim3(grid), dim3(T3Warps<KT>::THREADS), (size_t)smem_bytes, st, maps, p);
  if (le != cudaSuccess) { rsa_set_error("conv_tc3: launch: %s", cudaGetErrorString(le)); return RSA_ERR_CUDA; }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
template <int C, int KT>
int launch3(const Tc3Maps& maps, const Tc3Params& p, int smem_bytes, cudaStream_t st) {
  return p.stats ? launch3s<C, KT, true>(maps, p, smem_bytes, st) : launch3s<C, KT, false>(maps, p, smem_bytes, st);
}

}  // namespace

/* Shapes the thin-layer kernels accept: 32 or 64 channels in and out, H a multiple of 16, W a multiple of 32. */
extern "C" int rsa_conv_tc3_supported(int N, int H, int W, int C) {
  return (C == 32 || C == 64) && N >= 1 && H >= 16 && H % 16 == 0 && W >= 32 && W % 32 == 0;
}

/* out[n,h,w,:] = epi( sum_b sum_tap x_b[n, h+dy*dil_b, w+dx*dil_b, :] . wt_b[tap] + sum_b bias_b )
 *   epi: + residual, + out (accumulate), ReLU, mask (keep where mask > 0), in that order; bf16 NHWC throughout.
 * xs[b]: bf16 [N,H,W,C]; wts[b]: bf16 [9][C][C] K-major copies ([tap][co][ci] forward, [tap][ci][co] with a negative
 * dilation for the data gradient, as rsa_conv_tc2_fwd); biases[b] fp32[C] or NULL; nbr <= 4 branches (C = 32; one for
 * C = 64) accumulate into one TMEM tile (ResBlock-a branch sum, model2.py:23-31).  stats (double[2C], optional) +=
 * {sum, sum of squares} of the stored bf16 values (BatchNormalization batch statistics, model2.py:21).
 * bnr_x (optional, data-gradient use): the launch computes d(a) for a = [relu](BatchNorm(bnr_x)) and fuses that
 * BatchNormalization's backward reductions (FusedBatchNormGrad behind model2.py:17,21): the ReLU mask is recomputed from
 * bnr_x with the forward statistics bnr_stats/count/eps and gamma/beta, out receives g = d(a) * mask and `stats`
 * (double[2C], zeroed) += {sum g, sum g * xhat}. */
extern "C" int rsa_conv_tc3_fwd(const void* const* xs, const void* const* wts, const float* const* biases,
                                const int* dils, int nbr, void* out, const void* residual, const void* mask,
                                double* stats, int N, int H, int W, int C, int accumulate, int relu, const void* bnr_x,
                                const double* bnr_stats, double bnr_count, float bnr_eps, const float* bnr_gamma,
                                const float* bnr_beta, int bnr_relu, void* stream) {
  RSA_REQUIRE(xs && wts && dils && out && nbr >= 1 && nbr <= T3_MAXBR, RSA_ERR_SHAPE, "conv_tc3_fwd: bad arguments");
  RSA_REQUIRE(rsa_conv_tc3_supported(N, H, W, C), RSA_ERR_SHAPE, "conv_tc3_fwd: unsupported shape N=%d H=%d W=%d C=%d", N, H, W, C);
  RSA_REQUIRE(C == 32 || nbr == 1, RSA_ERR_SHAPE, "conv_tc3_fwd: fused branches need C = 32 (resident weights)");
  EncodeTiledFn enc = get_encode();
  RSA_REQUIRE(enc, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled not available from the driver");
  const int PITCH = C * 2, BOXB = 128 * PITCH;
  Tc3Params p;
  p.N = N; p.H = H; p.W = W; p.nbr = nbr;
  int max_ad = 0, any_box = 0;
  for (int b = 0; b < T3_MAXBR; ++b) {
    if (b < nbr) {
      RSA_REQUIRE(xs[b] && wts[b] && dils[b] != 0, RSA_ERR_SHAPE, "conv_tc3_fwd: branch %d: null pointer or zero dilation", b);
      const int ad = dils[b] < 0 ? -dils[b] : dils[b];
      p.dil[b] = dils[b];
      p.halo[b] = ad <= 3;
      if (p.halo[b]) max_ad = ad > max_ad ? ad : max_ad; else any_box = 1;
      p.bias[b] = biases ? biases[b] : nullptr;
    } else { p.dil[b] = 1; p.halo[b] = 1; p.bias[b] = nullptr; }
  }
  // item = 16 x (8*KT) pixels; wide items amortise the halo, narrow ones keep the ring deep when shared memory is
  // short (several resident branches, or 64 channels with a dilation-3 halo)
  static const int kt_env = getenv("RSA_TC3_KT") ? atoi(getenv("RSA_TC3_KT")) : 0;
  int KT = C == 32 ? (nbr > 2 ? 2 : 4) : (max_ad == 3 ? 1 : 2);
  if (C == 32 && (kt_env == 2 || kt_env == 4)) KT = kt_env;
  if (C == 64 && (kt_env == 1 || kt_env == 2)) KT = kt_env;
  // band mode (Tc3Params::band): only large dilations in the launch, rows of 128 pixels, band boxes within TMA's 256 limit
  static const int band_env = getenv("RSA_TC3_BAND") ? atoi(getenv("RSA_TC3_BAND")) : 1;
  int big_ad = 0;
  for (int b = 0; b < nbr; ++b) { const int ad = dils[b] < 0 ? -dils[b] : dils[b]; big_ad = ad > big_ad ? ad : big_ad; }
  // RSA_TC3_BAND: 0 off, 1 large dilations only, 2 every dilation >= RSA_TC3_BAND_MIN (experiment: bands against halo tiles)
  static const int band_min = getenv("RSA_TC3_BAND_MIN") ? atoi(getenv("RSA_TC3_BAND_MIN")) : 1;
  int small_ad = 1 << 30;
  for (int b = 0; b < nbr; ++b) { const int ad = dils[b] < 0 ? -dils[b] : dils[b]; small_ad = ad < small_ad ? ad : small_ad; }
  const bool band_ok = W % 128 == 0 && 128 + 2 * big_ad <= 256 && H % KT == 0;
  p.band = (band_ok && ((band_env == 1 && any_box && !max_ad) || (band_env == 2 && small_ad >= band_min))) ? 1 : 0;
  p.IH = p.band ? KT : 16;
  p.IWP = p.band ? 128 : 8 * KT;
  // at most one addend: the identity input of the first branch (residual) or the running sum (accumulate)
  RSA_REQUIRE(!(residual && accumulate), RSA_ERR_SHAPE, "conv_tc3_fwd: residual and accumulate are exclusive");
  p.has_add = (residual || accumulate) ? 1 : 0;
  p.has_mask = mask ? 1 : 0;
  p.has_bnx = bnr_x ? 1 : 0;
  RSA_REQUIRE(!bnr_x || (stats && bnr_stats && bnr_gamma && bnr_beta && !mask && bnr_count > 0), RSA_ERR_SHAPE,
              "conv_tc3_fwd: the fused BatchNorm backward needs stats (reduction output), forward statistics, gamma, beta and no mask");
  p.bnr_stats = bnr_stats; p.bnr_count = bnr_count; p.bnr_eps = bnr_eps; p.bnr_gamma = bnr_gamma; p.bnr_beta = bnr_beta;
  p.bnr_relu = bnr_relu;
  const int nside = p.has_add + p.has_mask + p.has_bnx;
  // side slices in flight per epilogue warp: HBM latency x the side stream's share of the bandwidth wants ~32 KB per SM
  // (two slices ahead on eight warps); 64 channels carry 72 KB of weights and keep >= 2 operand stages instead
  static const int sd_env = getenv("RSA_TC3_SDEPTH") ? atoi(getenv("RSA_TC3_SDEPTH")) : 0;
  p.sdepth = C == 64 ? (nside == 2 ? 1 : 2) : (nside == 1 ? 3 : 2);
  if (sd_env >= 1 && sd_env <= 4 && nside) p.sdepth = sd_env;
  // shared memory: leave ~8 KB of the SM to the bandwidth-bound kernels that run beside the persistent CTA (see
  // conv_tc3_kernel) unless that costs an operand stage
  // result slices: staged in shared memory + TMA store, or four 16-byte global stores per thread.  The shared-memory port is
  // the contended resource of the 32-channel halo and band launches (MMA operands, TMA fills, staging), where the direct
  // stores measure 4-7 % faster; box mode and 64 channels are 5 % slower with them - unless the 32 KB of staging are what
  // keeps a 64-channel launch with side inputs out of band mode (below).  RSA_TC3_DIRECT=0 / 1 forces one way.
  static const int direct_env = getenv("RSA_TC3_DIRECT") ? atoi(getenv("RSA_TC3_DIRECT")) : -1;
  p.direct = direct_env >= 0 ? direct_env : (C == 32 && (!any_box || p.band));
  auto plan = [&](int kt, int budget_kb) {
    int slot = any_box ? kt * BOXB : 0;
    if (p.band) slot = kt * (128 + 2 * big_ad) * PITCH;
    else if (max_ad) { const int hb = (16 + 2 * max_ad) * (8 * kt + 2 * max_ad) * PITCH; slot = hb > slot ? hb : slot; }
    slot = (slot + 1023) & ~1023;
    const Tc3Smem L0(C, nbr, nside, p.sdepth, 0, slot, p.direct);
    int ns = (budget_kb * 1024 - L0.total - 256) / (slot + 16);
    p.slot_bytes = slot;
    p.nstages = ns > 8 ? 8 : ns;
  };
  plan(KT, 227);
  if (p.nstages < 2 && C == 32 && KT == 4) { KT = 2; if (p.band) p.IH = KT; plan(KT, 227); }
  if (p.nstages < 2 && p.band && !p.direct && direct_env < 0) { p.direct = 1; plan(KT, 227); }   // bands instead of staging
  if (p.nstages < 2 && p.band) {                   // bands too large beside the side slices: boxes
    p.band = 0; p.IH = 16;
    if (direct_env < 0) p.direct = 0;
    plan(KT, 227);
  }
  { const int full = p.nstages; plan(KT, 219); if (p.nstages < full && p.nstages < 4) plan(KT, 227); }
  RSA_REQUIRE(p.nstages >= 2, RSA_ERR_SHAPE, "conv_tc3_fwd: shared memory budget allows only %d stage(s)", p.nstages);
  static const int alt_env = getenv("RSA_TC3_ALT") ? atoi(getenv("RSA_TC3_ALT")) : 1;
  p.alt = (!any_box && !p.band && alt_env && nbr == 1 && C == 32) ? 1 : 0;   // measured: C = 64 (two chains per item) is faster with both warps on one item
  if (p.alt) p.nstages &= ~1;                    // see conv_tc3_kernel: an even ring keeps every stage with one MMA warp
  if (!p.band) p.IWP = 8 * KT;
  p.tiles_w = W / p.IWP; p.tiles_h = H / p.IH;
  p.items = p.tiles_w * p.tiles_h * N;
  const Tc3Smem L(C, nbr, nside, p.sdepth, p.nstages, p.slot_bytes, p.direct);
  RSA_REQUIRE(L.total <= 227 * 1024, RSA_ERR_SHAPE, "conv_tc3_fwd: shared memory %d", L.total);
  p.stats = stats; p.relu = relu;
  p.out = (uint8_t*)out;
  p.debug = getenv("RSA_TC3_DEBUG") ? atoi(getenv("RSA_TC3_DEBUG")) : 0;
  const CUtensorMapSwizzle swz = C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  Tc3Maps maps;
  for (int b = 0; b < nbr; ++b) {
    const int ad = p.dil[b] < 0 ? -p.dil[b] : p.dil[b];
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)(p.halo[b] ? 8 * KT + 2 * ad : 8 * KT), (cuuint32_t)(p.halo[b] ? 16 + 2 * ad : 16), 1};
    if (p.band) { box[1] = (cuuint32_t)(128 + 2 * ad); box[2] = (cuuint32_t)KT; }
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&maps.a[b], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(xs[b]), gdim, gstr, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(x%d) failed (%d)", b, (int)r);
    cuuint64_t wdim[3] = {(cuuint64_t)C, (cuuint64_t)C, 9};
    cuuint64_t wstr[2] = {(cuuint64_t)C * 2, (cuuint64_t)C * C * 2};
    cuuint32_t wbox[3] = {(cuuint32_t)C, (cuuint32_t)C, 1};
    cuuint32_t wes[3] = {1, 1, 1};
    r = enc(&maps.w[b], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wts[b]), wdim, wstr, wbox, wes,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(w%d) failed (%d)", b, (int)r);
  }
  for (int b = nbr; b < T3_MAXBR; ++b) { maps.a[b] = maps.a[0]; maps.w[b] = maps.w[0]; }
  {
    // epilogue slices: 8 x 4 pixels x 32 channels, SWIZZLE_64B for both channel counts (a 64-channel tensor is stored /
    // loaded as two 32-channel halves by different warps); the TMA-stored output and the side inputs share the box shape
    auto enc_slice = [&](CUtensorMap* tm, const void* base) -> CUresult {
      cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[4] = {32, 8, 4, 1};
      if (p.band) { box[1] = 32; box[2] = 1; }      // band items: a slice is 32 consecutive pixels of one image row
      cuuint32_t es[4] = {1, 1, 1, 1};
      return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r = enc_slice(&maps.out, out);
    RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(out) failed (%d)", (int)r);
    maps.add = maps.out; maps.mask = maps.out; maps.bnx = maps.out;
    if (p.has_bnx) {
      r = enc_slice(&maps.bnx, bnr_x);
      RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(bnr_x) failed (%d)", (int)r);
    }
    if (p.has_add) {
      r = enc_slice(&maps.add, residual ? residual : out);
      RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(addend) failed (%d)", (int)r);
    }
    if (p.has_mask) {
      r = enc_slice(&maps.mask, mask);
      RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(mask) failed (%d)", (int)r);
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 32) return KT == 2 ? launch3<32, 2>(maps, p, L.total, st) : launch3<32, 4>(maps, p, L.total, st);
  return KT == 1 ? launch3<64, 1>(maps, p, L.total, st) : launch3<64, 2>(maps, p, L.total, st);
}

// =====================================================================================================
// Weight gradient of the thin-layer 3x3 convolution.
//   dW[tap][ci][co] += sum_pix x[pix + off(tap), ci] * dy[pix, co]          (Conv2D backward-filter, model2.py:19-24)
// GEMM view: K = pixels, both operands MN-major (channels contiguous).  conv_tc.cu's kernel loads nine shifted copies of
// every 64-pixel tile (655 MB of L2->SM traffic per launch at C = 32, one issuing thread); here an item is a 16 x 16
// pixel tile whose x halo is loaded ONCE and the taps of a tap row are the M atoms of one MMA: atom i starts i*d pixels
// further (LBO = d pixels).  C = 32: D_dy[128 x 32] = [tap(dy,-1) | tap(dy,0) | tap(dy,+1) | unused]; C = 64 (atoms of
// 64 channels): D_dy,0 = [tap(dy,-1) | tap(dy,0)], D_dy,1 = [tap(dy,+1) | unused].  Three warps (one per tap row) issue
// independent accumulation chains that run for the whole life of the persistent CTA; one red.global.add pass per CTA
// at the end.  Large dilations (C = 32 only): the two-dimensional halo does not fit, so an item is a flat 2 x 128 pixel strip
// and every tap row gets ONE band box 2 x (128 + 2d) pixels whose three column shifts are again the M atoms (LBO = d pixels):
// 300-350 bytes of L2 traffic per pixel.  Images narrower than 128 pixels keep the first scheme, nine 16 x 8 boxes per item
// with LBO = one box (640 bytes per pixel, 81-85 us per launch at 16 x 256 x 256 where the halo launches take 49-53).
// =====================================================================================================
namespace {

constexpr int W3_THREADS = 256;   // warp 0 TMA, 1..3 MMA (tap row dy = warp - 1), 4..7 final reduction

struct Wg3Params {
  int N, H, W, dil;
  int halo;               // 1: one two-dimensional halo box per item, 2: one band box per tap row, 0: one box per tap
  int IW, IH;             // item width / height in pixels: 16 x 16 (halo), 128 x 2 (bands), 8 x 16 (boxes)
  int items, tiles_w, tiles_h;
  int nstages, stage_bytes, a_bytes, a_tx;   // a_tx: bytes the halo box actually delivers
  int band_bytes;         // bands: shared-memory pitch between the tap rows' boxes
  float* dw;
};

template <int C>
__device__ __forceinline__ uint64_t w3_mndesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)T3<C>::LAYOUT << 61;
  return d;
}
__device__ __forceinline__ void w3_red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int C>
__global__ void __launch_bounds__(W3_THREADS, 1) conv_tc3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                       const __grid_constant__ CUtensorMap tmDY,
                                                                       const Wg3Params p) {
  constexpr int PITCH = T3<C>::PITCH;
  constexpr int NACC = C == 32 ? 1 : 2;              // accumulators (MMA chains) per tap row
  constexpr uint32_t TMEM_COLS = C == 32 ? 128 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.nstages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + p.nstages;
  uint64_t* done_bar = empty_bar + p.nstages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = p.dil;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX); prefetch_tmap(&tmDY);
    for (int s = 0; s < p.nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 3); }
    mbar_init(done_bar, 3);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  pdl_wait();
  if (threadIdx.x == 0) pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nmine = ((int)blockIdx.x < p.items) ? (p.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0, phase = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int r = item;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int n = r, h0 = th * p.IH, w0 = tw * p.IW;
        uint8_t* sa = smem + stage * p.stage_bytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (p.halo == 1) {
          mbar_expect_tx(&full_bar[stage], p.a_tx + 16 * p.IW * PITCH);
          tma_load_4d(sa, &tmX, &full_bar[stage], 0, w0 - d, h0 - d, n);
        } else if (p.halo == 2) {
          int nrow = 0;
          for (int dyi = 0; dyi < 3; ++dyi) { const int ch = h0 + (dyi - 1) * d; nrow += !(ch + p.IH <= 0 || ch >= p.H); }
          mbar_expect_tx(&full_bar[stage], nrow * p.a_tx + p.IH * p.IW * PITCH);
          for (int dyi = 0; dyi < 3; ++dyi) {
            const int ch = h0 + (dyi - 1) * d;
            if (ch + p.IH <= 0 || ch >= p.H) continue;
            tma_load_4d(sa + dyi * p.band_bytes, &tmX, &full_bar[stage], 0, w0 - d, ch, n);
          }
        } else {
          int nrow = 0;
          for (int dyi = 0; dyi < 3; ++dyi) { const int ch = h0 + (dyi - 1) * d; nrow += !(ch + 16 <= 0 || ch >= p.H); }
          mbar_expect_tx(&full_bar[stage], (3 * nrow + 1) * 16 * p.IW * PITCH);
          for (int dyi = 0; dyi < 3; ++dyi) {
            const int ch = h0 + (dyi - 1) * d;
            if (ch + 16 <= 0 || ch >= p.H) continue;
            for (int dxi = 0; dxi < 3; ++dxi)
              tma_load_4d(sa + (dyi * 3 + dxi) * 16 * p.IW * PITCH, &tmX, &full_bar[stage], 0, w0 + (dxi - 1) * d, ch, n);
          }
        }
        tma_load_4d(sa + p.a_bytes, &tmDY, &full_bar[stage], 0, w0, h0, n);
        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp <= 3) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, C) | (1u << 15) | (1u << 16);
      const int dyi = warp - 1;
      const uint32_t acc = tmem_base + (uint32_t)(dyi * NACC * C);
      const int Wh = p.IW + 2 * d;                        // halo mode: region row pitch in pixels
      const int ncb = p.IW / 8;                           // 8-pixel column blocks per item row
      const uint32_t boxb = 16 * p.IW * PITCH;
      int stage = 0, phase = 0;
      uint32_t accum = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int r = item;
        r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int ch = th * p.IH + (dyi - 1) * d;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
        const uint32_t sb = sa + p.a_bytes;
        if (p.halo == 1 || !(ch + p.IH <= 0 || ch >= p.H)) {
          // A: atom i = tap (dy, dx = i - 1) (C = 32: four atoms, C = 64: two per chain); two 8-pixel K groups per MMA
          // = rows (2rp, 2rp+1), column block cb
          const uint32_t a0 = p.halo == 1 ? sa + (uint32_t)((dyi * d * Wh) * PITCH)
                            : (p.halo == 2 ? sa + (uint32_t)(dyi * p.band_bytes) : sa + (uint32_t)(dyi * 3) * boxb);
          const uint32_t arow = p.halo ? Wh * PITCH : p.IW * PITCH;
          const uint32_t lbo = p.halo ? d * PITCH : boxb;
          const uint32_t brow = p.IW * PITCH;
          // the two 8-pixel K groups of an MMA: rows (2rp, 2rp+1) of one column block, or - items of a single row -
          // two neighbouring column blocks
          const int nrp = p.IH >= 2 ? p.IH / 2 : 1;
          const uint32_t cstep = p.IH >= 2 ? 8 * PITCH : 16 * PITCH;
          const uint32_t ksa = p.IH >= 2 ? arow : 8 * PITCH, ksb = p.IH >= 2 ? brow : 8 * PITCH;
          const int ncbk = p.IH >= 2 ? ncb : ncb / 2;
#pragma unroll 1
          for (int rp = 0; rp < nrp; ++rp) {
            for (int cb = 0; cb < ncbk; ++cb) {
              const uint64_t bdesc = w3_mndesc<C>(sb + 2 * rp * brow + cb * cstep, 0, ksb);
#pragma unroll
              for (int c = 0; c < NACC; ++c) {
                const uint64_t adesc = w3_mndesc<C>(a0 + 2 * rp * arow + cb * cstep + c * 2 * lbo, lbo, ksa);
                umma_bf16(acc + c * C, adesc, bdesc, idesc, accum);
              }
              accum = 1;
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.nstages) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else {
    // ===== final reduction: the lanes of D hold (tap, ci), the columns co =====
    const int q = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (nmine > 0) {
      for (int dyi = 0; dyi < 3; ++dyi) {
        // a tap row that was out of range for every item of this CTA never initialised its accumulator
        bool any = p.halo == 1;
        if (!any) {
          for (int item = blockIdx.x; item < p.items && !any; item += gridDim.x) {
            const int ch = ((item / p.tiles_w) % p.tiles_h) * p.IH + (dyi - 1) * d;
            any = !(ch + p.IH <= 0 || ch >= p.H);
          }
        }
        if (!any) continue;
#pragma unroll
        for (int c = 0; c < NACC; ++c) {
          // C = 32: lane quarter q = dx + 1 (q = 3 unused); C = 64: chain 0 holds dx = -1 (q < 2) and 0 (q >= 2), chain 1 dx = +1
          int dxi, ci;
          if (C == 32) { dxi = q; ci = lane; }
          else { dxi = c == 0 ? (q >> 1) : 2; ci = (q & 1) * 32 + lane; }
          if ((C == 32 && q == 3) || (C == 64 && c == 1 && q >= 2)) continue;
          float* dst = p.dw + ((size_t)(dyi * 3 + dxi) * C + ci) * C;
#pragma unroll
          for (int c0 = 0; c0 < C; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((dyi * NACC + c) * C + c0), v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              w3_red_add_v4(dst + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

template <int C>
int launch_wg3(const CUtensorMap& tmX, const CUtensorMap& tmDY, const Wg3Params& p, int smem_bytes, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc3_wgrad_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { rsa_set_error("conv_tc3_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RSA_ERR_CUDA; }
    configured = true;
  }
  const int grid = p.items < rsa_num_sms() ? p.items : rsa_num_sms();
  cudaError_t le = launch_pdl(conv_tc3_wgrad_kernel<C>, dim3(grid), dim3(W3_THREADS), (size_t)smem_bytes, st, tmX, tmDY, p);
  if (le != cudaSuccess) { rsa_set_error("conv_tc3_wgrad: launch: %s", cudaGetErrorString(le)); return RSA_ERR_CUDA; }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

}  // namespace

/* C = 32: any dilation; C = 64: dilations <= 3, and larger ones where the band boxes apply (W a multiple of 128, band of
 * 128 + 2 dil <= 256 pixels) - the nine boxes of a large dilation do not fit beside the dy tile. */
extern "C" int rsa_conv_tc3_wgrad_supported(int N, int H, int W, int C, int dil) {
  return rsa_conv_tc3_supported(N, H, W, C) && dil > 0 && (C == 32 || dil <= 3 || (W % 128 == 0 && 128 + 2 * dil <= 256));
}

/* dw[tap][ci][co] (fp32 HWIO, zeroed by the caller once per step) += sum_pix x[pix+off(tap), ci] * dy[pix, co] for the
 * thin (32- / 64-channel) layers; x, dy bf16 NHWC [N,H,W,C], dil > 0.  Same contract as rsa_conv_tc_wgrad. */
extern "C" int rsa_conv_tc3_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int dil, void* stream) {
  RSA_REQUIRE(x && dy && dw && dil > 0, RSA_ERR_SHAPE, "conv_tc3_wgrad: bad arguments");
  RSA_REQUIRE(rsa_conv_tc3_wgrad_supported(N, H, W, C, dil), RSA_ERR_SHAPE,
              "conv_tc3_wgrad: unsupported shape N=%d H=%d W=%d C=%d dil=%d", N, H, W, C, dil);
  EncodeTiledFn enc = get_encode();
  RSA_REQUIRE(enc, RSA_ERR_CUDA, "conv_tc3_wgrad: cuTensorMapEncodeTiled not available from the driver");
  const int PITCH = C * 2;
  Wg3Params p;
  p.N = N; p.H = H; p.W = W; p.dil = dil; p.dw = dw;
  static const int band_env = getenv("RSA_TC3_WG_BAND") ? atoi(getenv("RSA_TC3_WG_BAND")) : 1;
  p.halo = dil <= 3 ? 1 : ((W % 128 == 0 && 128 + 2 * dil <= 256 && (band_env || C == 64)) ? 2 : 0);
  p.IW = p.halo == 1 ? 16 : (p.halo == 2 ? 128 : 8);
  p.IH = p.halo == 2 ? (C == 32 ? 2 : 1) : 16;         // 64 channels: single rows keep two stages in shared memory
  p.tiles_w = W / p.IW; p.tiles_h = H / p.IH;
  p.items = p.tiles_w * p.tiles_h * N;
  // the unused trailing M atom reads up to 2*dil pixels (halo, bands) / one box past the A region: keep that inside the stage
  p.band_bytes = 0;
  int a_raw;
  if (p.halo == 2) {
    p.a_tx = p.IH * (p.IW + 2 * dil) * PITCH;                  // one band
    p.band_bytes = (p.a_tx + 1023) & ~1023;
    a_raw = 3 * p.band_bytes;
  } else {
    a_raw = p.halo ? (16 + 2 * dil) * (p.IW + 2 * dil) * PITCH : 9 * 16 * p.IW * PITCH;
    p.a_tx = a_raw;
  }
  p.a_bytes = (a_raw + 2 * dil * PITCH + 1023) & ~1023;
  p.stage_bytes = p.a_bytes + p.IH * p.IW * PITCH;
  p.stage_bytes = (p.stage_bytes + 1023) & ~1023;
  int ns = (227 * 1024 - 2048) / p.stage_bytes;
  if (ns > 6) ns = 6;
  RSA_REQUIRE(ns >= 2, RSA_ERR_SHAPE, "conv_tc3_wgrad: stage of %d bytes leaves %d stage(s)", p.stage_bytes, ns);
  p.nstages = ns;
  const int smem_bytes = ns * p.stage_bytes + (2 * ns + 1) * 8 + 16 + 1024;
  const CUtensorMapSwizzle swz = C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUtensorMap tmX, tmDY;
  auto encode = [&](CUtensorMap* tm, const void* base, int bw, int bh) -> CUresult {
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUresult r = p.halo == 1 ? encode(&tmX, x, p.IW + 2 * dil, 16 + 2 * dil)
             : (p.halo == 2 ? encode(&tmX, x, p.IW + 2 * dil, p.IH) : encode(&tmX, x, p.IW, 16));
  RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_wgrad: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
  r = encode(&tmDY, dy, p.IW, p.IH);
  RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_wgrad: cuTensorMapEncodeTiled(dy) failed (%d)", (int)r);
  if (C == 32) return launch_wg3<32>(tmX, tmDY, p, smem_bytes, (cudaStream_t)stream);
  return launch_wg3<64>(tmX, tmDY, p, smem_bytes, (cudaStream_t)stream);
}
