// conv_tc3.cu — thin-layer (C = 32) dilated 3x3 'same' convolution on tcgen05, built around what the ncu captures
// of conv_tc2 showed for these layers (profiles/r1b_*): the tensor pipe sat at 7 % because every 128-pixel tile paid
// nine TMA round trips and a ~5.7k-cycle register epilogue on 4 warps; a first version of this kernel then showed
// the single MMA-issuing thread (~100 cycles per 16-cycle N=32 MMA) and shared-memory operand bandwidth as the next
// limits.  Hence:
//
//   * weights of all branches stay resident in shared memory for the life of the (persistent) CTA;
//   * an item is 16 x (8*KT) pixels = KT accumulators; small dilations (|d| <= 3) load ONE halo box per item and
//     form the nine taps as shifted UMMA descriptors into it (measured in scripts/exp_desc.cu: with base_offset = 0
//     any row shift and any stride-byte-offset address a TMA-swizzled tile correctly); large dilations load one
//     16 x (8*KT) box per tap.  Either way one stage feeds all KT sub-tiles;
//   * KT warps issue the MMAs, one per sub-tile (independent accumulators), so the issue rate scales;
//   * up to four branches (ResBlock-a: dilations 1/3/15/31, model2.py:23-31) accumulate into the same TMEM tiles,
//     so the branch sum and the identity add happen once, in the epilogue;
//   * the epilogue is spread over 8 warps, prefetches its bf16 side inputs one sub-tile ahead, stages the bf16
//     tile in shared memory (swizzled, conflict free) and leaves the global write to a TMA store issued by a
//     dedicated warp; BatchNorm statistics of the stored values are reduced with a 16-wide shuffle butterfly.
//
// What bounds it now (scripts/exp_mma_rate.cu, scripts/trace_tc3.py): an SS-mode tcgen05.mma reads its operands from
// shared memory at 128 B/clk, i.e. 32 cycles for the 128x16 A slice + N/4 for B: 40 cycles per N=32 MMA against a 16-cycle
// tensor floor, whatever the swizzle.  A C=32 sub-tile (18 MMAs, 90 KB of operand reads) therefore cannot take less than
// ~700 cycles = 19.7 us per launch at batch 16; an experiment with the stores / TMEM loads removed runs at 35 us, the
// full kernel at 42-46 us.  (Direct global stores from the epilogue instead of the staged TMA store were measured slower.)
//
// Replaces cuDNN's Conv2D forward / backward-data behind keras Conv2D(32, 3, dilation_rate=d, padding='same') at
// model2.py:19-24,153-178 for the C = 32 layers (enc1, dec1, heads).
#include "tc_common.cuh"

namespace {

constexpr int T3_MAXBR = 4;
constexpr int T3_MAXSB = 4;

// compile-time geometry of one channel class (C = 32: SWIZZLE_64B rows, C = 64: SWIZZLE_128B rows)
template <int C>
struct T3 {
  static constexpr int PITCH = C * 2;               // bytes per pixel = swizzle span
  static constexpr int BOXB = 128 * PITCH;          // one 16x8-pixel sub-tile
  static constexpr int WBYTES = 9 * C * PITCH;      // one branch's weights
  static constexpr uint32_t LAYOUT = C == 32 ? 4u : 2u;   // UMMA layout code: SWIZZLE_64B / SWIZZLE_128B
};
// warp roles: 0 TMA producer, 1..KT MMA issuers (one per sub-tile), KT+1 TMA store + side loads, EPI0..EPI0+EW-1 epilogue.
// The epilogue is a chain of dependent latencies per warp (tcgen05.ld, shared loads, proxy fence), so it is spread over
// EW = 8 or 16 warps: lane quarter q = warp % 4 (the TMEM lanes a warp may read), column group (warp - EPI0) / 4.
template <int KT, int EW> struct T3Warps {
  static constexpr int STORE = KT + 1;
  static constexpr int EPI0 = (KT + 2 + 3) / 4 * 4;
  static constexpr int THREADS = (EPI0 + EW) * 32;
};

struct Tc3Params {
  int N, H, W;
  int nbr;
  int dil[T3_MAXBR];      // signed: negative = data gradient (taps mirrored)
  int halo[T3_MAXBR];     // 1: one halo box per item, 0: one box per tap
  int items, tiles_w, tiles_h;
  int nstages, slot_bytes, nsb;
  int nsb_log, nsi_log;         // nsb, nsi are powers of two
  int nsi, has_add, has_mask;   // side-input ring: slots, addend present (residual or previous out), ReLU mask present
  int has_bnx, bnr_relu;        // fused BatchNorm backward: BN input tile in the side ring; the BN was followed by ReLU
  const double* bnr_stats;      // {sum, sumsq} of the BN input (forward statistics)
  double bnr_count;
  float bnr_eps;
  const float* bnr_gamma;
  const float* bnr_beta;
  const float* bias[T3_MAXBR];
  double* stats;
  int relu;
  long long* trace;       // diagnostic: per-CTA cycles spent waiting on each barrier family (rsa_conv_tc3_set_trace)
};

// timed mbarrier wait for the diagnostic trace (plain wait when tracing is off)
__device__ __forceinline__ void twait(uint64_t* bar, uint32_t parity, long long* tr, int slot) {
  if (!tr) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  tr[slot] += clock64() - t0;
}

template <int N> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[N]);
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
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

struct Tc3Maps { CUtensorMap a[T3_MAXBR]; CUtensorMap w[T3_MAXBR]; CUtensorMap out; CUtensorMap add; CUtensorMap mask; CUtensorMap bnx; };

// shared-memory carve-up (offsets from the 1024-aligned base)
struct Tc3Smem {
  int w_off, st_off, side_off, ring_off, misc_off, bar_off, total;
  __host__ __device__ Tc3Smem(int C, int nbr, int nsb, int nside, int nsi, int nstages, int slot_bytes) {
    w_off = 0;
    st_off = (nbr * 9 * C * C * 2 + 1023) & ~1023;
    side_off = st_off + nsb * 128 * C * 2;                   // [nsi][nside] tiles of 128 pixels
    ring_off = side_off + nsi * nside * 128 * C * 2;
    misc_off = ring_off + nstages * slot_bytes;              // bias[C], csum[4][C], csq[4][C], BN coefficients [4][C]
    bar_off = misc_off + 13 * C * 4;
    total = bar_off + (2 * nstages + 4 * T3_MAXSB + 8) * 8 + 16 + 1024;
  }
};

template <int C, int KT, int EW>
__global__ void __launch_bounds__(T3Warps<KT, EW>::THREADS, 1) conv_tc3_kernel(const __grid_constant__ Tc3Maps maps, const Tc3Params p) {
  constexpr int PITCH = T3<C>::PITCH, BOXB = T3<C>::BOXB, WBYTES = T3<C>::WBYTES;
  constexpr int EPI0 = T3Warps<KT, EW>::EPI0;
  constexpr int NCT = C / (EW / 4);     // accumulator columns per epilogue thread
  constexpr int CHK = NCT < 16 ? NCT : 16;   // columns per tcgen05.ld
  static_assert(CHK == 8 || CHK == 16, "epilogue chunk");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nside = p.has_add + p.has_mask + p.has_bnx;
  const Tc3Smem L(C, p.nbr, p.nsb, nside, p.nsi, p.nstages, p.slot_bytes);
  uint8_t* wsm = smem + L.w_off;
  uint8_t* ysm = smem + L.st_off;
  uint8_t* sidesm = smem + L.side_off;
  uint8_t* ring = smem + L.ring_off;
  float* bias_s = reinterpret_cast<float*>(smem + L.misc_off);
  float* csum = bias_s + C;                  // [4][C]: one slot per lane quarter, summed in a fixed order (deterministic)
  float* csq = csum + 4 * C;
  float* bnc = csq + 4 * C;                  // [4][C]: invstd, -mean*invstd, gamma, beta of the fused BatchNorm backward
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L.bar_off);
  uint64_t* empty_bar = full_bar + p.nstages;
  uint64_t* tfull = empty_bar + p.nstages;   // [2]
  uint64_t* tempty = tfull + 2;              // [2]
  uint64_t* sready = tempty + 2;             // [nsb] staged tile written by the 8 epilogue warps
  uint64_t* sfree = sready + T3_MAXSB;       // [nsb] staged tile read by its TMA store
  uint64_t* ifull = sfree + T3_MAXSB;        // [nsi] side-input tiles landed
  uint64_t* wbar = ifull + T3_MAXSB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_stats = p.stats != nullptr;
  long long* tr = p.trace ? p.trace + (size_t)blockIdx.x * 16 : nullptr;
  constexpr uint32_t TMEM_COLS = 2 * KT * C <= 64 ? 64 : (2 * KT * C <= 128 ? 128 : 256);
  static_assert(2 * KT * C <= 256, "accumulator ring exceeds the TMEM allocation");

  if (threadIdx.x == 0) {
    for (int b = 0; b < p.nbr; ++b) { prefetch_tmap(&maps.a[b]); prefetch_tmap(&maps.w[b]); }
    prefetch_tmap(&maps.out);
    if (p.has_add) prefetch_tmap(&maps.add);
    if (p.has_mask) prefetch_tmap(&maps.mask);
    if (p.has_bnx) prefetch_tmap(&maps.bnx);
    for (int s = 0; s < p.nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], KT); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], KT); mbar_init(&tempty[s], EW); }
    for (int s = 0; s < T3_MAXSB; ++s) {
      mbar_init(&sready[s], EW); mbar_init(&sfree[s], 1);
      mbar_init(&ifull[s], 1);
    }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
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
      bnc[c] = inv; bnc[C + c] = -mean * inv; bnc[2 * C + c] = p.bnr_gamma[c]; bnc[3 * C + c] = p.bnr_beta[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
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
        const int n = r, h0 = th * 16, w0 = tw * 8 * KT;
        for (int b = 0; b < p.nbr; ++b) {
          const int d = p.dil[b], ad = d < 0 ? -d : d;
          if (p.halo[b]) {
            twait(&empty_bar[stage], phase ^ 1, tr, 0);
            mbar_expect_tx(&full_bar[stage], (16 + 2 * ad) * (8 * KT + 2 * ad) * PITCH);
            tma_load_4d(ring + stage * p.slot_bytes, &maps.a[b], &full_bar[stage], 0, w0 - ad, h0 - ad, n);
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          } else {
            for (int tap = 0; tap < 9; ++tap) {
              const int ch = h0 + (tap / 3 - 1) * d, cw = w0 + (tap % 3 - 1) * d;
              if (ch + 16 <= 0 || ch >= p.H || cw + 8 * KT <= 0 || cw >= p.W) continue;
              twait(&empty_bar[stage], phase ^ 1, tr, 0);
              mbar_expect_tx(&full_bar[stage], KT * BOXB);
              tma_load_4d(ring + stage * p.slot_bytes, &maps.a[b], &full_bar[stage], 0, cw, ch, n);
              if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp <= KT) {
    // ===== MMA issuers: warp 1+s owns sub-tile s of every item =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, C);
      const int s = warp - 1;
      mbar_wait(wbar, 0);
      tc_fence_after();
      const uint32_t wbase = smem_u32(wsm);
      const uint32_t bhi = t3_desc_hi<C>(8 * PITCH);
      int stage = 0, phase = 0, it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        int r = item;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h;
        const int h0 = th * 16, w0 = tw * 8 * KT;
        twait(&tempty[it & 1], ((it >> 1) & 1) ^ 1, s == 0 ? tr : nullptr, 2);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)(((it & 1) * KT + s) * C);
        uint32_t started = 0;
        for (int b = 0; b < p.nbr; ++b) {
          const int d = p.dil[b], ad = d < 0 ? -d : d;
          const uint32_t wb = wbase + b * WBYTES;
          if (p.halo[b]) {
            const int Wh = 8 * KT + 2 * ad;
            const uint32_t ahi = t3_desc_hi<C>(Wh * PITCH);
            twait(&full_bar[stage], phase, s == 0 ? tr : nullptr, 3);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + stage * p.slot_bytes) + (uint32_t)((ad * Wh + ad + 8 * s) * PITCH);
            const int rowb = d * Wh * PITCH, colb = d * PITCH;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint64_t adesc = t3_desc(ahi, sa + (tap / 3 - 1) * rowb + (tap % 3 - 1) * colb);
              const uint64_t bdesc = t3_desc(bhi, wb + tap * C * PITCH);
#pragma unroll
              for (int k = 0; k < C / 16; ++k)
                umma_bf16(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, started | (tap > 0) | (k > 0));
            }
            started = 1;
            umma_commit(&empty_bar[stage]);
            if (++stage == p.nstages) { stage = 0; phase ^= 1; }
          } else {
            const uint32_t ahi = t3_desc_hi<C>(8 * KT * PITCH);
            for (int tap = 0; tap < 9; ++tap) {
              const int ch = h0 + (tap / 3 - 1) * d, cw = w0 + (tap % 3 - 1) * d;
              if (ch + 16 <= 0 || ch >= p.H || cw + 8 * KT <= 0 || cw >= p.W) continue;
              twait(&full_bar[stage], phase, s == 0 ? tr : nullptr, 3);
              tc_fence_after();
              const uint64_t adesc = t3_desc(ahi, smem_u32(ring + stage * p.slot_bytes) + 8 * s * PITCH);
              const uint64_t bdesc = t3_desc(bhi, wb + tap * C * PITCH);
#pragma unroll
              for (int k = 0; k < C / 16; ++k)
                umma_bf16(acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, started | (k > 0));
              started = 1;
              umma_commit(&empty_bar[stage]);
              if (++stage == p.nstages) { stage = 0; phase ^= 1; }
            }
          }
        }
        umma_commit(&tfull[it & 1]);
      }
    }
  } else if (warp == T3Warps<KT, EW>::STORE) {
    // ===== TMA store of the staged tiles; the same thread keeps the side-input ring (addend, ReLU mask, BatchNorm
    // input: 16x8 boxes like the output tiles) nsi sub-tiles ahead of the epilogue.  sready[j] completing for sub-tile
    // `seq` also means the eight epilogue warps are done with that sub-tile's side slot, so the slot is refilled right
    // there - the main operand pipeline (warp 0) never waits for the epilogue =====
    if (lane == 0) {
      auto issue_side = [&](int sq) {
        const int item = blockIdx.x + (sq / KT) * (int)gridDim.x;
        if (item >= p.items) return;
        int r = item;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int n = r, h0 = th * 16, w0 = tw * 8 * KT + 8 * (sq % KT);
        const int k = sq & (p.nsi - 1);
        mbar_expect_tx(&ifull[k], nside * BOXB);
        uint8_t* dst = sidesm + k * nside * BOXB;
        if (p.has_add) tma_load_4d(dst, &maps.add, &ifull[k], 0, w0, h0, n);
        if (p.has_mask) tma_load_4d(dst + p.has_add * BOXB, &maps.mask, &ifull[k], 0, w0, h0, n);
        if (p.has_bnx) tma_load_4d(dst + (p.has_add + p.has_mask) * BOXB, &maps.bnx, &ifull[k], 0, w0, h0, n);
      };
      if (nside)
        for (int sq = 0; sq < p.nsi; ++sq) issue_side(sq);
      int seq = 0;
      const int lag = p.nsb >> 1;          // stores allowed in flight before a buffer is handed back
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        int r = item;
        const int tw = r % p.tiles_w; r /= p.tiles_w;
        const int th = r % p.tiles_h; r /= p.tiles_h;
        const int n = r, h0 = th * 16, w0 = tw * 8 * KT;
        for (int s = 0; s < KT; ++s, ++seq) {
          const int j = seq & (p.nsb - 1);
          twait(&sready[j], (seq >> p.nsb_log) & 1, tr, 5);
          if (nside) issue_side(seq + p.nsi);
          tma_store_4d(&maps.out, ysm + j * BOXB, 0, w0 + 8 * s, h0, n);
          // the store issued `lag` tiles ago has finished reading its buffer: hand that buffer back
          if (lag == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          if (seq >= lag) mbar_arrive(&sfree[(seq - lag) & (p.nsb - 1)]);
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp >= EPI0) {
    // ===== epilogue warps: lane quarter q, column group hs =====
    const int q = warp & 3, hs = (warp - EPI0) >> 2;
    const int rrow = q * 32 + lane;                 // accumulator row = pixel of the 16x8 sub-tile
    const uint32_t srow = (uint32_t)rrow * PITCH;
    const uint32_t sw = C == 32 ? (uint32_t)((rrow >> 1) & 3) : (uint32_t)(rrow & 7);   // swizzle phase of this row
    float acc_s[NCT], acc_q[NCT];                   // BatchNorm statistics of this thread's pixels (all its sub-tiles)
#pragma unroll
    for (int j = 0; j < NCT; ++j) { acc_s[j] = 0.f; acc_q[j] = 0.f; }
    float bias_r[C == 32 ? NCT : 1];                // 32 channels: the thread's biases live in registers
    if constexpr (C == 32) {
#pragma unroll
      for (int j = 0; j < NCT; ++j) bias_r[j] = bias_s[hs * NCT + j];
    }
    int it = 0, seq = 0;
    long long* etr = (warp == EPI0) ? tr : nullptr;
    const long long e_t0 = clock64();
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      if (lane == 0) twait(&tfull[it & 1], (it >> 1) & 1, etr, 6); else mbar_wait(&tfull[it & 1], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int s = 0; s < KT; ++s, ++seq) {
        const int j = seq & (p.nsb - 1);
        uint8_t* yb = ysm + j * BOXB + srow;
        const int ks = nside ? (seq & (p.nsi - 1)) : 0;
        const uint8_t* ib = sidesm + ks * nside * BOXB + srow;
        if (nside) {
          if (lane == 0) twait(&ifull[ks], (seq >> p.nsi_log) & 1, etr, 8);
          __syncwarp();
        }
#pragma unroll
        for (int cc = 0; cc < NCT; cc += CHK) {
          uint32_t v[CHK];
          const long long tl0 = etr ? clock64() : 0;
          tmem_ld<CHK>(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((it & 1) * KT + s) * C + hs * NCT + cc), v);
          if (etr && lane == 0) etr[11] += clock64() - tl0;
          if (s == KT - 1 && cc + CHK >= NCT) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[it & 1]);
          }
          float f[CHK];
          if constexpr (C == 32) {
#pragma unroll
            for (int i = 0; i < CHK; ++i) f[i] = __uint_as_float(v[i]) + bias_r[cc + i];
          } else {
#pragma unroll
            for (int i = 0; i < CHK; i += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(bias_s + hs * NCT + cc + i);
              f[i] = __uint_as_float(v[i]) + bv.x; f[i + 1] = __uint_as_float(v[i + 1]) + bv.y;
              f[i + 2] = __uint_as_float(v[i + 2]) + bv.z; f[i + 3] = __uint_as_float(v[i + 3]) + bv.w;
            }
          }
          const uint32_t c16 = (uint32_t)((hs * NCT + cc) >> 3);        // first 16-byte chunk of this group inside the row
          uint32_t o[CHK / 8];
#pragma unroll
          for (int k = 0; k < CHK / 8; ++k) o[k] = ((c16 + k) ^ sw) << 4;
          auto ld_side = [&](const uint8_t* base, float* t) {
#pragma unroll
            for (int k = 0; k < CHK / 8; ++k) unpack8(*reinterpret_cast<const uint4*>(base + o[k]), t + 8 * k);
          };
          if (p.has_add) {
            float t[CHK];
            ld_side(ib, t);
#pragma unroll
            for (int i = 0; i < CHK; ++i) f[i] += t[i];
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < CHK; ++i) f[i] = fmaxf(f[i], 0.f);
          }
          if (p.has_mask) {
            float t[CHK];
            ld_side(ib + p.has_add * BOXB, t);
#pragma unroll
            for (int i = 0; i < CHK; ++i) f[i] = t[i] > 0.f ? f[i] : 0.f;
          }
          if (p.has_bnx) {
            // fused BatchNorm(+ReLU) backward reductions: f is d(relu(bn(x))); recompute the ReLU mask from x like the
            // forward did, keep g = f * mask as the stored value and accumulate {sum g, sum g*xhat} below
            float xv[CHK], xh[4];
            ld_side(ib + (p.has_add + p.has_mask) * BOXB, xv);
            const float* cf = bnc + hs * NCT + cc;
#pragma unroll
            for (int i = 0; i < CHK; i += 4) {
              const float4 ca = *reinterpret_cast<const float4*>(cf + i), cb = *reinterpret_cast<const float4*>(cf + C + i);
              const float4 cg = *reinterpret_cast<const float4*>(cf + 2 * C + i), ct = *reinterpret_cast<const float4*>(cf + 3 * C + i);
              xh[0] = fmaf(xv[i], ca.x, cb.x); xh[1] = fmaf(xv[i + 1], ca.y, cb.y);
              xh[2] = fmaf(xv[i + 2], ca.z, cb.z); xh[3] = fmaf(xv[i + 3], ca.w, cb.w);
              if (p.bnr_relu) {
                f[i] = fmaf(cg.x, xh[0], ct.x) > 0.f ? f[i] : 0.f; f[i + 1] = fmaf(cg.y, xh[1], ct.y) > 0.f ? f[i + 1] : 0.f;
                f[i + 2] = fmaf(cg.z, xh[2], ct.z) > 0.f ? f[i + 2] : 0.f; f[i + 3] = fmaf(cg.w, xh[3], ct.w) > 0.f ? f[i + 3] : 0.f;
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) { acc_s[cc + i + e] += f[i + e]; acc_q[cc + i + e] = fmaf(f[i + e], xh[e], acc_q[cc + i + e]); }
            }
          }
          uint4 pk[CHK / 8];
#pragma unroll
          for (int k = 0; k < CHK / 8; ++k) {
            pk[k].x = pack_bf16x2(f[8 * k], f[8 * k + 1]); pk[k].y = pack_bf16x2(f[8 * k + 2], f[8 * k + 3]);
            pk[k].z = pack_bf16x2(f[8 * k + 4], f[8 * k + 5]); pk[k].w = pack_bf16x2(f[8 * k + 6], f[8 * k + 7]);
          }
          if (cc == 0) {
            if (lane == 0) twait(&sfree[j], ((seq >> p.nsb_log) & 1) ^ 1, etr, 7);
            __syncwarp();
          }
#pragma unroll
          for (int k = 0; k < CHK / 8; ++k) *reinterpret_cast<uint4*>(yb + o[k]) = pk[k];
          if (has_stats && !p.has_bnx) {
            // per-thread partial sums of the stored (bf16-rounded) values; reduced across lanes once per CTA
            float t[CHK];
#pragma unroll
            for (int k = 0; k < CHK / 8; ++k) unpack8(pk[k], t + 8 * k);
#pragma unroll
            for (int i = 0; i < CHK; ++i) { acc_s[cc + i] += t[i]; acc_q[cc + i] = fmaf(t[i], t[i], acc_q[cc + i]); }
          }
        }
        const long long tf0 = etr ? clock64() : 0;
        fence_proxy_async();
        __syncwarp();
        if (etr && lane == 0) etr[12] += clock64() - tf0;
        if (lane == 0) mbar_arrive(&sready[j]);
      }
    }
    if (etr && lane == 0) { etr[9] += clock64() - e_t0; etr[10] += seq; }
    if (has_stats) {
      // CHK-wide butterfly reduce-scatter over the warp's 32 pixels: after log2(CHK) halving exchanges every lane holds one
      // channel's partial sum; the remaining xor steps finish the sum over the lanes that share that channel
#pragma unroll
      for (int cc = 0; cc < NCT; cc += CHK) {
        int ch = 0;
#pragma unroll
        for (int off = 16, n = CHK / 2; n >= 1; off >>= 1, n >>= 1) {
          const bool upper = (lane & off) != 0;
          if (upper) ch += n;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const float send_s = upper ? acc_s[cc + i] : acc_s[cc + i + n], keep_s = upper ? acc_s[cc + i + n] : acc_s[cc + i];
            const float send_q = upper ? acc_q[cc + i] : acc_q[cc + i + n], keep_q = upper ? acc_q[cc + i + n] : acc_q[cc + i];
            acc_s[cc + i] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, off);
            acc_q[cc + i] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, off);
          }
        }
        constexpr int REST = 32 / CHK;            // lanes sharing one channel: 2 (CHK 16) or 4 (CHK 8)
#pragma unroll
        for (int off = REST / 2; off >= 1; off >>= 1) {
          acc_s[cc] += __shfl_xor_sync(0xffffffffu, acc_s[cc], off);
          acc_q[cc] += __shfl_xor_sync(0xffffffffu, acc_q[cc], off);
        }
        if ((lane & (REST - 1)) == 0) {      // exactly one warp (q, hs) owns slot [q][channel]: no atomics, fixed order below
          csum[q * C + hs * NCT + cc + ch] = acc_s[cc];
          csq[q * C + hs * NCT + cc + ch] = acc_q[cc];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
      if (it > 0 && warp < EPI0 + C / 32) {
        const int c = (warp - EPI0) * 32 + lane;
        const double s4 = (((double)csum[c] + (double)csum[C + c]) + (double)csum[2 * C + c]) + (double)csum[3 * C + c];
        const double q4 = (((double)csq[c] + (double)csq[C + c]) + (double)csq[2 * C + c]) + (double)csq[3 * C + c];
        atomicAdd(p.stats + c, s4);
        atomicAdd(p.stats + C + c, q4);
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

template <int C, int KT, int EW>
int launch3(const Tc3Maps& maps, const Tc3Params& p, int smem_bytes, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc3_kernel<C, KT, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { rsa_set_error("conv_tc3: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return RSA_ERR_CUDA; }
    configured = true;
  }
  const int grid = p.items < rsa_num_sms() ? p.items : rsa_num_sms();
  cudaError_t le = launch_pdl(conv_tc3_kernel<C, KT, EW>, dim3(grid), dim3(T3Warps<KT, EW>::THREADS), (size_t)smem_bytes, st, maps, p);
  if (le != cudaSuccess) { rsa_set_error("conv_tc3: launch: %s", cudaGetErrorString(le)); return RSA_ERR_CUDA; }
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

}  // namespace

static long long* g_t3_trace = nullptr;
/* Diagnostic hook: when buf != NULL (device memory, 148*16 int64, zeroed by the caller) every following rsa_conv_tc3_fwd
 * launch adds, per CTA, the cycles its roles spent blocked: [0] producer on the A ring, [1] producer on the side ring,
 * [2] MMA on the accumulators (epilogue behind), [3] MMA on A tiles (TMA behind), [5] store warp on staged tiles,
 * [6] epilogue on accumulators (MMA behind), [7] epilogue on staging buffers, [8] epilogue on side tiles,
 * [9] epilogue total cycles, [10] sub-tiles.  scripts/trace_tc3.py prints the breakdown. */
extern "C" int rsa_conv_tc3_set_trace(long long* buf) { g_t3_trace = buf; return RSA_OK; }

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
  // short (four resident branches, or 64 channels with a dilation-3 halo)
  int KT = C == 32 ? (nbr > 1 ? 2 : 4) : (max_ad == 3 ? 1 : 2);
  p.nsb = C == 32 ? 4 : 2;
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
  p.nsb_log = p.nsb == 4 ? 2 : 1;
  p.nsi = nside ? (C == 32 ? 4 : (nside == 2 ? 1 : 2)) : 0;     // 64 channels: shared memory is short, keep >= 2 A stages
  p.nsi_log = p.nsi == 4 ? 2 : (p.nsi == 2 ? 1 : 0);
  auto plan = [&](int kt) {
    int slot = any_box ? kt * BOXB : 0;
    if (max_ad) { const int hb = (16 + 2 * max_ad) * (8 * kt + 2 * max_ad) * PITCH; slot = hb > slot ? hb : slot; }
    slot = (slot + 1023) & ~1023;
    const Tc3Smem L0(C, nbr, p.nsb, nside, p.nsi, 0, slot);
    int ns = (227 * 1024 - L0.total - 256) / (slot + 16);
    p.slot_bytes = slot;
    p.nstages = ns > 8 ? 8 : ns;
  };
  plan(KT);
  RSA_REQUIRE(p.nstages >= 2, RSA_ERR_SHAPE, "conv_tc3_fwd: shared memory budget allows only %d stage(s)", p.nstages);
  p.tiles_w = W / (8 * KT); p.tiles_h = H / 16;
  p.items = p.tiles_w * p.tiles_h * N;
  const Tc3Smem L(C, nbr, p.nsb, nside, p.nsi, p.nstages, p.slot_bytes);
  RSA_REQUIRE(L.total <= 227 * 1024, RSA_ERR_SHAPE, "conv_tc3_fwd: shared memory %d", L.total);
  p.stats = stats; p.relu = relu; p.trace = g_t3_trace;
  const CUtensorMapSwizzle swz = C == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  Tc3Maps maps;
  for (int b = 0; b < nbr; ++b) {
    const int ad = p.dil[b] < 0 ? -p.dil[b] : p.dil[b];
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)(p.halo[b] ? 8 * KT + 2 * ad : 8 * KT), (cuuint32_t)(p.halo[b] ? 16 + 2 * ad : 16), 1};
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
    // 16x8-pixel tiles: the TMA-stored output and the TMA-loaded epilogue side inputs share one box shape
    auto enc_tile = [&](CUtensorMap* tm, const void* base) -> CUresult {
      cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      cuuint32_t box[4] = {(cuuint32_t)C, 8, 16, 1};
      cuuint32_t es[4] = {1, 1, 1, 1};
      return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    CUresult r = enc_tile(&maps.out, out);
    RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(out) failed (%d)", (int)r);
    maps.add = maps.out; maps.mask = maps.out; maps.bnx = maps.out;
    if (p.has_bnx) {
      r = enc_tile(&maps.bnx, bnr_x);
      RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(bnr_x) failed (%d)", (int)r);
    }
    if (p.has_add) {
      r = enc_tile(&maps.add, residual ? residual : out);
      RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(addend) failed (%d)", (int)r);
    }
    if (p.has_mask) {
      r = enc_tile(&maps.mask, mask);
      RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_fwd: cuTensorMapEncodeTiled(mask) failed (%d)", (int)r);
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  // eight epilogue warps: sixteen (EW = 16) measured 5-15 % slower on every shape - the epilogue's shared-memory accesses
  // queue behind the tensor core's operand reads, so more warps only add contention
  if (C == 32) return KT == 2 ? launch3<32, 2, 8>(maps, p, L.total, st) : launch3<32, 4, 8>(maps, p, L.total, st);
  return KT == 1 ? launch3<64, 1, 8>(maps, p, L.total, st) : launch3<64, 2, 8>(maps, p, L.total, st);
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
// at the end.  Large dilations (C = 32 only) use nine 16 x 8 boxes per item with LBO = one box.
// =====================================================================================================
namespace {

constexpr int W3_THREADS = 256;   // warp 0 TMA, 1..3 MMA (tap row dy = warp - 1), 4..7 final reduction

struct Wg3Params {
  int N, H, W, dil, halo;
  int IW;                 // item width in pixels: 16 (halo) or 8 (boxes); item height 16
  int items, tiles_w, tiles_h;
  int nstages, stage_bytes, a_bytes, a_tx;   // a_tx: bytes the halo box actually delivers
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
        const int n = r, h0 = th * 16, w0 = tw * p.IW;
        uint8_t* sa = smem + stage * p.stage_bytes;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (p.halo) {
          mbar_expect_tx(&full_bar[stage], p.a_tx + 16 * p.IW * PITCH);
          tma_load_4d(sa, &tmX, &full_bar[stage], 0, w0 - d, h0 - d, n);
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
        const int ch = th * 16 + (dyi - 1) * d;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * p.stage_bytes);
        const uint32_t sb = sa + p.a_bytes;
        if (p.halo || !(ch + 16 <= 0 || ch >= p.H)) {
          // A: atom i = tap (dy, dx = i - 1) (C = 32: four atoms, C = 64: two per chain); two 8-pixel K groups per MMA
          // = rows (2rp, 2rp+1), column block cb
          const uint32_t a0 = p.halo ? sa + (uint32_t)((dyi * d * Wh) * PITCH) : sa + (uint32_t)(dyi * 3) * boxb;
          const uint32_t arow = p.halo ? Wh * PITCH : p.IW * PITCH;
          const uint32_t lbo = p.halo ? d * PITCH : boxb;
          const uint32_t brow = p.IW * PITCH;
#pragma unroll 1
          for (int rp = 0; rp < 8; ++rp) {
            for (int cb = 0; cb < ncb; ++cb) {
              const uint64_t bdesc = w3_mndesc<C>(sb + 2 * rp * brow + cb * 8 * PITCH, 0, brow);
#pragma unroll
              for (int c = 0; c < NACC; ++c) {
                const uint64_t adesc = w3_mndesc<C>(a0 + 2 * rp * arow + cb * 8 * PITCH + c * 2 * lbo, lbo, arow);
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
        bool any = p.halo;
        if (!any) {
          for (int item = blockIdx.x; item < p.items && !any; item += gridDim.x) {
            const int ch = ((item / p.tiles_w) % p.tiles_h) * 16 + (dyi - 1) * d;
            any = !(ch + 16 <= 0 || ch >= p.H);
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

/* C = 32: any dilation; C = 64: dilations <= 3 (the nine boxes of a large dilation do not fit beside the dy tile). */
extern "C" int rsa_conv_tc3_wgrad_supported(int N, int H, int W, int C, int dil) {
  return rsa_conv_tc3_supported(N, H, W, C) && dil > 0 && (C == 32 || dil <= 3);
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
  p.halo = dil <= 3;
  p.IW = p.halo ? 16 : 8;
  p.tiles_w = W / p.IW; p.tiles_h = H / 16;
  p.items = p.tiles_w * p.tiles_h * N;
  // the unused trailing M atom reads up to 2*dil pixels (halo) / one box past the A region: keep that inside the stage
  const int a_raw = p.halo ? (16 + 2 * dil) * (p.IW + 2 * dil) * PITCH : 9 * 16 * p.IW * PITCH;
  p.a_tx = a_raw;
  p.a_bytes = (a_raw + 2 * dil * PITCH + 1023) & ~1023;
  p.stage_bytes = p.a_bytes + 16 * p.IW * PITCH;
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
  CUresult r = p.halo ? encode(&tmX, x, p.IW + 2 * dil, 16 + 2 * dil) : encode(&tmX, x, p.IW, 16);
  RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_wgrad: cuTensorMapEncodeTiled(x) failed (%d)", (int)r);
  r = encode(&tmDY, dy, p.IW, 16);
  RSA_REQUIRE(r == CUDA_SUCCESS, RSA_ERR_CUDA, "conv_tc3_wgrad: cuTensorMapEncodeTiled(dy) failed (%d)", (int)r);
  if (C == 32) return launch_wg3<32>(tmX, tmDY, p, smem_bytes, (cudaStream_t)stream);
  return launch_wg3<64>(tmX, tmDY, p, smem_bytes, (cudaStream_t)stream);
}
