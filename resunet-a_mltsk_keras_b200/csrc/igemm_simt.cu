// igemm_simt.cu — generic gather implicit GEMM on CUDA cores with fp32 accumulation.
//
// This is the fp32 *validation-mode* convolution engine (north_star: rel-L2 <= 1e-4 against the
// reference) and the engine of every gather-style 1x1 convolution (stride-2 down-sampling,
// nearest up-sampling, channel concat, combine's ReLU; model2.py:36-39,81-94,101-111).  The
// bf16 performance path of the 3x3 ResBlock-a convolutions is conv_tc.cu (tcgen05/TMA).
//
// GEMM view: M = N*Ho*Wo output pixels, N = Co, K = sum of segment channels.  A K-segment is a
// (tensor, spatial map) pair, so a dilated 3x3 conv is 9 segments of the same tensor with
// offsets (dy*d, dx*d) (model2.py:19-24), a concat is one segment per input (model2.py:83) and
// up/down-sampling is a shift / multiplier on the pixel coordinate.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct IgemmParams {
  rsa_seg_t seg[RSA_MAX_SEG];
  int nseg;
  const float* w;
  int ldw, transB;
  const float* bias;
  void* out;
  const void* residual;
  const void* mask;
  double* stats;
  int N, Ho, Wo, Co;
  int accumulate, relu;
};

// source pixel of segment s for output pixel (n,h,w); returns element offset or -1
__device__ __forceinline__ int64_t seg_src_offset(const rsa_seg_t& s, int n, int h, int w) {
  int hm = h * s.mult, wm = w * s.mult;
  if (s.aligned) {
    int msk = (1 << s.shift) - 1;
    if ((hm & msk) | (wm & msk)) return -1;
  }
  int hs = (hm >> s.shift) + s.off_h;
  int ws = (wm >> s.shift) + s.off_w;
  if (hs < 0 || hs >= s.Hs || ws < 0 || ws >= s.Ws) return -1;
  return (((int64_t)n * s.Hs + hs) * s.Ws + ws) * s.C;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(NT) igemm_fwd_kernel(const IgemmParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ double red_s[BN], red_q[BN];   // double: order-independent below fp32 resolution (reproducible statistics)

  const int tid = threadIdx.x;
  const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-load mapping: pixel am, 4 consecutive channels from ak
  const int am = tid >> 2, ak = (tid & 3) * 4;
  const int64_t m = m0 + am;
  const bool mvalid = m < M;
  int pn = 0, ph = 0, pw = 0;
  if (mvalid) {
    pw = (int)(m % p.Wo);
    int64_t t = m / p.Wo;
    ph = (int)(t % p.Ho);
    pn = (int)(t / p.Ho);
  }
  // B-load mapping
  const int bk_nt = tid >> 4, bn_nt = (tid & 15) * 4;   // non-transposed: row k, 4 cols
  const int bn_t = tid >> 2, bk_t = (tid & 3) * 4;      // transposed: col n, 4 consecutive k
  // compute mapping
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int si = 0; si < p.nseg; ++si) {
    const rsa_seg_t& s = p.seg[si];
    const int C = s.C;
    int64_t soff = mvalid ? seg_src_offset(s, pn, ph, pw) : -1;
    const TI* sp = soff >= 0 ? reinterpret_cast<const TI*>(s.src) + soff : nullptr;
    const bool vecA = (C & 3) == 0;
    const float* wseg = p.w + s.w_off;
    for (int kc = 0; kc < C; kc += BK) {
      // ---- A tile
      float a4[4] = {0.f, 0.f, 0.f, 0.f};
      if (sp) {
        int c = kc + ak;
        if (vecA) {
          if (c < C) ld4<TI>(sp + c, a4);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < C) a4[j] = ldf<TI>(sp + c + j);
        }
        if (s.relu_in) {
#pragma unroll
          for (int j = 0; j < 4; ++j) a4[j] = fmaxf(a4[j], 0.f);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[ak + j][am] = a4[j];
      // ---- B tile
      if (!p.transB) {
        int k = kc + bk_nt;
        float b4[4] = {0.f, 0.f, 0.f, 0.f};
        if (k < C) {
          const float* wp = wseg + (int64_t)k * p.ldw + n0 + bn_nt;
          if (((reinterpret_cast<uintptr_t>(wp) & 15) == 0) && n0 + bn_nt + 3 < p.Co) {
            float4 t = __ldg(reinterpret_cast<const float4*>(wp));
            b4[0] = t.x; b4[1] = t.y; b4[2] = t.z; b4[3] = t.w;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n0 + bn_nt + j < p.Co) b4[j] = __ldg(wp + j);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[bk_nt][bn_nt + j] = b4[j];
      } else {
        int nn = n0 + bn_t;
        float b4[4] = {0.f, 0.f, 0.f, 0.f};
        if (nn < p.Co) {
          const float* wp = wseg + (int64_t)nn * p.ldw + kc + bk_t;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (kc + bk_t + j < C) b4[j] = __ldg(wp + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) Bs[bk_t + j][bn_t] = b4[j];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        float a[4] = {av.x, av.y, av.z, av.w};
        float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue
  TO* out = reinterpret_cast<TO*>(p.out);
  const TO* res = reinterpret_cast<const TO*>(p.residual);
  const TO* msk = reinterpret_cast<const TO*>(p.mask);
  const int nb = n0 + tx * 4;
  float bias4[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (nb + j < p.Co) bias4[j] = __ldg(p.bias + nb + j);
  }
  float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
  const bool vecO = ((p.Co & 3) == 0) && (nb + 3 < p.Co);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
    int64_t o = mm * p.Co + nb;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias4[j];
    if (vecO) {
      float t[4];
      if (res) { ld4<TO>(res + o, t);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += t[j]; }
      if (p.accumulate) { ld4<TO>(out + o, t);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += t[j]; }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f); }
      if (msk) { ld4<TO>(msk + o, t);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = t[j] > 0.f ? v[j] : 0.f; }
      st4<TO>(out + o, v);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (nb + j >= p.Co) continue;
        if (res) v[j] += ldf<TO>(res + o + j);
        if (p.accumulate) v[j] += ldf<TO>(out + o + j);
        if (p.relu) v[j] = fmaxf(v[j], 0.f);
        if (msk) v[j] = ldf<TO>(msk + o + j) > 0.f ? v[j] : 0.f;
        stf<TO>(out + o + j, v[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { ssum[j] += v[j]; ssq[j] += v[j] * v[j]; }
  }
  if (p.stats) {
    if (tid < BN) { red_s[tid] = 0.0; red_q[tid] = 0.0; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&red_s[tx * 4 + j], (double)ssum[j]);
      atomicAdd(&red_q[tx * 4 + j], (double)ssq[j]);
    }
    __syncthreads();
    if (tid < BN && n0 + tid < p.Co) {
      atomicAdd(p.stats + n0 + tid, red_s[tid]);
      atomicAdd(p.stats + p.Co + n0 + tid, red_q[tid]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct WgradParams {
  rsa_seg_t seg[RSA_MAX_SEG];
  int nseg;
  const void* dy;
  float* dw;
  int ldw;
  float* dbias;
  int N, Ho, Wo, Co;
  int64_t chunk;   // pixels per blockIdx.z
};

constexpr int WK = 64, WN = 64, WP = 16;   // dW tile 64 (k) x 64 (n), 16 pixels per step

template <typename TI, typename TG>
__global__ void __launch_bounds__(NT) igemm_wgrad_kernel(const WgradParams p) {
  __shared__ float As[WP][WK + 4];
  __shared__ float Bs[WP][WN + 4];
  const int tid = threadIdx.x;
  // blockIdx.x -> (segment, k-tile)
  int si = 0, kt = blockIdx.x;
  for (; si < p.nseg; ++si) {
    int tiles = (p.seg[si].C + WK - 1) / WK;
    if (kt < tiles) break;
    kt -= tiles;
  }
  if (si >= p.nseg) return;
  const rsa_seg_t& s = p.seg[si];
  const int C = s.C;
  const int kc = kt * WK;
  const int n0 = blockIdx.y * WN;
  const int64_t M = (int64_t)p.N * p.Ho * p.Wo;
  const int64_t mbeg = (int64_t)blockIdx.z * p.chunk;
  const int64_t mend = min(M, mbeg + p.chunk);
  const TG* dy = reinterpret_cast<const TG*>(p.dy);

  const int lp = tid >> 4, lc = (tid & 15) * 4;   // load mapping: pixel lp, 4 consecutive k / n
  const int ty = tid >> 4, tx = tid & 15;
  const bool vecA = (C & 3) == 0;
  const bool vecB = (p.Co & 3) == 0;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = p.dbias != nullptr && blockIdx.x == 0 && ty == 0;

  // pixel coordinates of this thread's load row, advanced incrementally (no per-iteration 64-bit division)
  int pw, ph, pn;
  {
    int64_t m = mbeg + lp;
    pw = (int)(m % p.Wo);
    int64_t t = m / p.Wo;
    ph = (int)(t % p.Ho);
    pn = (int)(t / p.Ho);
  }
  for (int64_t mb = mbeg; mb < mend; mb += WP) {
    int64_t m = mb + lp;
    float a4[4] = {0.f, 0.f, 0.f, 0.f}, b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < mend) {
      int64_t soff = seg_src_offset(s, pn, ph, pw);
      int c = kc + lc;
      if (soff >= 0 && c < C) {
        const TI* sp = reinterpret_cast<const TI*>(s.src) + soff + c;
        if (vecA) ld4<TI>(sp, a4);
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < C) a4[j] = ldf<TI>(sp + j);
        }
        if (s.relu_in) {
#pragma unroll
          for (int j = 0; j < 4; ++j) a4[j] = fmaxf(a4[j], 0.f);
        }
      }
      int nn = n0 + lc;
      if (nn < p.Co) {
        const TG* gp = dy + m * p.Co + nn;
        if (vecB) ld4<TG>(gp, b4);
        else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (nn + j < p.Co) b4[j] = ldf<TG>(gp + j);
        }
      }
    }
    pw += WP;                       // advance WP pixels along the row-major (n, h, w) order
    while (pw >= p.Wo) { pw -= p.Wo; if (++ph == p.Ho) { ph = 0; ++pn; } }
#pragma unroll
    for (int j = 0; j < 4; ++j) { As[lp][lc + j] = a4[j]; Bs[lp][lc + j] = b4[j]; }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < WP; ++pp) {
      float4 av = *reinterpret_cast<const float4*>(&As[pp][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[pp][tx * 4]);
      float a[4] = {av.x, av.y, av.z, av.w};
      float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] += b[j];
      }
    }
    __syncthreads();
  }
  float* dwseg = p.dw + s.w_off;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int k = kc + ty * 4 + i;
    if (k >= C) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int nn = n0 + tx * 4 + j;
      if (nn < p.Co) atomicAdd(dwseg + (int64_t)k * p.ldw + nn, acc[i][j]);
    }
  }
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int nn = n0 + tx * 4 + j;
      if (nn < p.Co) atomicAdd(p.dbias + nn, bsum[j]);
    }
  }
}

int check_segs(const rsa_seg_t* segs, int nseg) {
  RSA_REQUIRE(nseg >= 1 && nseg <= RSA_MAX_SEG, RSA_ERR_SHAPE, "igemm: nseg=%d out of range [1,%d]", nseg,
              RSA_MAX_SEG);
  for (int i = 0; i < nseg; ++i) {
    RSA_REQUIRE(segs[i].src != nullptr && segs[i].C > 0 && segs[i].Hs > 0 && segs[i].Ws > 0, RSA_ERR_SHAPE,
                "igemm: bad segment %d", i);
    RSA_REQUIRE(segs[i].mult >= 1 && segs[i].shift >= 0 && segs[i].shift < 8, RSA_ERR_SHAPE,
                "igemm: bad gather map in segment %d", i);
  }
  return RSA_OK;
}

}  // namespace

extern "C" int rsa_igemm_fwd(const rsa_seg_t* segs, int nseg, int in_dtype, const float* w, int ldw,
                             int transB, const float* bias, void* out, int out_dtype, const void* residual,
                             const void* mask, double* stats, int N, int Ho, int Wo, int Co, int accumulate,
                             int relu, void* stream) {
  int rc = check_segs(segs, nseg);
  if (rc) return rc;
  RSA_REQUIRE(w && out && N > 0 && Ho > 0 && Wo > 0 && Co > 0, RSA_ERR_SHAPE, "igemm_fwd: bad shape/pointers");
  RSA_REQUIRE((in_dtype | 1) == 1 && (out_dtype | 1) == 1, RSA_ERR_DTYPE, "igemm_fwd: bad dtype");
  IgemmParams p;
  for (int i = 0; i < nseg; ++i) p.seg[i] = segs[i];
  p.nseg = nseg; p.w = w; p.ldw = ldw; p.transB = transB; p.bias = bias; p.out = out;
  p.residual = residual; p.mask = mask; p.stats = stats; p.N = N; p.Ho = Ho; p.Wo = Wo; p.Co = Co;
  p.accumulate = accumulate; p.relu = relu;
  int64_t M = (int64_t)N * Ho * Wo;
  dim3 grid((unsigned)ceil_div64(M, BM), (unsigned)((Co + BN - 1) / BN));
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == RSA_F32 && out_dtype == RSA_F32) igemm_fwd_kernel<float, float><<<grid, NT, 0, st>>>(p);
  else if (in_dtype == RSA_BF16 && out_dtype == RSA_BF16) igemm_fwd_kernel<bf16, bf16><<<grid, NT, 0, st>>>(p);
  else if (in_dtype == RSA_BF16 && out_dtype == RSA_F32) igemm_fwd_kernel<bf16, float><<<grid, NT, 0, st>>>(p);
  else igemm_fwd_kernel<float, bf16><<<grid, NT, 0, st>>>(p);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

extern "C" int rsa_igemm_wgrad(const rsa_seg_t* segs, int nseg, int in_dtype, const void* dy, int dy_dtype,
                               float* dw, int ldw, float* dbias, int N, int Ho, int Wo, int Co, void* stream) {
  int rc = check_segs(segs, nseg);
  if (rc) return rc;
  RSA_REQUIRE(dy && dw && N > 0 && Ho > 0 && Wo > 0 && Co > 0, RSA_ERR_SHAPE, "igemm_wgrad: bad shape/pointers");
  RSA_REQUIRE((in_dtype | 1) == 1 && (dy_dtype | 1) == 1, RSA_ERR_DTYPE, "igemm_wgrad: bad dtype");
  WgradParams p;
  int ktiles = 0;
  for (int i = 0; i < nseg; ++i) { p.seg[i] = segs[i]; ktiles += (segs[i].C + WK - 1) / WK; }
  p.nseg = nseg; p.dy = dy; p.dw = dw; p.ldw = ldw; p.dbias = dbias; p.N = N; p.Ho = Ho; p.Wo = Wo; p.Co = Co;
  int64_t M = (int64_t)N * Ho * Wo;
  int ntiles = (Co + WN - 1) / WN;
  // split the pixel reduction so that the grid covers ~4 waves of the 148 SMs
  int64_t want = ceil_div64(4 * rsa_num_sms(), (int64_t)ktiles * ntiles);
  int64_t maxsplit = ceil_div64(M, 4 * WP);
  int64_t split = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
  if (split > 65535) split = 65535;
  int64_t chunk = ceil_div64(ceil_div64(M, split), WP) * WP;
  split = ceil_div64(M, chunk);
  p.chunk = chunk;
  dim3 grid((unsigned)ktiles, (unsigned)ntiles, (unsigned)split);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == RSA_F32 && dy_dtype == RSA_F32) igemm_wgrad_kernel<float, float><<<grid, NT, 0, st>>>(p);
  else if (in_dtype == RSA_BF16 && dy_dtype == RSA_BF16) igemm_wgrad_kernel<bf16, bf16><<<grid, NT, 0, st>>>(p);
  else if (in_dtype == RSA_BF16 && dy_dtype == RSA_F32) igemm_wgrad_kernel<bf16, float><<<grid, NT, 0, st>>>(p);
  else igemm_wgrad_kernel<float, bf16><<<grid, NT, 0, st>>>(p);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
