// labels.cu — multitask label generation on the GPU.
//
// The reference derives the boundary / distance / colour targets of every patch with OpenCV on the CPU
// (multitasking_utils.py:6-34: cv2.Canny(mask,0,1) + 3x3 cross dilate; cv2.distanceTransform(DIST_L2, precise) + min-max
// normalisation; preprocess_save_patches_ISPRS.py:224-228: 8-bit RGB->HSV / [179,255,255]) and stores them as four extra
// .npy files per patch.  These kernels restate OpenCV's algorithms for exactly those argument values (see
// oracle/labels_oracle.py for the derivation and the cv2-generated known answers) so that the targets can be produced from
// the one-hot segmentation batch that is already resident in HBM:
//   * boundary and colour are integer computations: bit-exact against OpenCV;
//   * distance: exact integer squared Euclidean distance, correctly rounded float32 sqrt, OpenCV's min-max formula.
// One CTA owns one (patch, class) plane; planes are small (<= 64 K pixels for 256x256 patches) and live in L1/L2, so the
// kernels are written for clarity, not bandwidth: the whole batch costs far less than one training step.
#include "common.cuh"

namespace {

constexpr int LT = 512;

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// plane(n, c)[y][x] = (uint8) label[n, y, x, c]   (label.astype(np.uint8), multitasking_utils.py:11,29)
__global__ void __launch_bounds__(256) onehot_to_planes_kernel(const float* __restrict__ label, uint8_t* __restrict__ planes,
                                                                 int N, int H, int W, int C) {
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const int64_t hw = (int64_t)H * W;
    const int n = (int)(pix / hw);
    planes[((int64_t)n * C + c) * hw + pix % hw] = (uint8_t)label[i];
  }
}

// Sobel 3x3 with BORDER_REPLICATE on a uint8 plane
__device__ __forceinline__ void sobel_at(const uint8_t* __restrict__ m, int H, int W, int y, int x, int& dx, int& dy) {
  const int y0 = clampi(y - 1, 0, H - 1), y2 = clampi(y + 1, 0, H - 1);
  const int x0 = clampi(x - 1, 0, W - 1), x2 = clampi(x + 1, 0, W - 1);
  const int a = m[y0 * W + x0], b = m[y0 * W + x], c = m[y0 * W + x2];
  const int d = m[y * W + x0], f = m[y * W + x2];
  const int g = m[y2 * W + x0], h = m[y2 * W + x], i = m[y2 * W + x2];
  dx = (c - a) + 2 * (f - d) + (i - g);
  dy = (g - a) + 2 * (h - b) + (i - c);
}

// cv2.Canny(plane, 0, 1) -> dilate(cross 3x3) -> /255, written to out[n, y, x, c] (float32 NHWC)
// scratch: mag (uint8) and map (uint8) planes, 2 * H * W bytes per (n, c)
__global__ void __launch_bounds__(LT) boundary_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ scratch,
                                                      float* __restrict__ out, int N, int H, int W, int C) {
  const int plane = blockIdx.x, n = plane / C, c = plane % C;
  const int hw = H * W;
  const uint8_t* m = planes + (int64_t)plane * hw;
  uint8_t* mag = scratch + (int64_t)plane * 2 * hw;
  uint8_t* map = mag + hw;
  // 1. L1 gradient magnitude (<= 8 on a {0,1} mask; saturate for general uint8 input is not needed by the reference)
  for (int p = threadIdx.x; p < hw; p += LT) {
    int dx, dy;
    sobel_at(m, H, W, p / W, p % W, dx, dy);
    const int g = abs(dx) + abs(dy);
    mag[p] = (uint8_t)(g > 255 ? 255 : g);
  }
  __syncthreads();
  // 2. non-maximum suppression (zero magnitude outside the image); map: 2 = edge seed (mag > high = 1), 0 = candidate
  //    (mag > low = 0), 1 = not an edge
  for (int p = threadIdx.x; p < hw; p += LT) {
    const int y = p / W, x = p % W;
    const int mc = mag[p];
    uint8_t v = 1;
    if (mc > 0) {
      int dx, dy;
      sobel_at(m, H, W, y, x, dx, dy);
      auto M = [&](int yy, int xx) -> int { return (yy < 0 || yy >= H || xx < 0 || xx >= W) ? 0 : mag[yy * W + xx]; };
      const int ax = abs(dx), ay = abs(dy) << 15;
      const int tg22 = ax * 13573;
      bool ok;
      if (ay < tg22) ok = mc > M(y, x - 1) && mc >= M(y, x + 1);
      else {
        const int tg67 = tg22 + (ax << 16);
        if (ay > tg67) ok = mc > M(y - 1, x) && mc >= M(y + 1, x);
        else {
          const int s = ((dx ^ dy) < 0) ? -1 : 1;
          ok = mc > M(y - 1, x - s) && mc > M(y + 1, x + s);
        }
      }
      if (ok) v = mc > 1 ? 2 : 0;
    }
    map[p] = v;
  }
  // 3. hysteresis: grow the seeds through 8-connected candidates until nothing changes
  for (;;) {
    __syncthreads();
    int changed = 0;
    for (int p = threadIdx.x; p < hw; p += LT) {
      if (map[p] != 0) continue;
      const int y = p / W, x = p % W;
      bool hit = false;
      for (int dy = -1; dy <= 1 && !hit; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W && map[yy * W + xx] == 2) { hit = true; break; }
        }
      if (hit) { map[p] = 2; changed = 1; }
    }
    if (!__syncthreads_or(changed)) break;
  }
  // 4. dilate with the 3x3 cross (border ignored), /255
  for (int p = threadIdx.x; p < hw; p += LT) {
    const int y = p / W, x = p % W;
    bool e = map[p] == 2;
    if (y > 0) e = e || map[p - W] == 2;
    if (y < H - 1) e = e || map[p + W] == 2;
    if (x > 0) e = e || map[p - 1] == 2;
    if (x < W - 1) e = e || map[p + 1] == 2;
    out[((int64_t)n * hw + p) * C + c] = e ? 1.f : 0.f;
  }
}

// exact Euclidean distance transform + min-max normalisation of one plane
// scratch: int32 g[H*W] (vertical distance to the nearest zero of the column) + float32 d[H*W] per plane
__global__ void __launch_bounds__(LT) distance_kernel(const uint8_t* __restrict__ planes, int32_t* __restrict__ scratch,
                                                      float* __restrict__ out, int N, int H, int W, int C) {
  const int plane = blockIdx.x, n = plane / C, c = plane % C;
  const int hw = H * W;
  const uint8_t* m = planes + (int64_t)plane * hw;
  int32_t* g = scratch + (int64_t)plane * 2 * hw;
  float* d = reinterpret_cast<float*>(g + hw);
  const int INF = 1 << 20;
  __shared__ float s_red[2][LT / 32];
  __shared__ int s_any;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  // 1. column pass
  for (int x = threadIdx.x; x < W; x += LT) {
    int last = -INF;
    for (int y = 0; y < H; ++y) {
      if (m[y * W + x] == 0) last = y;
      g[y * W + x] = last == -INF ? INF : y - last;
    }
    int next = INF;
    for (int y = H - 1; y >= 0; --y) {
      if (m[y * W + x] == 0) next = y;
      const int dn = next == INF ? INF : next - y;
      if (dn < g[y * W + x]) g[y * W + x] = dn;
    }
    if (last != -INF) s_any = 1;
  }
  __syncthreads();
  const bool any_zero = s_any != 0;
  // 2. row pass: exact minimum of (x - x')^2 + g(y, x')^2, correctly rounded sqrt
  float mx = 0.f, mn = 3.0e38f;
  for (int p = threadIdx.x; p < hw; p += LT) {
    float dist = 0.f;
    if (any_zero) {
      const int y = p / W, x = p % W;
      const int32_t* gr = g + y * W;
      long long best = (long long)gr[x] * gr[x];
      for (int r = 1; r < W && (long long)r * r < best; ++r) {
        if (x - r >= 0) { const long long v = (long long)r * r + (long long)gr[x - r] * gr[x - r]; if (v < best) best = v; }
        if (x + r < W) { const long long v = (long long)r * r + (long long)gr[x + r] * gr[x + r]; if (v < best) best = v; }
      }
      dist = __fsqrt_rn((float)best);
    }
    d[p] = dist;
    mx = fmaxf(mx, dist); mn = fminf(mn, dist);
  }
  // 3. block min / max
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); }
  if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = mx; s_red[1][threadIdx.x >> 5] = mn; }
  __syncthreads();
  mx = s_red[0][0]; mn = s_red[1][0];
  for (int i = 1; i < LT / 32; ++i) { mx = fmaxf(mx, s_red[0][i]); mn = fminf(mn, s_red[1][i]); }
  // 4. cv2.normalize(NORM_MINMAX, 0, 1): dst = src * (float)scale + (float)shift, scale/shift in double; zeros if flat
  const double range = (double)mx - (double)mn;
  const bool flat = !(range > 2.220446049250313e-16);
  const double scale = flat ? 0.0 : 1.0 / range;
  const float a = (float)scale, b = (float)(0.0 - (double)mn * scale);
  for (int p = threadIdx.x; p < hw; p += LT)
    out[((int64_t)n * hw + p) * C + c] = flat ? 0.f : __fadd_rn(__fmul_rn(d[p], a), b);
}

// 8-bit RGB -> HSV (OpenCV fixed point, H in [0,180)) / [179, 255, 255]
__global__ void __launch_bounds__(256) hsv_kernel(const uint8_t* __restrict__ rgb, float* __restrict__ out, int64_t npix) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < npix; i += (int64_t)gridDim.x * 256) {
    const int r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2];
    const int v = max(max(r, g), b), vmin = min(min(r, g), b), diff = v - vmin;
    // sdiv_table[v] = cvRound((255 << 12) / v), hdiv_table180[d] = cvRound((180 << 12) / (6 d)); rint = round-half-even
    const int sdiv = v ? (int)rint((double)(255 << 12) / (double)v) : 0;
    const int hdiv = diff ? (int)rint((double)(180 << 12) / (6.0 * (double)diff)) : 0;
    const int s = (diff * sdiv + (1 << 11)) >> 12;
    int h = v == r ? g - b : (v == g ? b - r + 2 * diff : r - g + 4 * diff);
    h = (h * hdiv + (1 << 11)) >> 12;
    if (h < 0) h += 180;
    out[3 * i] = __fdiv_rn((float)(h & 255), 179.f);
    out[3 * i + 1] = __fdiv_rn((float)(s & 255), 255.f);
    out[3 * i + 2] = __fdiv_rn((float)v, 255.f);
  }
}

}  // namespace

/* Bytes of scratch rsa_label_boundary / rsa_label_distance need for a [N,H,W,C] one-hot batch (planes + per-plane work). */
extern "C" int64_t rsa_label_workspace_bytes(int N, int H, int W, int C) {
  const int64_t planes = (int64_t)N * C * H * W;
  return planes + planes * 8 + 256;          // uint8 planes + max(2 uint8, int32 + float32) per pixel
}

/* bound[n,y,x,c] = dilate_cross3(Canny(uint8(label[n,:,:,c]), 0, 1)) / 255  — multitasking_utils.py:6-22.
 * label, bound: fp32 [N,H,W,C] device tensors; workspace: rsa_label_workspace_bytes() bytes. */
extern "C" int rsa_label_boundary(const float* label, float* bound, void* workspace, int N, int H, int W, int C, void* stream) {
  RSA_REQUIRE(label && bound && workspace && N > 0 && H > 0 && W > 0 && C > 0, RSA_ERR_SHAPE, "label_boundary: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* planes = (uint8_t*)workspace;
  const int64_t np = (int64_t)N * C * H * W;
  onehot_to_planes_kernel<<<rsa_num_sms() * 8, 256, 0, st>>>(label, planes, N, H, W, C);
  boundary_kernel<<<N * C, LT, 0, st>>>(planes, planes + ((np + 255) & ~(int64_t)255), bound, N, H, W, C);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* dist[n,y,x,c] = minmax01(distanceTransform(uint8(label[n,:,:,c]), DIST_L2, precise))  — multitasking_utils.py:25-34. */
extern "C" int rsa_label_distance(const float* label, float* dist, void* workspace, int N, int H, int W, int C, void* stream) {
  RSA_REQUIRE(label && dist && workspace && N > 0 && H > 0 && W > 0 && C > 0, RSA_ERR_SHAPE, "label_distance: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* planes = (uint8_t*)workspace;
  const int64_t np = (int64_t)N * C * H * W;
  onehot_to_planes_kernel<<<rsa_num_sms() * 8, 256, 0, st>>>(label, planes, N, H, W, C);
  distance_kernel<<<N * C, LT, 0, st>>>(planes, (int32_t*)(planes + ((np + 255) & ~(int64_t)255)), dist, N, H, W, C);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* color[p, 0:3] = cvtColor(rgb, RGB2HSV) / [179, 255, 255] for uint8 RGB pixels — preprocess_save_patches_ISPRS.py:89-94,224-228. */
extern "C" int rsa_label_hsv(const uint8_t* rgb, float* color, int64_t npix, void* stream) {
  RSA_REQUIRE(rgb && color && npix > 0, RSA_ERR_SHAPE, "label_hsv: bad arguments");
  hsv_kernel<<<rsa_num_sms() * 8, 256, 0, (cudaStream_t)stream>>>(rgb, color, npix);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
