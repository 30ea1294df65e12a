// postproc.cu — Amazon deforestation evaluation post-processing.
//
// The reference evaluates a reconstructed prediction map with skimage.morphology.area_opening(img, area_threshold,
// connectivity=1) — on a binary map: drop every 4-connected component of ones smaller than the threshold — followed by mask
// arithmetic and a confusion matrix over the pixels that remain under consideration (utils.py:505-548,
// utils2.py:312-356).  Here the component filter is a union-find labelling on the GPU (one atomicMin-based merge pass over
// the right/down neighbours, one flatten + size-count pass, one filter pass) and the mask pipeline + histogram is one
// fused pass; scenes are tens of megapixels, the work is HBM/atomic bound and takes milliseconds.
#include "common.cuh"

namespace {

constexpr int PT = 256;

__device__ __forceinline__ int uf_find(const int* __restrict__ parent, int x) {
  int p = parent[x];
  while (p != x) { x = p; p = parent[x]; }
  return x;
}
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&parent[b], a);        // hook the larger root under the smaller one
    if (old == b) return;
    b = old;                                         // somebody re-rooted b in the meantime: retry from there
  }
}

__global__ void __launch_bounds__(PT) cc_init_kernel(const uint8_t* __restrict__ img, int* __restrict__ parent, int* __restrict__ size, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) {
    parent[i] = img[i] ? (int)i : -1;
    size[i] = 0;
  }
}
__global__ void __launch_bounds__(PT) cc_merge_kernel(const uint8_t* __restrict__ img, int* __restrict__ parent, int H, int W) {
  const int64_t n = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) {
    if (!img[i]) continue;
    const int x = (int)(i % W), y = (int)(i / W);
    if (x + 1 < W && img[i + 1]) uf_union(parent, (int)i, (int)i + 1);
    if (y + 1 < H && img[i + W]) uf_union(parent, (int)i, (int)i + W);
  }
}
__global__ void __launch_bounds__(PT) cc_flatten_count_kernel(int* __restrict__ parent, int* __restrict__ size, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) {
    if (parent[i] < 0) continue;
    const int r = uf_find(parent, (int)i);
    parent[i] = r;                                   // benign race: every writer stores the component's unique root
    atomicAdd(&size[r], 1);
  }
}
__global__ void __launch_bounds__(PT) cc_filter_kernel(const int* __restrict__ parent, const int* __restrict__ size, uint8_t* __restrict__ out,
                                                       int64_t n, int area_threshold) {
  for (int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) {
    const int p = parent[i];
    out[i] = (p >= 0 && size[uf_find(parent, p)] >= area_threshold) ? 1 : 0;
  }
}

// utils.py:527-545: which pixels count and with which (reference, prediction) value; cm[3][3] over the counted pixels
__global__ void __launch_bounds__(PT) amazon_consider_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ opened,
                                                             const uint8_t* __restrict__ ref_clip, const uint8_t* __restrict__ clip_mask,
                                                             uint8_t* __restrict__ ref_consider, uint8_t* __restrict__ pred_consider,
                                                             uint8_t* __restrict__ selected, int64_t n, unsigned long long* __restrict__ cm) {
  __shared__ unsigned int bins[9];
  if (threadIdx.x < 9) bins[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * PT + threadIdx.x; i < n; i += (int64_t)gridDim.x * PT) {
    const int p = pred[i], o = opened[i], r = ref_clip[i];
    const int mask_areas = (p - o == 1) ? 0 : 1;     // prediction pixels of components below the area threshold
    const int mask_borders = (r == 2) ? 0 : 1;       // past deforestation is not evaluated
    const int m = mask_areas * mask_borders;
    const int rv = m * r, pv = m * p;
    const int sel = (clip_mask[i] * m == 1) ? 1 : 0;
    if (ref_consider) ref_consider[i] = (uint8_t)rv;
    if (pred_consider) pred_consider[i] = (uint8_t)pv;
    if (selected) selected[i] = (uint8_t)sel;
    if (sel && rv < 3 && pv < 3) atomicAdd(&bins[rv * 3 + pv], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 9 && bins[threadIdx.x]) atomicAdd(cm + threadIdx.x, (unsigned long long)bins[threadIdx.x]);
}

inline int grid_for_n(int64_t n) {
  int64_t b = (n + PT - 1) / PT;
  const int64_t cap = (int64_t)rsa_num_sms() * 16;
  return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}

}  // namespace

/* out = area_opening(img, area_threshold, connectivity=1) for a binary uint8 [H,W] map: ones that belong to a 4-connected
 * component of at least area_threshold pixels (skimage.morphology.area_opening at utils.py:531, utils2.py:323,400).
 * workspace: 2*H*W int32.  H*W < 2^31. */
extern "C" int rsa_area_opening_binary(const uint8_t* img, uint8_t* out, int H, int W, int area_threshold, void* workspace, void* stream) {
  RSA_REQUIRE(img && out && workspace && H > 0 && W > 0 && (int64_t)H * W < 2147483647LL, RSA_ERR_SHAPE, "area_opening_binary: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)H * W;
  int* parent = (int*)workspace;
  int* size = parent + n;
  const int g = grid_for_n(n);
  cc_init_kernel<<<g, PT, 0, st>>>(img, parent, size, n);
  cc_merge_kernel<<<g, PT, 0, st>>>(img, parent, H, W);
  cc_flatten_count_kernel<<<g, PT, 0, st>>>(parent, size, n);
  cc_filter_kernel<<<g, PT, 0, st>>>(parent, size, out, n, area_threshold);
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}

/* The mask pipeline of utils.py:527-545 on uint8 maps of n pixels: pred (0/1 reconstruction), opened (its area opening),
 * ref_clip (0 / 1 deforestation / 2 past deforestation), clip_mask (1 = inside the evaluated tiles).  Optional outputs:
 * ref_consider, pred_consider, selected (the boolean index of utils.py:544-545); cm (int64[9], zeroed by the caller) +=
 * confusion counts [reference][prediction] over the selected pixels. */
extern "C" int rsa_amazon_consider(const uint8_t* pred, const uint8_t* opened, const uint8_t* ref_clip, const uint8_t* clip_mask,
                                   uint8_t* ref_consider, uint8_t* pred_consider, uint8_t* selected, int64_t n, int64_t* cm, void* stream) {
  RSA_REQUIRE(pred && opened && ref_clip && clip_mask && cm && n > 0, RSA_ERR_SHAPE, "amazon_consider: bad arguments");
  amazon_consider_kernel<<<grid_for_n(n), PT, 0, (cudaStream_t)stream>>>(pred, opened, ref_clip, clip_mask, ref_consider, pred_consider,
                                                                      selected, n, reinterpret_cast<unsigned long long*>(cm));
  RSA_CHECK_LAUNCH();
  return RSA_OK;
}
