// exp_desc.cu — hardware experiment: which UMMA shared-memory descriptor forms address a *shifted* window of a
// TMA-swizzled tile correctly?  Decides whether a halo tile can feed all nine taps of a 3x3 convolution.
//   A: R rows x KC bf16 (row pitch = KC*2 bytes = swizzle span), loaded by one/two TMA boxes with SWIZZLE_{128,64,32}B.
//   MMA row m reads smem row  shift + (m/8)*G + (m%8)   (SBO = G*pitch), K = KC.
//   bo mode 0: base_offset field 0; mode 1: base_offset = (start_address >> 7) & 7.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o exp_desc scripts/exp_desc.cu -I resunet-a_mltsk_keras_b200/csrc
#include "tc_common.cuh"
#include <vector>
#include <cstdlib>
#include <cmath>

void rsa_set_error(const char*, ...) {}

__global__ void __launch_bounds__(128) exp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                  int KC, int R, int shift, int G, int bo_mode, float* D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int pitch = KC * 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 256 * 128;      // after the largest A (256 rows x 128 B)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32 * 128);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(32));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, R * pitch + 32 * pitch);
    for (int r0 = 0; r0 < R; r0 += 128)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(sA + r0 * pitch)), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(bar)), "r"(0), "r"(r0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sB)), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA) + shift * pitch;
    uint64_t adesc = 0;
    adesc |= (uint64_t)((a0 >> 4) & 0x3FFF);
    adesc |= (uint64_t)1 << 16;
    adesc |= (uint64_t)(((G * pitch) >> 4) & 0x3FFF) << 32;
    adesc |= (uint64_t)1 << 46;
    if (bo_mode == 1) adesc |= (uint64_t)((a0 >> 7) & 7) << 49;
    adesc |= (uint64_t)(pitch == 128 ? 2 : (pitch == 64 ? 4 : 6)) << 61;
    const uint64_t bdesc = make_kmajor_desc_any(smem_u32(sB), pitch);
    const uint32_t idesc = make_idesc(128, 32);
    for (int k = 0; k < KC / 16; ++k) umma_bf16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k > 0);
    umma_commit(mbar);
  }
  mbar_wait(mbar, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32)); }
}

static float bf(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
  EncodeTiledFn enc = get_encode();
  if (!enc) { printf("no encode fn\n"); return 1; }
  const int R = 256;
  cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int KC : {64, 32, 16}) {
    const int pitch = KC * 2;
    std::vector<uint16_t> hA(R * KC), hB(32 * KC);
    srand(1);
    // small integers: exactly representable, products and sums exact in fp32
    auto rnd = []() { float f = (float)(rand() % 17 - 8); uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); };
    for (auto& x : hA) x = rnd();
    for (auto& x : hB) x = rnd();
    uint16_t *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * 32 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    const CUtensorMapSwizzle swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUtensorMap tmA, tmB;
    {
      cuuint64_t gdim[2] = {(cuuint64_t)KC, (cuuint64_t)R}; cuuint64_t gstr[1] = {(cuuint64_t)pitch};
      cuuint32_t box[2] = {(cuuint32_t)KC, 128}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      cuuint64_t gdimb[2] = {(cuuint64_t)KC, 32}; cuuint32_t boxb[2] = {(cuuint32_t)KC, 32};
      CUresult r2 = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, gdimb, gstr, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r || r2) { printf("encode failed %d %d\n", (int)r, (int)r2); return 1; }
    }
    std::vector<float> hD(128 * 32);
    for (int G : {8, 10, 14}) {
      for (int bo = 0; bo < 2; ++bo) {
        printf("KC=%d pitch=%dB G=%d bo_mode=%d :", KC, pitch, G, bo);
        for (int shift = 0; shift <= 20; ++shift) {
          cudaMemset(dD, 0, 128 * 32 * 4);
          exp_kernel<<<1, 128, 64 * 1024>>>(tmA, tmB, KC, R, shift, G, bo, dD);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf(" [shift %d: %s]\n", shift, cudaGetErrorString(e)); return 2; }
          cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
          double maxerr = 0;
          for (int m = 0; m < 128; ++m) {
            const int row = shift + (m / 8) * G + (m % 8);
            for (int n = 0; n < 32; ++n) {
              double acc = 0;
              for (int k = 0; k < KC; ++k) acc += (double)bf(hA[row * KC + k]) * bf(hB[n * KC + k]);
              maxerr = fmax(maxerr, fabs(acc - hD[m * 32 + n]));
            }
          }
          printf(" %d:%s", shift, maxerr < 1e-3 ? "ok" : "BAD");
        }
        printf("\n");
      }
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
