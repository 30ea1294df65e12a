// exp_mma_align.cu — hardware experiment (round 2): does a tcgen05.mma whose K-major SWIZZLE_64B A operand is a SHIFTED window
// of a halo tile (start address not aligned to the 512-byte swizzle atom, stride between 8-row groups = halo row pitch, not a
// multiple of 512 bytes) read shared memory slower than an aligned operand?  conv_tc3's halo mode issues exactly such MMAs and
// sustains ~59 cycles per N=32 MMA where scripts/exp_mma_rate.cu measured 40 for aligned operands.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/bin/exp_mma_align scripts/exp_mma_align.cu -I resunet-a_mltsk_keras_b200/csrc -cudart static
#include "tc_common.cuh"
#include <vector>
void rsa_set_error(const char*, ...) {}

__global__ void __launch_bounds__(128) rate_kernel(int N, int a_shift, int sbo, int layout, int iters, int nissue, int chains, int lanes, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (lanes > 1) {
    // several issuing LANES of one warp (diverged), each with its own accumulator: is the ~100-cycle issue interval a
    // property of the thread or of the warp?
    if (lane < lanes && warp < nissue) {
      const uint32_t idesc = make_idesc(128, N);
      const uint32_t sa = smem_u32(smem) + a_shift, sb = smem_u32(smem) + 144 * 1024;
      const int id = warp * lanes + lane;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < 2; ++k) {
          const uint32_t a0 = sa + (uint32_t)((it & 3) * 64 + id * 512);
          uint64_t ad = 0, bd = 0;
          ad |= (uint64_t)(((a0 + k * 32) >> 4) & 0x3FFF); ad |= (uint64_t)1 << 16; ad |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; ad |= (uint64_t)1 << 46; ad |= (uint64_t)layout << 61;
          bd |= (uint64_t)(((sb + k * 32) >> 4) & 0x3FFF); bd |= (uint64_t)1 << 16; bd |= (uint64_t)((512 >> 4) & 0x3FFF) << 32; bd |= (uint64_t)1 << 46; bd |= (uint64_t)layout << 61;
          umma_bf16(tmem + id * 32, ad, bd, idesc, 1);
        }
      }
      umma_commit(&bar[warp * 0 + (id & 3)]);
      mbar_wait(&bar[id & 3], 0);
      if (lane == 0) out[blockIdx.x * 4 + warp] = clock64() - t0;
    }
  } else if (lane == 0 && warp < nissue) {
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t sa = smem_u32(smem) + a_shift, sb = smem_u32(smem) + 144 * 1024;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < 2; ++k) {               // C = 32: two K = 16 steps per 64-byte row
        for (int c = 0; c < chains; ++c) {
          // operand window c of the tile: 8 pixels further along the row, like the sub-tiles of an item
          const uint32_t a0 = sa + (uint32_t)((it & 3) * 64 + c * 512);
          uint64_t ad = 0, bd = 0;
          ad |= (uint64_t)(((a0 + k * 32) >> 4) & 0x3FFF); ad |= (uint64_t)1 << 16; ad |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; ad |= (uint64_t)1 << 46; ad |= (uint64_t)layout << 61;
          bd |= (uint64_t)(((sb + k * 32) >> 4) & 0x3FFF); bd |= (uint64_t)1 << 16; bd |= (uint64_t)((512 >> 4) & 0x3FFF) << 32; bd |= (uint64_t)1 << 46; bd |= (uint64_t)layout << 61;
          umma_bf16(tmem + (warp * chains + c) * 32, ad, bd, idesc, 1);
        }
      }
    }
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    out[blockIdx.x * 4 + warp] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512)); }
}

int main() {
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  long long* d; cudaMalloc(&d, 148 * 4 * 8);
  std::vector<long long> h(148 * 4);
  const int iters = 2000;
  struct Case { const char* name; int shift, sbo, nissue, chains, lanes; };
  const Case cases[] = {
    {"aligned window, SBO 512 (box mode)       4 threads x 1 chain ", 0, 512, 4, 1, 1},
    {"aligned window, SBO 512                  2 threads x 2 chains", 0, 512, 2, 2, 1},
    {"aligned window, SBO 512                  1 thread  x 4 chains", 0, 512, 1, 4, 1},
    {"aligned window, SBO 2048 (32-px rows)    4 threads x 1 chain ", 0, 2048, 4, 1, 1},
    {"aligned start,  SBO 2176 (34-px halo)    4 threads x 1 chain ", 0, 2176, 4, 1, 1},
    {"start + 64 B,   SBO 2048                 4 threads x 1 chain ", 64, 2048, 4, 1, 1},
    {"start + 64 B,   SBO 2176 (halo, d = 1)   4 threads x 1 chain ", 64, 2176, 4, 1, 1},
    {"start + 192 B,  SBO 2432 (halo, d = 3)   4 threads x 1 chain ", 192, 2432, 4, 1, 1},
    {"start + 64 B,   SBO 2176 (halo, d = 1)   2 threads x 2 chains", 64, 2176, 2, 2, 1},
    {"start + 64 B,   SBO 2176 (halo, d = 1)   1 thread  x 4 chains", 64, 2176, 1, 4, 1},
    {"aligned window, SBO 512      2 warps x 2 issuing lanes x 1 chain", 0, 512, 2, 1, 2},
    {"aligned window, SBO 512      1 warp  x 4 issuing lanes x 1 chain", 0, 512, 1, 1, 4},
  };
  for (const Case& c : cases) {
    cudaMemset(d, 0, 148 * 4 * 8);
    rate_kernel<<<148, 128, 170 * 1024>>>(32, c.shift, c.sbo, 4, iters, c.nissue, c.chains, c.lanes, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int b = 0; b < 148; ++b) for (int w = 0; w < c.nissue; ++w) mx = h[b * 4 + w] > mx ? h[b * 4 + w] : mx;
    const double mmas = (double)iters * 2 * c.nissue * c.chains * c.lanes;
    printf("%s: %6.1f cycles per N=32 MMA per SM (tensor floor 16, aligned operand floor 40)\n", c.name, mx / mmas);
  }
  return 0;
}
