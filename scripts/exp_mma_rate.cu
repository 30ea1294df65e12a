// exp_mma_rate.cu — hardware experiment: sustained tcgen05.mma (kind::f16, M=128, K=16, SS mode) issue rate as a function
// of N and of the shared-memory operand layout (K-major SWIZZLE_128B / 64B / 32B rows).  Answers: is a C=32 layer
// (64-byte rows, SWIZZLE_64B) paying bank conflicts on its half-row K slices, and would 32-byte planes (SWIZZLE_32B) fix it?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/bin/exp_mma_rate scripts/exp_mma_rate.cu -I resunet-a_mltsk_keras_b200/csrc -cudart static
#include "tc_common.cuh"
#include <vector>
void rsa_set_error(const char*, ...) {}

// pitch: bytes per operand row (= swizzle span); layout code 2/4/6; ksteps per row = pitch/32; nissue MMAs cycling over rows blocks
__global__ void __launch_bounds__(128) rate_kernel(int N, int pitch, int layout, int iters, int nthreads_issue, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (lane == 0 && warp < nthreads_issue) {
    const uint32_t idesc = make_idesc(128, N);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 64 * 1024;
    const int ks = pitch / 32;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // A: 128 rows x pitch; B: N rows x pitch; walk the k-steps of the row, a different A block every iteration
      const uint32_t a0 = sa + (uint32_t)((it & 3) * 128 * pitch);
      for (int k = 0; k < ks; ++k) {
        uint64_t ad = 0, bd = 0;
        ad |= (uint64_t)(((a0 + k * 32) >> 4) & 0x3FFF); ad |= (uint64_t)1 << 16; ad |= (uint64_t)(((8 * pitch) >> 4) & 0x3FFF) << 32; ad |= (uint64_t)1 << 46; ad |= (uint64_t)layout << 61;
        bd |= (uint64_t)(((sb + k * 32) >> 4) & 0x3FFF); bd |= (uint64_t)1 << 16; bd |= (uint64_t)(((8 * pitch) >> 4) & 0x3FFF) << 32; bd |= (uint64_t)1 << 46; bd |= (uint64_t)layout << 61;
        umma_bf16(tmem + warp * 128, ad, bd, idesc, 1);
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
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long* d; cudaMalloc(&d, 148 * 4 * 8);
  std::vector<long long> h(148 * 4);
  for (int nthr : {1, 4}) for (int N : {32, 64, 128}) for (int pitch : {128, 64, 32}) {
    const int layout = pitch == 128 ? 2 : (pitch == 64 ? 4 : 6);
    const int iters = 2000;
    cudaMemset(d, 0, 148 * 4 * 8);
    rate_kernel<<<148, 128, 100 * 1024>>>(N, pitch, layout, iters, nthr, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int b = 0; b < 148; ++b) for (int w = 0; w < nthr; ++w) mx = h[b * 4 + w] > mx ? h[b * 4 + w] : mx;
    const double mmas = (double)iters * (pitch / 32) * nthr;
    printf("issuers=%d N=%3d row pitch %3dB (SW%-3d): %6.1f cycles per MMA per SM  (tensor floor %d)\n", nthr, N, pitch, pitch, mx / mmas, N / 2);
  }
  return 0;
}
