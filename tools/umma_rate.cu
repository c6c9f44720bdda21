// Microbenchmark: back-to-back tcgen05.mma throughput, cta_group::1 (M=128) vs cta_group::2 (M=256, B split over the CTA
// pair), bf16 K-major SW128 operands resident in shared memory.  Prints cycles per MMA (N=256, K=16).
#include <cstdio>
#include <cuda_runtime.h>
#include "../scldm_b200/csrc/sm100.cuh"

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int n, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < (64 + 32) * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  sm100::fence_proxy_async_smem();
  if (threadIdx.x == 0) { sm100::mbar_init(&bar, 1); sm100::fence_barrier_init(); }
  if (threadIdx.x < 32) { if (PAIR) sm100::tmem_alloc2(&tptr, 512); else sm100::tmem_alloc(&tptr, 512); }
  sm100::tc_fence_before();
  if (PAIR) sm100::cluster_sync_all(); else __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tm = tptr;
  const uint32_t rank = PAIR ? sm100::cluster_ctarank() : 0;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = sm100::make_idesc_bf16(PAIR ? 256 : 128, n);
    const uint64_t a = sm100::make_kmajor_sw128_desc(sm100::smem_u32(smem));
    const uint64_t b = sm100::make_kmajor_sw128_desc(sm100::smem_u32(smem + 64 * 1024));
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (PAIR) sm100::umma_bf16_ss2(tm, a + 2ull * k, b + 2ull * k, idesc, 1u);
        else sm100::umma_bf16_ss(tm, a + 2ull * k, b + 2ull * k, idesc, 1u);
      }
    }
    if (PAIR) sm100::umma_commit2(&bar); else sm100::umma_commit(&bar);
  }
  if (threadIdx.x == 0) {
    sm100::mbar_wait(&bar, 0);
    t1 = clock64();
    if (rank == 0) out[blockIdx.x] = t1 - t0;
  }
  sm100::tc_fence_before();
  if (PAIR) sm100::cluster_sync_all(); else __syncthreads();
  if (threadIdx.x < 32) { if (PAIR) sm100::tmem_dealloc2(tm, 512); else sm100::tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 256 * 8);
  long long h[256];
  const int iters = 2000;
  for (int n : {128, 192, 256}) {
    for (int pair = 0; pair < 2; ++pair) {
      cudaMemset(d, 0, 256 * 8);
      const size_t smem = 96 * 1024;
      if (pair) {
        cudaFuncSetAttribute(rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, rate_kernel<true>, iters, n, d);
      } else {
        cudaFuncSetAttribute(rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        rate_kernel<false><<<148, 128, smem>>>(iters, n, d);
      }
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("N=%d cta_group::%d : %.1f cycles per MMA (M=%d x N x K16) -> %.0f flop/clk/SM  [%s]\n", n, pair + 1, (double)mx / (iters * 4), pair ? 256 : 128,
             2.0 * 128 * n * 16 * (iters * 4) / mx, cudaGetErrorString(e));
    }
  }
  return 0;
}
