// Microbenchmark: L2 -> SM bulk-TMA bandwidth when every SM streams the SAME L2-resident buffer (weights pattern).
#include <cstdio>
#include <cuda_runtime.h>
#include "../scldm_b200/csrc/sm100.cuh"

template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(128, 1) stream_kernel(const uint8_t* src, size_t src_bytes, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) sm100::mbar_init(&full[i], 1);
    sm100::fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    const size_t nchunks = src_bytes / CHUNK;
    int issued = 0;
    for (; issued < STAGES && issued < iters; ++issued) {
      sm100::mbar_arrive_expect_tx(&full[issued], CHUNK);
      sm100::bulk_g2s(smem + issued * CHUNK, src + ((size_t)(issued + blockIdx.x * 7) % nchunks) * CHUNK, CHUNK, &full[issued]);
    }
    for (int i = 0; i < iters; ++i) {
      const int st = i % STAGES;
      sm100::mbar_wait(&full[st], (i / STAGES) & 1);
      if (issued < iters) {
        sm100::mbar_arrive_expect_tx(&full[st], CHUNK);
        sm100::bulk_g2s(smem + st * CHUNK, src + ((size_t)(issued + blockIdx.x * 7) % nchunks) * CHUNK, CHUNK, &full[st]);
        ++issued;
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  uint8_t* src; long long* d;
  const size_t bytes = 2u << 20;  // 2 MB of "weights"
  cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes); cudaMalloc(&d, 1024 * 8);
  long long h[148];
  const int iters = 2000;
  auto run = [&](auto kern, int stages, int chunk, int grid) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * chunk);
    kern<<<grid, 128, stages * chunk>>>(src, bytes, iters, d);
    kern<<<grid, 128, stages * chunk>>>(src, bytes, iters, d);
    cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("grid %3d stages %d chunk %5d: %.1f B/clk/SM (%.0f B/clk total) %s\n", grid, stages, chunk, (double)iters * chunk / mx,
           (double)iters * chunk / mx * grid, cudaGetErrorString(cudaGetLastError()));
  };
  for (int grid : {1, 37, 74, 148}) {
    run(stream_kernel<3, 32768>, 3, 32768, grid);
    run(stream_kernel<6, 32768>, 6, 32768, grid);
    run(stream_kernel<4, 16384>, 4, 16384, grid);
  }
  return 0;
}
