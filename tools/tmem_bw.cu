// Microbenchmark: tcgen05.ld throughput per SM vs number of reading warps / vector width.
#include <cstdio>
#include <cuda_runtime.h>
#include "../scldm_b200/csrc/sm100.cuh"

template <int WIDTH>
__global__ void tmem_ld_bw(long long* out, int iters, float* sink) {
  __shared__ uint32_t tptr;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) sm100::tmem_alloc(&tptr, 512);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t base = tptr + (((warp & 3) * 32u) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < 512; c += 128) {
      if constexpr (WIDTH == 32) {
        uint32_t v[32];
        sm100::tmem_ld_32x32b_x32(base + ((c + (warp >> 2) * 32) & 511), v);
        sm100::tmem_ld_wait();
        acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
      } else {
        uint32_t v[16];
        sm100::tmem_ld_32x32b_x16(base + ((c + (warp >> 2) * 16) & 511), v);
        sm100::tmem_ld_wait();
        acc += __uint_as_float(v[0]) + __uint_as_float(v[15]);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  __syncthreads();
  if (warp == 0) sm100::tmem_dealloc(tptr, 512);
}

// MUFU / FMA mixed throughput of candidate SiLU formulations (per SM, 16 warps)
template <int MODE>
__global__ void silu_bw(long long* out, int iters, float* sink, float seed) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = seed + threadIdx.x * 1e-3f + j;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = x[j];
      if (MODE == 0) v = __fdividef(v, 1.0f + __expf(-v));
      else if (MODE == 1) { float h = 0.5f * v, th; asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h)); v = h * th + h; }
      x[j] = v + 1e-3f;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0;
  for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 123.456f) sink[0] = s;
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 1024 * 8); cudaMalloc(&sink, 4);
  long long h[4];
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    tmem_ld_bw<32><<<1, warps * 32>>>(d, iters, sink);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    double bytes = (double)iters * 4 * warps * 32 * 32 * 4;
    printf("tcgen05.ld x32: %2d warps  %8lld cycles  %.1f B/clk/SM\n", warps, h[0], bytes / h[0]);
    tmem_ld_bw<16><<<1, warps * 32>>>(d, iters, sink);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    bytes = (double)iters * 4 * warps * 32 * 16 * 4;
    printf("tcgen05.ld x16: %2d warps  %8lld cycles  %.1f B/clk/SM\n", warps, h[0], bytes / h[0]);
  }
  for (int mode = 0; mode < 2; ++mode) {
    if (mode == 0) silu_bw<0><<<1, 512>>>(d, iters, sink, 0.5f); else silu_bw<1><<<1, 512>>>(d, iters, sink, 0.5f);
    cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
    printf("silu mode %d (0=exp+rcp, 1=tanh.approx): %lld cycles for %d elems/thread x 512 thr => %.2f elem/clk/SM\n", mode, h[0],
           iters * 8, (double)iters * 8 * 512 / h[0]);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
