// Microbenchmark: SM -> L2 write bandwidth, all SMs active, each CTA rewriting its own 128 KB tile (L2 resident).
#include <cstdio>
#include <cuda_runtime.h>
#include "../scldm_b200/csrc/sm100.cuh"

// MODE 0: st.global.v4  1: red.global.add.v4.f32  2: bulk store 32 KB pieces  3: bulk reduce-add 1 KB rows
template <int MODE>
__global__ void __launch_bounds__(512, 1) store_kernel(float* dst, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* tile0 = dst + (size_t)blockIdx.x * 4 * 32768;  // 4 x 128 KB per CTA, cycled so that L1 never hits
  for (int i = threadIdx.x; i < 32768; i += 512) reinterpret_cast<float*>(smem)[i] = 1.0f;
  sm100::fence_proxy_async_smem();
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float* tile = tile0 + (it & 3) * 32768;
    if (MODE == 0) {
#pragma unroll
      for (int k = 0; k < 16; ++k) reinterpret_cast<float4*>(tile)[k * 512 + threadIdx.x] = make_float4(1.f, 2.f, 3.f, (float)it);
    } else if (MODE == 1) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float* a = tile + (k * 512 + threadIdx.x) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
      }
    } else if (MODE == 4) {   // RMW, fully coalesced (512 B per warp instruction), all 16 loads in flight
      float4 v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = reinterpret_cast<float4*>(tile)[k * 512 + threadIdx.x];
#pragma unroll
      for (int k = 0; k < 16; ++k) { v[k].x += 1.f; reinterpret_cast<float4*>(tile)[k * 512 + threadIdx.x] = v[k]; }
    } else if (MODE == 5) {   // RMW in the residual-epilogue pattern: warp (q, sub) owns 32 rows x 64 cols; 8 rows x 64 B per instruction
      const int w = threadIdx.x >> 5, l = threadIdx.x & 31, q = w & 3, sub = w >> 2;
      float* xp = tile + (q * 32 + (l >> 2)) * 256 + sub * 64 + (l & 3) * 4;
#pragma unroll
      for (int sc = 0; sc < 4; ++sc) {
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = *reinterpret_cast<float4*>(xp + i * 8 * 256 + sc * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i].x += 1.f; *reinterpret_cast<float4*>(xp + i * 8 * 256 + sc * 16) = v[i]; }
      }
    } else if (MODE == 6) {   // loads only, coalesced (sum kept live)
      float4 v[16]; float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) v[k] = reinterpret_cast<float4*>(tile)[k * 512 + threadIdx.x];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc += v[k].x;
      if (acc == 123.456f) tile[0] = acc;
    } else if (MODE == 2) {
      if (threadIdx.x < 4) {
        sm100::bulk_s2g(tile + threadIdx.x * 8192, smem + threadIdx.x * 32768, 32768);
        sm100::bulk_commit();
        sm100::bulk_wait_read<0>();
      }
      __syncthreads();
    } else {
      if (threadIdx.x < 128) {
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(tile + threadIdx.x * 256),
                     "r"(sm100::smem_u32(smem + threadIdx.x * 1024)), "r"(1024)
                     : "memory");
        sm100::bulk_commit();
        sm100::bulk_wait_read<0>();
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (MODE >= 2 && threadIdx.x < 128) sm100::bulk_wait<0>();
  __threadfence();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

int main() {
  float* dst; long long* d;
  cudaMalloc(&dst, (size_t)148 * 4 * 131072); cudaMemset(dst, 0, (size_t)148 * 4 * 131072); cudaMalloc(&d, 1024 * 8);
  long long h[148];
  const int iters = 200;
  auto run = [&](auto kern, const char* name, int grid) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    kern<<<grid, 512, 131072>>>(dst, iters, d);
    kern<<<grid, 512, 131072>>>(dst, iters, d);
    cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-28s grid %3d: %.1f B/clk/SM (%.0f B/clk total) %s\n", name, grid, (double)iters * 131072 / mx, (double)iters * 131072 / mx * grid,
           cudaGetErrorString(cudaGetLastError()));
  };
  for (int grid : {1, 148}) {
    run(store_kernel<0>, "st.global.v4", grid);
    run(store_kernel<1>, "red.global.add.v4.f32", grid);
    run(store_kernel<2>, "bulk store 4x32KB", grid);
    run(store_kernel<3>, "bulk reduce-add 128x1KB", grid);
    run(store_kernel<4>, "RMW coalesced ld+st", grid);
    run(store_kernel<5>, "RMW 8rows x 64B pattern", grid);
    run(store_kernel<6>, "ld.global.v4 only", grid);
  }
  return 0;
}
