// Probe of the tcgen05.ld / tcgen05.st .16x256b register <-> (TMEM lane, column) mapping against .32x32b.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tmem_shape_probe tools/tmem_shape_probe.cu && tools/tmem_shape_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(float* out, float* out2) {
  __shared__ uint32_t tptr;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((warp * 32u) << 16);
  // write with 32x32b: lane l of quadrant `warp`, columns 0..15: value = (32 warp + l) * 100 + col
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = __float_as_uint((float)((warp * 32 + lane) * 100 + c));
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(base),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
               "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  // read with 16x256b.x2 at lane offsets 0 and 16 of the quadrant
  for (int h = 0; h < 2; ++h) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(base + ((uint32_t)(16 * h) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) out[((warp * 2 + h) * 32 + lane) * 8 + i] = __uint_as_float(r[i]);
    // write back through 16x256b st with +0.5 and read again with 32x32b to check the st mapping is the same
    for (int i = 0; i < 8; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + 0.5f);
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base + 16 + ((uint32_t)(16 * h) << 16)),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                 "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(base + 16) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 16; ++c) out2[(warp * 32 + lane) * 16 + c] = __uint_as_float(v[c]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tptr) : "memory");
}
int main() {
  float *d, *d2;
  cudaMalloc(&d, 4 * 2 * 32 * 8 * 4); cudaMalloc(&d2, 128 * 16 * 4);
  probe<<<1, 128>>>(d, d2);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  static float h[4 * 2 * 32 * 8], h2[128 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(h2, d2, sizeof(h2), cudaMemcpyDeviceToHost);
  int bad = 0, bad2 = 0;
  for (int w = 0; w < 4; ++w) for (int hh = 0; hh < 2; ++hh) for (int l = 0; l < 32; ++l) for (int i = 0; i < 8; ++i) {
    const int g = l >> 2, t = l & 3, n = i >> 2, rr = (i >> 1) & 1, e = i & 1;
    const float expect = (float)((w * 32 + hh * 16 + g + 8 * rr) * 100 + 8 * n + 2 * t + e);
    const float got = h[((w * 2 + hh) * 32 + l) * 8 + i];
    if (got != expect) { if (bad < 12) printf("ld mismatch w%d h%d lane%d reg%d: got %.1f expected %.1f\n", w, hh, l, i, got, expect); ++bad; }
  }
  for (int r = 0; r < 128; ++r) for (int c = 0; c < 16; ++c) {
    const float expect = (float)(r * 100 + c) + 0.5f;
    if (h2[r * 16 + c] != expect) { if (bad2 < 12) printf("st mismatch row%d col%d: got %.1f expected %.1f\n", r, c, h2[r * 16 + c], expect); ++bad2; }
  }
  printf("16x256b ld mapping: %s (%d mismatches); st round trip: %s (%d)\n", bad ? "DIFFERENT" : "as assumed", bad, bad2 ? "DIFFERENT" : "ok", bad2);
  if (bad) { printf("lane 0..7 regs of w0 h0:\n"); for (int l = 0; l < 8; ++l) { for (int i = 0; i < 8; ++i) printf(" %6.0f", h[l * 8 + i]); printf("\n"); } }
  return 0;
}
