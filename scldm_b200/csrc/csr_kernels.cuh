// Device-side sparsification of the generated count matrix.
// The reference copies the dense (rows, G) fp32 counts to the host and builds scipy.sparse.csr_matrix per batch
// (src/scldm/_utils.py:186-200); counts are > 90 % zeros, so building the CSR arrays on the GPU cuts the D2H
// transfer roughly 10x.  Output is exactly what scipy.sparse.csr_matrix(dense) holds: indptr (rows+1), column indices
// ascending within a row, data = the non-zero values (fp32).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace csr {

constexpr int THREADS = 256;

// nnz per row: one CTA per row, coalesced 4-byte loads (rows are only 4-byte aligned in general: G need not be even)
__global__ void __launch_bounds__(THREADS) count_kernel(const float* __restrict__ dense, int G, int* __restrict__ row_nnz) {
  __shared__ int s_part[THREADS / 32];
  const float* r = dense + (size_t)blockIdx.x * G;
  int c = 0;
  for (int g = threadIdx.x; g < G; g += THREADS) c += r[g] != 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) t += s_part[i];
    row_nnz[blockIdx.x] = t;
  }
}

// indptr[0] = 0, indptr[i+1] = sum_{j<=i} row_nnz[j]: one CTA, rows walked in chunks of 1024 with a carried total
__global__ void __launch_bounds__(1024) scan_kernel(const int* __restrict__ row_nnz, int rows, long long* __restrict__ indptr) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  if (threadIdx.x == 0) { s_carry = 0; indptr[0] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < rows; base += 1024) {
    const int i = base + threadIdx.x;
    long long v = i < rows ? (long long)row_nnz[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
      long long w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const long long incl = v + (warp > 0 ? s_warp[warp - 1] : 0) + s_carry;
    if (i < rows) indptr[i + 1] = incl;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = incl;
    __syncthreads();
  }
}

// ordered compaction of one row per CTA: chunks of 256 genes, ballot + warp prefix + block prefix keep ascending order
__global__ void __launch_bounds__(THREADS) fill_kernel(const float* __restrict__ dense, int G, const long long* __restrict__ indptr,
                                                       int* __restrict__ indices, float* __restrict__ data) {
  __shared__ int s_warp[THREADS / 32];
  const float* r = dense + (size_t)blockIdx.x * G;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long out = indptr[blockIdx.x];
  for (int base = 0; base < G; base += THREADS) {
    const int g = base + threadIdx.x;
    const float v = g < G ? r[g] : 0.0f;
    const bool nz = v != 0.0f;
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    const int before = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int warp_off = 0, total = 0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) {
      const int c = s_warp[i];
      if (i < warp) warp_off += c;
      total += c;
    }
    if (nz) {
      const long long o = out + warp_off + before;
      indices[o] = g;
      data[o] = v;
    }
    out += total;
    __syncthreads();
  }
}

// "expressed" tokenizer (reference: tokenize_cells(sample_genes="expressed"), src/scldm/datamodule.py:708-731): per cell the
// expressed genes (count > 0) packed left in gene order into (rows, S) token / count arrays, padded with the mask token and
// zeros; library size = row sum.  One CTA per cell; a cell with more than S expressed genes sets *overflow (the reference
// raises) and is truncated.
__global__ void __launch_bounds__(THREADS) tokenize_expressed_kernel(const float* __restrict__ dense, int G, const long long* __restrict__ gene_ids,
                                                                    int S, long long mask_idx, long long* __restrict__ genes_out,
                                                                    float* __restrict__ counts_out, float* __restrict__ library,
                                                                    int* __restrict__ overflow) {
  __shared__ int s_warp[THREADS / 32];
  __shared__ float s_sum[THREADS / 32];
  const float* r = dense + (size_t)blockIdx.x * G;
  long long* go = genes_out + (size_t)blockIdx.x * S;
  float* co = counts_out + (size_t)blockIdx.x * S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int out = 0;
  float sum = 0.0f;
  for (int base = 0; base < G; base += THREADS) {
    const int g = base + threadIdx.x;
    const float v = g < G ? r[g] : 0.0f;
    sum += v;
    const bool nz = v > 0.0f;
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    const int before = __popc(m & ((1u << lane) - 1u));
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int warp_off = 0, total = 0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) {
      const int c = s_warp[i];
      if (i < warp) warp_off += c;
      total += c;
    }
    if (nz) {
      const int o = out + warp_off + before;
      if (o < S) { go[o] = gene_ids[g]; co[o] = v; }
    }
    out += total;
    __syncthreads();
  }
  for (int o = min(out, S) + threadIdx.x; o < S; o += THREADS) { go[o] = mask_idx; co[o] = 0.0f; }
  // library size: the row sum in a fixed (lane-strided, then tree) order
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) s_sum[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) t += s_sum[i];
    library[blockIdx.x] = t;
    if (out > S) atomicMax(overflow, out);
  }
}

}  // namespace csr
