// Generation-evaluation and SDE-sampler kernels.
//   * pairwise statistics of two count / feature matrices for the MMD kernels of src/scldm/evaluations.py:10-82 (RBF, Bray-Curtis,
//     Tanimoto, Ruzicka) and the cost matrix of `wasserstein` (:85-108): one pass computes, per pair (i, j), sum x*y, sum |x - y|,
//     sum |x + y| and sum min(x, y); every kernel value is an O(1) function of those and of the row sums
//   * Sinkhorn-Knopp scaling iterations (POT's `sinkhorn2`, the solver `wasserstein(method="sinkhorn")` calls)
//   * the Euler-Maruyama / Heun steps of the SiT SDE sampler (src/scldm/transport/integrators.py:7-75, transport.py:226-322)
#pragma once

#include "rng.cuh"
#include "sm100.cuh"

namespace evk {

// out[q][i][j] for q in {dot, l1, abs_sum, min_sum}; x [nx][D], y [ny][D] row-major fp32.  64 x 64 pairs per CTA, 4 x 4 per thread.
constexpr int PT = 64, PK = 32;
__global__ void __launch_bounds__(256) pair_stats_kernel(const float* __restrict__ x, int nx, const float* __restrict__ y, int ny, int D, float* __restrict__ out) {
  __shared__ float sx[PK][PT + 1], sy[PK][PT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i0 = blockIdx.y * PT, j0 = blockIdx.x * PT;
  float dot[4][4] = {}, l1[4][4] = {}, as[4][4] = {}, mn[4][4] = {};
  for (int k0 = 0; k0 < D; k0 += PK) {
    for (int e = threadIdx.x; e < PT * PK; e += 256) {
      const int r = e / PK, k = e % PK;
      sx[k][r] = (i0 + r < nx && k0 + k < D) ? x[(size_t)(i0 + r) * D + k0 + k] : 0.f;
      sy[k][r] = (j0 + r < ny && k0 + k < D) ? y[(size_t)(j0 + r) * D + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < PK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = sx[k][ty * 4 + u]; b[u] = sy[k][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          dot[u][v] += a[u] * b[v];
          l1[u][v] += fabsf(a[u] - b[v]);
          as[u][v] += fabsf(a[u] + b[v]);
          mn[u][v] += fminf(a[u], b[v]);
        }
    }
    __syncthreads();
  }
  const size_t plane = (size_t)nx * ny;
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < nx && j < ny) {
        const size_t o = (size_t)i * ny + j;
        out[o] = dot[u][v]; out[plane + o] = l1[u][v]; out[2 * plane + o] = as[u][v]; out[3 * plane + o] = mn[u][v];
      }
    }
}

// Sinkhorn-Knopp (POT `sinkhorn_knopp`): v = b / (K^T u) ; u = a / (K v).  K [n][m] row-major.
__global__ void __launch_bounds__(256) sinkhorn_ktu_kernel(const float* __restrict__ K, const float* __restrict__ u, const float* __restrict__ b, int n, int m,
                                                           float* __restrict__ v) {   // one thread per column j: coalesced over j
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= m) return;
  float acc = 0.f;
  for (int i = 0; i < n; ++i) acc += K[(size_t)i * m + j] * u[i];
  v[j] = b[j] / acc;
}
__global__ void __launch_bounds__(256) sinkhorn_kv_kernel(const float* __restrict__ K, const float* __restrict__ v, const float* __restrict__ a, int n, int m,
                                                          float* __restrict__ u) {    // one warp per row i
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= n) return;
  float acc = 0.f;
  for (int j = lane; j < m; j += 32) acc += K[(size_t)i * m + j] * v[j];
  acc = sm100::warp_sum(acc);
  if (lane == 0) u[i] = a[i] / acc;
}
// res[0] = sum_ij u_i K_ij v_j M_ij (transport cost) ; res[1] = sum_j | v_j sum_i u_i K_ij - b_j | (marginal violation, the stopping criterion)
__global__ void __launch_bounds__(256) sinkhorn_eval_kernel(const float* __restrict__ K, const float* __restrict__ M, const float* __restrict__ u, const float* __restrict__ v,
                                                            const float* __restrict__ b, int n, int m, float* __restrict__ res) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  float cost = 0.f, err = 0.f;
  if (j < m) {
    float col = 0.f;
    for (int i = 0; i < n; ++i) {
      const float k = K[(size_t)i * m + j] * u[i];
      col += k;
      cost += k * M[(size_t)i * m + j];
    }
    cost *= v[j];
    err = fabsf(col * v[j] - b[j]);
  }
  cost = sm100::warp_sum(cost);
  err = sm100::warp_sum(err);
  if ((threadIdx.x & 31) == 0) { atomicAdd(res, cost); atomicAdd(res + 1, err); }
}

// ---- SDE sampler pieces, Linear path (alpha = t, sigma = 1 - t: path.py:24-50) with a velocity model ----
//   score = (t v - x) / (1 - t)                           (get_score_from_velocity, path.py:79-95: var = (1-t)^2 + t (1-t) = 1 - t)
//   diffusion D(t): "constant" norm | "SBDM" norm (1-t)/t | "sigma" norm (1-t) | "linear" norm (1-t) | "decreasing" 0.25 (norm cos(pi t) + 1)^2 |
//                   "inccreasing-decreasing" norm sin^2(pi t)                        (compute_diffusion, path.py:52-77)
//   sde drift = v + D score                               (transport.py:231-233)
__device__ __forceinline__ float diffusion_of(float t, int form, float norm) {
  switch (form) {
    case 0: return norm;
    case 1: return norm * (1.0f - t) / t;
    case 2: case 3: return norm * (1.0f - t);
    case 4: { const float c = norm * cospif(t) + 1.0f; return 0.25f * c * c; }
    default: { const float s = sinpif(t); return norm * s * s; }
  }
}
// drift[i] = v + D(t) (t v - x) / (1 - t)
__global__ void __launch_bounds__(256) sde_drift_kernel(const float* __restrict__ v, const float* __restrict__ x, float t, int form, float norm, float* __restrict__ drift,
                                                        long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float d = diffusion_of(t, form, norm);
  drift[i] = v[i] + d * (t * v[i] - x[i]) / (1.0f - t);
}
// xhat = x + sqrt(2 D(t)) sqrt(dt) w ; w = `noise` when given, else Philox N(0,1) keyed by (seed, global cell, element, step)
__global__ void __launch_bounds__(256) sde_kick_kernel(const float* __restrict__ x, const float* __restrict__ noise, float t, float dt, int form, float norm,
                                                       unsigned long long seed, long long cell_offset, int per_cell, unsigned int step, float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float w;
  if (noise != nullptr) w = noise[i];
  else {
    const long long cell = cell_offset + i / per_cell;
    rng::Philox g(seed, (uint32_t)(i % per_cell), (uint32_t)cell, ((uint32_t)(cell >> 32) ^ 0x5DE00000u) + step);
    w = g.normal();
  }
  out[i] = x[i] + sqrtf(2.0f * diffusion_of(t, form, norm)) * sqrtf(dt) * w;
}
// out = a + c1 * d1 (+ c2 * d2)
__global__ void __launch_bounds__(256) axpy2_kernel(const float* __restrict__ a, float c1, const float* __restrict__ d1, float c2, const float* __restrict__ d2,
                                                    float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float r = a[i] + c1 * d1[i];
  if (d2 != nullptr) r += c2 * d2[i];
  out[i] = r;
}

}  // namespace evk
