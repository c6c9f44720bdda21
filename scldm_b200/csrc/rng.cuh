// Counter-based RNG (Philox4x32-10) and the samplers the generation path needs:
// standard normal (Box-Muller), Gamma (Marsaglia-Tsang), Poisson (multiplication method for
// small rates, Hormann's PTRS transformed rejection for large rates).
//
// Streams are keyed by (seed, global cell index, element index) so that results do not depend
// on how cells are sharded over GPUs or chunked inside one GPU.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace rng {

struct Philox {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t out[4];
  int have;

  __device__ __forceinline__ Philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2) {
    key[0] = (uint32_t)seed;
    key[1] = (uint32_t)(seed >> 32);
    ctr[0] = c0; ctr[1] = c1; ctr[2] = c2; ctr[3] = 0;
    have = 0;
  }

  __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }

  __device__ __forceinline__ void refill() {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    ++ctr[3];
    have = 4;
  }

  __device__ __forceinline__ uint32_t next() {
    if (have == 0) refill();
    return out[--have];
  }
  // uniform in (0, 1]
  __device__ __forceinline__ float uniform() { return ((float)(next() >> 8) + 1.0f) * (1.0f / 16777216.0f); }
  __device__ __forceinline__ float normal() {
    const float u1 = uniform(), u2 = uniform();
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
  }
};

// Gamma(shape=alpha, scale=1)
__device__ __forceinline__ float gamma_mt(Philox& g, float alpha) {
  const bool boost = alpha < 1.0f;
  const float a = boost ? alpha + 1.0f : alpha;
  const float d = a - (1.0f / 3.0f);
  const float c = rsqrtf(9.0f * d);
  float v, x;
  for (int it = 0; it < 64; ++it) {
    x = g.normal();
    v = 1.0f + c * x;
    if (v <= 0.0f) continue;
    v = v * v * v;
    const float u = g.uniform();
    const float x2 = x * x;
    if (u < 1.0f - 0.0331f * x2 * x2) break;
    if (logf(u) < 0.5f * x2 + d * (1.0f - v + logf(v))) break;
  }
  float r = d * v;
  if (boost) r *= powf(g.uniform(), 1.0f / alpha);
  return r;
}

// Poisson(lam)
__device__ __forceinline__ float poisson(Philox& g, float lam) {
  if (!(lam > 0.0f)) return 0.0f;
  if (lam < 10.0f) {
    const float enlam = expf(-lam);
    float prod = g.uniform();
    int k = 0;
    while (prod > enlam && k < 200) {
      prod *= g.uniform();
      ++k;
    }
    return (float)k;
  }
  // PTRS (W. Hormann, 1993)
  const float slam = sqrtf(lam), loglam = logf(lam);
  const float b = 0.931f + 2.53f * slam;
  const float a = -0.059f + 0.02483f * b;
  const float invalpha = 1.1239f + 1.1328f / (b - 3.4f);
  const float vr = 0.9277f - 3.6224f / (b - 2.0f);
  for (int it = 0; it < 256; ++it) {
    const float U = g.uniform() - 0.5f;
    const float V = g.uniform();
    const float us = 0.5f - fabsf(U);
    const float k = floorf((2.0f * a / us + b) * U + lam + 0.43f);
    if (us >= 0.07f && V <= vr) return k;
    if (k < 0.0f || (us < 0.013f && V > us)) continue;
    if (logf(V) + logf(invalpha) - logf(a / (us * us) + b) <= -lam + k * loglam - lgammaf(k + 1.0f)) return k;
  }
  return floorf(lam);
}

// NB(mu, theta) as Poisson(Gamma(theta, rate = theta/mu)), gamma draw clamped to 1e8 (scvi semantics).
// Small means (the bulk of a count matrix: mean count per gene << 1) take the Gamma-Poisson MIXTURE's own law instead of
// the two-stage draw: pmf(0) = (theta/(theta+mu))^theta, pmf(k+1) = pmf(k) (k+theta)/(k+1) mu/(theta+mu), inverted with one
// uniform (same distribution, one Philox block instead of two or more and no rejection loops).
constexpr float NB_INVERSION_MAX_MU = 4.0f;
// inversion of the NB(mu, theta) CDF at u in (0, 1] (the small-mean path of `negative_binomial`; exposed for the u = 1.0 unit test)
__device__ __forceinline__ float nb_invert(float u, float mu, float theta) {
  const float q = mu / (theta + mu);
  float p = __expf(-theta * log1pf(mu / theta));     // pmf(0)
  float c = p;
  int k = 0;
  // The fp32 CDF saturates just below 1 (e.g. 0.99999994 for mu = 0.01, theta = 1) while u may be exactly 1.0: once a term no
  // longer changes c the tail is exhausted in fp32, and the scan must stop THERE - running on to the loop cap would turn
  // u in (c_final, 1] (probability ~1e-7 per draw, several draws per 4736 x 17002 step) into spurious counts of 256.
  while (u > c && k < 256) {
    p *= ((float)k + theta) / (float)(k + 1) * q;
    const float c_next = c + p;
    if (c_next == c) break;
    c = c_next;
    ++k;
  }
  return (float)k;
}
__device__ __forceinline__ float negative_binomial(Philox& g, float mu, float theta) {
  if (mu < NB_INVERSION_MAX_MU && theta > 1e-3f) {
    if (!(mu > 0.0f)) return 0.0f;
    return nb_invert(g.uniform(), mu, theta);          // u in (0, 1]
  }
  const float lam = fminf(gamma_mt(g, theta) * (mu / theta), 1e8f);
  return poisson(g, lam);
}

}  // namespace rng
