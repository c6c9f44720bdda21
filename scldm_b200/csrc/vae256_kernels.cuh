// Census-scale VAE (n_embed = 256: 8 heads x 32 in the latent Blocks, 4 cross heads x 64 in the MCAB, SwiGLU hidden 684) for sm_100a.
//
// At E = 256 every contraction of the MCAB (layers.py:267-330) is a tensor-core GEMM: the rows are gene tokens (decode: cells x G
// rows, the dominant cost at the census vocabulary) or count tokens (encode: cells x S rows).  The path composes
//   * the latent Blocks (nnets.py:204-205 / :140-141): `dit::dit_blocks_kernel` - an E = 256 Block without adaLN is a DiT block
//     whose "modulation" row is constant (mul = ln.weight - 1, add = ln.bias, gate = 1) and whose biases are zero,
//   * the slab GEMM `trn::gemm_kernel` (tcgen05 / TMEM, bulk-TMA fed) for K/V, Q-side, c_proj and [w1|w2], the last with a
//     SwiGLU . v epilogue: the NB-head Linear(E -> 1) is folded through mlp.c_proj (logit = w.x + (W3^T w).s + b), so the hidden
//     activations and the (rows, E) MLP output never exist in memory,
//   * the kernels below: 16-key cross attention on mma.sync (decode), LN2 + head pre-dot, online-softmax pooling (encode).
#pragma once

#include "train_kernels.cuh"

namespace v256 {

using dit::bf16;
constexpr int E = 256;
constexpr int TOK = 16;          // latent tokens / inducing points
constexpr int LAT = 16;          // latent channels
constexpr int XH = 4;            // cross-attention heads
constexpr int XHD = 64;          // cross head dim
using trn::load8;
using trn::pack8;
using trn::slab_chunk;
using trn::store8;

// decoder_latent_input (nnets.py:170-173): LN over the 16 latent channels (no affine) -> Linear(16 -> 256, no bias)
__global__ void __launch_bounds__(256) lat_in_kernel(const float* __restrict__ z, const float* __restrict__ w /*[256][16]*/, float eps,
                                                     float* __restrict__ X) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  __shared__ float zn[TOK][LAT];
  const int cell = blockIdx.x, tid = threadIdx.x;
  {
    const int tok = tid >> 4, k = tid & 15;
    const float v = z[(size_t)cell * TOK * LAT + tid];
    float s = v;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / LAT);
    const float d = v - mean;
    float q = d * d;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    zn[tok][k] = d * rsqrtf(q * (1.0f / LAT) + eps);
  }
  __syncthreads();
  float wr[LAT];
#pragma unroll
  for (int k = 0; k < LAT; ++k) wr[k] = w[tid * LAT + k];
  for (int tok = 0; tok < TOK; ++tok) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < LAT; ++k) acc += zn[tok][k] * wr[k];
    X[((size_t)cell * TOK + tok) * E + tid] = acc;
  }
}

// fp32 [n][256] -> bf16 [n][256] row-major (the cached Q-side table)
__global__ void __launch_bounds__(256) to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, long long n8) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n8) return;
  float v[8];
  load8(src + i * 8, v);
  *reinterpret_cast<uint4*>(dst + i * 8) = pack8(v);
}

// gather + LN (affine) of the gene-embedding rows of a vocabulary -> slab tensor: the A operand of the Q-side GEMM
// qp = c_attn_q(ln_1q(emb)) (layers.py:253,326), cell invariant because use_adaln = false
__global__ void __launch_bounds__(256) emb_ln_kernel(const float* __restrict__ emb, const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
                                                     int n_ids, bf16* __restrict__ out) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp, c0 = lane * 8;
  float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, w[8], b[8], h[8];
  if (row < n_ids) load8(emb + (size_t)row * E + c0, x);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  const float mean = sm100::warp_sum(s) * (1.0f / E);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] -= mean; sq += x[j] * x[j]; }
  const float rs = rsqrtf(sm100::warp_sum(sq) * (1.0f / E) + eps);
  load8(ln_w + c0, w);
  load8(ln_b + c0, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = row < n_ids ? x[j] * rs * w[j] + b[j] : 0.f;
  *reinterpret_cast<uint4*>(slab_chunk(out, E, row, c0)) = pack8(h);
}

// ------------------------------------------------------------------------------------------
// Decoder MCAB attention (layers.py:229-264 via :321-327): per (cell, gene) softmax over the cell's 16 latent keys, 4 heads x 64.
//   qp   bf16 [n_ids][256]   cached Q-side table          kv fp32 [cells*16][512]  (k | v: "k first", layers.py:252)
//   ao   slab tensor [cells * g_pad][256]: row = cell * g_pad + gene position (positions >= G use gene id 0 and are never read back)
// One CTA = 128 gene rows of one cell, 8 warps x 16 rows; scores / PV on mma.sync.m16n8k16 with the probabilities kept in registers
// (the C fragment of S is the A fragment of P V).
// ------------------------------------------------------------------------------------------
constexpr int KS_LD = E + 8;     // bf16 row stride of the K tile (conflict-free fragment loads)
constexpr int VT_LD = 24;        // bf16 row stride of the transposed V tile [256][16]
__global__ void __launch_bounds__(256) mcab_attn_kernel(const bf16* __restrict__ qp, const long long* __restrict__ genes, int G, int g_pad,
                                                        const float* __restrict__ kv, bf16* __restrict__ ao) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  __shared__ __align__(16) bf16 sK[TOK * KS_LD];
  __shared__ __align__(16) bf16 sVt[E * VT_LD];
  const int cell = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < TOK * E; i += 256) {
    const int key = i >> 8, c = i & 255;
    const float* src = kv + (size_t)(cell * TOK + key) * (2 * E);
    sK[key * KS_LD + c] = __float2bfloat16(src[c]);
    sVt[c * VT_LD + key] = __float2bfloat16(src[E + c]);
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const int pos0 = tile * 128 + warp * 16 + g, pos1 = pos0 + 8;     // gene positions of this thread's two fragment rows
  const long long id0 = pos0 < G ? genes[pos0] : 0, id1 = pos1 < G ? genes[pos1] : 0;
  const uint32_t* q0 = reinterpret_cast<const uint32_t*>(qp + (size_t)id0 * E);
  const uint32_t* q1 = reinterpret_cast<const uint32_t*>(qp + (size_t)id1 * E);
  const size_t row0 = (size_t)cell * g_pad + pos0, row1 = row0 + 8;
  const float scale_log2 = 0.125f * 1.4426950408889634f;     // 1/sqrt(64) * log2(e)
#pragma unroll 1
  for (int h = 0; h < XH; ++h) {
    float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};   // scores vs keys [0,8) and [8,16)
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int w = h * 32 + ks * 8 + t;             // 32-bit word index of columns (h*64 + ks*16 + 2t, +1)
      uint32_t a[4] = {q0[w], q1[w], q0[w + 4], q1[w + 4]};
      const uint32_t* k0 = reinterpret_cast<const uint32_t*>(sK + g * KS_LD) + w;
      const uint32_t* k1 = reinterpret_cast<const uint32_t*>(sK + (8 + g) * KS_LD) + w;
      dit::mma_bf16_16816(s0, a, k0[0], k0[4]);
      dit::mma_bf16_16816(s1, a, k1[0], k1[4]);
    }
    // softmax over the 16 keys of rows g (c0,c1) and g + 8 (c2,c3): the quad holds a full row
    float m0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1])), m1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float p0[4], p1[4];
    p0[0] = sm100::ex2_approx((s0[0] - m0) * scale_log2); p0[1] = sm100::ex2_approx((s0[1] - m0) * scale_log2);
    p1[0] = sm100::ex2_approx((s1[0] - m0) * scale_log2); p1[1] = sm100::ex2_approx((s1[1] - m0) * scale_log2);
    p0[2] = sm100::ex2_approx((s0[2] - m1) * scale_log2); p0[3] = sm100::ex2_approx((s0[3] - m1) * scale_log2);
    p1[2] = sm100::ex2_approx((s1[2] - m1) * scale_log2); p1[3] = sm100::ex2_approx((s1[3] - m1) * scale_log2);
    float l0 = p0[0] + p0[1] + p1[0] + p1[1], l1 = p0[2] + p0[3] + p1[2] + p1[3];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float r0 = 1.0f / l0, r1 = 1.0f / l1;
    // A fragment of P (16 x 16, k = key): {row g keys 2t..; row g+8 keys 2t..; row g keys 8+2t..; row g+8 keys 8+2t..}
    const uint32_t pa[4] = {sm100::pack_bf16x2(p0[0] * r0, p0[1] * r0), sm100::pack_bf16x2(p0[2] * r1, p0[3] * r1),
                            sm100::pack_bf16x2(p1[0] * r0, p1[1] * r0), sm100::pack_bf16x2(p1[2] * r1, p1[3] * r1)};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t* vt = reinterpret_cast<const uint32_t*>(sVt + (h * XHD + nt * 8 + g) * VT_LD);
      dit::mma_bf16_16816(o, pa, vt[t], vt[4 + t]);      // B fragment: V[key 2t, 2t+1][d] and V[key 8+2t, +1][d], d = h*64 + nt*8 + g
      const int col = h * XHD + nt * 8 + 2 * t;
      *reinterpret_cast<uint32_t*>(slab_chunk(ao, E, (int)row0, col & ~7) + (col & 7)) = sm100::pack_bf16x2(o[0], o[1]);
      *reinterpret_cast<uint32_t*>(slab_chunk(ao, E, (int)row1, col & ~7) + (col & 7)) = sm100::pack_bf16x2(o[2], o[3]);
    }
  }
}

// x = q + attn (the residual is the RAW gene embedding, layers.py:327) ; h = LN2(x) (affine) -> slab tensor (A operand of [w1|w2]);
// logit0[row] = head_w . x + head_b: the part of the NB-head Linear (stochastic_layers.py:104-109) that does not pass through the MLP
__global__ void __launch_bounds__(256) mcab_ln2_kernel(const float* __restrict__ y, const float* __restrict__ emb, const long long* __restrict__ genes, int G, int g_pad,
                                                       long long rows, const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
                                                       const float* __restrict__ head_w, float head_b, bf16* __restrict__ h, float* __restrict__ logit) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const int pos = (int)(row % g_pad), c0 = lane * 8;
  const long long id = pos < G ? genes[pos] : 0;
  float x[8], e[8], w[8], b[8], hw[8], o[8];
  load8(y + (size_t)row * E + c0, x);
  load8(emb + (size_t)id * E + c0, e);
  load8(head_w + c0, hw);
  float s = 0.f, dot = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] += e[j]; s += x[j]; dot += hw[j] * x[j]; }
  dot = sm100::warp_sum(dot);
  const float mean = sm100::warp_sum(s) * (1.0f / E);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] -= mean; sq += x[j] * x[j]; }
  const float rs = rsqrtf(sm100::warp_sum(sq) * (1.0f / E) + eps);
  load8(ln_w + c0, w);
  load8(ln_b + c0, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = x[j] * rs * w[j] + b[j];
  *reinterpret_cast<uint4*>(slab_chunk(h, E, (int)row, c0)) = pack8(o);
  if (lane == 0) logit[row] = dot + head_b;
}

// per (cell, 1024-gene block): (max, sum exp(l - max)) of the logits: the partials nb_finalize_kernel merges (softmax over genes)
__global__ void __launch_bounds__(256) logit_partials_kernel(const float* __restrict__ logit, int G, int g_pad, float2* __restrict__ partials) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  __shared__ float rm[8], rs[8];
  const int cell = blockIdx.y, tid = threadIdx.x;
  const float* l = logit + (size_t)cell * g_pad;
  float v[4], m = -INFINITY;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = blockIdx.x * 1024 + i * 256 + tid;
    v[i] = gi < G ? l[gi] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = sm100::warp_max(m);
  if ((tid & 31) == 0) rm[tid >> 5] = m;
  __syncthreads();
  float bm = rm[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) bm = fmaxf(bm, rm[i]);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += (v[i] == -INFINITY) ? 0.f : __expf(v[i] - bm);
  s = sm100::warp_sum(s);
  if ((tid & 31) == 0) rs[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += rs[i];
    partials[(size_t)cell * gridDim.x + blockIdx.x] = make_float2(bm, t);
  }
}

// ------------------------------------------------------------------------------------------
// Encoder side (layers.py:97-118, nnets.py:137-144)
// ------------------------------------------------------------------------------------------
// token rows: x = emb[gene] * log1p(count) (layers.py:28-31) -> LN1 (affine) -> slab tensor [cells * s_pad][256]; rows s >= S are zero
__global__ void __launch_bounds__(256) tok_ln_kernel(const float* __restrict__ emb, const long long* __restrict__ genes, const float* __restrict__ counts, int S,
                                                     int s_pad, long long rows, const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
                                                     bf16* __restrict__ out) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const long long cell = row / s_pad;
  const int s = (int)(row % s_pad), c0 = lane * 8;
  float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, w[8], b[8], h[8];
  const bool live = s < S;
  if (live) {
    const float sc = log1pf(counts[cell * S + s]);
    load8(emb + (size_t)genes[cell * S + s] * E + c0, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= sc;
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) sum += x[j];
  const float mean = sm100::warp_sum(sum) * (1.0f / E);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] -= mean; sq += x[j] * x[j]; }
  const float rs = rsqrtf(sm100::warp_sum(sq) * (1.0f / E) + eps);
  load8(ln_w + c0, w);
  load8(ln_b + c0, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = live ? x[j] * rs * w[j] + b[j] : 0.f;
  *reinterpret_cast<uint4*>(slab_chunk(out, E, (int)row, c0)) = pack8(h);
}

// MCAB pooling: 16 inducing-point queries attend over the S (unmasked: padding tokens included, as the reference) tokens of a cell.
// grid (4 heads, cells), 256 threads = 16 queries x 16 key slices; online softmax per thread, slices merged through shared memory.
//   q_tbl fp32 [16][256] = c_attn_q(ln_1q(inducing_points)) (cell invariant) ; kv fp32 [cells * s_pad][512] (k | v)
constexpr size_t pool_smem_bytes() { return (size_t)TOK * 16 * (XHD + 2) * sizeof(float); }
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ q_tbl, const float* __restrict__ kv, int S, int s_pad, float* __restrict__ out) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  extern __shared__ __align__(16) float sm_pool[];   // [16 q][16 slices][66]: m, l, acc[64]
  const int h = blockIdx.x, cell = blockIdx.y, tid = threadIdx.x;
  const int qi = tid >> 4, sl = tid & 15;
  float q[XHD];
#pragma unroll
  for (int d = 0; d < XHD; d += 4) {
    const float4 t = *reinterpret_cast<const float4*>(q_tbl + qi * E + h * XHD + d);
    q[d] = t.x * 0.125f; q[d + 1] = t.y * 0.125f; q[d + 2] = t.z * 0.125f; q[d + 3] = t.w * 0.125f;   // 1 / sqrt(64)
  }
  float m = -INFINITY, l = 0.f, acc[XHD];
#pragma unroll
  for (int d = 0; d < XHD; ++d) acc[d] = 0.f;
  const float* base = kv + (size_t)cell * s_pad * (2 * E) + h * XHD;
  for (int s = sl; s < S; s += 16) {
    const float* kr = base + (size_t)s * (2 * E);
    float sc = 0.f;
#pragma unroll
    for (int d = 0; d < XHD; d += 4) {
      const float4 k = *reinterpret_cast<const float4*>(kr + d);
      sc += q[d] * k.x + q[d + 1] * k.y + q[d + 2] * k.z + q[d + 3] * k.w;
    }
    const float mn = fmaxf(m, sc);
    const float corr = __expf(m - mn), p = __expf(sc - mn);
    l = l * corr + p;
    const float* vr = kr + E;
#pragma unroll
    for (int d = 0; d < XHD; d += 4) {
      const float4 v = *reinterpret_cast<const float4*>(vr + d);
      acc[d] = acc[d] * corr + p * v.x; acc[d + 1] = acc[d + 1] * corr + p * v.y;
      acc[d + 2] = acc[d + 2] * corr + p * v.z; acc[d + 3] = acc[d + 3] * corr + p * v.w;
    }
    m = mn;
  }
  float* mine = sm_pool + (size_t)(qi * 16 + sl) * (XHD + 2);
  mine[0] = m; mine[1] = l;
#pragma unroll
  for (int d = 0; d < XHD; ++d) mine[2 + d] = acc[d];
  __syncthreads();
  // merge: thread (qi, sl) produces output dims [4 sl, 4 sl + 4) of query qi
  const float* qs = sm_pool + (size_t)qi * 16 * (XHD + 2);
  float gm = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; ++j) gm = fmaxf(gm, qs[j * (XHD + 2)]);
  float gl = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float* sj = qs + j * (XHD + 2);
    const float w = (sj[0] == -INFINITY) ? 0.f : __expf(sj[0] - gm);
    gl += sj[1] * w;
#pragma unroll
    for (int d = 0; d < 4; ++d) o[d] += sj[2 + sl * 4 + d] * w;
  }
  const float inv = 1.0f / gl;
  *reinterpret_cast<float4*>(out + ((size_t)cell * TOK + qi) * E + h * XHD + sl * 4) = make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
}

// out[row][c] = a[row][c] + base[row % 16][c]   (x = inducing_points + attn, layers.py:312-313,327; + pos_embed, nnets.py:139)
__global__ void __launch_bounds__(256) add_rows_kernel(const float* __restrict__ a, const float* __restrict__ base, float* __restrict__ out, long long n4) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const long long row = i >> 6;
  const int c = (int)(i & 63) * 4;
  const float4 x = *reinterpret_cast<const float4*>(a + row * E + c);
  const float4 b = *reinterpret_cast<const float4*>(base + (row & 15) * E + c);
  *reinterpret_cast<float4*>(out + row * E + c) = make_float4(x.x + b.x, x.y + b.y, x.z + b.z, x.w + b.w);
}

// encoder_latent_input (nnets.py:132-135): Linear(256 -> 16, no bias) -> LN over the 16 latent channels (no affine); warp per row
__global__ void __launch_bounds__(256) enc_out_kernel(const float* __restrict__ X, const float* __restrict__ w /*[16][256]*/, float eps, float* __restrict__ z,
                                                      long long rows) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  float x[8];
  load8(X + row * E + lane * 8, x);
  float mine = 0.f;
#pragma unroll
  for (int o = 0; o < LAT; ++o) {
    float wv[8];
    load8(w + (size_t)o * E + lane * 8, wv);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += x[j] * wv[j];
    acc = sm100::warp_sum(acc);
    if (lane == o) mine = acc;
  }
  // lanes 0..15 hold the 16 outputs
  float s = lane < LAT ? mine : 0.f;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / LAT);
  const float d = mine - mean;
  float q = lane < LAT ? d * d : 0.f;
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if (lane < LAT) z[row * LAT + lane] = d * rsqrtf(q * (1.0f / LAT) + eps);
}

}  // namespace v256
