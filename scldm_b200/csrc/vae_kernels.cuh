// Transformer-VAE decoder kernels (scLDM generation hot path), fp32 CUDA-core first version.
//
//   dec_latent_multi_kernel   per cell: LN16 -> Linear(16->32) -> n_layer Blocks (E=32, 8 heads of 4) ->
//                       K,V = c_attn(LN1(x)) of the MCAB (16 keys x 64)      nnets.py:203-205, layers.py:252
//   qside_kernel        per gene id: Qp = c_attn_q(LN1q(emb[g]))  -- cell-invariant when use_adaln=false,
//                       computed once per vocabulary                            layers.py:253,326
//   mcab_decode_kernel  per (cell, gene): 4-head attention over the 16 latent keys, c_proj, x = q + attn,
//                       LN2, SwiGLU MLP, residual, NB-head logit; never materialises (G,E) activations
//                       or the (cells,4,G,16) score tensor                       layers.py:325-330
//   nb_finalize_kernel  softmax over genes * library size -> mu; theta = exp(table[g]); optional
//                       Gamma-Poisson draw                                      stochastic_layers.py:102-116
#pragma once

#include "rng.cuh"
#include "sm100.cuh"

namespace vae {

constexpr int E = 32;        // n_embed
constexpr int LAT = 16;      // n_embed_latent
constexpr int TOK = 16;      // latent tokens
constexpr int NH = 8;        // self-attention heads (head_dim 4)
constexpr int NHC = 4;       // cross-attention heads (head_dim 8)
constexpr int HID = 88;      // SwiGLU hidden
constexpr int KV = 2 * E;    // 64

// ---- packed fp32 weights of one non-adaLN Block (all matrices transposed to [in][out]) ----
struct BlockW {
  const float* ln1_w; const float* ln1_b; const float* ln2_w; const float* ln2_b;
  const float* wqkv_t;   // [32][96]
  const float* wproj_t;  // [32][32]
  const float* w1_t;     // [32][88]
  const float* w2_t;     // [32][88]
  const float* w3_t;     // [88][32]
};

struct DecLatentParams {
  const float* z;        // [cells][16][16]
  const float* win_t;    // decoder_latent_input.1.weight^T [16][32]
  const float* blocks;   // n_layer x BLOCK_STRIDE floats (see pack.py)
  int n_layer;
  const float* ca_ln1_w; const float* ca_ln1_b;
  const float* ca_wkv_t; // [32][64]  (k | v)
  float eps;
  float* kv;             // [cells][16][64] fp32 (k | v), or nullptr
  __nv_bfloat16* kvb;    // [cells][ K: 16x32 | V^T: 32x16 ] bf16 for the tensor-core MCAB kernel, or nullptr
};

constexpr int BLK_LN1W = 0, BLK_LN1B = 32, BLK_LN2W = 64, BLK_LN2B = 96, BLK_WQKV = 128, BLK_WPROJ = BLK_WQKV + 32 * 96,
              BLK_W1 = BLK_WPROJ + 32 * 32, BLK_W2 = BLK_W1 + 32 * HID, BLK_W3 = BLK_W2 + 32 * HID,
              BLOCK_STRIDE = BLK_W3 + HID * 32;

// LayerNorm over 32 channels for 16 tokens; 128 threads: thread = (token = tid/8, part = tid%8 -> 4 channels)
__device__ __forceinline__ void ln32_tokens(const float (*x)[E], float (*y)[E], const float* w, const float* b, float eps, int tid) {
  const int tok = tid >> 3, part = tid & 7;
  float v[4];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[j] = x[tok][part * 4 + j]; s += v[j]; }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float mean = s * (1.0f / E);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) { v[j] -= mean; ss += v[j] * v[j]; }
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);
  const float rstd = rsqrtf(ss * (1.0f / E) + eps);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = part * 4 + j;
    y[tok][c] = v[j] * rstd * (w ? w[c] : 1.f) + (b ? b[c] : 0.f);
  }
}

// One non-adaLN Block (layers.py:222-226) on a 16 x 32 tile in shared memory; nt = threads working on this tile (>= 128), tid
// their index; `w` = the block's packed weights (global or shared memory).  Every barrier is CTA-wide: all tiles of a CTA run
// in lock-step.
// out[4 tokens][NC columns] += h[4 tokens][K] x w[K][ldw] for one thread: the four tokens' activations come as float4 along k
// (broadcast reads: the token group is warp-uniform), the weights as one scalar per (k, column): 4 + 4 NC shared-memory reads per
// 16 NC FMAs (the plain one-output-per-thread loop issued two reads per FMA and was bound by that).  K % 4 == 0; rows 16-byte aligned.
template <int NC>
__device__ __forceinline__ void tile_gemm4(const float* h0, int ldh, int K, const float* w, int ldw, const int (&col)[NC], float (&acc)[4][NC]) {
  for (int k = 0; k < K; k += 4) {
    float4 hv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) hv[t] = *reinterpret_cast<const float4*>(h0 + t * ldh + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float wv[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) wv[c] = w[(k + kk) * ldw + col[c]];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float hk = kk == 0 ? hv[t].x : (kk == 1 ? hv[t].y : (kk == 2 ? hv[t].z : hv[t].w));
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[t][c] = fmaf(hk, wv[c], acc[t][c]);
      }
    }
  }
}

__device__ __forceinline__ void vae_block_layer(float (*x)[E], float (*h)[E], float (*qkv)[3 * E], float (*hid)[HID], const float* w, float eps,
                                                int tid, int nt) {
  (void)nt;   // the first 128 threads of the tile's group do the work: thread = (token group tg = 4 tokens, column lane cl)
  const int tg = (tid >> 5) & 3, cl = tid & 31, t0 = tg * 4;
  {
    if (tid < 128) ln32_tokens(x, h, w + BLK_LN1W, w + BLK_LN1B, eps, tid);
    __syncthreads();
    if (tid < 128) {   // q | k | v = h x Wqkv: columns cl, cl + 32, cl + 64
      float acc[4][3] = {};
      const int col[3] = {cl, cl + 32, cl + 64};
      tile_gemm4<3>(&h[t0][0], E, E, w + BLK_WQKV, 3 * E, col, acc);
#pragma unroll
      for (int t = 0; t < 4; ++t) { qkv[t0 + t][cl] = acc[t][0]; qkv[t0 + t][cl + 32] = acc[t][1]; qkv[t0 + t][cl + 64] = acc[t][2]; }
    }
    __syncthreads();
    if (tid < 128) {
      // one thread per (query token, head): head_dim 4 = one float4, softmax over 16 keys, scale 1/sqrt(4)
      const int tok = tid >> 3, hd = tid & 7;
      const float4 q = *reinterpret_cast<const float4*>(&qkv[tok][hd * 4]);
      float s[TOK];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < TOK; ++k) {
        const float4 kk = *reinterpret_cast<const float4*>(&qkv[k][E + hd * 4]);
        s[k] = (q.x * kk.x + q.y * kk.y + q.z * kk.z + q.w * kk.w) * 0.5f;
        mx = fmaxf(mx, s[k]);
      }
      float den = 0.f;
#pragma unroll
      for (int k = 0; k < TOK; ++k) { s[k] = __expf(s[k] - mx); den += s[k]; }
      const float inv = 1.0f / den;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < TOK; ++k) {
        const float4 vv = *reinterpret_cast<const float4*>(&qkv[k][2 * E + hd * 4]);
        o.x += s[k] * vv.x; o.y += s[k] * vv.y; o.z += s[k] * vv.z; o.w += s[k] * vv.w;
      }
      *reinterpret_cast<float4*>(&h[tok][hd * 4]) = make_float4(o.x * inv, o.y * inv, o.z * inv, o.w * inv);
    }
    __syncthreads();
    if (tid < 128) {   // x += attn x Wproj
      float acc[4][1] = {};
      const int col[1] = {cl};
      tile_gemm4<1>(&h[t0][0], E, E, w + BLK_WPROJ, E, col, acc);
#pragma unroll
      for (int t = 0; t < 4; ++t) x[t0 + t][cl] += acc[t][0];
    }
    __syncthreads();
    if (tid < 128) ln32_tokens(x, h, w + BLK_LN2W, w + BLK_LN2B, eps, tid);
    __syncthreads();
    if (tid < 128) {   // SwiGLU hidden units cl, cl + 32, cl + 64 (HID = 88: the third only for cl < 24)
      float a[4][3] = {}, b[4][3] = {};
      const int col[3] = {cl, cl + 32, cl + 64 < HID ? cl + 64 : cl};   // (an out-of-range column re-reads column cl; its result is dropped)
      tile_gemm4<3>(&h[t0][0], E, E, w + BLK_W1, HID, col, a);
      tile_gemm4<3>(&h[t0][0], E, E, w + BLK_W2, HID, col, b);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        hid[t0 + t][cl] = sm100::silu(a[t][0]) * b[t][0];
        hid[t0 + t][cl + 32] = sm100::silu(a[t][1]) * b[t][1];
        if (cl + 64 < HID) hid[t0 + t][cl + 64] = sm100::silu(a[t][2]) * b[t][2];
      }
    }
    __syncthreads();
    if (tid < 128) {   // x += hid x W3
      float acc[4][1] = {};
      const int col[1] = {cl};
      tile_gemm4<1>(&hid[t0][0], HID, HID, w + BLK_W3, E, col, acc);
#pragma unroll
      for (int t = 0; t < 4; ++t) x[t0 + t][cl] += acc[t][0];
    }
    __syncthreads();
  }
}

// n_layer Blocks with the weights read from global memory (one tile per CTA)
__device__ __forceinline__ void vae_block_stack(float (*x)[E], float (*h)[E], float (*qkv)[3 * E], float (*hid)[HID],
                                                const float* blocks, int n_layer, float eps, int tid, int nt) {
  for (int l = 0; l < n_layer; ++l) vae_block_layer(x, h, qkv, hid, blocks + (size_t)l * BLOCK_STRIDE, eps, tid, nt);
}

// Decoder front (nnets.py:200-205): LN16 -> Linear(16 -> 32) -> n_layer Blocks -> K/V of the MCAB, DL_CELLS cells per CTA (one
// 128-thread group each) with every layer's weights staged in shared memory once per CTA.  (A one-cell-per-CTA version re-read 50 KB
// of weights per layer and cell through L2 and was bound by that latency: 0.44 ms for 1184 cells.)
constexpr int DL_CELLS = 4;
constexpr int DL_ACT_FLOATS = TOK * (E + E + 3 * E + HID);   // x | h | qkv | hid of one cell
constexpr size_t dec_latent_multi_smem_bytes() { return (size_t)(BLOCK_STRIDE + DL_CELLS * DL_ACT_FLOATS) * sizeof(float); }

__global__ void __launch_bounds__(128 * DL_CELLS) dec_latent_multi_kernel(const DecLatentParams p, int n_cells) {
  extern __shared__ __align__(16) float dl_smem[];
  float* sw = dl_smem;                                   // one Block's weights
  const int grp = threadIdx.x >> 7, tid = threadIdx.x & 127;
  float* act = dl_smem + BLOCK_STRIDE + grp * DL_ACT_FLOATS;
  float (*x)[E] = reinterpret_cast<float (*)[E]>(act);
  float (*h)[E] = reinterpret_cast<float (*)[E]>(act + TOK * E);
  float (*qkv)[3 * E] = reinterpret_cast<float (*)[3 * E]>(act + 2 * TOK * E);
  float (*hid)[HID] = reinterpret_cast<float (*)[HID]>(act + 5 * TOK * E);
  const int cell = min(blockIdx.x * DL_CELLS + grp, n_cells - 1);   // a surplus group recomputes the last cell (no divergent barriers)
  const bool live = blockIdx.x * DL_CELLS + grp < n_cells;
  {
    float (*zs)[LAT] = reinterpret_cast<float (*)[LAT]>(&qkv[0][0]);
    for (int i = tid; i < TOK * LAT; i += 128) zs[i / LAT][i % LAT] = p.z[(size_t)cell * TOK * LAT + i];
    __syncthreads();
    if (tid < TOK) {
      float m = 0.f;
      for (int j = 0; j < LAT; ++j) m += zs[tid][j];
      m *= (1.0f / LAT);
      float var = 0.f;
      for (int j = 0; j < LAT; ++j) { const float dlt = zs[tid][j] - m; var += dlt * dlt; }
      const float rstd = rsqrtf(var * (1.0f / LAT) + p.eps);
      for (int j = 0; j < LAT; ++j) zs[tid][j] = (zs[tid][j] - m) * rstd;
    }
    __syncthreads();
    for (int i = tid; i < TOK * E; i += 128) {
      const int tok = i / E, c = i % E;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < LAT; ++k) acc += zs[tok][k] * p.win_t[k * E + c];
      x[tok][c] = acc;
    }
  }
  for (int l = 0; l < p.n_layer; ++l) {
    __syncthreads();   // the previous layer's readers of `sw` are done (and x is complete)
    const float4* src = reinterpret_cast<const float4*>(p.blocks + (size_t)l * BLOCK_STRIDE);
    for (int i = threadIdx.x; i < BLOCK_STRIDE / 4; i += 128 * DL_CELLS) reinterpret_cast<float4*>(sw)[i] = src[i];
    __syncthreads();
    vae_block_layer(x, h, qkv, hid, sw, p.eps, tid, 128);
  }
  ln32_tokens(x, h, p.ca_ln1_w, p.ca_ln1_b, p.eps, tid);
  __syncthreads();
  if (!live) return;
  for (int i = tid; i < TOK * KV; i += 128) {
    const int tok = i / KV, j = i % KV;
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < E; ++k) acc += h[tok][k] * p.ca_wkv_t[k * KV + j];
    if (p.kv) p.kv[(size_t)cell * TOK * KV + i] = acc;
    if (p.kvb) {
      if (j < E) p.kvb[(size_t)cell * 1024 + tok * E + j] = __float2bfloat16(acc);
      else p.kvb[(size_t)cell * 1024 + 512 + (j - E) * TOK + tok] = __float2bfloat16(acc);
    }
  }
}
static_assert(BLOCK_STRIDE % 4 == 0, "float4 staging of a Block's weights");

// Qp[g] = Wq * LN1q(emb[g]) for every vocabulary id (incl. the mask id 0)
__global__ void __launch_bounds__(128) qside_kernel(const float* __restrict__ emb, const float* __restrict__ ln_w,
                                                     const float* __restrict__ ln_b, const float* __restrict__ wq /*[out][in]*/,
                                                     float eps, int n_ids, float* __restrict__ qp,
                                                     __nv_bfloat16* __restrict__ qp_bf16) {
  __shared__ float s_wq[E * E];
  for (int i = threadIdx.x; i < E * E; i += 128) s_wq[i] = wq[i];
  __syncthreads();
  const int g = blockIdx.x * 128 + threadIdx.x;
  if (g >= n_ids) return;
  float v[E];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < E; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(emb + (size_t)g * E + c);
    v[c] = t.x; v[c + 1] = t.y; v[c + 2] = t.z; v[c + 3] = t.w;
    s += t.x + t.y + t.z + t.w;
  }
  const float mean = s * (1.0f / E);
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < E; ++c) { v[c] -= mean; ss += v[c] * v[c]; }
  const float rstd = rsqrtf(ss * (1.0f / E) + eps);
#pragma unroll
  for (int c = 0; c < E; ++c) v[c] = v[c] * rstd * ln_w[c] + ln_b[c];
#pragma unroll 4
  for (int j = 0; j < E; j += 4) {
    float o[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < E; ++c) acc += v[c] * s_wq[(j + jj) * E + c];
      o[jj] = acc;
    }
    if (qp) *reinterpret_cast<float4*>(qp + (size_t)g * E + j) = make_float4(o[0], o[1], o[2], o[3]);
    if (qp_bf16) {
      uint2 pk;
      pk.x = sm100::pack_bf16x2(o[0], o[1]);
      pk.y = sm100::pack_bf16x2(o[2], o[3]);
      *reinterpret_cast<uint2*>(qp_bf16 + (size_t)g * E + j) = pk;
    }
  }
}

// ---- MCAB decode + NB-head logit ---------------------------------------------------------
struct McabParams {
  const float* emb;       // gene embedding table [n_ids][32]
  const float* qp;        // Q-side table [n_ids][32]
  const long long* genes; // [G] vocabulary ids shared by all cells (datamodule.py:689 tiles one row)
  int G;
  const float* kv;        // [cells][16][64]
  int n_cells;
  int cells_per_block;
  // smem-resident weights, one contiguous fp32 blob:
  //   wproj [32][32] (out,in) | ln2_w[32] | ln2_b[32] | w1 [88][32] | w2 [88][32] | w3_t [88][32] | head_w[32] | head_b |
  //   pad[3] | theta-head w[32] | theta-head b | pad[3]     (second output channel of an unshared-theta head, else zeros)
  const float* wblob;
  float eps;
  float* log_theta;       // [cells][G] log theta of an unshared-theta head (stochastic_layers.py:111-113), or nullptr
  float* logits;          // [cells][G]
  float2* partials;       // [cells][gene_tiles] (max, sum exp(l - max))
  int gene_tiles;
};
constexpr int MW_PROJ = 0, MW_LN2W = 1024, MW_LN2B = 1056, MW_W1 = 1088, MW_W2 = MW_W1 + HID * E, MW_W3T = MW_W2 + HID * E,
              MW_HW = MW_W3T + HID * E, MW_HB = MW_HW + E, MW_HW2 = MW_HB + 4, MW_HB2 = MW_HW2 + E, MW_TOTAL = MW_HB2 + 4;

__global__ void __launch_bounds__(128) mcab_decode_kernel(const McabParams p) {
  extern __shared__ __align__(16) float smf[];
  float* sw = smf;                      // MW_TOTAL
  float* skv = smf + MW_TOTAL;          // [16][64]
  __shared__ float red_m[4], red_s[4];
  const int tid = threadIdx.x;
  for (int i = tid; i < MW_TOTAL; i += 128) sw[i] = p.wblob[i];
  const int gi = blockIdx.x * 128 + tid;
  const bool valid = gi < p.G;
  float emb[E], qv[E];
  {
    const long long gid = valid ? p.genes[gi] : 0;
#pragma unroll
    for (int c = 0; c < E; c += 4) {
      const float4 a = *reinterpret_cast<const float4*>(p.emb + (size_t)gid * E + c);
      const float4 b = *reinterpret_cast<const float4*>(p.qp + (size_t)gid * E + c);
      emb[c] = a.x; emb[c + 1] = a.y; emb[c + 2] = a.z; emb[c + 3] = a.w;
      qv[c] = b.x; qv[c + 1] = b.y; qv[c + 2] = b.z; qv[c + 3] = b.w;
    }
  }
  const int cell0 = blockIdx.y * p.cells_per_block;
  const int cell1 = min(cell0 + p.cells_per_block, p.n_cells);
  const float sc = 0.35355339059327373f * 1.4426950408889634f;  // 1/sqrt(8) * log2(e)
  for (int cell = cell0; cell < cell1; ++cell) {
    __syncthreads();
    for (int i = tid; i < TOK * KV / 4; i += 128)
      reinterpret_cast<float4*>(skv)[i] = reinterpret_cast<const float4*>(p.kv + (size_t)cell * TOK * KV)[i];
    __syncthreads();
    // ---- 4-head cross attention over the 16 latent keys (head_dim 8) ----
    float o[E];
#pragma unroll
    for (int hh = 0; hh < NHC; ++hh) {
      float s[TOK];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < TOK; ++k) {
        const float4 k0 = *reinterpret_cast<const float4*>(skv + k * KV + hh * 8);
        const float4 k1 = *reinterpret_cast<const float4*>(skv + k * KV + hh * 8 + 4);
        float a = qv[hh * 8 + 0] * k0.x + qv[hh * 8 + 1] * k0.y + qv[hh * 8 + 2] * k0.z + qv[hh * 8 + 3] * k0.w +
                  qv[hh * 8 + 4] * k1.x + qv[hh * 8 + 5] * k1.y + qv[hh * 8 + 6] * k1.z + qv[hh * 8 + 7] * k1.w;
        s[k] = a;
        mx = fmaxf(mx, a);
      }
      float den = 0.f;
#pragma unroll
      for (int k = 0; k < TOK; ++k) { s[k] = exp2f((s[k] - mx) * sc); den += s[k]; }
      const float inv = 1.0f / den;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < TOK; ++k) {
        const float4 v0 = *reinterpret_cast<const float4*>(skv + k * KV + E + hh * 8);
        const float4 v1 = *reinterpret_cast<const float4*>(skv + k * KV + E + hh * 8 + 4);
        acc[0] += s[k] * v0.x; acc[1] += s[k] * v0.y; acc[2] += s[k] * v0.z; acc[3] += s[k] * v0.w;
        acc[4] += s[k] * v1.x; acc[5] += s[k] * v1.y; acc[6] += s[k] * v1.z; acc[7] += s[k] * v1.w;
      }
#pragma unroll
      for (int d = 0; d < 8; ++d) o[hh * 8 + d] = acc[d] * inv;
    }
    // ---- x = q + c_proj(attn) ----
    float x[E];
#pragma unroll
    for (int c = 0; c < E; ++c) {
      float acc = emb[c];
#pragma unroll
      for (int j = 0; j < E; j += 4) {
        const float4 w = *reinterpret_cast<const float4*>(sw + MW_PROJ + c * E + j);
        acc += o[j] * w.x + o[j + 1] * w.y + o[j + 2] * w.z + o[j + 3] * w.w;
      }
      x[c] = acc;
    }
    // ---- LN2 ----
    float hl[E];
    {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < E; ++c) s += x[c];
      const float mean = s * (1.0f / E);
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < E; ++c) { hl[c] = x[c] - mean; ss += hl[c] * hl[c]; }
      const float rstd = rsqrtf(ss * (1.0f / E) + p.eps);
#pragma unroll
      for (int c = 0; c < E; ++c) hl[c] = hl[c] * rstd * sw[MW_LN2W + c] + sw[MW_LN2B + c];
    }
    // ---- SwiGLU MLP, hidden unit by hidden unit (rank-1 updates of x) ----
#pragma unroll 2
    for (int j = 0; j < HID; ++j) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int c = 0; c < E; c += 4) {
        const float4 w1 = *reinterpret_cast<const float4*>(sw + MW_W1 + j * E + c);
        const float4 w2 = *reinterpret_cast<const float4*>(sw + MW_W2 + j * E + c);
        a += hl[c] * w1.x + hl[c + 1] * w1.y + hl[c + 2] * w1.z + hl[c + 3] * w1.w;
        b += hl[c] * w2.x + hl[c + 1] * w2.y + hl[c + 2] * w2.z + hl[c + 3] * w2.w;
      }
      const float hv = sm100::silu(a) * b;
#pragma unroll
      for (int c = 0; c < E; c += 4) {
        const float4 w3 = *reinterpret_cast<const float4*>(sw + MW_W3T + j * E + c);
        x[c] += hv * w3.x; x[c + 1] += hv * w3.y; x[c + 2] += hv * w3.z; x[c + 3] += hv * w3.w;
      }
    }
    // ---- NB head logit + per-tile softmax partials ----
    float logit = sw[MW_HB];
#pragma unroll
    for (int c = 0; c < E; ++c) logit += x[c] * sw[MW_HW + c];
    if (valid) p.logits[(size_t)cell * p.G + gi] = logit;
    if (p.log_theta != nullptr) {   // unshared theta: second output channel of the head
      float lt = sw[MW_HB2];
#pragma unroll
      for (int c = 0; c < E; ++c) lt += x[c] * sw[MW_HW2 + c];
      if (valid) p.log_theta[(size_t)cell * p.G + gi] = lt;
    }
    const float lm = valid ? logit : -INFINITY;
    const float wm = sm100::warp_max(lm);
    if ((tid & 31) == 0) red_m[tid >> 5] = wm;
    __syncthreads();
    const float bm = fmaxf(fmaxf(red_m[0], red_m[1]), fmaxf(red_m[2], red_m[3]));
    const float ev = valid ? __expf(logit - bm) : 0.f;
    const float ws = sm100::warp_sum(ev);
    if ((tid & 31) == 0) red_s[tid >> 5] = ws;
    __syncthreads();
    if (tid == 0) p.partials[(size_t)cell * p.gene_tiles + blockIdx.x] = make_float2(bm, red_s[0] + red_s[1] + red_s[2] + red_s[3]);
  }
}

// ---- MCAB decode on tensor cores (mma.sync bf16, fp32 accumulate) -------------------------------
// One warp owns 16 genes (one m16 tile) and keeps the whole per-gene chain in registers:
//   S_h = Q_h K_h^T (m16n8k8 x2 per head) -> softmax over 16 keys (quad shuffles) -> O_h = P_h V_h (m16n8k16)
//   -> x = emb + O Wproj^T (8 mma) -> LN2 -> for each 16-wide hidden chunk: [w1|w2] (8 mma) -> silu*mul ->
//   x += h W3^T (4 mma) -> logit = x . w_head + b.   C fragments feed the next GEMM's A fragments directly.
// Weights arrive pre-arranged in B-fragment order (pack.py: mma_b_frags), 64 x u32 per (k-step, n-tile).
struct McabTcParams {
  const float* emb;            // [n_ids][32] fp32 (residual uses the unrounded embedding)
  const __nv_bfloat16* qp;     // [n_ids][32] bf16 Q-side table
  const long long* genes;
  int G;
  const __nv_bfloat16* kvb;    // [cells][ K: 16 keys x 32 | V^T: 32 dims x 16 keys ] bf16
  int n_cells;
  int cells_per_block;
  const uint32_t* wfrag;       // 80 fragments x 64 u32: proj (2x4) | per hidden chunk c<6: w1 (2x2) | w2 (2x2) | w3 (4)
  const float* small;          // ln2_w[32] | ln2_b[32] | head_w[32] | head_b | pad[3] | theta-head w[32] | theta-head b | pad[3]
  float eps;
  float* log_theta;            // [cells][G] (unshared-theta head) or nullptr
  float* logits;
  float2* partials;
  int gene_tiles;
};
constexpr int TC_NFRAG = 8 + 6 * 12;

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movmatrix_trans_b16(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ void mma_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS) mcab_decode_tc_kernel(const McabTcParams p) {
  __shared__ uint2 s_frag[TC_NFRAG * 32];   // 20 KB
  __shared__ float s_small[136];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < TC_NFRAG * 32; i += 256) s_frag[i] = reinterpret_cast<const uint2*>(p.wfrag)[i];
  if (tid < 136) s_small[tid] = p.small[tid];
  __syncthreads();

  // ---- gene-side operands of this warp's 16 genes (cell invariant) ----
  const int gene0 = blockIdx.x * 128 + warp * 16;
  const int gi0 = gene0 + g, gi1 = gene0 + g + 8;
  const bool v0 = gi0 < p.G, v1 = gi1 < p.G;
  const long long id0 = v0 ? p.genes[gi0] : 0, id1 = v1 ? p.genes[gi1] : 0;
  float embf[4][4];   // C-fragment layout: [nt]{(row g: cols 8nt+2t,+1), (row g+8: ...)}
  uint32_t qa[4][2];  // per head h: a0 = (row g, dims 8h+2t,+1), a1 = (row g+8, ...)
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 e0 = *reinterpret_cast<const float2*>(p.emb + (size_t)id0 * E + nt * 8 + 2 * t);
    const float2 e1 = *reinterpret_cast<const float2*>(p.emb + (size_t)id1 * E + nt * 8 + 2 * t);
    embf[nt][0] = e0.x; embf[nt][1] = e0.y; embf[nt][2] = e1.x; embf[nt][3] = e1.y;
    qa[nt][0] = *reinterpret_cast<const uint32_t*>(p.qp + (size_t)id0 * E + nt * 8 + 2 * t);
    qa[nt][1] = *reinterpret_cast<const uint32_t*>(p.qp + (size_t)id1 * E + nt * 8 + 2 * t);
  }
  // LN2 affine + head weights in C-fragment column order
  float lw[4][2], lb[4][2], hw[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = nt * 8 + 2 * t + j;
      lw[nt][j] = s_small[c]; lb[nt][j] = s_small[32 + c]; hw[nt][j] = s_small[64 + c];
    }
  }
  const float head_b = s_small[96];
  const float sc = 0.35355339059327373f * 1.4426950408889634f;  // 1/sqrt(8) * log2(e)

  const int cell0 = blockIdx.y * p.cells_per_block;
  const int cell1 = min(cell0 + p.cells_per_block, p.n_cells);
  // K/V B-fragments of a cell: kb[h][nt2] = K[key 8nt2+g][dims 8h+2t,+1]; vb[h][half] = V^T[dim 8h+g][keys 2t+8half,+1]
  uint32_t kb[4][2], vb[4][2];
  auto load_kv = [&](int cell, uint32_t (&kk)[4][2], uint32_t (&vv)[4][2]) {
    const __nv_bfloat16* base = p.kvb + (size_t)cell * 1024;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        kk[h][j] = *reinterpret_cast<const uint32_t*>(base + (8 * j + g) * 32 + 8 * h + 2 * t);
        vv[h][j] = *reinterpret_cast<const uint32_t*>(base + 512 + (8 * h + g) * 16 + 2 * t + 8 * j);
      }
    }
  };
  if (cell0 < cell1) load_kv(cell0, kb, vb);

  for (int cell = cell0; cell < cell1; ++cell) {
    uint32_t kn[4][2], vn[4][2];
    if (cell + 1 < cell1) load_kv(cell + 1, kn, vn);   // prefetch the next cell's fragments

    // ---- attention over the 16 latent keys, head by head; O stays in C fragments ----
    float o[4][4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_1688(s0, qa[h][0], qa[h][1], kb[h][0]);   // keys 0-7
      mma_1688(s1, qa[h][0], qa[h][1], kb[h][1]);   // keys 8-15
      float m0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));   // row g
      float m1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));   // row g+8
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      const float n0 = -m0 * sc, n1 = -m1 * sc;      // exp2(s*sc - m*sc): one FMA + one SFU op per score
      s0[0] = sm100::ex2_approx(fmaf(s0[0], sc, n0)); s0[1] = sm100::ex2_approx(fmaf(s0[1], sc, n0));
      s1[0] = sm100::ex2_approx(fmaf(s1[0], sc, n0)); s1[1] = sm100::ex2_approx(fmaf(s1[1], sc, n0));
      s0[2] = sm100::ex2_approx(fmaf(s0[2], sc, n1)); s0[3] = sm100::ex2_approx(fmaf(s0[3], sc, n1));
      s1[2] = sm100::ex2_approx(fmaf(s1[2], sc, n1)); s1[3] = sm100::ex2_approx(fmaf(s1[3], sc, n1));
      float l0 = (s0[0] + s0[1]) + (s1[0] + s1[1]), l1 = (s0[2] + s0[3]) + (s1[2] + s1[3]);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float r0 = sm100::rcp_approx(l0), r1 = sm100::rcp_approx(l1);
      // unnormalised probabilities (<= 1, exact in the softmax sense) go through the PV MMA; the 1/l scaling is applied to the
      // four output values of a row instead of its eight probabilities
      o[h][0] = o[h][1] = o[h][2] = o[h][3] = 0.f;
      mma_16816(o[h], sm100::pack_bf16x2(s0[0], s0[1]), sm100::pack_bf16x2(s0[2], s0[3]),
                sm100::pack_bf16x2(s1[0], s1[1]), sm100::pack_bf16x2(s1[2], s1[3]), vb[h][0], vb[h][1]);
      o[h][0] *= r0; o[h][1] *= r0; o[h][2] *= r1; o[h][3] *= r1;
    }
    // ---- x = q + c_proj(attn): A k-step 0 = heads (0,1), k-step 1 = heads (2,3) ----
    float x[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { x[nt][0] = embf[nt][0]; x[nt][1] = embf[nt][1]; x[nt][2] = embf[nt][2]; x[nt][3] = embf[nt][3]; }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t a0 = sm100::pack_bf16x2(o[2 * ks][0], o[2 * ks][1]), a1 = sm100::pack_bf16x2(o[2 * ks][2], o[2 * ks][3]);
      const uint32_t a2 = sm100::pack_bf16x2(o[2 * ks + 1][0], o[2 * ks + 1][1]), a3 = sm100::pack_bf16x2(o[2 * ks + 1][2], o[2 * ks + 1][3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint2 b = s_frag[(ks * 4 + nt) * 32 + lane];
        mma_16816(x[nt], a0, a1, a2, a3, b.x, b.y);
      }
    }
    // ---- LN2 (rows g and g+8 live in the 4 lanes of a quad) -> A fragments of the MLP ----
    uint32_t ha[2][4];
    {
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { sa += x[nt][0] + x[nt][1]; sb += x[nt][2] + x[nt][3]; }
      sa += __shfl_xor_sync(0xffffffffu, sa, 1); sa += __shfl_xor_sync(0xffffffffu, sa, 2);
      sb += __shfl_xor_sync(0xffffffffu, sb, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
      const float ma = sa * (1.0f / E), mb = sb * (1.0f / E);
      float qa2 = 0.f, qb2 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float d0 = x[nt][0] - ma, d1 = x[nt][1] - ma, d2 = x[nt][2] - mb, d3 = x[nt][3] - mb;
        qa2 += d0 * d0 + d1 * d1; qb2 += d2 * d2 + d3 * d3;
      }
      qa2 += __shfl_xor_sync(0xffffffffu, qa2, 1); qa2 += __shfl_xor_sync(0xffffffffu, qa2, 2);
      qb2 += __shfl_xor_sync(0xffffffffu, qb2, 1); qb2 += __shfl_xor_sync(0xffffffffu, qb2, 2);
      const float ra = rsqrtf(qa2 * (1.0f / E) + p.eps), rb = rsqrtf(qb2 * (1.0f / E) + p.eps);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int nt = 2 * ks + j;
          ha[ks][2 * j] = sm100::pack_bf16x2((x[nt][0] - ma) * ra * lw[nt][0] + lb[nt][0], (x[nt][1] - ma) * ra * lw[nt][1] + lb[nt][1]);
          ha[ks][2 * j + 1] = sm100::pack_bf16x2((x[nt][2] - mb) * rb * lw[nt][0] + lb[nt][0], (x[nt][3] - mb) * rb * lw[nt][1] + lb[nt][1]);
        }
      }
    }
    // ---- SwiGLU MLP streamed over 16-wide hidden chunks (hidden 88 zero-padded to 96) ----
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const uint2* fr = s_frag + (8 + c * 12) * 32 + lane;
      float a1[2][4] = {}, a2[2][4] = {};
#pragma unroll
      for (int n2 = 0; n2 < 2; ++n2) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint2 b1 = fr[(n2 * 2 + ks) * 32], b2 = fr[(4 + n2 * 2 + ks) * 32];
          mma_16816(a1[n2], ha[ks][0], ha[ks][1], ha[ks][2], ha[ks][3], b1.x, b1.y);
          mma_16816(a2[n2], ha[ks][0], ha[ks][1], ha[ks][2], ha[ks][3], b2.x, b2.y);
        }
      }
      const uint32_t h0 = sm100::pack_bf16x2(sm100::silu_from_half(a1[0][0]) * a2[0][0], sm100::silu_from_half(a1[0][1]) * a2[0][1]);
      const uint32_t h1 = sm100::pack_bf16x2(sm100::silu_from_half(a1[0][2]) * a2[0][2], sm100::silu_from_half(a1[0][3]) * a2[0][3]);
      const uint32_t h2 = sm100::pack_bf16x2(sm100::silu_from_half(a1[1][0]) * a2[1][0], sm100::silu_from_half(a1[1][1]) * a2[1][1]);
      const uint32_t h3 = sm100::pack_bf16x2(sm100::silu_from_half(a1[1][2]) * a2[1][2], sm100::silu_from_half(a1[1][3]) * a2[1][3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint2 b3 = fr[(8 + nt) * 32];
        mma_16816(x[nt], h0, h1, h2, h3, b3.x, b3.y);
      }
    }
    // ---- NB-head logit + softmax-over-genes partials ----
    float la = 0.f, lb2 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      la += x[nt][0] * hw[nt][0] + x[nt][1] * hw[nt][1];
      lb2 += x[nt][2] * hw[nt][0] + x[nt][3] * hw[nt][1];
    }
    la += __shfl_xor_sync(0xffffffffu, la, 1); la += __shfl_xor_sync(0xffffffffu, la, 2);
    lb2 += __shfl_xor_sync(0xffffffffu, lb2, 1); lb2 += __shfl_xor_sync(0xffffffffu, lb2, 2);
    la += head_b; lb2 += head_b;
    if (t == 0) {
      if (v0) p.logits[(size_t)cell * p.G + gi0] = la;
      if (v1) p.logits[(size_t)cell * p.G + gi1] = lb2;
    }
    if (p.log_theta != nullptr) {   // unshared theta: second output channel of the head (weights read from shared memory: rare path)
      float ta = 0.f, tb = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float w0 = s_small[100 + nt * 8 + 2 * t], w1 = s_small[100 + nt * 8 + 2 * t + 1];
        ta += x[nt][0] * w0 + x[nt][1] * w1;
        tb += x[nt][2] * w0 + x[nt][3] * w1;
      }
      ta += __shfl_xor_sync(0xffffffffu, ta, 1); ta += __shfl_xor_sync(0xffffffffu, ta, 2);
      tb += __shfl_xor_sync(0xffffffffu, tb, 1); tb += __shfl_xor_sync(0xffffffffu, tb, 2);
      if (t == 0) {
        if (v0) p.log_theta[(size_t)cell * p.G + gi0] = ta + s_small[132];
        if (v1) p.log_theta[(size_t)cell * p.G + gi1] = tb + s_small[132];
      }
    }
    const float lm = fmaxf(v0 ? la : -INFINITY, v1 ? lb2 : -INFINITY);
    const float wm = sm100::warp_max(lm);
    const float ev = (t == 0) ? ((v0 ? __expf(la - wm) : 0.f) + (v1 ? __expf(lb2 - wm) : 0.f)) : 0.f;
    const float wsum = sm100::warp_sum(ev);
    // one (max, sum) partial per warp (16 genes): no CTA barrier in the per-cell loop, so the eight warps drift freely and hide
    // each other's latencies; nb_finalize_kernel merges 8 x gene_tiles partials per cell
    if (lane == 0) p.partials[(size_t)cell * p.gene_tiles * 8 + blockIdx.x * 8 + warp] = make_float2(wm, wsum);
#pragma unroll
    for (int h = 0; h < 4; ++h) { kb[h][0] = kn[h][0]; kb[h][1] = kn[h][1]; vb[h][0] = vn[h][0]; vb[h][1] = vn[h][1]; }
  }
}

// ---- NB reconstruction loss (distributions.py:6-42 log_nb_positive; models.py:233-247 VAE.loss) ---------------------
// nll[cell] = -sum_g log NB(x | mu, theta), the per-cell term of `recon_loss.sum(dim=1)`; eps = 1e-8 as the reference.
// One CTA per cell, coalesced float4 rows; HBM-bound: 12 B per (cell, gene) when theta is per cell, 8 B when it is the
// shared (G,) row.  Zero counts (the bulk of a count matrix) skip the three lgamma terms, which cancel exactly there.
// log(k!) for k <= 16: lgamma(x + 1) of the small integer counts that make up almost all non-zero entries
__constant__ float c_log_factorial[17] = {0.f, 0.f, 0.6931471806f, 1.7917594692f, 3.1780538303f, 4.7874917428f, 6.5792512120f,
                                          8.5251613611f, 10.6046029027f, 12.8018274801f, 15.1044125730f, 17.5023078459f,
                                          19.9872144957f, 22.5521638531f, 25.1912211827f, 27.8992713838f, 30.6718601061f};
__device__ __forceinline__ float nb_logp(float x, float mu, float th) {
  const float eps = 1e-8f;
  const float l_tm = __logf(th + mu + eps);
  float r = th * (__logf(th + eps) - l_tm);
  if (x != 0.f) {
    r += x * (__logf(mu + eps) - l_tm);
    const int k = (int)x;
    if ((float)k == x && k <= 16) {
      // integer count: lgamma(x + theta) - lgamma(theta) = sum_{i<x} log(theta + i) (exact identity), lgamma(x + 1) = log(x!)
      float acc = 0.f;
      for (int i = 0; i < k; ++i) acc += __logf(th + (float)i);
      r += acc - c_log_factorial[k];
    } else {
      r += lgammaf(x + th) - lgammaf(th) - lgammaf(x + 1.f);
    }
  }
  return r;
}
__global__ void __launch_bounds__(256) nb_nll_kernel(const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ theta,
                                                      long long theta_row_stride, int G, float* __restrict__ nll) {
  __shared__ float red[8];
  const size_t cell = blockIdx.x;
  const float* xr = x + cell * (size_t)G;
  const float* mr = mu + cell * (size_t)G;
  const float* tr = theta + cell * (size_t)theta_row_stride;
  float acc = 0.f;
  if ((G & 3) == 0 && (theta_row_stride & 3) == 0) {
    for (int g = threadIdx.x * 4; g < G; g += 1024) {
      const float4 a = *reinterpret_cast<const float4*>(xr + g), b = *reinterpret_cast<const float4*>(mr + g), c = *reinterpret_cast<const float4*>(tr + g);
      acc += (nb_logp(a.x, b.x, c.x) + nb_logp(a.y, b.y, c.y)) + (nb_logp(a.z, b.z, c.z) + nb_logp(a.w, b.w, c.w));
    }
  } else {
    for (int g = threadIdx.x; g < G; g += 256) acc += nb_logp(xr[g], mr[g], tr[g]);
  }
  acc = sm100::warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    nll[cell] = -t;
  }
}

// ---- MCAB encode (layers.py:97-118, 305-330; nnets.py:137-144) ------------------------------------------------
// One CTA (8 warps) per cell.  Phase 1, flash-style pooling on mma.sync: every warp walks 16-token blocks of the
// cell's S gene tokens: gather emb[gene]*log1p(count) -> LN1 (quad shuffles) -> K,V = c_attn (16 mma) -> per-head
// scores against the cached, cell-invariant Q = c_attn_q(LN1q(inducing points)) (m16n8k8) -> online softmax over
// tokens (NO key masking: padding tokens keep their softmax mass) -> P V (V^T via movmatrix).
// Phase 2 merges the warps' (max, sum, acc) states; phase 3 is the fp32 tail on the 16 x 32 latent tile:
// c_proj, + inducing points (raw-query residual), LN2 + SwiGLU, + pos_embed, n_layer Blocks, Linear(32->16), LN.
// multiplicative count transforms of InputTransformerVAE (layers.py:28-44, PROJ_FUNC): token = emb[gene] * count_scale(count)
__device__ __forceinline__ float count_scale(float c, int agg) {
  switch (agg) {
    case 1: return c == 0.f ? -1.f : log1pf(c);       // "log1pzero"
    case 2: return asinhf(sqrtf(c + 1.f));            // "anscombe"
    case 3: return sqrtf(c + 1.f);                    // "sqrt"
    default: return log1pf(c);                        // "log1p"
  }
}

struct EncParams {
  int agg;                      // count transform, see count_scale
  const float* emb;             // gene embedding table [n_ids][32]
  const long long* genes;       // [cells][S]
  const float* counts;          // [cells][S]
  int S;
  int n_cells;
  const uint32_t* wkv_frag;     // c_attn (k|v) in mma B-fragment order: [2 ks][8 nt][32 lanes][2 u32]
  const __nv_bfloat16* q_tbl;   // [16][32] bf16: c_attn_q(LN1q(inducing_points))
  const float* ln1_w; const float* ln1_b;
  const float* inducing;        // [16][32]
  const float* wproj_t;         // ca_layer.attn.c_proj^T [32][32]
  const float* ln2_w; const float* ln2_b;
  const float* w1_t; const float* w2_t;   // [32][88]
  const float* w3_t;            // [88][32]
  const float* pos;             // encoder.pos_embed [16][32] or nullptr
  const float* blocks; int n_layer;
  const float* wlat_t;          // encoder_latent_input.0.weight^T [32][16]
  float eps;
  float* z;                     // [cells][16][16]
};

__global__ void __launch_bounds__(256) mcab_encode_kernel(const EncParams p) {
  __shared__ uint2 s_wkv[16 * 32];              // 4 KB
  __shared__ float s_state[8][4][16][10];       // per warp, head, query: m, l, acc[8]   (20 KB)
  __shared__ __align__(16) float x[TOK][E];
  __shared__ __align__(16) float h[TOK][E];
  __shared__ __align__(16) float qkv[TOK][3 * E];
  __shared__ __align__(16) float hid[TOK][HID];
  const int cell = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 16 * 32; i += 256) s_wkv[i] = reinterpret_cast<const uint2*>(p.wkv_frag)[i];
  __syncthreads();

  // Q fragments per head (m16n8k8 A operand): a0 = (query g, dims 8h+2t,+1), a1 = (query g+8, ...)
  uint32_t qa[4][2];
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
    qa[hh][0] = *reinterpret_cast<const uint32_t*>(p.q_tbl + g * E + 8 * hh + 2 * t);
    qa[hh][1] = *reinterpret_cast<const uint32_t*>(p.q_tbl + (g + 8) * E + 8 * hh + 2 * t);
  }
  // LN1 affine in A-fragment column order: columns {2t, 2t+1, 2t+8, 2t+9} + 16*ks
  float lw[2][4], lb[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = 16 * ks + 2 * t + (j & 1) + 8 * (j >> 1);
      lw[ks][j] = p.ln1_w[c]; lb[ks][j] = p.ln1_b[c];
    }
  }
  const float sc = 0.35355339059327373f * 1.4426950408889634f;  // 1/sqrt(8) * log2(e)
  float m_run[4][2], l_run[4][2], o_acc[4][4];
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
    m_run[hh][0] = m_run[hh][1] = -INFINITY; l_run[hh][0] = l_run[hh][1] = 0.f;
    o_acc[hh][0] = o_acc[hh][1] = o_acc[hh][2] = o_acc[hh][3] = 0.f;
  }
  const long long* gp = p.genes + (size_t)cell * p.S;
  const float* cp = p.counts + (size_t)cell * p.S;
  const int n_blk = (p.S + 15) >> 4;
  // Software pipeline over this warp's token blocks: the gene ids / counts are fetched two blocks ahead and the embedding rows
  // they point to one block ahead, so the two dependent global loads (id -> embedding row) of a block are in flight while the
  // previous block is normalised, projected and pooled (the kernel is latency-bound: one CTA walks a cell's S tokens).
  struct TokIds { long long id0, id1; float c0, c1; };
  auto load_ids = [&](int blk) {
    TokIds r;
    const int t0 = blk * 16 + g, t1 = t0 + 8;
    const bool ok0 = blk < n_blk && t0 < p.S, ok1 = blk < n_blk && t1 < p.S;
    r.id0 = ok0 ? gp[t0] : 0; r.id1 = ok1 ? gp[t1] : 0;
    r.c0 = ok0 ? cp[t0] : 0.f; r.c1 = ok1 ? cp[t1] : 0.f;
    return r;
  };
  auto gather = [&](const TokIds& r, float2 (&e)[2][4]) {   // [ks]{row g cols 2t, row g cols 2t+8, row g+8 cols 2t, row g+8 cols 2t+8}
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      e[ks][0] = *reinterpret_cast<const float2*>(p.emb + (size_t)r.id0 * E + 16 * ks + 2 * t);
      e[ks][1] = *reinterpret_cast<const float2*>(p.emb + (size_t)r.id0 * E + 16 * ks + 2 * t + 8);
      e[ks][2] = *reinterpret_cast<const float2*>(p.emb + (size_t)r.id1 * E + 16 * ks + 2 * t);
      e[ks][3] = *reinterpret_cast<const float2*>(p.emb + (size_t)r.id1 * E + 16 * ks + 2 * t + 8);
    }
  };
  TokIds ids_cur = load_ids(warp), ids_nxt = load_ids(warp + 8);
  float2 emb_cur[2][4], emb_nxt[2][4];
  gather(ids_cur, emb_cur);
  for (int blk = warp; blk < n_blk; blk += 8) {
    gather(ids_nxt, emb_nxt);                       // rows of block blk + 8 (row 0 of the table when that block does not exist)
    const TokIds ids_nn = load_ids(blk + 16);
    // ---- tokens g and g+8 of this block: scale by log1p(count), LayerNorm over 32 channels ----
    const float c0 = count_scale(ids_cur.c0, p.agg), c1 = count_scale(ids_cur.c1, p.agg);   // (tokens beyond S are masked below)
    float xv[2][2][4];  // [row g / g+8][ks][4 cols]
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      xv[0][ks][0] = emb_cur[ks][0].x * c0; xv[0][ks][1] = emb_cur[ks][0].y * c0; xv[0][ks][2] = emb_cur[ks][1].x * c0; xv[0][ks][3] = emb_cur[ks][1].y * c0;
      xv[1][ks][0] = emb_cur[ks][2].x * c1; xv[1][ks][1] = emb_cur[ks][2].y * c1; xv[1][ks][2] = emb_cur[ks][3].x * c1; xv[1][ks][3] = emb_cur[ks][3].y * c1;
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int j = 0; j < 4; ++j) emb_cur[ks][j] = emb_nxt[ks][j];
    ids_cur = ids_nxt;
    ids_nxt = ids_nn;
    uint32_t xa[2][4];  // A fragments of LN1(x): [ks]{a0 (row g, cols 2t..), a1 (row g+8), a2 (row g, cols 2t+8..), a3}
    {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s0 += xv[0][ks][j]; s1 += xv[1][ks][j]; }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float mu0 = s0 * (1.0f / E), mu1 = s1 * (1.0f / E);
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          xv[0][ks][j] -= mu0; q0 += xv[0][ks][j] * xv[0][ks][j];
          xv[1][ks][j] -= mu1; q1 += xv[1][ks][j] * xv[1][ks][j];
        }
      q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
      const float r0 = rsqrtf(q0 * (1.0f / E) + p.eps), r1 = rsqrtf(q1 * (1.0f / E) + p.eps);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        xa[ks][0] = sm100::pack_bf16x2(xv[0][ks][0] * r0 * lw[ks][0] + lb[ks][0], xv[0][ks][1] * r0 * lw[ks][1] + lb[ks][1]);
        xa[ks][1] = sm100::pack_bf16x2(xv[1][ks][0] * r1 * lw[ks][0] + lb[ks][0], xv[1][ks][1] * r1 * lw[ks][1] + lb[ks][1]);
        xa[ks][2] = sm100::pack_bf16x2(xv[0][ks][2] * r0 * lw[ks][2] + lb[ks][2], xv[0][ks][3] * r0 * lw[ks][3] + lb[ks][3]);
        xa[ks][3] = sm100::pack_bf16x2(xv[1][ks][2] * r1 * lw[ks][2] + lb[ks][2], xv[1][ks][3] * r1 * lw[ks][3] + lb[ks][3]);
      }
    }
    // ---- K (n-tiles 0-3 = heads) and V (n-tiles 4-7) of the 16 tokens: C rows = tokens g / g+8 ----
    float kv[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      kv[nt][0] = kv[nt][1] = kv[nt][2] = kv[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint2 b = s_wkv[(ks * 8 + nt) * 32 + lane];
        mma_16816(kv[nt], xa[ks][0], xa[ks][1], xa[ks][2], xa[ks][3], b.x, b.y);
      }
    }
    // ---- per head: scores (16 queries x 16 tokens), online softmax over tokens, O += P V ----
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
      // B operand of the score MMA: B[k=dim][n=token] = K[token][dim]; the K C-fragment is already in that layout
      const uint32_t kb0 = sm100::pack_bf16x2(kv[hh][0], kv[hh][1]);   // tokens 0-7  of the block
      const uint32_t kb1 = sm100::pack_bf16x2(kv[hh][2], kv[hh][3]);   // tokens 8-15
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_1688(s0, qa[hh][0], qa[hh][1], kb0);   // rows = queries g / g+8, cols = tokens 2t,2t+1
      mma_1688(s1, qa[hh][0], qa[hh][1], kb1);   // cols = tokens 8+2t, 8+2t+1
      // tokens beyond S do not exist (ragged tail of the last block) -> no softmax mass
      const int tb = blk * 16 + 2 * t;
      if (tb >= p.S) { s0[0] = s0[2] = -INFINITY; }
      if (tb + 1 >= p.S) { s0[1] = s0[3] = -INFINITY; }
      if (tb + 8 >= p.S) { s1[0] = s1[2] = -INFINITY; }
      if (tb + 9 >= p.S) { s1[1] = s1[3] = -INFINITY; }
      float bm0 = fmaxf(fmaxf(s0[0], s0[1]), fmaxf(s1[0], s1[1]));   // query g
      float bm1 = fmaxf(fmaxf(s0[2], s0[3]), fmaxf(s1[2], s1[3]));   // query g+8
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float nm0 = fmaxf(m_run[hh][0], bm0), nm1 = fmaxf(m_run[hh][1], bm1);
      // nm0 / nm1 are finite (every block a warp visits holds at least one real token), so the folded form is safe
      const float n0 = -nm0 * sc, n1 = -nm1 * sc;
      const float f0 = sm100::ex2_approx(fmaf(m_run[hh][0], sc, n0)), f1 = sm100::ex2_approx(fmaf(m_run[hh][1], sc, n1));   // ex2(-inf) = 0 on first use
      m_run[hh][0] = nm0; m_run[hh][1] = nm1;
      s0[0] = sm100::ex2_approx(fmaf(s0[0], sc, n0)); s0[1] = sm100::ex2_approx(fmaf(s0[1], sc, n0));
      s1[0] = sm100::ex2_approx(fmaf(s1[0], sc, n0)); s1[1] = sm100::ex2_approx(fmaf(s1[1], sc, n0));
      s0[2] = sm100::ex2_approx(fmaf(s0[2], sc, n1)); s0[3] = sm100::ex2_approx(fmaf(s0[3], sc, n1));
      s1[2] = sm100::ex2_approx(fmaf(s1[2], sc, n1)); s1[3] = sm100::ex2_approx(fmaf(s1[3], sc, n1));
      l_run[hh][0] = l_run[hh][0] * f0 + (s0[0] + s0[1]) + (s1[0] + s1[1]);   // per-lane partial sums (quad-reduced at the end)
      l_run[hh][1] = l_run[hh][1] * f1 + (s0[2] + s0[3]) + (s1[2] + s1[3]);
      o_acc[hh][0] *= f0; o_acc[hh][1] *= f0; o_acc[hh][2] *= f1; o_acc[hh][3] *= f1;
      // B operand of P V: B[k=token][n=dim] = V[token][dim] -> transpose the V C-fragment 8x8 tiles in registers
      const uint32_t vb0 = movmatrix_trans_b16(sm100::pack_bf16x2(kv[4 + hh][0], kv[4 + hh][1]));   // tokens 0-7
      const uint32_t vb1 = movmatrix_trans_b16(sm100::pack_bf16x2(kv[4 + hh][2], kv[4 + hh][3]));   // tokens 8-15
      mma_16816(o_acc[hh], sm100::pack_bf16x2(s0[0], s0[1]), sm100::pack_bf16x2(s0[2], s0[3]), sm100::pack_bf16x2(s1[0], s1[1]),
                sm100::pack_bf16x2(s1[2], s1[3]), vb0, vb1);
    }
  }
  // ---- phase 2: merge the 8 warps' online-softmax states ----
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
    float l0 = l_run[hh][0], l1 = l_run[hh][1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    if (t == 0) {
      s_state[warp][hh][g][0] = m_run[hh][0]; s_state[warp][hh][g][1] = l0;
      s_state[warp][hh][g + 8][0] = m_run[hh][1]; s_state[warp][hh][g + 8][1] = l1;
    }
    s_state[warp][hh][g][2 + 2 * t] = o_acc[hh][0]; s_state[warp][hh][g][3 + 2 * t] = o_acc[hh][1];
    s_state[warp][hh][g + 8][2 + 2 * t] = o_acc[hh][2]; s_state[warp][hh][g + 8][3 + 2 * t] = o_acc[hh][3];
  }
  __syncthreads();
  for (int i = tid; i < TOK * E; i += 256) {   // pooled[q][8h+d] -> h
    const int qi = i / E, c = i % E, hh = c >> 3, d = c & 7;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) M = fmaxf(M, s_state[w][hh][qi][0]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const float mw = s_state[w][hh][qi][0];
      const float f = (mw == -INFINITY) ? 0.f : exp2f((mw - M) * sc);
      L += s_state[w][hh][qi][1] * f;
      O += s_state[w][hh][qi][2 + d] * f;
    }
    h[qi][c] = O / L;
  }
  __syncthreads();
  // ---- phase 3: x = inducing + c_proj(pooled); x += MLP(LN2(x)); x += pos; Blocks; Linear; LN ----
  for (int i = tid; i < TOK * E; i += 256) {
    const int qi = i / E, c = i % E;
    float acc = p.inducing[i];
#pragma unroll 8
    for (int k = 0; k < E; ++k) acc += h[qi][k] * p.wproj_t[k * E + c];
    x[qi][c] = acc;
  }
  __syncthreads();
  if (tid < 128) ln32_tokens(x, h, p.ln2_w, p.ln2_b, p.eps, tid);
  __syncthreads();
  for (int i = tid; i < TOK * HID; i += 256) {
    const int qi = i / HID, j = i % HID;
    float a = 0.f, b = 0.f;
#pragma unroll 8
    for (int k = 0; k < E; ++k) { a += h[qi][k] * p.w1_t[k * HID + j]; b += h[qi][k] * p.w2_t[k * HID + j]; }
    hid[qi][j] = sm100::silu(a) * b;
  }
  __syncthreads();
  for (int i = tid; i < TOK * E; i += 256) {
    const int qi = i / E, c = i % E;
    float acc = p.pos ? p.pos[i] : 0.f;
#pragma unroll 8
    for (int k = 0; k < HID; ++k) acc += hid[qi][k] * p.w3_t[k * E + c];
    x[qi][c] += acc;
  }
  __syncthreads();
  vae_block_stack(x, h, qkv, hid, p.blocks, p.n_layer, p.eps, tid, 256);
  // Linear(32 -> 16) + LayerNorm(16, no affine)
  float (*zl)[LAT] = reinterpret_cast<float (*)[LAT]>(&qkv[0][0]);
  for (int i = tid; i < TOK * LAT; i += 256) {
    const int qi = i / LAT, c = i % LAT;
    float acc = 0.f;
#pragma unroll 8
    for (int k = 0; k < E; ++k) acc += x[qi][k] * p.wlat_t[k * LAT + c];
    zl[qi][c] = acc;
  }
  __syncthreads();
  if (tid < TOK) {
    float m = 0.f;
    for (int j = 0; j < LAT; ++j) m += zl[tid][j];
    m *= (1.0f / LAT);
    float var = 0.f;
    for (int j = 0; j < LAT; ++j) { const float dlt = zl[tid][j] - m; var += dlt * dlt; }
    const float rstd = rsqrtf(var * (1.0f / LAT) + p.eps);
    for (int j = 0; j < LAT; ++j) p.z[((size_t)cell * TOK + tid) * LAT + j] = (zl[tid][j] - m) * rstd;
  }
}

// ---- softmax-over-genes finalisation + optional NB sampling --------------------------------
struct NbParams {
  const float* logits;     // [cells][G] (row stride logit_stride when that is non-zero)
  long long logit_stride;
  const float2* partials;  // [cells][gene_tiles] (max, sum) pairs; gene_tiles = partials per cell (8 per 128-gene tile on the tensor-core path)
  int gene_tiles;
  int G;
  int n_cells;
  const float* lib;        // [cells] library size
  const float* theta_tbl;  // decoder_head.theta.weight [n_ids] (log theta); nullptr for an unshared-theta head, whose log theta
                           // the MCAB kernel left in `theta` [cells][G] (exponentiated in place here)
  const long long* genes;  // [G]
  float* mu;               // [cells][G] or nullptr
  float* theta;            // [G] or nullptr (written by cell 0 blocks)
  float* counts;           // [cells][G] or nullptr
  unsigned long long seed;
  long long cell_offset;   // global index of cell 0 (sharding / chunking invariance)
};

__global__ void __launch_bounds__(256) nb_finalize_kernel(const NbParams p) {
  __shared__ float s_m, s_s;
  __shared__ float rm[8], rs[8];
  const int cell = blockIdx.x, tid = threadIdx.x;
  // combine the per-tile (max, sum) pairs of this cell
  float m = -INFINITY;
  for (int i = tid; i < p.gene_tiles; i += 256) m = fmaxf(m, p.partials[(size_t)cell * p.gene_tiles + i].x);
  m = sm100::warp_max(m);
  if ((tid & 31) == 0) rm[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) {
    float mm = rm[0];
    for (int i = 1; i < 8; ++i) mm = fmaxf(mm, rm[i]);
    s_m = mm;
  }
  __syncthreads();
  const float gm = s_m;
  float s = 0.f;
  for (int i = tid; i < p.gene_tiles; i += 256) {
    const float2 pr = p.partials[(size_t)cell * p.gene_tiles + i];
    s += pr.y * __expf(pr.x - gm);
  }
  s = sm100::warp_sum(s);
  if ((tid & 31) == 0) rs[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += rs[i];
    s_s = t;
  }
  __syncthreads();
  const float scale = p.lib[cell] / s_s;
  const long long gcell = p.cell_offset + cell;
  for (int gi = blockIdx.y * 256 + tid; gi < p.G; gi += gridDim.y * 256) {
    const float muv = __expf(p.logits[(size_t)cell * (p.logit_stride ? p.logit_stride : (long long)p.G) + gi] - gm) * scale;
    float th;
    if (p.theta_tbl != nullptr) {
      th = __expf(p.theta_tbl[p.genes[gi]]);
      if (p.theta && cell == 0) p.theta[gi] = th;
    } else {
      th = __expf(p.theta[(size_t)cell * p.G + gi]);
      p.theta[(size_t)cell * p.G + gi] = th;
    }
    if (p.mu) p.mu[(size_t)cell * p.G + gi] = muv;
    if (p.counts) {
      rng::Philox g(p.seed, (uint32_t)gi, (uint32_t)gcell, (uint32_t)(gcell >> 32) ^ 0x4E42u);
      p.counts[(size_t)cell * p.G + gi] = rng::negative_binomial(g, muv, th);
    }
  }
}

// standard normal noise / log-normal size factors keyed by global cell index
__global__ void __launch_bounds__(256) randn_cells_kernel(float* out, int n_cells, int per_cell, unsigned long long seed,
                                                           long long cell_offset, unsigned int stream) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)n_cells * per_cell) return;
  const long long cell = cell_offset + i / per_cell;
  rng::Philox g(seed, (uint32_t)(i % per_cell), (uint32_t)cell, (uint32_t)(cell >> 32) ^ stream);
  out[i] = g.normal();
}

}  // namespace vae
