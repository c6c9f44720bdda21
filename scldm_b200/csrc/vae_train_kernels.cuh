// VAE training step (n_embed = 32) on the device: forward with the activations the backward pass needs, NB loss and its
// gradient, backward of decoder / encoder into a flat gradient buffer.
// Reference: VAE.training_step src/scldm/models.py:249-287, TransformerVAE.forward vae.py:29-56, VAE.loss models.py:233-247,
// log_nb_positive distributions.py:6-42, Encoder / Decoder nnets.py:82-208, CrossAttentionBlock layers.py:267-330,
// Block layers.py:177-226, InputTransformerVAE layers.py:97-118, NegativeBinomialTransformerLayer stochastic_layers.py:102-116.
//
// Where the work is: the decoder's cross-attention block runs on every (cell, gene) token - B x G of them (4.6 M for 128 cells
// of the census vocabulary), 28 k multiply-adds each for forward + dgrad + wgrad.  dec_mcab_train_kernel keeps a 32-token tile of
// one cell in shared memory, recomputes its forward and differentiates it there; every GEMM of the tile (attention products, c_proj,
// [w1|w2], their dgrads and the weight gradients, which accumulate in registers over all cells a CTA visits) is TF32 mma.sync fed
// from shared memory - the precision the reference trains in (torch.set_float32_matmul_precision("high"), scripts/train.py:18);
// mlp.c_proj never runs: the head is linear in the block's output, which folds it into two dot products (forward) and rank-1
// updates (backward).  The encoder's token side (K / V projection, pooling, their backward) runs on the same mma helpers over
// 128-token tiles; what is per cell (16 latent tokens x 32 channels through 2 x n_layer Blocks) is fp32 CUDA-core code with the
// Block weights staged in shared memory: < 1 % of the step's arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstddef>

#include "vae_kernels.cuh"

namespace vtr {

constexpr int E = 32, M = 16, LAT = 16, H = 88;
constexpr int NHB = 8, HDB = 4;   // Block self attention: 8 heads x 4
constexpr int NHC = 4, HDC = 8;   // cross attention: 4 heads x 8
// parameter group of a Block (element offsets inside the flat buffer, from the block's base)
constexpr int B_LN1W = 0, B_LN1B = 32, B_CATTN = 64, B_CPROJ = B_CATTN + 96 * 32, B_LN2W = B_CPROJ + 1024, B_LN2B = B_LN2W + 32,
              B_W1 = B_LN2B + 32, B_W2 = B_W1 + H * 32, B_W3 = B_W2 + H * 32, B_SIZE = B_W3 + 32 * H;
// parameter group of a CrossAttentionBlock
constexpr int C_LN1W = 0, C_LN1B = 32, C_LN1QW = 64, C_LN1QB = 96, C_LN2W = 128, C_LN2B = 160, C_CATTN = 192,
              C_CATTNQ = C_CATTN + 64 * 32, C_CPROJ = C_CATTNQ + 1024, C_W1 = C_CPROJ + 1024, C_W2 = C_W1 + H * 32,
              C_W3 = C_W2 + H * 32, C_SIZE = C_W3 + 32 * H;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
// 16-byte vector reduction into global memory (sm_90+): one L2 transaction instead of four
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------------------
// CTA-cooperative fp32 helpers on 16-row tiles in shared memory (the per-cell latent chain)
// ---------------------------------------------------------------------------------------------------------------
// y[r][o] (=, +=) sum_k x[r][k] W[o][k]   (W: global, row-major [O][K], K % 4 == 0)
template <int ADD>
__device__ __forceinline__ void lin16(const float* x, int ldx, int K, const float* __restrict__ W, int ldw, int O, float* y, int ldy) {
  for (int idx = threadIdx.x; idx < 16 * O; idx += blockDim.x) {
    const int r = idx / O, o = idx - r * O;
    const float* w = W + (size_t)o * ldw;
    const float* xr = x + r * ldx;
    float acc = 0.f;
    for (int k = 0; k < K; k += 4) {
      const float4 wv = *reinterpret_cast<const float4*>(w + k);
      acc += wv.x * xr[k] + wv.y * xr[k + 1] + wv.z * xr[k + 2] + wv.w * xr[k + 3];
    }
    if (ADD) y[r * ldy + o] += acc; else y[r * ldy + o] = acc;
  }
}
// dx[r][k] (=, +=) sum_o dy[r][o] W[o][k]
template <int ADD>
__device__ __forceinline__ void lin16_t(const float* dy, int ldy, int O, const float* __restrict__ W, int ldw, int K, float* dx, int ldx) {
  for (int idx = threadIdx.x; idx < 16 * K; idx += blockDim.x) {
    const int r = idx / K, k = idx - r * K;
    const float* d = dy + r * ldy;
    float acc = 0.f;
    for (int o = 0; o < O; ++o) acc += d[o] * W[(size_t)o * ldw + k];
    if (ADD) dx[r * ldx + k] += acc; else dx[r * ldx + k] = acc;
  }
}
// gW[o][k] += sum_{r < R} dy[r][o] x[r][k]   (atomics into the flat gradient buffer)
__device__ __forceinline__ void wgrad_rows(const float* dy, int ldy, int O, const float* x, int ldx, int K, int R, float* gW) {
  for (int idx = threadIdx.x; idx < O * K; idx += blockDim.x) {
    const int o = idx / K, k = idx - o * K;
    float acc = 0.f;
    for (int r = 0; r < R; ++r) acc += dy[r * ldy + o] * x[r * ldx + k];
    atomicAdd(gW + idx, acc);
  }
}
// LayerNorm over `dim` <= 32 channels of 16 rows, one warp per row (w / b nullable: no affine)
__device__ __forceinline__ void ln16(const float* x, int ldx, int dim, const float* __restrict__ w, const float* __restrict__ b,
                                     float eps, float* y, int ldy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float inv = 1.f / (float)dim;
  for (int r = warp; r < 16; r += nw) {
    const float v = lane < dim ? x[r * ldx + lane] : 0.f;
    const float mean = warp_sum(v) * inv;
    const float d = lane < dim ? v - mean : 0.f;
    const float rstd = rsqrtf(warp_sum(d * d) * inv + eps);
    if (lane < dim) y[r * ldy + lane] = d * rstd * (w ? w[lane] : 1.f) + (b ? b[lane] : 0.f);
  }
}
// backward of the above: dy = gradient w.r.t. the LayerNorm output, x = its input; dx (=, +=); gw / gb nullable
template <int ADD>
__device__ __forceinline__ void ln16_bwd(const float* dy, int ldd, const float* x, int ldx, int dim, const float* __restrict__ w,
                                         float eps, float* dx, int lddx, float* gw, float* gb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const float inv = 1.f / (float)dim;
  float accw = 0.f, accb = 0.f;
  for (int r = warp; r < 16; r += nw) {
    const float v = lane < dim ? x[r * ldx + lane] : 0.f;
    const float mean = warp_sum(v) * inv;
    const float d = lane < dim ? v - mean : 0.f;
    const float rstd = rsqrtf(warp_sum(d * d) * inv + eps);
    const float xhat = d * rstd;
    const float g = lane < dim ? dy[r * ldd + lane] : 0.f;
    accw += g * xhat;
    accb += g;
    const float gy = g * (w ? (lane < dim ? w[lane] : 0.f) : 1.f);
    const float m1 = warp_sum(gy) * inv, m2 = warp_sum(gy * xhat) * inv;
    const float dxv = rstd * (gy - m1 - xhat * m2);
    if (lane < dim) { if (ADD) dx[r * lddx + lane] += dxv; else dx[r * lddx + lane] = dxv; }
  }
  if (gw && lane < dim) { atomicAdd(gw + lane, accw); atomicAdd(gb + lane, accb); }
}

// shared-memory working set of the latent chain (floats)
struct LatS {
  float x[16 * 32], xn[16 * 32], qkv[16 * 96], ao[16 * 32], u[16 * H], v[16 * H];
  // backward only
  float xm[16 * 32], hh[16 * H], P[NHB * 16 * 16], dS[NHB * 16 * 16], dx[16 * 32], dt[16 * 96], dao[16 * 32], dhh[16 * H];
};
constexpr size_t LAT_FWD_SMEM = offsetof(LatS, xm);

// Block self attention on the 16 tokens: thread (head, query)
__device__ __forceinline__ void attn16_fwd(const float* qkv, float* ao, float* P) {
  if (threadIdx.x < 128) {
    const int h = threadIdx.x >> 4, i = threadIdx.x & 15;
    float q[HDB], s[16], mx = -1e30f;
#pragma unroll
    for (int d = 0; d < HDB; ++d) q[d] = qkv[i * 96 + h * HDB + d] * 0.5f;   // 1 / sqrt(head_dim = 4)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HDB; ++d) a += q[d] * qkv[j * 96 + 32 + h * HDB + d];
      s[j] = a;
      mx = fmaxf(mx, a);
    }
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { s[j] = __expf(s[j] - mx); l += s[j]; }
    const float il = 1.f / l;
    float o[HDB] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float p = s[j] * il;
      if (P) P[(h * 16 + i) * 16 + j] = p;
#pragma unroll
      for (int d = 0; d < HDB; ++d) o[d] += p * qkv[j * 96 + 64 + h * HDB + d];
    }
#pragma unroll
    for (int d = 0; d < HDB; ++d) ao[i * 32 + h * HDB + d] = o[d];
  }
}
// backward: dao -> dqkv (q | k | v gradients, [16][96]); dS is scratch.  Contains its own CTA barriers.
__device__ __forceinline__ void attn16_bwd(const float* qkv, const float* P, const float* dao, float* dS, float* dqkv) {
  if (threadIdx.x < 128) {
    const int h = threadIdx.x >> 4, i = threadIdx.x & 15;
    float g[HDB], dp[16], D = 0.f;
#pragma unroll
    for (int d = 0; d < HDB; ++d) g[d] = dao[i * 32 + h * HDB + d];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HDB; ++d) a += g[d] * qkv[j * 96 + 64 + h * HDB + d];
      dp[j] = a;
      D += P[(h * 16 + i) * 16 + j] * a;
    }
    float dq[HDB] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float ds = P[(h * 16 + i) * 16 + j] * (dp[j] - D) * 0.5f;   // includes the 1 / sqrt(head_dim) of the scores
      dS[(h * 16 + i) * 16 + j] = ds;
#pragma unroll
      for (int d = 0; d < HDB; ++d) dq[d] += ds * qkv[j * 96 + 32 + h * HDB + d];
    }
#pragma unroll
    for (int d = 0; d < HDB; ++d) dqkv[i * 96 + h * HDB + d] = dq[d];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const int h = threadIdx.x >> 4, j = threadIdx.x & 15;
    float dk[HDB] = {0.f, 0.f, 0.f, 0.f}, dv[HDB] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float ds = dS[(h * 16 + i) * 16 + j], p = P[(h * 16 + i) * 16 + j];
#pragma unroll
      for (int d = 0; d < HDB; ++d) {
        dk[d] += ds * qkv[i * 96 + h * HDB + d];
        dv[d] += p * dao[i * 32 + h * HDB + d];
      }
    }
#pragma unroll
    for (int d = 0; d < HDB; ++d) { dqkv[j * 96 + 32 + h * HDB + d] = dk[d]; dqkv[j * 96 + 64 + h * HDB + d] = dv[d]; }
  }
  __syncthreads();
}

// x <- x + c_proj(silu(w1 LN2(x)) * (w2 LN2(x)))   (SwiGLU MLP of Block / CrossAttentionBlock, layers.py:161-174)
__device__ __forceinline__ void mlp16_fwd(LatS& s, const float* ln2w, const float* ln2b, const float* w1, const float* w2, const float* w3,
                                          int l32, int l88, float eps) {
  ln16(s.x, 32, 32, ln2w, ln2b, eps, s.xn, 32);
  __syncthreads();
  lin16<0>(s.xn, 32, 32, w1, l32, H, s.u, H);
  lin16<0>(s.xn, 32, 32, w2, l32, H, s.v, H);
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * H; i += blockDim.x) { const float a = s.u[i]; s.u[i] = a * sigmoidf_(a) * s.v[i]; }
  __syncthreads();
  lin16<1>(s.u, H, H, w3, l88, 32, s.x, 32);
  __syncthreads();
}
// backward of the MLP half: xin = its input (LN2 input), s.dx = gradient of its output on entry, of its input on exit
__device__ __forceinline__ void mlp16_bwd(LatS& s, const float* xin, const float* ln2w, const float* ln2b, const float* w1, const float* w2,
                                          const float* w3, int l32, int l88, float* g_ln2w, float* g_ln2b, float* g_w1, float* g_w2, float* g_w3,
                                          float eps) {
  ln16(xin, 32, 32, ln2w, ln2b, eps, s.xn, 32);
  __syncthreads();
  lin16<0>(s.xn, 32, 32, w1, l32, H, s.u, H);
  lin16<0>(s.xn, 32, 32, w2, l32, H, s.v, H);
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * H; i += blockDim.x) { const float a = s.u[i]; s.hh[i] = a * sigmoidf_(a) * s.v[i]; }
  __syncthreads();
  wgrad_rows(s.dx, 32, 32, s.hh, H, H, 16, g_w3);
  lin16_t<0>(s.dx, 32, 32, w3, l88, H, s.dhh, H);
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * H; i += blockDim.x) {
    const float a = s.u[i], sg = sigmoidf_(a), dh = s.dhh[i];
    s.u[i] = dh * s.v[i] * sg * (1.f + a * (1.f - sg));   // du
    s.v[i] = dh * a * sg;                                 // dv
  }
  __syncthreads();
  wgrad_rows(s.u, H, H, s.xn, 32, 32, 16, g_w1);
  wgrad_rows(s.v, H, H, s.xn, 32, 32, 16, g_w2);
  lin16_t<0>(s.u, H, H, w1, l32, 32, s.dao, 32);
  __syncthreads();
  lin16_t<1>(s.v, H, H, w2, l32, 32, s.dao, 32);
  __syncthreads();
  ln16_bwd<1>(s.dao, 32, xin, 32, 32, ln2w, eps, s.dx, 32, g_ln2w, g_ln2b);
  __syncthreads();
}

// A Block's weights staged in shared memory with padded rows (16-byte row reads of 32 different rows are then conflict-free per
// quarter warp): one bulk of coalesced L2 reads per block instead of latency-bound weight reads inside every small GEMM
constexpr int SLD32 = 36, SLD88 = 92;
constexpr int S_LN1W = 0, S_LN1B = 32, S_CATTN = 64, S_CPROJ = S_CATTN + 96 * SLD32, S_LN2W = S_CPROJ + 32 * SLD32, S_LN2B = S_LN2W + 32,
              S_W1 = S_LN2B + 32, S_W2 = S_W1 + H * SLD32, S_W3 = S_W2 + H * SLD32, S_SIZE = S_W3 + 32 * SLD88;
__device__ __forceinline__ void stage_block(const float* __restrict__ bp, float* sw) {
  for (int i = threadIdx.x; i < B_SIZE / 4; i += blockDim.x) {
    const int e = i * 4;
    const float4 v = *reinterpret_cast<const float4*>(bp + e);
    int dst;
    if (e < B_CATTN) dst = e;                                                                   // ln_1 weight | bias
    else if (e < B_CPROJ) dst = S_CATTN + ((e - B_CATTN) >> 5) * SLD32 + ((e - B_CATTN) & 31);
    else if (e < B_LN2W) dst = S_CPROJ + ((e - B_CPROJ) >> 5) * SLD32 + ((e - B_CPROJ) & 31);
    else if (e < B_W1) dst = S_LN2W + (e - B_LN2W);
    else if (e < B_W2) dst = S_W1 + ((e - B_W1) >> 5) * SLD32 + ((e - B_W1) & 31);
    else if (e < B_W3) dst = S_W2 + ((e - B_W2) >> 5) * SLD32 + ((e - B_W2) & 31);
    else dst = S_W3 + ((e - B_W3) / H) * SLD88 + ((e - B_W3) % H);
    *reinterpret_cast<float4*>(sw + dst) = v;
  }
  __syncthreads();
}
__device__ __forceinline__ void block16_fwd(LatS& s, const float* bp, float* sw, float eps) {
  stage_block(bp, sw);
  ln16(s.x, 32, 32, sw + S_LN1W, sw + S_LN1B, eps, s.xn, 32);
  __syncthreads();
  lin16<0>(s.xn, 32, 32, sw + S_CATTN, SLD32, 96, s.qkv, 96);
  __syncthreads();
  attn16_fwd(s.qkv, s.ao, nullptr);
  __syncthreads();
  lin16<1>(s.ao, 32, 32, sw + S_CPROJ, SLD32, 32, s.x, 32);
  __syncthreads();
  mlp16_fwd(s, sw + S_LN2W, sw + S_LN2B, sw + S_W1, sw + S_W2, sw + S_W3, SLD32, SLD88, eps);
}
// s.x = the block's input, s.dx = gradient of its output -> s.dx = gradient of its input; weight gradients into gp
__device__ __forceinline__ void block16_bwd(LatS& s, const float* bp, float* sw, float* gp, float eps) {
  stage_block(bp, sw);
  ln16(s.x, 32, 32, sw + S_LN1W, sw + S_LN1B, eps, s.xn, 32);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) s.xm[i] = s.x[i];
  __syncthreads();
  lin16<0>(s.xn, 32, 32, sw + S_CATTN, SLD32, 96, s.qkv, 96);
  __syncthreads();
  attn16_fwd(s.qkv, s.ao, s.P);
  __syncthreads();
  lin16<1>(s.ao, 32, 32, sw + S_CPROJ, SLD32, 32, s.xm, 32);   // xm = x + attention
  __syncthreads();
  mlp16_bwd(s, s.xm, sw + S_LN2W, sw + S_LN2B, sw + S_W1, sw + S_W2, sw + S_W3, SLD32, SLD88, gp + B_LN2W, gp + B_LN2B, gp + B_W1, gp + B_W2,
            gp + B_W3, eps);
  // s.dx = d xm.  attention half (s.xn was overwritten by the MLP's LN2 output: recompute LN1)
  ln16(s.x, 32, 32, sw + S_LN1W, sw + S_LN1B, eps, s.xn, 32);
  wgrad_rows(s.dx, 32, 32, s.ao, 32, 32, 16, gp + B_CPROJ);
  lin16_t<0>(s.dx, 32, 32, sw + S_CPROJ, SLD32, 32, s.dao, 32);
  __syncthreads();
  attn16_bwd(s.qkv, s.P, s.dao, s.dS, s.dt);
  wgrad_rows(s.dt, 96, 96, s.xn, 32, 32, 16, gp + B_CATTN);
  lin16_t<0>(s.dt, 96, 96, sw + S_CATTN, SLD32, 32, s.dao, 32);
  __syncthreads();
  ln16_bwd<1>(s.dao, 32, s.x, 32, 32, sw + S_LN1W, eps, s.dx, 32, gp + B_LN1W, gp + B_LN1B);
  __syncthreads();
}

struct LatParams {
  const float* params; float* grads;
  long long enc_ca, dec_ca, inducing, enc_blocks, dec_blocks, enc_lat, dec_lat;   // element offsets
  const float* pos;         // encoder.pos_embed [16][32] or nullptr (frozen in the reference, nnets.py:103-106)
  int n_layer; float eps;
  const float* ao_enc;      // [B][16][32] pooled attention output of the encoder MCAB
  float* x1_enc;            // [B][16][32] inducing + c_proj(ao)
  float* xe;                // [n_layer + 1][B][16][32] encoder block inputs / output
  float* hlat;              // [B][16][16] encoder_latent_input Linear output (before the LayerNorm)
  float* z;                 // [B][16][16] latents
  float* xd;                // [n_layer + 1][B][16][32]
  float* kdec; float* vdec; // [B][16][32] keys / values of the decoder MCAB
  const float* dkdec; const float* dvdec;
  float* dao_enc;           // [B][16][32]
  int B;
};

__device__ __forceinline__ void tile_load(float* dst, const float* __restrict__ src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
__device__ __forceinline__ void tile_store(float* __restrict__ dst, const float* src, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// One CTA per cell: encoder MCAB tail -> encoder Blocks -> latent projection + LN -> decoder front -> decoder Blocks -> K, V
__global__ void __launch_bounds__(256) latent_fwd_kernel(const LatParams p) {
  extern __shared__ float4 lat_smem4[];
  LatS& s = *reinterpret_cast<LatS*>(lat_smem4);
  float* sw = reinterpret_cast<float*>(lat_smem4) + sizeof(LatS) / 4;
  const int b = blockIdx.x;
  const size_t cell = (size_t)b * 512, lay = (size_t)p.B * 512;
  const float* eca = p.params + p.enc_ca;
  tile_load(s.ao, p.ao_enc + cell, 512);
  tile_load(s.x, p.params + p.inducing, 512);
  __syncthreads();
  lin16<1>(s.ao, 32, 32, eca + C_CPROJ, 32, 32, s.x, 32);   // x1 = q + attention (the residual is the raw query, layers.py:327)
  __syncthreads();
  tile_store(p.x1_enc + cell, s.x, 512);
  mlp16_fwd(s, eca + C_LN2W, eca + C_LN2B, eca + C_W1, eca + C_W2, eca + C_W3, 32, H, p.eps);
  if (p.pos) { for (int i = threadIdx.x; i < 512; i += blockDim.x) s.x[i] += p.pos[i]; __syncthreads(); }
  tile_store(p.xe + cell, s.x, 512);
  for (int l = 0; l < p.n_layer; ++l) {
    block16_fwd(s, p.params + p.enc_blocks + (size_t)l * B_SIZE, sw, p.eps);
    tile_store(p.xe + (size_t)(l + 1) * lay + cell, s.x, 512);
  }
  lin16<0>(s.x, 32, 32, p.params + p.enc_lat, 32, LAT, s.ao, LAT);
  __syncthreads();
  tile_store(p.hlat + (size_t)b * 256, s.ao, 256);
  ln16(s.ao, LAT, LAT, nullptr, nullptr, p.eps, s.xn, LAT);
  __syncthreads();
  tile_store(p.z + (size_t)b * 256, s.xn, 256);
  ln16(s.xn, LAT, LAT, nullptr, nullptr, p.eps, s.ao, LAT);     // the decoder normalises the latents again (nnets.py:203)
  __syncthreads();
  lin16<0>(s.ao, LAT, LAT, p.params + p.dec_lat, LAT, 32, s.x, 32);
  __syncthreads();
  tile_store(p.xd + cell, s.x, 512);
  for (int l = 0; l < p.n_layer; ++l) {
    block16_fwd(s, p.params + p.dec_blocks + (size_t)l * B_SIZE, sw, p.eps);
    tile_store(p.xd + (size_t)(l + 1) * lay + cell, s.x, 512);
  }
  const float* dca = p.params + p.dec_ca;
  ln16(s.x, 32, 32, dca + C_LN1W, dca + C_LN1B, p.eps, s.xn, 32);
  __syncthreads();
  lin16<0>(s.xn, 32, 32, dca + C_CATTN, 32, 64, s.qkv, 64);   // k first, then v (layers.py:252)
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const int j = i >> 5, c = i & 31;
    p.kdec[cell + i] = s.qkv[j * 64 + c];
    p.vdec[cell + i] = s.qkv[j * 64 + 32 + c];
  }
}

__global__ void __launch_bounds__(256) latent_bwd_kernel(const LatParams p) {
  extern __shared__ float4 lat_smem4[];
  LatS& s = *reinterpret_cast<LatS*>(lat_smem4);
  float* sw = reinterpret_cast<float*>(lat_smem4) + sizeof(LatS) / 4;
  const int b = blockIdx.x;
  const size_t cell = (size_t)b * 512, lay = (size_t)p.B * 512;
  const float* dca = p.params + p.dec_ca;
  float* gdca = p.grads + p.dec_ca;
  // K, V of the decoder MCAB = c_attn(LN1(latents))
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    const int j = i >> 5, c = i & 31;
    s.dt[j * 64 + c] = p.dkdec[cell + i];
    s.dt[j * 64 + 32 + c] = p.dvdec[cell + i];
  }
  tile_load(s.x, p.xd + (size_t)p.n_layer * lay + cell, 512);
  __syncthreads();
  ln16(s.x, 32, 32, dca + C_LN1W, dca + C_LN1B, p.eps, s.xn, 32);
  __syncthreads();
  wgrad_rows(s.dt, 64, 64, s.xn, 32, 32, 16, gdca + C_CATTN);
  lin16_t<0>(s.dt, 64, 64, dca + C_CATTN, 32, 32, s.dao, 32);
  __syncthreads();
  ln16_bwd<0>(s.dao, 32, s.x, 32, 32, dca + C_LN1W, p.eps, s.dx, 32, gdca + C_LN1W, gdca + C_LN1B);
  __syncthreads();
  for (int l = p.n_layer - 1; l >= 0; --l) {
    tile_load(s.x, p.xd + (size_t)l * lay + cell, 512);
    __syncthreads();
    block16_bwd(s, p.params + p.dec_blocks + (size_t)l * B_SIZE, sw, p.grads + p.dec_blocks + (size_t)l * B_SIZE, p.eps);
  }
  // decoder front: x0 = W_d LN(z)
  tile_load(s.qkv, p.z + (size_t)b * 256, 256);                 // z
  __syncthreads();
  ln16(s.qkv, LAT, LAT, nullptr, nullptr, p.eps, s.qkv + 256, LAT);   // LN(z)
  __syncthreads();
  wgrad_rows(s.dx, 32, 32, s.qkv + 256, LAT, LAT, 16, p.grads + p.dec_lat);
  lin16_t<0>(s.dx, 32, 32, p.params + p.dec_lat, LAT, LAT, s.qkv + 512, LAT);   // d LN(z)
  __syncthreads();
  ln16_bwd<0>(s.qkv + 512, LAT, s.qkv, LAT, LAT, nullptr, p.eps, s.qkv + 768, LAT, nullptr, nullptr);   // dz
  tile_load(s.ao, p.hlat + (size_t)b * 256, 256);
  __syncthreads();
  ln16_bwd<0>(s.qkv + 768, LAT, s.ao, LAT, LAT, nullptr, p.eps, s.qkv + 1024, LAT, nullptr, nullptr);   // d hlat
  tile_load(s.x, p.xe + (size_t)p.n_layer * lay + cell, 512);
  __syncthreads();
  wgrad_rows(s.qkv + 1024, LAT, LAT, s.x, 32, 32, 16, p.grads + p.enc_lat);
  lin16_t<0>(s.qkv + 1024, LAT, LAT, p.params + p.enc_lat, 32, 32, s.dx, 32);
  __syncthreads();
  for (int l = p.n_layer - 1; l >= 0; --l) {
    tile_load(s.x, p.xe + (size_t)l * lay + cell, 512);
    __syncthreads();
    block16_bwd(s, p.params + p.enc_blocks + (size_t)l * B_SIZE, sw, p.grads + p.enc_blocks + (size_t)l * B_SIZE, p.eps);
  }
  // encoder MCAB tail: x2 = x1 + MLP(LN2(x1)), x1 = inducing + c_proj(ao)
  const float* eca = p.params + p.enc_ca;
  float* geca = p.grads + p.enc_ca;
  tile_load(s.xm, p.x1_enc + cell, 512);
  __syncthreads();
  mlp16_bwd(s, s.xm, eca + C_LN2W, eca + C_LN2B, eca + C_W1, eca + C_W2, eca + C_W3, 32, H, geca + C_LN2W, geca + C_LN2B, geca + C_W1, geca + C_W2,
            geca + C_W3, p.eps);
  for (int i = threadIdx.x; i < 512; i += blockDim.x) atomicAdd(p.grads + p.inducing + i, s.dx[i]);
  tile_load(s.ao, p.ao_enc + cell, 512);
  __syncthreads();
  wgrad_rows(s.dx, 32, 32, s.ao, 32, 32, 16, geca + C_CPROJ);
  lin16_t<0>(s.dx, 32, 32, eca + C_CPROJ, 32, 32, s.dao, 32);
  __syncthreads();
  tile_store(p.dao_enc + cell, s.dao, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// TF32 mma.sync helpers (shared by the encoder-token backward and the decoder tile kernel)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// Operand convention of the tile kernel: with EXACT every product is 3 x TF32 on raw fp32 operands (hi / lo split at load time); without,
// every mma operand is rounded to TF32 ONCE, where it is stored to shared memory (rt<EXACT>), and the fragment loads feed the bits through.
template <bool EXACT>
__device__ __forceinline__ float rt(float x) { return EXACT ? x : __uint_as_float(to_tf32(x)); }

// one m16n8k8 product on fp32 fragments (TF32, or 3 x TF32 when EXACT)
template <bool EXACT>
__device__ __forceinline__ void mma_f(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
  uint32_t ah[4], bh[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) ah[i] = EXACT ? to_tf32(a[i]) : __float_as_uint(a[i]);
  bh[0] = EXACT ? to_tf32(b[0]) : __float_as_uint(b[0]); bh[1] = EXACT ? to_tf32(b[1]) : __float_as_uint(b[1]);
  if (EXACT) {
    uint32_t al[4], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) al[i] = to_tf32(a[i] - __uint_as_float(ah[i]));
    bl[0] = to_tf32(b[0] - __uint_as_float(bh[0])); bl[1] = to_tf32(b[1] - __uint_as_float(bh[1]));
    mma_tf32(c, al, bh);
    mma_tf32(c, ah, bl);
  }
  mma_tf32(c, ah, bh);
}

// acc[i] (16 x 8 tile i) += A[16 x 8 KS] * B[8 KS x 8] for `nt` column tiles; element strides: A(r, k) = A[r sar + k sac],
// B(k, n) = Bm[k sbr + n sbc]; column tile i starts at n = i * nstep.  EXACT: 3 x TF32 (hi / lo split), fp32-grade products.
template <bool EXACT, int NT, int KS>
__device__ __forceinline__ void warp_gemm(float (*acc)[4], const float* A, int sar, int sac, const float* Bm, int sbr, int sbc, int nt,
                                          int nstep) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int k0 = ks * 8;
    float af[4];
    af[0] = A[g * sar + (k0 + t) * sac];
    af[1] = A[(g + 8) * sar + (k0 + t) * sac];
    af[2] = A[g * sar + (k0 + t + 4) * sac];
    af[3] = A[(g + 8) * sar + (k0 + t + 4) * sac];
    uint32_t ah[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ah[i] = EXACT ? to_tf32(af[i]) : __float_as_uint(af[i]); if (EXACT) al[i] = to_tf32(af[i] - __uint_as_float(ah[i])); }
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      if (i < nt) {
        const float* bp = Bm + (i * nstep + g) * sbc;
        const float b0 = bp[(k0 + t) * sbr], b1 = bp[(k0 + t + 4) * sbr];
        uint32_t bh[2] = {EXACT ? to_tf32(b0) : __float_as_uint(b0), EXACT ? to_tf32(b1) : __float_as_uint(b1)};
        if (EXACT) {
          uint32_t bl[2] = {to_tf32(b0 - __uint_as_float(bh[0])), to_tf32(b1 - __uint_as_float(bh[1]))};
          mma_tf32(acc[i], al, bh);
          mma_tf32(acc[i], ah, bl);
        }
        mma_tf32(acc[i], ah, bh);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Query side of a CrossAttentionBlock: Q = c_attn_q(LN1q(row)); rows = inducing points (encoder) or emb[gene] (decoder).
// Cell-invariant, so it is computed once per step and its backward once on the gradients summed over cells.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) qside_fwd_kernel(const float* __restrict__ rows, const long long* __restrict__ ids, int n,
                                                        const float* __restrict__ ca, float eps, float* __restrict__ Q) {
  __shared__ float sW[1024], sw[32], sb[32];
  for (int i = threadIdx.x; i < 1024; i += 128) sW[i] = ca[C_CATTNQ + i];
  if (threadIdx.x < 32) { sw[threadIdx.x] = ca[C_LN1QW + threadIdx.x]; sb[threadIdx.x] = ca[C_LN1QB + threadIdx.x]; }
  __syncthreads();
  const int r = blockIdx.x * 128 + threadIdx.x;
  if (r >= n) return;
  const float* src = rows + (size_t)(ids ? ids[r] : r) * 32;
  float x[32];
#pragma unroll
  for (int k = 0; k < 32; k += 4) { const float4 v = *reinterpret_cast<const float4*>(src + k); x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w; }
  float mean = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) mean += x[k];
  mean *= (1.f / 32.f);
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) { x[k] -= mean; var += x[k] * x[k]; }
  const float rstd = rsqrtf(var * (1.f / 32.f) + eps);
#pragma unroll
  for (int k = 0; k < 32; ++k) x[k] = x[k] * rstd * sw[k] + sb[k];
#pragma unroll 4
  for (int o = 0; o < 32; ++o) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) a += sW[o * 32 + k] * x[k];
    Q[(size_t)r * 32 + o] = a;
  }
}

// dQ [n][32] (summed over cells) -> gradients of c_attn_q, ln_1q and of the rows themselves (+ dres, the residual path of the
// block: x = q + attention); d_rows[id] += ...
__global__ void __launch_bounds__(128) qside_bwd_kernel(const float* __restrict__ rows, const long long* __restrict__ ids, int n,
                                                        const float* __restrict__ ca, float eps, const float* __restrict__ dQ,
                                                        const float* __restrict__ dres, float* d_rows, float* gca) {
  __shared__ float sW[1024], sw[32];
  __shared__ float t1[128 * 33], t2[128 * 33];
  for (int i = threadIdx.x; i < 1024; i += 128) sW[i] = ca[C_CATTNQ + i];
  if (threadIdx.x < 32) sw[threadIdx.x] = ca[C_LN1QW + threadIdx.x];
  __syncthreads();
  const int r = blockIdx.x * 128 + threadIdx.x;
  const bool valid = r < n;
  const long long id = valid ? (ids ? ids[r] : r) : 0;
  const float* src = rows + (size_t)id * 32;
  float x[32], dq[32];
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(src + k);
    x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w;
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) dq[k] = valid ? dQ[(size_t)r * 32 + k] : 0.f;
  float mean = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) mean += x[k];
  mean *= (1.f / 32.f);
  float var = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) { x[k] -= mean; var += x[k] * x[k]; }
  const float rstd = rsqrtf(var * (1.f / 32.f) + eps);
#pragma unroll
  for (int k = 0; k < 32; ++k) x[k] *= rstd;                      // xhat
  // d(LN output) = Wq^T dq
  float m1 = 0.f, m2 = 0.f;
  float g[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    float a = 0.f;
#pragma unroll
    for (int o = 0; o < 32; ++o) a += sW[o * 32 + k] * dq[o];
    t1[threadIdx.x * 33 + k] = a * x[k];     // -> d ln_1q.weight
    t2[threadIdx.x * 33 + k] = a;            // -> d ln_1q.bias
    g[k] = a * sw[k];
    m1 += g[k];
    m2 += g[k] * x[k];
  }
  m1 *= (1.f / 32.f); m2 *= (1.f / 32.f);
  if (valid) {
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
      float dxv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) dxv[q] = rstd * (g[k + q] - m1 - x[k + q] * m2) + (dres ? dres[(size_t)r * 32 + k + q] : 0.f);
      red_add_v4(d_rows + (size_t)id * 32 + k, dxv[0], dxv[1], dxv[2], dxv[3]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int k = threadIdx.x & 31;
    const float* t = threadIdx.x < 32 ? t1 : t2;
    float a = 0.f;
    for (int i = 0; i < 128; ++i) a += t[i * 33 + k];
    atomicAdd(gca + (threadIdx.x < 32 ? C_LN1QW : C_LN1QB) + k, a);
  }
  __syncthreads();
  // c_attn_q weight gradient: sum over rows of dq (x) LN1q(row)
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    t1[threadIdx.x * 33 + k] = dq[k];
    t2[threadIdx.x * 33 + k] = valid ? x[k] * sw[k] + ca[C_LN1QB + k] : 0.f;
  }
  __syncthreads();
  wgrad_rows(t1, 33, 32, t2, 33, 32, 128, gca + C_CATTNQ);
}

// ---------------------------------------------------------------------------------------------------------------
// Encoder tokens: x = emb[gene] * f(count) -> LN1 -> K, V = c_attn -> pooling by the inducing-point queries (layers.py:97-118, 248-264)
// ---------------------------------------------------------------------------------------------------------------
struct EncTokParams {
  const float* emb; const long long* genes; const float* counts; int agg; long long n_tok;
  const float* ca; float eps;
  float* K; float* V;      // [n_tok][32]
  float* stats;            // [n_tok][2] mean, rstd of LN1
};
// Merge of the per-CTA pooling states (max, sum, acc[8] per (query, head)) of a cell left by enc_fused_fwd_kernel: the 16 inducing-point
// queries x 4 heads attend to ALL S tokens of a cell (no key masking); also leaves the log-sum-exp for the backward.
__global__ void __launch_bounds__(64) enc_pool_merge_kernel(const float* __restrict__ part, int n_parts, float* __restrict__ AO,
                                                            float* __restrict__ lse) {
  const int b = blockIdx.x, pair = threadIdx.x, m = pair >> 2, h = pair & 3;
  const float* src = part + ((size_t)b * n_parts * 64 + pair) * 10;
  float gm = -1e30f;
  for (int i = 0; i < n_parts; ++i) gm = fmaxf(gm, src[i * 640]);
  float gl = 0.f, ga[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < n_parts; ++i) {
    const float c = __expf(src[i * 640] - gm);
    gl += src[i * 640 + 1] * c;
#pragma unroll
    for (int d = 0; d < 8; ++d) ga[d] += src[i * 640 + 2 + d] * c;
  }
  const float il = 1.f / gl;
#pragma unroll
  for (int d = 0; d < 8; ++d) AO[(size_t)b * 512 + m * 32 + h * 8 + d] = ga[d] * il;
  lse[b * 64 + pair] = gm + __logf(gl);
}


// Fused forward of the encoder's token side on mma.sync: a CTA walks ENC_FWD_TILES consecutive 128-token tiles of its cell (4: 0.37 ms, 8: 0.39, 16: 0.39 at the census shape); per tile
//   emb x f(count) -> LN1 (thread per token) -> [K | V] = xn Wkv^T (TF32 mma; written to HBM for the backward AND kept in shared memory)
//   -> S^T = K_h Q_h^T, running max / sum per (query, head) across the tiles, P tile -> AO_h += P_h^T V_h (mma over the tokens).
// It leaves one unnormalised (max, sum, acc) state per (query, head) and CTA; enc_pool_merge_kernel merges the CTAs of a cell.
constexpr int ENC_FWD_TILES = 4;
constexpr int EF_LD = 40, EF_LDP = 68;
constexpr int ENC_FWD_SMEM_FLOATS = 3 * 128 * EF_LD + 128 * EF_LDP + 64 * EF_LD + 16 * 36 + 64 + 4 * 64 + 4 * 64 + 3 * 64;
template <bool EXACT>
__global__ void __launch_bounds__(128) enc_fused_fwd_kernel(const EncTokParams p, const float* __restrict__ Q, int S, float* __restrict__ part) {
  extern __shared__ float4 encf_smem4[];
  float* sm = reinterpret_cast<float*>(encf_smem4);
  float* tXN = sm;                      // [128][40] LN1 output
  float* tK = tXN + 128 * EF_LD;        // [128][40]
  float* tV = tK + 128 * EF_LD;         // [128][40]
  float* tP = tV + 128 * EF_LD;         // [128][68] softmax numerators, column = query * 4 + head
  float* sW = tP + 128 * EF_LDP;        // c_attn [64][40]
  float* sQ = sW + 64 * EF_LD;          // [16][36] scaled queries
  float* sLn = sQ + 16 * 36;            // ln_1 weight | bias
  float* sTMax = sLn + 64;              // [4 warps][64] column maxima of the warps' tokens
  float* sTSum = sTMax + 256;           // [4 warps][64] column sums
  float* sMax = sTSum + 256;            // [64] running maximum per (query, head)
  float* sSum = sMax + 64;              // [64] running sum
  float* sScale = sSum + 64;            // [64] exp(old max - new max) of this tile
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 2048; i += 128) sW[(i >> 5) * EF_LD + (i & 31)] = rt<EXACT>(p.ca[C_CATTN + i]);
  for (int i = tid; i < 512; i += 128) sQ[(i >> 5) * 36 + (i & 31)] = rt<EXACT>(Q[i] * 0.35355339059327373f);
  if (tid < 64) { sLn[tid] = p.ca[C_LN1W + tid]; sMax[tid] = -1e30f; sSum[tid] = 0.f; }
  float acc_o[1][4] = {};               // warp h: AO_h[query g (+8)][d = 2 t (+1)], unnormalised
  const int n_tiles = (S + 127) / 128;
  __syncthreads();
  for (int tile = blockIdx.x * ENC_FWD_TILES; tile < min(n_tiles, (blockIdx.x + 1) * ENC_FWD_TILES); ++tile) {
    // ---- token phase: LN1 of emb x f(count) ----
    {
      const int s_ = tile * 128 + tid;
      const bool valid = s_ < S;
      const long long tk = (long long)b * S + (valid ? s_ : 0);
      const float f = vae::count_scale(valid ? p.counts[tk] : 0.f, p.agg);
      const float* src = p.emb + (size_t)(valid ? p.genes[tk] : 0) * 32;
      float x[32];
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(src + k);
        x[k] = v.x * f; x[k + 1] = v.y * f; x[k + 2] = v.z * f; x[k + 3] = v.w * f;
      }
      float mean = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) mean += x[k];
      mean *= (1.f / 32.f);
      float var = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) { x[k] -= mean; var += x[k] * x[k]; }
      const float rstd = rsqrtf(var * (1.f / 32.f) + p.eps);
      if (valid) { p.stats[tk * 2] = mean; p.stats[tk * 2 + 1] = rstd; }
#pragma unroll
      for (int k = 0; k < 32; ++k) tXN[tid * EF_LD + k] = valid ? rt<EXACT>(x[k] * rstd * sLn[k] + sLn[32 + k]) : 0.f;
    }
    __syncthreads();
    // ---- [K | V] of the warp's 32 tokens ----
#pragma unroll 1
    for (int mtile = 0; mtile < 2; ++mtile) {
      const int r0 = warp * 32 + mtile * 16;
      float acc[8][4] = {};
      warp_gemm<EXACT, 8, 4>(acc, tXN + r0 * EF_LD, EF_LD, 1, sW, 1, EF_LD, 8, 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int row = r0 + g + hh * 8, s_ = tile * 128 + row, col = (i & 3) * 8 + 2 * t;
          const bool valid = s_ < S;
          float* dt = (i < 4 ? tK : tV) + row * EF_LD + col;
          const float a0 = acc[i][hh * 2], a1 = acc[i][hh * 2 + 1];
          dt[0] = valid ? rt<EXACT>(a0) : 0.f; dt[1] = valid ? rt<EXACT>(a1) : 0.f;
          if (valid) *reinterpret_cast<float2*>((i < 4 ? p.K : p.V) + ((long long)b * S + s_) * 32 + col) = make_float2(a0, a1);
        }
      }
    }
    __syncwarp();
    // ---- scores of the warp's tokens against the 16 queries, per head; column maxima ----
    float sc[2][4][2][4];     // [row tile][head][query half][fragment]
#pragma unroll
    for (int mtile = 0; mtile < 2; ++mtile) {
      const int r0 = warp * 32 + mtile * 16;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float* kr = tK + r0 * EF_LD + h * 8;
        const float ak[4] = {kr[g * EF_LD + t], kr[(g + 8) * EF_LD + t], kr[g * EF_LD + t + 4], kr[(g + 8) * EF_LD + t + 4]};
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const float bq[2] = {sQ[(8 * n + g) * 36 + h * 8 + t], sQ[(8 * n + g) * 36 + h * 8 + t + 4]};
          sc[mtile][h][n][0] = sc[mtile][h][n][1] = sc[mtile][h][n][2] = sc[mtile][h][n][3] = 0.f;
          mma_f<EXACT>(sc[mtile][h][n], ak, bq);
        }
      }
    }
    const bool v00 = tile * 128 + warp * 32 + g < S, v01 = tile * 128 + warp * 32 + g + 8 < S;
    const bool v10 = tile * 128 + warp * 32 + 16 + g < S, v11 = tile * 128 + warp * 32 + 24 + g < S;
#pragma unroll
    for (int h = 0; h < 4; ++h) {
#pragma unroll
      for (int n = 0; n < 2; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {     // column m = 8 n + 2 t + e: maximum over this thread's four rows, then over the eight row groups g
          float mx = fmaxf(fmaxf(v00 ? sc[0][h][n][e] : -1e30f, v01 ? sc[0][h][n][2 + e] : -1e30f),
                           fmaxf(v10 ? sc[1][h][n][e] : -1e30f, v11 ? sc[1][h][n][2 + e] : -1e30f));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
          if (g == 0) sTMax[warp * 64 + (8 * n + 2 * t + e) * 4 + h] = mx;
        }
      }
    }
    __syncthreads();
    if (tid < 64) {
      const float old = sMax[tid];
      const float nm = fmaxf(fmaxf(old, fmaxf(sTMax[tid], sTMax[64 + tid])), fmaxf(sTMax[128 + tid], sTMax[192 + tid]));
      sScale[tid] = __expf(old - nm);
      sMax[tid] = nm;
    }
    __syncthreads();
    // ---- numerators -> P tile, column sums ----
#pragma unroll
    for (int h = 0; h < 4; ++h) {
#pragma unroll
      for (int n = 0; n < 2; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = (8 * n + 2 * t + e) * 4 + h;
          const float nm = sMax[col];
          const float p00 = v00 ? __expf(sc[0][h][n][e] - nm) : 0.f, p01 = v01 ? __expf(sc[0][h][n][2 + e] - nm) : 0.f;
          const float p10 = v10 ? __expf(sc[1][h][n][e] - nm) : 0.f, p11 = v11 ? __expf(sc[1][h][n][2 + e] - nm) : 0.f;
          const int r0 = warp * 32 + g;
          tP[r0 * EF_LDP + col] = rt<EXACT>(p00); tP[(r0 + 8) * EF_LDP + col] = rt<EXACT>(p01);
          tP[(r0 + 16) * EF_LDP + col] = rt<EXACT>(p10); tP[(r0 + 24) * EF_LDP + col] = rt<EXACT>(p11);
          float sum = p00 + p01 + p10 + p11;
          sum += __shfl_xor_sync(0xffffffffu, sum, 4); sum += __shfl_xor_sync(0xffffffffu, sum, 8); sum += __shfl_xor_sync(0xffffffffu, sum, 16);
          if (g == 0) sTSum[warp * 64 + col] = sum;
        }
      }
    }
    __syncthreads();
    if (tid < 64) sSum[tid] = sSum[tid] * sScale[tid] + sTSum[tid] + sTSum[64 + tid] + sTSum[128 + tid] + sTSum[192 + tid];
    // ---- AO_h = AO_h * scale + P_h^T V_h over the tile's 128 tokens: warp h ----
    {
      const float s0 = sScale[g * 4 + warp], s1 = sScale[(g + 8) * 4 + warp];
      acc_o[0][0] *= s0; acc_o[0][1] *= s0; acc_o[0][2] *= s1; acc_o[0][3] *= s1;
      warp_gemm<EXACT, 1, 16>(acc_o, tP + warp, 4, EF_LDP, tV + warp * 8, EF_LD, 1, 1, 8);
    }
    __syncthreads();   // the tiles are rewritten by the next trip
  }
  // ---- this CTA's state per (query, head): max | sum | acc[8] ----
  {
    float* dst = part + (((size_t)b * gridDim.x + blockIdx.x) * 64) * 10;
    const int h = warp;
#pragma unroll
    for (int e = 0; e < 4; ++e) dst[((g + (e >> 1) * 8) * 4 + h) * 10 + 2 + 2 * t + (e & 1)] = acc_o[0][e];
    if (tid < 64) { dst[tid * 10] = sMax[tid]; dst[tid * 10 + 1] = sSum[tid]; }
  }
}

// Backward of the pooling and of the token side, one thread per token of a 128-token tile of one cell:
// dAO -> (dK, dV of the token) -> c_attn^T -> LN1 backward -> d emb[gene] (scatter) ; weight gradients of c_attn / ln_1 and
// the query gradient dQ (summed over tokens and cells) through shared-memory tiles.
struct EncBwdParams {
  const float* emb; const long long* genes; const float* counts; int agg; int S;
  const float* ca; float* gca; float eps;
  const float* Q;            // [16][32]
  const float* K; const float* V; const float* stats; const float* lse;   // saved by the forward
  const float* AO; const float* dAO;     // [B][16][32]
  float* dQ;                 // [16][32] accumulated (atomics)
  float* g_emb;              // gradient of the embedding table
};
constexpr int ELD64 = 68, ELD32 = 36;   // row strides of the 64- / 32-wide token tiles (with the weights staged in a dead tile: 112 KB, 2 CTAs / SM)
constexpr int ENC_TILES_PER_CTA = 8;
constexpr int EQ = 36;                  // row stride of the 16 x 32 query / dAO tiles (conflict-free B-fragment reads)
constexpr int ENC_BWD_SMEM_FLOATS = 128 * ELD64 + 128 * ELD32 + 128 * ELD64 + 128 * ELD32 + 16 * EQ * 2 + 64 + 64 + 64;
template <bool EXACT>
__global__ void __launch_bounds__(128) enc_tokens_bwd_kernel(const EncBwdParams p) {
  extern __shared__ float4 enc_smem4[];
  float* sm = reinterpret_cast<float*>(enc_smem4);
  float* tDKV = sm;                     // [128][68]  dK | dV of the tile's tokens
  float* tXN = tDKV + 128 * ELD64;      // [128][40]  LN1 output
  float* tDS = tXN + 128 * ELD32;       // [128][68]  dS[(query, head)] of the tile's tokens (scaled)
  float* tK = tDS + 128 * ELD64;        // [128][40]
  float* sQ = tK + 128 * ELD32;         // [16][36] scaled queries
  float* sDAO = sQ + 16 * EQ;           // [16][36]
  float* sD = sDAO + 16 * EQ;           // [64] rowsum(dAO * AO) per (query, head)
  float* sLse = sD + 64;                // [64] log-sum-exp of the pooling softmax per (query, head)
  float* sW = tK;                       // c_attn [64][40]: staged into the key tile once the dQ product has consumed it (every trip)
  float* tV = tXN;                      // values of the tile: dead before the LN1 output is written
  float* sLn = sLse + 64;               // ln_1 weight | bias
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 512; i += 128) {
    sQ[(i >> 5) * EQ + (i & 31)] = rt<EXACT>(p.Q[i] * 0.35355339059327373f);
    sDAO[(i >> 5) * EQ + (i & 31)] = rt<EXACT>(p.dAO[(size_t)b * 512 + i]);
  }
  if (threadIdx.x < 64) sLn[threadIdx.x] = p.ca[C_LN1W + threadIdx.x];   // ln_1.weight, ln_1.bias are adjacent
  if (threadIdx.x < 64) {
    const int m = threadIdx.x >> 2, h = threadIdx.x & 3;
    float a = 0.f;
    for (int d = 0; d < 8; ++d) a += p.dAO[(size_t)b * 512 + m * 32 + h * 8 + d] * p.AO[(size_t)b * 512 + m * 32 + h * 8 + d];
    sD[threadIdx.x] = a;
    sLse[threadIdx.x] = p.lse[b * 64 + threadIdx.x];
  }
  // A CTA walks ENC_TILES_PER_CTA consecutive 128-token tiles of its cell; the gradients that are sums over tokens (c_attn, dQ, ln_1)
  // accumulate in registers across the tiles and reach global memory once per CTA (8 x fewer same-address atomics)
  const int warp = threadIdx.x >> 5;
  float acc_w[4][4] = {}, acc_q[1][4] = {}, acc_ln = 0.f;
  const int n_tiles = (p.S + 127) / 128;
  for (int tile = blockIdx.x * ENC_TILES_PER_CTA; tile < min(n_tiles, (blockIdx.x + 1) * ENC_TILES_PER_CTA); ++tile) {
  const int s = tile * 128 + threadIdx.x;
  const bool valid = s < p.S;
  const long long tk = (long long)b * p.S + (valid ? s : 0);
  float kk[32];
  // the tile's keys / values into shared memory (one token row per thread; rows beyond S are zero and masked below)
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    const float4 a = *reinterpret_cast<const float4*>(p.K + tk * 32 + k), c = *reinterpret_cast<const float4*>(p.V + tk * 32 + k);
    float* dk_ = tK + threadIdx.x * ELD32 + k;
    float* dv_ = tV + threadIdx.x * ELD32 + k;
    dk_[0] = valid ? rt<EXACT>(a.x) : 0.f; dk_[1] = valid ? rt<EXACT>(a.y) : 0.f; dk_[2] = valid ? rt<EXACT>(a.z) : 0.f; dk_[3] = valid ? rt<EXACT>(a.w) : 0.f;
    dv_[0] = valid ? rt<EXACT>(c.x) : 0.f; dv_[1] = valid ? rt<EXACT>(c.y) : 0.f; dv_[2] = valid ? rt<EXACT>(c.z) : 0.f; dv_[3] = valid ? rt<EXACT>(c.w) : 0.f;
  }
  __syncthreads();
  // Attention backward of the tile on mma.sync, warp w = tokens 32 w .. 32 w + 31 (two row tiles), per head:
  //   S^T = K_h Q_h^T (16 tokens x 16 queries), P = exp(S - lse), dP^T = V_h dAO_h^T, dS = P (dP - D);
  //   dK_h = dS Q_h, dV_h = P dAO_h (the accumulator fragments feed the A operand through the permuted contraction order)
  {
    const int warp_ = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int mtile = 0; mtile < 2; ++mtile) {
      const int r0 = warp_ * 32 + mtile * 16;
      const bool v_lo = tile * 128 + r0 + g < p.S, v_hi = tile * 128 + r0 + g + 8 < p.S;
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float* kr = tK + r0 * ELD32 + h * 8;
        const float* vr = tV + r0 * ELD32 + h * 8;
        const float ak[4] = {kr[g * ELD32 + t], kr[(g + 8) * ELD32 + t], kr[g * ELD32 + t + 4], kr[(g + 8) * ELD32 + t + 4]};
        const float av[4] = {vr[g * ELD32 + t], vr[(g + 8) * ELD32 + t], vr[g * ELD32 + t + 4], vr[(g + 8) * ELD32 + t + 4]};
        float pr[2][4], ds[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const float bq[2] = {sQ[(8 * n + g) * EQ + h * 8 + t], sQ[(8 * n + g) * EQ + h * 8 + t + 4]};
          const float bd[2] = {sDAO[(8 * n + g) * EQ + h * 8 + t], sDAO[(8 * n + g) * EQ + h * 8 + t + 4]};
          pr[n][0] = pr[n][1] = pr[n][2] = pr[n][3] = 0.f;
          ds[n][0] = ds[n][1] = ds[n][2] = ds[n][3] = 0.f;
          mma_f<EXACT>(pr[n], ak, bq);      // scores (the queries carry the 1 / sqrt(head_dim))
          mma_f<EXACT>(ds[n], av, bd);      // dP
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int m = 8 * n + 2 * t + (e & 1);
            const float pe = ((e >> 1) ? v_hi : v_lo) ? __expf(pr[n][e] - sLse[m * 4 + h]) : 0.f;
            const float de = pe * (ds[n][e] - sD[m * 4 + h]);
            pr[n][e] = pe;
            ds[n][e] = de;
            tDS[(r0 + g + (e >> 1) * 8) * ELD64 + m * 4 + h] = rt<EXACT>(de * 0.35355339059327373f);
          }
        }
        float dk4[4] = {0.f, 0.f, 0.f, 0.f}, dv4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const float a1[4] = {rt<EXACT>(ds[ks][0]), rt<EXACT>(ds[ks][2]), rt<EXACT>(ds[ks][1]), rt<EXACT>(ds[ks][3])};
          const float a2[4] = {rt<EXACT>(pr[ks][0]), rt<EXACT>(pr[ks][2]), rt<EXACT>(pr[ks][1]), rt<EXACT>(pr[ks][3])};
          const float b1[2] = {sQ[(8 * ks + 2 * t) * EQ + h * 8 + g], sQ[(8 * ks + 2 * t + 1) * EQ + h * 8 + g]};
          const float b2[2] = {sDAO[(8 * ks + 2 * t) * EQ + h * 8 + g], sDAO[(8 * ks + 2 * t + 1) * EQ + h * 8 + g]};
          mma_f<EXACT>(dk4, a1, b1);
          mma_f<EXACT>(dv4, a2, b2);
        }
        float* o = tDKV + (r0 + g) * ELD64 + h * 8 + 2 * t;
        o[0] = rt<EXACT>(dk4[0]); o[1] = rt<EXACT>(dk4[1]); o[8 * ELD64] = rt<EXACT>(dk4[2]); o[8 * ELD64 + 1] = rt<EXACT>(dk4[3]);
        o[32] = rt<EXACT>(dv4[0]); o[33] = rt<EXACT>(dv4[1]); o[8 * ELD64 + 32] = rt<EXACT>(dv4[2]); o[8 * ELD64 + 33] = rt<EXACT>(dv4[3]);
      }
    }
  }
  __syncthreads();    // every warp is done with the value tile before the LN1 output overwrites it
  // token side: recompute x, LN1
  const float cnt = valid ? p.counts[tk] : 0.f;
  const float f = vae::count_scale(cnt, p.agg);
  const long long gid = valid ? p.genes[tk] : 0;
  const float mean = p.stats[tk * 2], rstd = p.stats[tk * 2 + 1];
  float xh[32];
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    const float4 v = *reinterpret_cast<const float4*>(p.emb + (size_t)gid * 32 + k);
    xh[k] = (v.x * f - mean) * rstd; xh[k + 1] = (v.y * f - mean) * rstd; xh[k + 2] = (v.z * f - mean) * rstd; xh[k + 3] = (v.w * f - mean) * rstd;
  }
#pragma unroll
  for (int k = 0; k < 32; ++k) tXN[threadIdx.x * ELD32 + k] = valid ? rt<EXACT>(xh[k] * sLn[k] + sLn[32 + k]) : 0.f;
  __syncthreads();
  // c_attn weight gradient (64 x 32 = dKV^T xn over the 128 tokens) and dQ (per head 16 queries x 8 = dS_h^T K_h) on mma.sync:
  // warp w owns weight rows 16 w .. 16 w + 15 and head w
  warp_gemm<EXACT, 4, 16>(acc_w, tDKV + warp * 16, 1, ELD64, tXN, ELD32, 1, 4, 8);
  warp_gemm<EXACT, 1, 16>(acc_q, tDS + warp, 4, ELD64, tK + warp * 8, ELD32, 1, 1, 8);
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += 128) sW[(i >> 5) * 40 + (i & 31)] = rt<EXACT>(p.ca[C_CATTN + i]);
  __syncthreads();
  // d(LN1 output)[token][32] = [dk | dv] Wkv on mma.sync: warp w computes the rows of its own 32 tokens into the (now dead) dS tile
  {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mtile = 0; mtile < 2; ++mtile) {
      const int r0 = warp * 32 + mtile * 16;
      float acc[4][4] = {};
      warp_gemm<EXACT, 4, 8>(acc, tDKV + r0 * ELD64, ELD64, 1, sW, 40, 1, 4, 8);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) tDS[(r0 + g + (e >> 1) * 8) * ELD64 + i * 8 + 2 * t + (e & 1)] = acc[i][e];
      }
    }
  }
  __syncwarp();
  float g[32], m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float a = tDS[threadIdx.x * ELD64 + k];
    kk[k] = a;   // gradient w.r.t. the LN1 output
    g[k] = a * sLn[k];
    m1 += g[k];
    m2 += g[k] * xh[k];
  }
  m1 *= (1.f / 32.f); m2 *= (1.f / 32.f);
  if (valid && f != 0.f) {
    const float sf = rstd * f;
#pragma unroll
    for (int k = 0; k < 32; k += 4)
      red_add_v4(p.g_emb + (size_t)gid * 32 + k, sf * (g[k] - m1 - xh[k] * m2), sf * (g[k + 1] - m1 - xh[k + 1] * m2),
                 sf * (g[k + 2] - m1 - xh[k + 2] * m2), sf * (g[k + 3] - m1 - xh[k + 3] * m2));
  }
  __syncthreads();
  // ln_1 affine gradients: stage (d out * xhat, d out) in the two 33-wide tiles
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    tXN[threadIdx.x * ELD32 + k] = valid ? kk[k] * xh[k] : 0.f;
    tK[threadIdx.x * ELD32 + k] = valid ? kk[k] : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int k = threadIdx.x & 31;
    const float* t = threadIdx.x < 32 ? tXN : tK;
    float a = 0.f;
    for (int i = 0; i < 128; ++i) a += t[i * ELD32 + k];
    acc_ln += a;
  }
  __syncthreads();     // the tiles are rewritten by the next trip
  }
  {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) atomicAdd(p.gca + C_CATTN + (warp * 16 + g + (e >> 1) * 8) * 32 + i * 8 + 2 * t + (e & 1), acc_w[i][e]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) atomicAdd(p.dQ + (g + (e >> 1) * 8) * 32 + warp * 8 + 2 * t + (e & 1), acc_q[0][e]);
    if (threadIdx.x < 64) atomicAdd(p.gca + (threadIdx.x < 32 ? C_LN1W : C_LN1B) + (threadIdx.x & 31), acc_ln);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// NB head: mu = softmax_genes(logit) * library, theta = exp(table[gene]); loss = sum_cells nll * loss_scale; d loss / d logit
// (stochastic_layers.py:102-116, distributions.py:6-42, models.py:233-247).  One CTA per cell.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float digammaf_(float x) {
  float r = 0.f;
  while (x < 6.f) { r -= 1.f / x; x += 1.f; }
  const float i = 1.f / x, i2 = i * i;
  return r + __logf(x) - 0.5f * i - i2 * (1.f / 12.f - i2 * (1.f / 120.f - i2 * (1.f / 252.f)));
}
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) { const float t = __shfl_xor_sync(0xffffffffu, v, o); v = is_max ? fmaxf(v, t) : v + t; }
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}
struct NbLossParams {
  const float* logits; const float* counts; const float* library; const float* theta_tbl; const long long* genes; int G;
  float loss_scale; int backward;
  float* nll;       // [B]
  float* mu;        // nullable [B][G]
  float* dlogit;    // [B][G]
  float* g_theta;   // gradient of the theta table
};
__global__ void __launch_bounds__(512) nb_loss_kernel(const NbLossParams p) {
  __shared__ float red[16];
  const int b = blockIdx.x;
  const float* lg = p.logits + (size_t)b * p.G;
  const float* xc = p.counts + (size_t)b * p.G;
  float mx = -1e30f;
  for (int g = threadIdx.x; g < p.G; g += 512) mx = fmaxf(mx, lg[g]);
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int g = threadIdx.x; g < p.G; g += 512) sum += __expf(lg[g] - mx);
  sum = block_reduce(sum, red, false);
  const float lib = p.library[b], inv = 1.f / sum, eps = 1e-8f;
  float nll = 0.f, sgm = 0.f;
  for (int g = threadIdx.x; g < p.G; g += 512) {
    const float pr = __expf(lg[g] - mx) * inv, mu = lib * pr, x = xc[g];
    const long long gid = p.genes[g];
    const float th = __expf(p.theta_tbl[gid]);
    nll -= vae::nb_logp(x, mu, th);
    if (p.mu) p.mu[(size_t)b * p.G + g] = mu;
    if (p.backward) {
      const float itm = 1.f / (th + mu + eps);
      const float dmu = -(x / (mu + eps) - (th + x) * itm) * p.loss_scale;     // d(-ll)/dmu
      const float gm = dmu * mu;
      p.dlogit[(size_t)b * p.G + g] = gm;
      sgm += gm;
      float dps;   // digamma(x + theta) - digamma(theta)
      const int k = (int)x;
      if ((float)k == x && k <= 16) { dps = 0.f; for (int i = 0; i < k; ++i) dps += 1.f / (th + (float)i); }
      else dps = digammaf_(x + th) - digammaf_(th);
      const float dth = -(__logf(th + eps) - __logf(th + mu + eps) + th / (th + eps) - (th + x) * itm + dps) * p.loss_scale;
      atomicAdd(p.g_theta + gid, dth * th);
    }
  }
  nll = block_reduce(nll, red, false);
  if (threadIdx.x == 0) p.nll[b] = nll;
  if (p.backward) {
    sgm = block_reduce(sgm, red, false);
    for (int g = threadIdx.x; g < p.G; g += 512) {
      const float pr = __expf(lg[g] - mx) * inv;
      p.dlogit[(size_t)b * p.G + g] -= pr * sgm;
    }
  }
}

// Autograd bridge: caller-provided dLoss/dmu (and dLoss/dtheta of the expanded theta) -> dLoss/dlogit, d theta table.
// mu = library * softmax(logit): dlogit = mu dmu - p sum_genes(mu dmu)
struct MuGradParams {
  const float* logits; const float* library; const float* theta_tbl; const long long* genes; int G;
  const float* dmu; const float* dtheta;   // [B][G]; dtheta nullable
  float* dlogit; float* g_theta;
};
__global__ void __launch_bounds__(512) mu_grad_kernel(const MuGradParams p) {
  __shared__ float red[16];
  const int b = blockIdx.x;
  const float* lg = p.logits + (size_t)b * p.G;
  const float* dm = p.dmu + (size_t)b * p.G;
  float mx = -1e30f;
  for (int g = threadIdx.x; g < p.G; g += 512) mx = fmaxf(mx, lg[g]);
  mx = block_reduce(mx, red, true);
  float sum = 0.f;
  for (int g = threadIdx.x; g < p.G; g += 512) sum += __expf(lg[g] - mx);
  sum = block_reduce(sum, red, false);
  const float lib = p.library[b], inv = 1.f / sum;
  float sgm = 0.f;
  for (int g = threadIdx.x; g < p.G; g += 512) {
    const float gm = lib * __expf(lg[g] - mx) * inv * dm[g];
    p.dlogit[(size_t)b * p.G + g] = gm;
    sgm += gm;
    if (p.dtheta) {
      const long long gid = p.genes[g];
      atomicAdd(p.g_theta + gid, p.dtheta[(size_t)b * p.G + g] * __expf(p.theta_tbl[gid]));
    }
  }
  sgm = block_reduce(sgm, red, false);
  for (int g = threadIdx.x; g < p.G; g += 512) p.dlogit[(size_t)b * p.G + g] -= __expf(lg[g] - mx) * inv * sgm;
}

// ---------------------------------------------------------------------------------------------------------------
// Decoder MCAB on (cell, gene) tokens: forward (logits) and, with BWD, the whole backward of the tile.
// ---------------------------------------------------------------------------------------------------------------
// two products that share the A operand (u = n2 w1^T, v = n2 w2^T): the A fragments of a K step are loaded once
template <bool EXACT, int NT, int KS>
__device__ __forceinline__ void warp_gemm2(float (*acc1)[4], float (*acc2)[4], const float* A, int sar, int sac, const float* B1, const float* B2,
                                           int sbr, int sbc, int nt, int nstep) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    const int k0 = ks * 8;
    const float af[4] = {A[g * sar + (k0 + t) * sac], A[(g + 8) * sar + (k0 + t) * sac], A[g * sar + (k0 + t + 4) * sac],
                         A[(g + 8) * sar + (k0 + t + 4) * sac]};
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      if (i < nt) {
        const int o = (i * nstep + g) * sbc;
        const float b1[2] = {B1[o + (k0 + t) * sbr], B1[o + (k0 + t + 4) * sbr]};
        const float b2[2] = {B2[o + (k0 + t) * sbr], B2[o + (k0 + t + 4) * sbr]};
        mma_f<EXACT>(acc1[i], af, b1);
        mma_f<EXACT>(acc2[i], af, b2);
      }
    }
  }
}

struct DecTrainParams {
  const float* ca; float* gca;          // decoder_cross_attention parameter group / its gradients
  const float* emb; const long long* genes; int G;
  const float* Q;                       // [G][32] query-side table (qside_fwd_kernel)
  const float* Kc; const float* Vc;     // [B][16][32]
  const float* head_w; const float* head_b; float* g_head_w; float* g_head_b;
  float* logits;                        // [B][G]   (forward)
  const float* dlogit;                  // [B][G]   (backward)
  float* dK; float* dV;                 // [B][16][32] accumulated (atomics)
  float* dQ; float* dXsum;              // [G][32]   accumulated (atomics): gradient of Q, of the residual path
  int B, cells_per_chunk; float eps;
};
constexpr int DT = 32, LD32 = 36, LD88 = 100, LDK = 36;
// du / dv tiles: 96-float rows with an XOR swizzle of the column inside its 32-column group, so that BOTH fragment orientations are free of
// bank conflicts: token-major rows as the A operand of the d n2 dgrad (rows g, columns k0 + t) and hidden-major columns as the A operand
// of the weight-gradient products (rows k0 + t, columns j0 + g).  Padding alone serves only one of the two (4 (mod 32) vs 8 (mod 32)).
constexpr int LDU = 96;
__device__ __forceinline__ int usw(int r) { return ((r & 3) << 3) | (((r >> 2) & 1) << 2); }
__device__ __forceinline__ int uidx(int r, int c) { return r * LDU + (c ^ usw(r)); }
// 6 32-wide tiles + du / dv (h in the forward-only kernel) + keys / values + small vectors + the block's weights: 103 KB, 2 CTAs per SM
constexpr int DEC_SMEM_FLOATS = 6 * DT * LD32 + 2 * DT * LD88 + 2 * 16 * LDK + 2 * DT + DT + 4 * DT + DT + 32 * LD32 + 2 * H * LD32 + 32 * LD88 + 96 + 96;
// the forward-only launch needs neither the gradient tiles nor mlp.c_proj: 54 KB, three CTAs per SM
constexpr int DEC_FWD_SMEM_FLOATS = 4 * DT * LD32 + 2 * 16 * LDK + 2 * DT + DT + 4 * DT + DT + 32 * LD32 + 2 * H * LD32 + 96 + 96;

// Cross attention of one head for the 16 tokens of a warp's row tile, on mma.sync: S = Q K^T (two 8-key column tiles) -> softmax
// over the 16 keys on the accumulator fragments (a row lives in one quad) -> p[n][e] = P[row g (+8 for e >= 2)][key 8 n + 2 t + (e & 1)].
template <bool EXACT>
__device__ __forceinline__ void attn_probs(const float* sQrow, const float* sK, int h, float scale, float (&p)[2][4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float a[4] = {sQrow[g * LD32 + h * 8 + t], sQrow[(g + 8) * LD32 + h * 8 + t], sQrow[g * LD32 + h * 8 + t + 4], sQrow[(g + 8) * LD32 + h * 8 + t + 4]};
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const float b[2] = {sK[(8 * n + g) * LDK + h * 8 + t], sK[(8 * n + g) * LDK + h * 8 + t + 4]};
    p[n][0] = p[n][1] = p[n][2] = p[n][3] = 0.f;
    mma_f<EXACT>(p[n], a, b);
  }
  float m0 = fmaxf(fmaxf(p[0][0], p[0][1]), fmaxf(p[1][0], p[1][1])), m1 = fmaxf(fmaxf(p[0][2], p[0][3]), fmaxf(p[1][2], p[1][3]));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    p[n][0] = __expf((p[n][0] - m0) * scale); p[n][1] = __expf((p[n][1] - m0) * scale);
    p[n][2] = __expf((p[n][2] - m1) * scale); p[n][3] = __expf((p[n][3] - m1) * scale);
    l0 += p[n][0] + p[n][1];
    l1 += p[n][2] + p[n][3];
  }
  l0 = 1.f / quad_sum(l0); l1 = 1.f / quad_sum(l1);
#pragma unroll
  for (int n = 0; n < 2; ++n) { p[n][0] *= l0; p[n][1] *= l0; p[n][2] *= l1; p[n][3] *= l1; }
}
// out[16 x 8] = W[16 x 16 keys] * X[16 keys][h*8 .. h*8+7], W given as accumulator fragments of attn_probs' layout: the contraction
// index is visited in the permuted order (k = t <-> key 8 ks + 2 t, k = t + 4 <-> key 8 ks + 2 t + 1), so the fragments feed the
// A operand directly
template <bool EXACT>
__device__ __forceinline__ void attn_apply(const float (&w)[2][4], const float* sX, int h, float (&out)[4]) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  out[0] = out[1] = out[2] = out[3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const float a[4] = {rt<EXACT>(w[ks][0]), rt<EXACT>(w[ks][2]), rt<EXACT>(w[ks][1]), rt<EXACT>(w[ks][3])};
    const float b[2] = {sX[(8 * ks + 2 * t) * LDK + h * 8 + g], sX[(8 * ks + 2 * t + 1) * LDK + h * 8 + g]};
    mma_f<EXACT>(out, a, b);
  }
}

__device__ __forceinline__ float oct_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// One CTA = a tile of 32 gene tokens; it walks the cells of its chunk.  8 warps = (row tile of 16 tokens) x (quarter of the N dimension /
// attention head); two CTAs are resident per SM (103 KB of shared memory, <= 128 registers), so that the barrier-separated phases of one
// tile overlap with those of the other.
template <bool BWD, bool EXACT>
__global__ void __launch_bounds__(256, BWD ? 2 : 3) dec_mcab_train_kernel(const DecTrainParams p) {
  extern __shared__ float4 dec_smem4[];
  float* sm = reinterpret_cast<float*>(dec_smem4);
  float* sQ = sm;                    float* sAO = sQ + DT * LD32;    float* sX1 = sAO + DT * LD32;   float* sN2 = sX1 + DT * LD32;
  float* sD1 = sN2 + DT * LD32;      float* sD2 = sD1 + DT * LD32;   float* sU = sD2 + DT * LD32;    float* sV = sU + DT * LDU;    // backward only
  float* sK = BWD ? sV + DT * LDU : sN2 + DT * LD32;
  float* sVc = sK + 16 * LDK;        float* sStat = sVc + 16 * LDK;  float* sDl = sStat + 2 * DT;
  float* sLog = sDl + DT;            int* sGid = reinterpret_cast<int*>(sLog + 4 * DT);
  float* sWp = sLog + 5 * DT;        float* sW1 = sWp + 32 * LD32;   float* sW2 = sW1 + H * LD32;    float* sW3 = sW2 + H * LD32;    // sW3: backward only
  float* sLn = BWD ? sW3 + 32 * LD88 : sW3;
  float* sWh = sLn + 64;             float* sR = sLn + 96;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int mt = warp & 1, nq = warp >> 1;
  const int g0 = blockIdx.x * DT;
  const int b_begin = blockIdx.y * p.cells_per_chunk, b_end = min(p.B, b_begin + p.cells_per_chunk);
  const float scale = 0.35355339059327373f;   // 1 / sqrt(head_dim = 8)

  // ---- weights of the block into shared memory (padded rows: conflict-free fragment reads) ----
  for (int i = tid; i < 1024; i += 256) sWp[(i >> 5) * LD32 + (i & 31)] = rt<EXACT>(p.ca[C_CPROJ + i]);
  for (int i = tid; i < H * 32; i += 256) { sW1[(i >> 5) * LD32 + (i & 31)] = rt<EXACT>(p.ca[C_W1 + i]); sW2[(i >> 5) * LD32 + (i & 31)] = rt<EXACT>(p.ca[C_W2 + i]); }
  if (BWD) { for (int i = tid; i < 32 * H; i += 256) sW3[(i / H) * LD88 + (i % H)] = rt<EXACT>(p.ca[C_W3 + i]); }
  if (tid < 64) sLn[tid] = p.ca[C_LN2W + tid];   // ln_2.weight | ln_2.bias
  if (tid < 32) sWh[tid] = p.head_w[tid];
  if (tid < DT) sGid[tid] = (g0 + tid < p.G) ? (int)p.genes[g0 + tid] : -1;
  // ---- the tile's query rows (cell-invariant) ----
  for (int i = tid; i < DT * 32; i += 256) {
    const int tok = i >> 5, c = i & 31, gi = g0 + tok;
    sQ[tok * LD32 + c] = gi < p.G ? rt<EXACT>(p.Q[(size_t)gi * 32 + c]) : 0.f;
  }
  const float head_b = p.head_b[0];
  __syncthreads();
  // The head is linear in x2 = x1 + w3 h, so its backward never needs the mlp.c_proj GEMMs: with r = w3^T w_head,
  // dh[tok] = dlogit[tok] r, d w3 = w_head (x) sum_tok dlogit h, d w_head = sum_tok dlogit x1 + w3 sum_tok dlogit h.
  // (forward: logit = w_head . x1 + r . h + b - no mlp.c_proj GEMM either.  r from the fp32 weights, not the TF32-rounded tile.)
  if (tid < H) {
    float a = 0.f;
    for (int o = 0; o < 32; ++o) a += sWh[o] * p.ca[C_W3 + o * H + tid];
    sR[tid] = a;
  }
  // the residual rows q_in = emb[gene] of this thread's accumulator positions (rows g, g + 8 of its row tile; columns nq * 8 + 2 t, + 1)
  float qin[4] = {0.f, 0.f, 0.f, 0.f};
  {
    const int ga = sGid[mt * 16 + g], gb = sGid[mt * 16 + g + 8], col = nq * 8 + 2 * t;
    if (ga >= 0) { const float2 v = *reinterpret_cast<const float2*>(p.emb + (size_t)ga * 32 + col); qin[0] = v.x; qin[1] = v.y; }
    if (gb >= 0) { const float2 v = *reinterpret_cast<const float2*>(p.emb + (size_t)gb * 32 + col); qin[2] = v.x; qin[3] = v.y; }
  }
  // persistent gradient accumulators
  float acc_w[6][4] = {}, acc_wp[2][4] = {};
  float acc_lnw[4] = {}, acc_lnb[4] = {}, acc_xs[4] = {}, acc_dq[4] = {}, acc_sh[3][2] = {}, acc_s = 0.f;
  float hx_lo = 0.f, hx_hi = 0.f;                 // forward: partial head dot products of this thread's rows
  const int tok_a = tid >> 3, part = tid & 7;     // (token, 4-channel part) of the row phases
  const int nt0 = nq * 3, ntn = nq < 3 ? 3 : 2;   // this warp's column tiles of the 88 hidden units
  __syncthreads();

  // software pipeline over the cells: the next cell's keys / values / dlogit are fetched into registers while this one is processed
  // (otherwise every trip starts with an exposed L2 round trip in front of its first barrier)
  float nk[2], nv[2], ndl = 0.f;
  {
    const int b = b_begin < b_end ? b_begin : 0;
    nk[0] = p.Kc[(size_t)b * 512 + tid]; nk[1] = p.Kc[(size_t)b * 512 + tid + 256];
    nv[0] = p.Vc[(size_t)b * 512 + tid]; nv[1] = p.Vc[(size_t)b * 512 + tid + 256];
    if (BWD && tid < DT && g0 + tid < p.G) ndl = p.dlogit[(size_t)b * p.G + g0 + tid];
  }
  for (int b = b_begin; b < b_end; ++b) {
    // (1) the cell's keys / values and, for the backward, d loss / d logit of the tile
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = tid + r * 256;
      sK[(i >> 5) * LDK + (i & 31)] = rt<EXACT>(nk[r]);
      sVc[(i >> 5) * LDK + (i & 31)] = rt<EXACT>(nv[r]);
    }
    if (BWD && tid < DT) sDl[tid] = ndl;
    __syncthreads();
    if (b + 1 < b_end) {
      nk[0] = p.Kc[(size_t)(b + 1) * 512 + tid]; nk[1] = p.Kc[(size_t)(b + 1) * 512 + tid + 256];
      nv[0] = p.Vc[(size_t)(b + 1) * 512 + tid]; nv[1] = p.Vc[(size_t)(b + 1) * 512 + tid + 256];
      if (BWD && tid < DT && g0 + tid < p.G) ndl = p.dlogit[(size_t)(b + 1) * p.G + g0 + tid];
    }
    // (2) cross attention over the 16 latent keys: warp (row tile, head)
    {
      float pr[2][4], o[4];
      attn_probs<EXACT>(sQ + mt * 16 * LD32, sK, nq, scale, pr);
      attn_apply<EXACT>(pr, sVc, nq, o);
      float* dst = sAO + (mt * 16 + g) * LD32 + nq * 8 + 2 * t;
      dst[0] = rt<EXACT>(o[0]); dst[1] = rt<EXACT>(o[1]); dst[8 * LD32] = rt<EXACT>(o[2]); dst[8 * LD32 + 1] = rt<EXACT>(o[3]);
    }
    __syncthreads();
    // (3) x1 = q_in + c_proj(ao)
    {
      float acc[1][4] = {};
      warp_gemm<EXACT, 1, 4>(acc, sAO + mt * 16 * LD32, LD32, 1, sWp + (nq * 8) * LD32, 1, LD32, 1, 8);
      float* dst = sX1 + (mt * 16 + g) * LD32 + nq * 8 + 2 * t;
      dst[0] = qin[0] + acc[0][0]; dst[1] = qin[1] + acc[0][1]; dst[8 * LD32] = qin[2] + acc[0][2]; dst[8 * LD32 + 1] = qin[3] + acc[0][3];
      if (!BWD) {   // this thread's share of w_head . x1 for its two rows
        const float w0 = sWh[nq * 8 + 2 * t], w1 = sWh[nq * 8 + 2 * t + 1];
        hx_lo = dst[0] * w0 + dst[1] * w1;
        hx_hi = dst[8 * LD32] * w0 + dst[8 * LD32 + 1] * w1;
      }
    }
    __syncthreads();
    // (4) LN2: eight threads per token
    {
      float x[4], sum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { x[i] = sX1[tok_a * LD32 + part * 4 + i]; sum += x[i]; }
      const float mean = oct_sum(sum) * (1.f / 32.f);
      float var = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { x[i] -= mean; var += x[i] * x[i]; }
      const float rstd = rsqrtf(oct_sum(var) * (1.f / 32.f) + p.eps);
#pragma unroll
      for (int i = 0; i < 4; ++i) sN2[tok_a * LD32 + part * 4 + i] = rt<EXACT>(x[i] * rstd * sLn[part * 4 + i] + sLn[32 + part * 4 + i]);
      if (part == 0) { sStat[tok_a * 2] = mean; sStat[tok_a * 2 + 1] = rstd; }
    }
    __syncthreads();
    // (5) u = w1 n2, v = w2 n2, h = silu(u) v; backward: du, dv right away (dh = dlogit r)
    {
      float au[3][4] = {}, av[3][4] = {};
      warp_gemm2<EXACT, 3, 4>(au, av, sN2 + mt * 16 * LD32, LD32, 1, sW1 + (nt0 * 8) * LD32, sW2 + (nt0 * 8) * LD32, 1, LD32, ntn, 8);
      const float dl0 = BWD ? sDl[mt * 16 + g] : 0.f, dl1 = BWD ? sDl[mt * 16 + g + 8] : 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i < ntn) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = mt * 16 + g + (e >> 1) * 8, col = (nt0 + i) * 8 + 2 * t + (e & 1);
            const float u = au[i][e], v = av[i][e], sg = sigmoidf_(u), hv = u * sg * v;
            if (BWD) {
              const float dl = (e >> 1) ? dl1 : dl0, dh = dl * sR[col];
              acc_sh[i][e & 1] += dl * hv;
              sU[uidx(r, col)] = rt<EXACT>(dh * v * sg * (1.f + u * (1.f - sg)));
              sV[uidx(r, col)] = rt<EXACT>(dh * u * sg);
            } else {
              if (e >> 1) hx_hi += sR[col] * hv; else hx_lo += sR[col] * hv;
            }
          }
        }
      }
    }
    if (BWD) __syncthreads();
    if (!BWD) {
      // (6) the four column-quarter warps of a row tile leave their shares of w_head . x1 + r . h in four slots per token
      {
        const float lo = quad_sum(hx_lo), hi = quad_sum(hx_hi);
        if (t == 0) { sLog[nq * DT + mt * 16 + g] = lo; sLog[nq * DT + mt * 16 + g + 8] = hi; }
      }
      __syncthreads();
      // (7) head logit
      if (tid < DT && g0 + tid < p.G)
        p.logits[(size_t)b * p.G + g0 + tid] = head_b + sLog[tid] + sLog[DT + tid] + sLog[2 * DT + tid] + sLog[3 * DT + tid];
      continue;     // (the next writes of sLog / sX1 sit behind the barriers of the next iteration)
    }
    // (8) sums for the head gradient (threads 96..127: sum_tok dlogit x1; 128: sum_tok dlogit)
    if (tid >= 96 && tid < 128) {
      float a = 0.f;
#pragma unroll 8
      for (int r = 0; r < DT; ++r) a += sDl[r] * sX1[r * LD32 + tid - 96];
      acc_s += a;
    } else if (tid == 128) {
      float a = 0.f;
      for (int r = 0; r < DT; ++r) a += sDl[r];
      acc_s += a;
    }
    // (10) d w1 += du^T n2, d w2 += dv^T n2 ;  d n2 = du w1 + dv w2
    {
      // weight-gradient tiles of a warp: one matrix (du | dv), 3 row tiles x 2 column tiles: per token step 12 + 4 fragment loads feed 6 mma
      const float* sG = (warp & 1) ? sV : sU;
      const int nw0 = ((warp >> 1) & 1) * 2, mw0 = (warp >> 2) * 3;
#pragma unroll
      for (int ks = 0; ks < DT / 8; ++ks) {
        const int k0 = ks * 8;
        float bf[2][2];
#pragma unroll
        for (int n = 0; n < 2; ++n) { bf[n][0] = sN2[(k0 + t) * LD32 + (nw0 + n) * 8 + g]; bf[n][1] = sN2[(k0 + t + 4) * LD32 + (nw0 + n) * 8 + g]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int j0 = (mw0 + i) * 16 + g;
          const float a[4] = {sG[uidx(k0 + t, j0)], sG[uidx(k0 + t, j0 + 8)], sG[uidx(k0 + t + 4, j0)], sG[uidx(k0 + t + 4, j0 + 8)]};
          mma_f<EXACT>(acc_w[i * 2], a, bf[0]);
          mma_f<EXACT>(acc_w[i * 2 + 1], a, bf[1]);
        }
      }
      // d n2: warp (row tile, column half, du w1 | dv w2): the two K halves land in sD2 / sD1 and are summed by the row phase below
      {
        const int nh2 = (warp >> 1) & 1, kh = warp >> 2;
        float acc[2][4] = {};
        {
          const float* A = kh ? sV : sU;
          const float* Bw = (kh ? sW2 : sW1) + nh2 * 16;
          const int ra = (mt * 16 + g) * LDU, rb = ra + 8 * LDU, sw = usw(mt * 16 + g);      // rows g and g + 8 share the swizzle
#pragma unroll
          for (int ks = 0; ks < 11; ++ks) {
            const int k0 = ks * 8;
            const float a[4] = {A[ra + ((k0 + t) ^ sw)], A[rb + ((k0 + t) ^ sw)], A[ra + ((k0 + t + 4) ^ sw)], A[rb + ((k0 + t + 4) ^ sw)]};
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float bb[2] = {Bw[(k0 + t) * LD32 + i * 8 + g], Bw[(k0 + t + 4) * LD32 + i * 8 + g]};
              mma_f<EXACT>(acc[i], a, bb);
            }
          }
        }
        float* dst = (kh ? sD1 : sD2) + (mt * 16 + g) * LD32 + nh2 * 16 + 2 * t;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          dst[i * 8] = acc[i][0]; dst[i * 8 + 1] = acc[i][1]; dst[8 * LD32 + i * 8] = acc[i][2]; dst[8 * LD32 + i * 8 + 1] = acc[i][3];
        }
      }
    }
    __syncthreads();
    // (11) LN2 backward + residual: d x1
    {
      const float mean = sStat[tok_a * 2], rstd = sStat[tok_a * 2 + 1], dl = sDl[tok_a];
      float xh[4], gy[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = part * 4 + i;
        xh[i] = (sX1[tok_a * LD32 + c] - mean) * rstd;
        const float dn = sD2[tok_a * LD32 + c] + sD1[tok_a * LD32 + c];
        acc_lnw[i] += dn * xh[i];
        acc_lnb[i] += dn;
        gy[i] = dn * sLn[c];
        s1 += gy[i];
        s2 += gy[i] * xh[i];
      }
      const float m1 = oct_sum(s1) * (1.f / 32.f), m2 = oct_sum(s2) * (1.f / 32.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = part * 4 + i;
        const float dx = rstd * (gy[i] - m1 - xh[i] * m2) + dl * sWh[c];
        acc_xs[i] += dx;
        sD1[tok_a * LD32 + c] = rt<EXACT>(dx);
      }
    }
    __syncthreads();
    // (12) warps 0-3: d ao = dx1 c_proj (row tile, column half); warps 4-7: d c_proj += dx1^T ao (row tile of the weight, column half)
    if (warp < 4) {
      const int nh2 = warp >> 1;
      float acc[2][4] = {};
      warp_gemm<EXACT, 2, 4>(acc, sD1 + mt * 16 * LD32, LD32, 1, sWp + nh2 * 16, LD32, 1, 2, 8);
      float* dst = sD2 + (mt * 16 + g) * LD32 + nh2 * 16 + 2 * t;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        dst[i * 8] = rt<EXACT>(acc[i][0]); dst[i * 8 + 1] = rt<EXACT>(acc[i][1]);
        dst[8 * LD32 + i * 8] = rt<EXACT>(acc[i][2]); dst[8 * LD32 + i * 8 + 1] = rt<EXACT>(acc[i][3]);
      }
    } else {
      warp_gemm<EXACT, 2, DT / 8>(acc_wp, sD1 + (warp & 1) * 16, 1, LD32, sAO + ((warp >> 1) & 1) * 16, LD32, 1, 2, 8);
    }
    __syncthreads();
    // (13) attention backward, warp (row tile, head); dS and P tiles for the key / value gradients alias the dead du, dv tiles
    float* tDS = sU;   // [DT][68]
    float* tP = sV;    // [DT][68]
    {
      const int h = nq;
      float pr[2][4], dp[2][4];
      attn_probs<EXACT>(sQ + mt * 16 * LD32, sK, h, scale, pr);
      {
        const float* dr = sD2 + mt * 16 * LD32;
        const float a[4] = {dr[g * LD32 + h * 8 + t], dr[(g + 8) * LD32 + h * 8 + t], dr[g * LD32 + h * 8 + t + 4], dr[(g + 8) * LD32 + h * 8 + t + 4]};
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const float bb[2] = {sVc[(8 * n + g) * LDK + h * 8 + t], sVc[(8 * n + g) * LDK + h * 8 + t + 4]};
          dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
          mma_f<EXACT>(dp[n], a, bb);
        }
      }
      const float D0 = quad_sum(pr[0][0] * dp[0][0] + pr[0][1] * dp[0][1] + pr[1][0] * dp[1][0] + pr[1][1] * dp[1][1]);
      const float D1 = quad_sum(pr[0][2] * dp[0][2] + pr[0][3] * dp[0][3] + pr[1][2] * dp[1][2] + pr[1][3] * dp[1][3]);
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        dp[n][0] = pr[n][0] * (dp[n][0] - D0) * scale; dp[n][1] = pr[n][1] * (dp[n][1] - D0) * scale;
        dp[n][2] = pr[n][2] * (dp[n][2] - D1) * scale; dp[n][3] = pr[n][3] * (dp[n][3] - D1) * scale;
        const int r0 = mt * 16 + g, col = h * 16 + 8 * n + 2 * t;
        tDS[r0 * 68 + col] = rt<EXACT>(dp[n][0]); tDS[r0 * 68 + col + 1] = rt<EXACT>(dp[n][1]);
        tDS[(r0 + 8) * 68 + col] = rt<EXACT>(dp[n][2]); tDS[(r0 + 8) * 68 + col + 1] = rt<EXACT>(dp[n][3]);
        tP[r0 * 68 + col] = rt<EXACT>(pr[n][0]); tP[r0 * 68 + col + 1] = rt<EXACT>(pr[n][1]);
        tP[(r0 + 8) * 68 + col] = rt<EXACT>(pr[n][2]); tP[(r0 + 8) * 68 + col + 1] = rt<EXACT>(pr[n][3]);
      }
      float dq[4];
      attn_apply<EXACT>(dp, sK, h, dq);      // dQ[tok][h*8 + d] = sum_keys dS K
#pragma unroll
      for (int e = 0; e < 4; ++e) acc_dq[e] += dq[e];
    }
    __syncthreads();
    // (14) dK[key][c] = sum_tok dS[tok][h(c)][key] Q[tok][c] ;  dV[key][c] = sum_tok P[tok][h(c)][key] dAO[tok][c]: warp (head, K | V)
    {
      const int h = warp & 3, which = warp >> 2;
      float acc[1][4] = {};
      warp_gemm<EXACT, 1, DT / 8>(acc, (which ? tP : tDS) + h * 16, 1, 68, (which ? sD2 : sQ) + h * 8, LD32, 1, 1, 8);
      float* dst = (which ? p.dV : p.dK) + (size_t)b * 512 + h * 8 + 2 * t;
      atomicAdd(dst + g * 32, acc[0][0]); atomicAdd(dst + g * 32 + 1, acc[0][1]);
      atomicAdd(dst + (g + 8) * 32, acc[0][2]); atomicAdd(dst + (g + 8) * 32 + 1, acc[0][3]);
    }
    __syncthreads();
  }

  if (!BWD) return;
  // ---- flush the gradient accumulators ----
  {
    float* gw = p.gca + ((warp & 1) ? C_W2 : C_W1);
    const int nw0 = ((warp >> 1) & 1) * 2, mw0 = (warp >> 2) * 3;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = (mw0 + (i >> 1)) * 16 + g + (e >> 1) * 8, c = (nw0 + (i & 1)) * 8 + 2 * t + (e & 1);
        if (j < H) atomicAdd(gw + j * 32 + c, acc_w[i][e]);
      }
    }
  }
  if (warp >= 4) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int o = (warp & 1) * 16 + g + (e >> 1) * 8, c = ((warp >> 1) & 1) * 16 + i * 8 + 2 * t + (e & 1);
        atomicAdd(p.gca + C_CPROJ + o * 32 + c, acc_wp[i][e]);
      }
    }
  }
  {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gi = g0 + mt * 16 + g + (e >> 1) * 8;
      if (gi < p.G) atomicAdd(p.dQ + (size_t)gi * 32 + nq * 8 + 2 * t + (e & 1), acc_dq[e]);
    }
    const int gi = g0 + tok_a;
    if (gi < p.G) {
      red_add_v4(p.dXsum + (size_t)gi * 32 + part * 4, acc_xs[0], acc_xs[1], acc_xs[2], acc_xs[3]);
    }
  }
  __syncthreads();
  // s[j] = sum_tok dlogit h[j] (per-thread partials over rows g, g + 8 -> lanes of equal t -> the two row-tile warps) in sD2[0..87];
  // ln_2 partials in sD1 / sU
  if (tid < H) sD2[tid] = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { sD1[tok_a * LD32 + part * 4 + i] = acc_lnw[i]; sU[tok_a * LD32 + part * 4 + i] = acc_lnb[i]; }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float v = acc_sh[i][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0 && i < ntn) atomicAdd(sD2 + (nt0 + i) * 8 + 2 * t + e, v);
    }
  }
  __syncthreads();
  if (tid < H) {
    const float sj = sD2[tid];
    for (int o = 0; o < 32; ++o) atomicAdd(p.gca + C_W3 + o * H + tid, sWh[o] * sj);
  } else if (tid >= 96 && tid < 128) {
    const int c = tid - 96;
    float a = acc_s;
    for (int j = 0; j < H; ++j) a += sW3[c * LD88 + j] * sD2[j];
    atomicAdd(p.g_head_w + c, a);
  } else if (tid == 128) {
    atomicAdd(p.g_head_b, acc_s);
  } else if (tid >= 160 && tid < 224) {
    const int c = tid & 31;
    const float* tsrc = tid < 192 ? sD1 : sU;
    float a = 0.f;
    for (int r = 0; r < DT; ++r) a += tsrc[r * LD32 + c];
    atomicAdd(p.gca + (tid < 192 ? C_LN2W : C_LN2B) + c, a);
  }
}

}  // namespace vtr
