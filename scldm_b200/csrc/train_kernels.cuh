// DiT training kernels for sm_100a: forward with saved activations, backward, AdamW.
//
// Reference: LatentDiffusion.training_step (src/scldm/models.py:634-666) -> Transport.training_losses
// (src/scldm/transport/transport.py:110-150) -> DiT.forward (src/scldm/nnets.py:273-297) + torch autograd + torch.optim.AdamW
// (experiments/configs/model/ldm_base.yaml:36-40) + gradient clipping by norm (experiments/configs/training/default.yaml:15).
//
// Every dense contraction (forward, dgrad, wgrad) runs on tcgen05 / TMEM through ONE kernel, `gemm_kernel`, whose operands are
// "slab tensors": an activation T[rows][C] in bf16 stored as [row_tile][C/64][128 x 64] tiles with the 128-byte swizzle
// (sm100.cuh) - the same physical tile is a K-major UMMA operand when the contraction runs over C (forward, dgrad) and an
// MN-major operand when it runs over the rows (wgrad), so no transposed copy of an activation or of a weight is ever made:
//   forward  Y[rows][N]  = A[rows][K] . W[N][K]^T     A: slab tensor, K-major      B: packed weight tiles, K-major
//   dgrad    dA[rows][K] = dY[rows][N] . W[N][K]      A: dY slab tensor, K-major    B: the SAME packed tiles, MN-major
//   wgrad    dW[N][K]   += dY[rows][N]^T . A[rows][K] A: dY slab tensor, MN-major   B: A slab tensor, MN-major
// All three are expressed as a stream of 48 KB pipeline stages of 1-D bulk-TMA copies (GemmOp) plus UMMA descriptors.
// The pointwise / normalisation / attention pieces between the GEMMs are plain fp32 CUDA kernels that read fp32 row-major
// GEMM outputs and write the bf16 slab tensors the next GEMM consumes.
#pragma once

#include "dit_kernels.cuh"

namespace trn {

using dit::bf16;
constexpr int D = dit::D;            // 256
constexpr int TOK = dit::TOK;        // 16
constexpr int LAT = dit::LAT;        // 16
constexpr int NHEAD = dit::NHEAD;    // 8
constexpr int HD = dit::HD;          // 32
constexpr int SLAB_ELEMS = dit::A_SLAB_ELEMS;   // 128 x 64
constexpr int SLAB_BYTES = dit::A_SLAB_BYTES;   // 16 KB
constexpr int WSLAB_ELEMS = dit::B_SLAB_ELEMS;  // 256 x 64 (packed weight tile)

// element (row, col8..col8+7) of a slab tensor with C columns: pointer to the 16-byte chunk
__device__ __forceinline__ bf16* slab_chunk(bf16* base, int C, int row, int col8) {
  const int rt = row >> 7, r = row & 127;
  return base + ((size_t)rt * (C >> 6) + (col8 >> 6)) * SLAB_ELEMS + (sm100::swz_chunk_offset(r, (col8 & 63) >> 3) >> 1);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 o;
  o.x = sm100::pack_bf16x2(v[0], v[1]); o.y = sm100::pack_bf16x2(v[2], v[3]);
  o.z = sm100::pack_bf16x2(v[4], v[5]); o.w = sm100::pack_bf16x2(v[6], v[7]);
  return o;
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ==========================================================================================
// Generic slab GEMM on tcgen05: D[128 x 256] (fp32, TMEM) = sum over k-steps of A_step . B_step
// ==========================================================================================
struct GemmOp {
  const bf16* base;
  long long tile_stride;          // elements per CTA tile index (blockIdx.x for A, blockIdx.y for B)
  long long step_hi, step_lo;     // k-step s reads at (s / step_div) * step_hi + (s % step_div) * step_lo
  long long copy_stride;          // elements between the bulk copies of one step (they land back to back in shared memory)
  int step_div, n_copies, copy_bytes;
  int major;                      // 0: K-major, 1: MN-major (UMMA instruction descriptor bits 15 / 16)
  int lbo, sbo;                   // bytes (shared-memory matrix descriptor)
  int mma_adv;                    // bytes the descriptor start address advances per K = 16 MMA inside a step
};
struct GemmParams {
  GemmOp a, b;
  int n_steps;                    // k-steps in total; blockIdx.z takes a contiguous share (split-K)
  int mmas_per_step;
  float* out;                     // fp32 row-major [.][out_ld]; row = 128 blockIdx.x + r, col = 256 blockIdx.y + c
  long long out_ld;
  int valid_rows, valid_cols;
  const float* bias;              // [valid_cols] or nullptr (added by the blockIdx.z == 0 share only)
  int atomic;                     // 1: accumulate into `out` with vector atomics (split-K, gradient accumulation)
  int n_loop;                     // > 1: the CTA walks n_loop consecutive N tiles (two TMEM accumulators: epilogue t overlaps MMAs t + 1)
  int swap_xy;                    // 1: blockIdx.x walks the N tiles and blockIdx.y the M tiles (CTAs that share an A tile are scheduled together: L2 reuse)
  int epi;                        // 0: fp32 store / atomic accumulate;  1: SwiGLU . v epilogue (no `out`):
  const float* vdot;              //    the tile's columns are [w1 h (128) | w2 h (128)] of hidden units [128 n, 128 n + 128);
  float* dot_out;                 //    dot_out[row] += sum_i silu(a_i) * b_i * vdot[128 n + i]   (atomic = 0 and n_loop = all tiles: one plain update per row)
};

constexpr int G_STAGE_BYTES = 48 * 1024;
constexpr int G_NSTAGE = 3;
constexpr int G_OFF_STG = G_NSTAGE * G_STAGE_BYTES;                     // 144 KB
constexpr int G_OFF_BARS = G_OFF_STG + dit::RESID_STG_BYTES;            // + 64 KB epilogue staging
constexpr size_t gemm_smem_bytes() { return G_OFF_BARS + 128; }

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr, int lbo, int sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((uint32_t)lbo >> 4) << 16;
  d |= static_cast<uint64_t>((uint32_t)sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(dit::NUM_THREADS, 1) gemm_kernel(const GemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (sm100::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smStg = smem + G_OFF_STG;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_OFF_BARS);
  uint64_t* full = bars;
  uint64_t* empty = bars + G_NSTAGE;
  uint64_t* tmem_full = bars + 2 * G_NSTAGE;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_loop = p.n_loop > 1 ? p.n_loop : 1;
  const int bm = p.swap_xy ? blockIdx.y : blockIdx.x, bn0 = (p.swap_xy ? blockIdx.x : blockIdx.y) * n_loop;
  const int per = (p.n_steps + (int)gridDim.z - 1) / (int)gridDim.z;
  const int s0 = (int)blockIdx.z * per;
  const int s1 = min(p.n_steps, s0 + per);
  if (s0 >= s1) return;   // empty split-K share (uniform for the whole CTA)

  const uint32_t a_bytes = (uint32_t)p.a.n_copies * p.a.copy_bytes, b_bytes = (uint32_t)p.b.n_copies * p.b.copy_bytes;
  if (threadIdx.x == 0) {
    if (a_bytes + b_bytes > (uint32_t)G_STAGE_BYTES) __trap();
    for (int i = 0; i < G_NSTAGE; ++i) { sm100::mbar_init(&full[i], 1); sm100::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { sm100::mbar_init(&tmem_full[i], 1); sm100::mbar_init(&tmem_empty[i], dit::EPI_WARPS); }
    sm100::fence_barrier_init();
  }
  if (warp == 1) sm100::tmem_alloc(tmem_ptr_smem, 512);   // two 256-column accumulators: the epilogue of N tile t overlaps the MMAs of tile t + 1
  sm100::grid_dep_launch();   // barrier init / TMEM allocation above overlap the tail of the preceding kernel (PDL)
  sm100::grid_dep_wait();
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      dit::RingState rs;
      const bf16* a0 = p.a.base + (long long)bm * p.a.tile_stride;
      for (int t = 0; t < n_loop; ++t) {
        const bf16* b0 = p.b.base + (long long)(bn0 + t) * p.b.tile_stride;
        for (int s = s0; s < s1; ++s) {
          sm100::mbar_wait(&empty[rs.stage], rs.phase ^ 1);
          sm100::mbar_arrive_expect_tx(&full[rs.stage], a_bytes + b_bytes);
          uint8_t* st = smem + rs.stage * G_STAGE_BYTES;
          const bf16* as = a0 + (long long)(s / p.a.step_div) * p.a.step_hi + (long long)(s % p.a.step_div) * p.a.step_lo;
          for (int c = 0; c < p.a.n_copies; ++c) sm100::bulk_g2s(st + c * p.a.copy_bytes, as + c * p.a.copy_stride, p.a.copy_bytes, &full[rs.stage]);
          const bf16* bs = b0 + (long long)(s / p.b.step_div) * p.b.step_hi + (long long)(s % p.b.step_div) * p.b.step_lo;
          for (int c = 0; c < p.b.n_copies; ++c) sm100::bulk_g2s(st + a_bytes + c * p.b.copy_bytes, bs + c * p.b.copy_stride, p.b.copy_bytes, &full[rs.stage]);
          rs.advance(G_NSTAGE);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = sm100::make_idesc_bf16(128, 256) | ((uint32_t)(p.a.major & 1) << 15) | ((uint32_t)(p.b.major & 1) << 16);
      dit::RingState rs;
      for (int t = 0; t < n_loop; ++t) {
        const uint32_t acc = t & 1;
        sm100::mbar_wait(&tmem_empty[acc], ((t >> 1) & 1) ^ 1);
        sm100::tc_fence_after();
        for (int s = s0; s < s1; ++s) {
          sm100::mbar_wait(&full[rs.stage], rs.phase);
          sm100::tc_fence_after();
          const uint32_t st = sm100::smem_u32(smem + rs.stage * G_STAGE_BYTES);
          const uint64_t ad = make_sw128_desc(st, p.a.lbo, p.a.sbo), bd = make_sw128_desc(st + a_bytes, p.b.lbo, p.b.sbo);
          for (int k = 0; k < p.mmas_per_step; ++k)
            sm100::umma_bf16_ss(tmem_base + acc * 256, ad + (uint64_t)((k * p.a.mma_adv) >> 4), bd + (uint64_t)((k * p.b.mma_adv) >> 4), idesc,
                                (s == s0 && k == 0) ? 0u : 1u);
          sm100::umma_commit(&empty[rs.stage]);
          rs.advance(G_NSTAGE);
        }
        sm100::umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // 16 epilogue warps: lane quadrant q (TMEM lanes 32q..), column quarter sub (64 columns)
    const uint32_t ew = warp - 2, q = warp & 3, sub = ew >> 2;
    uint8_t* stg = smStg + ew * dit::RESID_WARP_STG;
    const uint32_t rg = lane >> 3, cchunk = lane & 7;
    float part = 0.f;   // epi 1: this thread's share of the row's SwiGLU . v dot product, accumulated over the N tiles of the CTA
    for (int t = 0; t < n_loop; ++t) {
      const uint32_t acc = t & 1;
      const int bn = bn0 + t;
      sm100::mbar_wait(&tmem_full[acc], (t >> 1) & 1);
      sm100::tc_fence_after();
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * 256;
      if (p.epi == 1) {
        // SwiGLU . v: this warp owns rows [32 q, 32 q + 32) (one per lane) and hidden units [32 sub, 32 sub + 32) of the tile
        uint32_t va[32], vb[32];
        sm100::tmem_ld_32x32b_x32(taddr + sub * 32, va);
        sm100::tmem_ld_32x32b_x32(taddr + 128 + sub * 32, vb);
        sm100::tmem_ld_wait();
        const float* vv = p.vdot + bn * 128 + sub * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float a = __uint_as_float(va[i]);
          part += a * sigmoidf_(a) * __uint_as_float(vb[i]) * vv[i];
        }
      } else {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int col0 = bn * 256 + sub * 64 + half * 32;
          if (col0 < p.valid_cols) {   // warp-uniform
            uint32_t v[32];
            sm100::tmem_ld_32x32b_x32(taddr + sub * 64 + half * 32, v);
            sm100::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<float4*>(stg + sm100::swz_chunk_offset(lane, c)) =
                  make_float4(__uint_as_float(v[c * 4 + 0]), __uint_as_float(v[c * 4 + 1]), __uint_as_float(v[c * 4 + 2]), __uint_as_float(v[c * 4 + 3]));
            __syncwarp();
            const int col = col0 + cchunk * 4;
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias != nullptr && blockIdx.z == 0 && col < p.valid_cols) bb = *reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const uint32_t r = it * 4 + rg;
              const int row = bm * 128 + q * 32 + r;
              float4 a = *reinterpret_cast<const float4*>(stg + sm100::swz_chunk_offset(r, cchunk));
              a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w;
              if (row < p.valid_rows && col < p.valid_cols) {
                float* dst = p.out + (long long)row * p.out_ld + col;
                if (p.atomic) atomicAdd(reinterpret_cast<float4*>(dst), a);
                else *reinterpret_cast<float4*>(dst) = a;
              }
            }
            __syncwarp();
          }
        }
      }
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&tmem_empty[acc]);     // the MMA warp may overwrite this accumulator (tile t + 2)
    }
    if (p.epi == 1) {
      float* red = reinterpret_cast<float*>(smStg);            // [4 sub][128 rows]
      red[sub * 128 + q * 32 + lane] = part;
      sm100::named_bar_sync(1, dit::EPI_THREADS);
      if (sub == 0) {
        const int r = q * 32 + lane, row = bm * 128 + r;
        if (row < p.valid_rows) {
          const float tot = (red[r] + red[128 + r]) + (red[256 + r] + red[384 + r]);
          if (p.atomic) atomicAdd(p.dot_out + row, tot);
          else p.dot_out[row] += tot;                          // single writer per row (all N tiles in this CTA): deterministic
        }
      }
    }
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) sm100::tmem_dealloc(tmem_base, 512);
}

// ==========================================================================================
// pointwise / normalisation kernels (fp32 math, bf16 slab outputs)
// ==========================================================================================

// h = LN(x) * (1 + mul[cell]) + add[cell]   (layers.py:91-94,217,219; final layer: layers.py:398-399) -> slab tensor [rows][256]
// optionally also fp32 row-major (the final layer's Linear 256->16 runs on CUDA cores)
struct LnModParams {
  const float* X; const float* mod; long long mod_stride; int off_mul, off_add; float eps;
  bf16* h; float* h_f32; int rows;
};
__global__ void __launch_bounds__(256) lnmod_fwd_kernel(const LnModParams p) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= p.rows) return;
  const int c0 = lane * 8;
  float x[8];
  load8(p.X + (size_t)row * D + c0, x);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  const float mean = sm100::warp_sum(s) * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] -= mean; sq += x[j] * x[j]; }
  const float rs = rsqrtf(sm100::warp_sum(sq) * (1.0f / D) + p.eps);
  const float* mrow = p.mod + (size_t)(row >> 4) * p.mod_stride;
  float mul[8], add[8], h[8];
  load8(mrow + p.off_mul + c0, mul);
  load8(mrow + p.off_add + c0, add);
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = x[j] * rs * (1.0f + mul[j]) + add[j];
  if (p.h != nullptr) *reinterpret_cast<uint4*>(slab_chunk(p.h, D, row, c0)) = pack8(h);
  if (p.h_f32 != nullptr) store8(p.h_f32 + (size_t)row * D + c0, h);
}

// backward of the above for one cell (16 rows): dX (+)= LN-backward(dh * (1 + mul)); dmul = sum_tok dh * xhat; dadd = sum_tok dh
struct LnModBwdParams {
  const float* X; const float* mod; long long mod_stride; int off_mul, off_add; float eps;
  float* dh; float* dX; int accumulate; float* dmod; long long dmod_stride;
  int clear_dh;   // 1: zero dh after reading it (the next split-K dgrad accumulates into it with atomics)
};
__global__ void __launch_bounds__(512) lnmod_bwd_kernel(const LnModBwdParams p) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  __shared__ float red[2][TOK][D];   // 32 KB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = blockIdx.x, row = cell * TOK + warp, c0 = lane * 8;
  float x[8], dh[8], mul[8];
  load8(p.X + (size_t)row * D + c0, x);
  load8(p.dh + (size_t)row * D + c0, dh);
  if (p.clear_dh) { const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; store8(p.dh + (size_t)row * D + c0, z8); }
  load8(p.mod + (size_t)cell * p.mod_stride + p.off_mul + c0, mul);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  const float mean = sm100::warp_sum(s) * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] -= mean; sq += x[j] * x[j]; }
  const float rs = rsqrtf(sm100::warp_sum(sq) * (1.0f / D) + p.eps);
  float g[8], m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    x[j] *= rs;                               // xhat
    g[j] = dh[j] * (1.0f + mul[j]);           // d xhat
    m1 += g[j];
    m2 += g[j] * x[j];
    red[0][warp][c0 + j] = dh[j] * x[j];
    red[1][warp][c0 + j] = dh[j];
  }
  m1 = sm100::warp_sum(m1) * (1.0f / D);
  m2 = sm100::warp_sum(m2) * (1.0f / D);
  float dx[8];
  float* dst = p.dX + (size_t)row * D + c0;
  if (p.accumulate) load8(dst, dx);
  else {
#pragma unroll
    for (int j = 0; j < 8; ++j) dx[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) dx[j] += rs * (g[j] - m1 - x[j] * m2);
  store8(dst, dx);
  __syncthreads();
  {
    const int which = threadIdx.x >> 8, c = threadIdx.x & 255;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < TOK; ++t) acc += red[which][t][c];
    p.dmod[(size_t)cell * p.dmod_stride + (which == 0 ? p.off_mul : p.off_add) + c] = acc;
  }
}

// x_out = x_in + gate[cell] * y   (layers.py:218,221)
__global__ void __launch_bounds__(256) resid_fwd_kernel(const float* __restrict__ Xin, const float* __restrict__ y, const float* __restrict__ mod,
                                                        long long mod_stride, int off_gate, float* __restrict__ Xout, int rows) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)rows * (D / 4)) return;
  const int row = (int)(idx >> 6), c = (int)(idx & 63) * 4;
  const float4 x = *reinterpret_cast<const float4*>(Xin + (size_t)row * D + c);
  const float4 yy = *reinterpret_cast<const float4*>(y + (size_t)row * D + c);
  const float4 g = *reinterpret_cast<const float4*>(mod + (size_t)(row >> 4) * mod_stride + off_gate + c);
  *reinterpret_cast<float4*>(Xout + (size_t)row * D + c) = make_float4(x.x + g.x * yy.x, x.y + g.y * yy.y, x.z + g.z * yy.z, x.w + g.w * yy.w);
}

// backward of the gated residual for one cell: dy = gate * dX (bf16 slab tensor), dgate = sum_tok dX * y
__global__ void __launch_bounds__(512) resid_bwd_kernel(const float* __restrict__ dX, const float* __restrict__ y, const float* __restrict__ mod,
                                                        long long mod_stride, int off_gate, bf16* __restrict__ dy, float* __restrict__ dmod,
                                                        long long dmod_stride) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  __shared__ float red[TOK][D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cell = blockIdx.x, row = cell * TOK + warp, c0 = lane * 8;
  float dx[8], yy[8], g[8], o[8];
  load8(dX + (size_t)row * D + c0, dx);
  load8(y + (size_t)row * D + c0, yy);
  load8(mod + (size_t)cell * mod_stride + off_gate + c0, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j] = g[j] * dx[j]; red[warp][c0 + j] = dx[j] * yy[j]; }
  *reinterpret_cast<uint4*>(slab_chunk(dy, D, row, c0)) = pack8(o);
  __syncthreads();
  if (threadIdx.x < D) {
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < TOK; ++t) acc += red[t][threadIdx.x];
    dmod[(size_t)cell * dmod_stride + off_gate + threadIdx.x] = acc;
  }
}

// SwiGLU (layers.py:161-174).  ab [rows][T*256]: per 128-hidden tile j the columns [256j, +128) = w1 x, [256j+128, +128) = w2 x.
// s[rows][T*128] = silu(a) * b as a slab tensor.
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const float* __restrict__ ab, int n_tiles, bf16* __restrict__ s, int rows) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const int per_row = n_tiles * 16;   // 8-wide groups per row
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)rows * per_row) return;
  const int row = (int)(idx / per_row), g8 = (int)(idx % per_row) * 8;
  const int j = g8 >> 7, off = g8 & 127;
  float a[8], b[8], o[8];
  load8(ab + (size_t)row * n_tiles * 256 + j * 256 + off, a);
  load8(ab + (size_t)row * n_tiles * 256 + j * 256 + 128 + off, b);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = a[k] * sigmoidf_(a[k]) * b[k];
  *reinterpret_cast<uint4*>(slab_chunk(s, n_tiles * 128, row, g8)) = pack8(o);
}
// ds [rows][T*128] -> dab slab tensor [rows][T*256] in the same interleaved column order as `ab`
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(float* __restrict__ ds, const float* __restrict__ ab, int n_tiles,
                                                         bf16* __restrict__ dab, int rows) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const int per_row = n_tiles * 16;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)rows * per_row) return;
  const int row = (int)(idx / per_row), g8 = (int)(idx % per_row) * 8;
  const int j = g8 >> 7, off = g8 & 127;
  float a[8], b[8], d[8], da[8], db[8];
  load8(ab + (size_t)row * n_tiles * 256 + j * 256 + off, a);
  load8(ab + (size_t)row * n_tiles * 256 + j * 256 + 128 + off, b);
  load8(ds + (size_t)row * n_tiles * 128 + g8, d);
  { const float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}; store8(ds + (size_t)row * n_tiles * 128 + g8, z8); }   // consumed: the next split-K dgrad accumulates into zeros
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float sg = sigmoidf_(a[k]);
    da[k] = d[k] * b[k] * sg * (1.0f + a[k] * (1.0f - sg));
    db[k] = d[k] * a[k] * sg;
  }
  *reinterpret_cast<uint4*>(slab_chunk(dab, n_tiles * 256, row, j * 256 + off)) = pack8(da);
  *reinterpret_cast<uint4*>(slab_chunk(dab, n_tiles * 256, row, j * 256 + 128 + off)) = pack8(db);
}

// ------------------------------------------------------------------------------------------
// 16-token self attention per (cell, head), fp32 (layers.py:143-158: softmax(q k^T / sqrt(hd)) v, unmasked).
// One CTA per cell, 4 warps; lane = (head-in-warp, query row).
// ------------------------------------------------------------------------------------------
constexpr int ATT_LD = D + 4;   // padded row stride of the shared tiles (floats)
__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* __restrict__ qkv, bf16* __restrict__ ao) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  __shared__ float sK[TOK][ATT_LD];
  __shared__ float sV[TOK][ATT_LD];
  const int cell = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < TOK * D / 4; i += 128) {
    const int r = i >> 6, c = (i & 63) * 4;
    const float* src = qkv + (size_t)(cell * TOK + r) * (3 * D);
    *reinterpret_cast<float4*>(&sK[r][c]) = *reinterpret_cast<const float4*>(src + D + c);
    *reinterpret_cast<float4*>(&sV[r][c]) = *reinterpret_cast<const float4*>(src + 2 * D + c);
  }
  const int h = warp * 2 + (lane >> 4), i = lane & 15, row = cell * TOK + i;
  float q[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 t = *reinterpret_cast<const float4*>(qkv + (size_t)row * (3 * D) + h * HD + d);
    q[d] = t.x; q[d + 1] = t.y; q[d + 2] = t.z; q[d + 3] = t.w;
  }
  __syncthreads();
  const float scale = 0.17677669529663687f;   // 1 / sqrt(32)
  float s[TOK], m = -1e30f;
#pragma unroll
  for (int j = 0; j < TOK; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc += q[d] * sK[j][h * HD + d];
    s[j] = acc * scale;
    m = fmaxf(m, s[j]);
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < TOK; ++j) { s[j] = __expf(s[j] - m); l += s[j]; }
  const float inv = 1.0f / l;
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
#pragma unroll
  for (int j = 0; j < TOK; ++j) {
    const float pj = s[j] * inv;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] += pj * sV[j][h * HD + d];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float v8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v8[k] = o[c * 8 + k];
    *reinterpret_cast<uint4*>(slab_chunk(ao, D, row, h * HD + c * 8)) = pack8(v8);
  }
}

// backward: dao [rows][256] fp32, qkv [rows][768] fp32 -> dqkv slab tensor [rows][768]
constexpr int ATT_PS = 17;
constexpr size_t attn_bwd_smem_bytes() { return (4 * TOK * ATT_LD + 2 * NHEAD * TOK * ATT_PS) * sizeof(float); }
__global__ void __launch_bounds__(128) attn_bwd_kernel(const float* __restrict__ qkv, float* __restrict__ dao, bf16* __restrict__ dqkv) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  extern __shared__ __align__(16) float sm_att[];
  float (*sQ)[ATT_LD] = reinterpret_cast<float (*)[ATT_LD]>(sm_att);
  float (*sK)[ATT_LD] = sQ + TOK;
  float (*sV)[ATT_LD] = sK + TOK;
  float (*sO)[ATT_LD] = sV + TOK;
  float* sP = sm_att + 4 * TOK * ATT_LD;          // [head][i][17]
  float* sS = sP + NHEAD * TOK * ATT_PS;
  const int cell = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < TOK * D / 4; i += 128) {
    const int r = i >> 6, c = (i & 63) * 4;
    const float* src = qkv + (size_t)(cell * TOK + r) * (3 * D);
    *reinterpret_cast<float4*>(&sQ[r][c]) = *reinterpret_cast<const float4*>(src + c);
    *reinterpret_cast<float4*>(&sK[r][c]) = *reinterpret_cast<const float4*>(src + D + c);
    *reinterpret_cast<float4*>(&sV[r][c]) = *reinterpret_cast<const float4*>(src + 2 * D + c);
    *reinterpret_cast<float4*>(&sO[r][c]) = *reinterpret_cast<const float4*>(dao + (size_t)(cell * TOK + r) * D + c);
    *reinterpret_cast<float4*>(dao + (size_t)(cell * TOK + r) * D + c) = make_float4(0.f, 0.f, 0.f, 0.f);   // consumed (see lnmod_bwd_kernel)
  }
  __syncthreads();
  const int h = warp * 2 + (lane >> 4), i = lane & 15, row = cell * TOK + i;
  const float scale = 0.17677669529663687f;
  float q[HD], dO[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) { q[d] = sQ[i][h * HD + d]; dO[d] = sO[i][h * HD + d]; }
  float pr[TOK], m = -1e30f;
#pragma unroll
  for (int j = 0; j < TOK; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc += q[d] * sK[j][h * HD + d];
    pr[j] = acc * scale;
    m = fmaxf(m, pr[j]);
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < TOK; ++j) { pr[j] = __expf(pr[j] - m); l += pr[j]; }
  const float inv = 1.0f / l;
  float dP[TOK], delta = 0.f;
#pragma unroll
  for (int j = 0; j < TOK; ++j) {
    pr[j] *= inv;
    float acc = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc += dO[d] * sV[j][h * HD + d];
    dP[j] = acc;
    delta += pr[j] * acc;
  }
  float dq[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) dq[d] = 0.f;
#pragma unroll
  for (int j = 0; j < TOK; ++j) {
    const float dS = pr[j] * (dP[j] - delta);
    sP[(h * TOK + i) * ATT_PS + j] = pr[j];
    sS[(h * TOK + i) * ATT_PS + j] = dS;
#pragma unroll
    for (int d = 0; d < HD; ++d) dq[d] += dS * sK[j][h * HD + d];
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float v8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v8[k] = dq[c * 8 + k] * scale;
    *reinterpret_cast<uint4*>(slab_chunk(dqkv, 3 * D, row, h * HD + c * 8)) = pack8(v8);
  }
  __syncwarp();
  // column phase: this lane now owns key/value row j = i of head h
  float dk[HD], dv[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
#pragma unroll
  for (int r = 0; r < TOK; ++r) {
    const float ds = sS[(h * TOK + r) * ATT_PS + i], pp = sP[(h * TOK + r) * ATT_PS + i];
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      dk[d] += ds * sQ[r][h * HD + d];
      dv[d] += pp * sO[r][h * HD + d];
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float k8[8], v8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { k8[k] = dk[c * 8 + k] * scale; v8[k] = dv[c * 8 + k]; }
    *reinterpret_cast<uint4*>(slab_chunk(dqkv, 3 * D, row, D + h * HD + c * 8)) = pack8(k8);
    *reinterpret_cast<uint4*>(slab_chunk(dqkv, 3 * D, row, 2 * D + h * HD + c * 8)) = pack8(v8);
  }
}

// ------------------------------------------------------------------------------------------
// conditioning: c = t_embedder(t) + sum_class emb[label]  (layers.py:351-364, nnets.py:389-456); A operand of the adaLN GEMM
// is SiLU(c) (layers.py:203-209).  One CTA per cell, 8 warps; a warp computes 32 outputs with lanes along K.
// ------------------------------------------------------------------------------------------
struct CondParams {
  const float* t; const int* cls_idx; int n_class; int n_cells;      // cls_idx [n_class][n_cells]
  const float* tables[8];
  const float* w0; const float* b0; const float* w2; const float* b2;   // [256][256] (out, in), [256]
  float* feat; float* h0; float* a0; float* c;                        // saved fp32 [n_cells][256]: sinusoid, pre-SiLU hidden, SiLU(hidden), c
  bf16* sc;                                                           // SiLU(c) slab tensor [cells_pad][256]
};
__device__ __forceinline__ void matvec256(const float* __restrict__ W, const float* __restrict__ bias, const float* x_sm, float* y_sm, int warp, int lane) {
  float xr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) xr[j] = x_sm[lane + 32 * j];
  for (int o = warp * 32; o < warp * 32 + 32; ++o) {
    const float* wr = W + (size_t)o * D;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += wr[lane + 32 * j] * xr[j];
    acc = sm100::warp_sum(acc);
    if (lane == 0) y_sm[o] = acc + bias[o];
  }
}
__global__ void __launch_bounds__(256) cond_fwd_kernel(const CondParams p) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  __shared__ float f[D], h[D], te[D];
  const int cell = blockIdx.x, d = threadIdx.x, warp = d >> 5, lane = d & 31;
  const float tv = p.t[cell];
  {
    const int k = d & 127;
    const float freq = expf(-9.210340371976184f * (float)k / 128.0f);
    const float arg = tv * freq;
    f[d] = (d < 128) ? cosf(arg) : sinf(arg);
    p.feat[(size_t)cell * D + d] = f[d];
  }
  __syncthreads();
  matvec256(p.w0, p.b0, f, h, warp, lane);
  __syncthreads();
  const float h0 = h[d];
  p.h0[(size_t)cell * D + d] = h0;
  __syncthreads();
  h[d] = h0 * sigmoidf_(h0);
  p.a0[(size_t)cell * D + d] = h[d];
  __syncthreads();
  matvec256(p.w2, p.b2, h, te, warp, lane);
  __syncthreads();
  float c = te[d];
  for (int k = 0; k < p.n_class; ++k) c += p.tables[k][(size_t)p.cls_idx[k * p.n_cells + cell] * D + d];
  p.c[(size_t)cell * D + d] = c;
  f[d] = c * sigmoidf_(c);
  __syncthreads();
  if (d < 32) {
    float v8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v8[j] = f[d * 8 + j];
    *reinterpret_cast<uint4*>(slab_chunk(p.sc, D, cell, d * 8)) = pack8(v8);
  }
}
// dsc = grad wrt SiLU(c) [cells][256] -> dc (saved), class-embedding grads (atomics), dh0 = silu'(h0) * (W2^T dc)
struct CondBwdParams {
  const float* dsc; const float* c; const float* h0; const int* cls_idx; int n_class; int n_cells;
  float* dtables[8]; const float* w2;
  float* dc; float* dh0;
};
__global__ void __launch_bounds__(256) cond_bwd_kernel(const CondBwdParams p) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  __shared__ float sdc[D];
  const int cell = blockIdx.x, d = threadIdx.x;
  const float c = p.c[(size_t)cell * D + d];
  const float sg = sigmoidf_(c);
  const float dc = p.dsc[(size_t)cell * D + d] * sg * (1.0f + c * (1.0f - sg));
  p.dc[(size_t)cell * D + d] = dc;
  sdc[d] = dc;
  for (int k = 0; k < p.n_class; ++k) atomicAdd(p.dtables[k] + (size_t)p.cls_idx[k * p.n_cells + cell] * D + d, dc);
  __syncthreads();
  float acc = 0.f;
#pragma unroll 8
  for (int o = 0; o < D; ++o) acc += p.w2[(size_t)o * D + d] * sdc[o];
  const float h0 = p.h0[(size_t)cell * D + d];
  const float s0 = sigmoidf_(h0);
  p.dh0[(size_t)cell * D + d] = acc * s0 * (1.0f + h0 * (1.0f - s0));
}

// dW[o][k] += sum_n dy[n][o] * x[n][k]   (small Linear layers: t_embedder, input_proj, final linear); grid (ceil(O*K/256), n chunks)
__global__ void __launch_bounds__(256) small_wgrad_kernel(const float* __restrict__ dy, int O, const float* __restrict__ x, int K, int n, int n_chunk,
                                                          float* __restrict__ dW) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= O * K) return;
  const int o = idx / K, k = idx % K;
  const int n0 = blockIdx.y * n_chunk, n1 = min(n, n0 + n_chunk);
  float acc = 0.f;
  for (int i = n0; i < n1; ++i) acc += dy[(size_t)i * O + o] * x[(size_t)i * K + k];
  atomicAdd(dW + idx, acc);
}
// db[c] += sum_n y[n][c]; grid (ceil(C/256), n chunks)
__global__ void __launch_bounds__(256) colsum_f32_kernel(const float* __restrict__ y, int C, long long ld, int n, int n_chunk, float* __restrict__ db) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const int n0 = blockIdx.y * n_chunk, n1 = min(n, n0 + n_chunk);
  float acc = 0.f;
  for (int i = n0; i < n1; ++i) acc += y[(size_t)i * ld + c];
  atomicAdd(db + c, acc);
}
// db[c] += sum over the rows of a slab tensor [rows][C] (bf16); grid (C/64, row tiles), 64 threads... one thread per column
__global__ void __launch_bounds__(64) colsum_slab_kernel(const bf16* __restrict__ t, int C, float* __restrict__ db) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const int slab = blockIdx.x, rt = blockIdx.y, c = threadIdx.x;
  const bf16* base = t + ((size_t)rt * (C >> 6) + slab) * SLAB_ELEMS;
  float acc = 0.f;
#pragma unroll 8
  for (int r = 0; r < 128; ++r) acc += __bfloat162float(base[(sm100::swz_chunk_offset(r, c >> 3) >> 1) + (c & 7)]);
  atomicAdd(db + slab * 64 + c, acc);
}
// fp32 [n][ld] columns [col0, col0 + ncols) -> slab tensor [.][C_dst] columns [dcol0, ...); one thread per 8 columns
__global__ void __launch_bounds__(256) f32_to_slab_kernel(const float* __restrict__ src, long long ld, int col0, int ncols, int n, bf16* __restrict__ dst,
                                                          int C_dst, int dcol0) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const int per_row = ncols >> 3;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)n * per_row) return;
  const int row = (int)(idx / per_row), g8 = (int)(idx % per_row) * 8;
  float v[8];
  load8(src + (size_t)row * ld + col0 + g8, v);
  *reinterpret_cast<uint4*>(slab_chunk(dst, C_dst, row, dcol0 + g8)) = pack8(v);
}

// input projection (nnets.py:290-291): X0[row][c] = sum_k x[row][k] w_in[c][k] + b_in[c] + pos[row % 16][c]
__global__ void __launch_bounds__(256) inproj_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w_in, const float* __restrict__ b_in,
                                                         const float* __restrict__ pos, float* __restrict__ X0) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  __shared__ float xs[LAT];
  const int row = blockIdx.x, c = threadIdx.x;
  if (c < LAT) xs[c] = x[(size_t)row * LAT + c];
  __syncthreads();
  float acc = b_in[c] + pos[(row & 15) * D + c];
#pragma unroll
  for (int k = 0; k < LAT; ++k) acc += xs[k] * w_in[c * LAT + k];
  X0[(size_t)row * D + c] = acc;
}
// dx[row][k] = sum_c dX0[row][c] w_in[c][k]   (only when the caller wants the gradient wrt the noisy latents)
__global__ void __launch_bounds__(256) inproj_dx_kernel(const float* __restrict__ dX0, const float* __restrict__ w_in, float* __restrict__ dx) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  __shared__ float part[16][LAT];
  const int row = blockIdx.x, k = threadIdx.x & 15, g = threadIdx.x >> 4;
  float acc = 0.f;
  for (int c = g * 16; c < g * 16 + 16; ++c) acc += dX0[(size_t)row * D + c] * w_in[c * LAT + k];
  part[g][k] = acc;
  __syncthreads();
  if (threadIdx.x < LAT) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += part[i][threadIdx.x];
    dx[(size_t)row * LAT + threadIdx.x] = s;
  }
}

// final Linear 256 -> 16 (layers.py:401): v[row][o] = sum_c hf[row][c] w_out[o][c] + b_out[o]; warp per row
__global__ void __launch_bounds__(256) final_lin_fwd_kernel(const float* __restrict__ hf, const float* __restrict__ w_out, const float* __restrict__ b_out,
                                                            float* __restrict__ v, int rows) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  float h[8];
  load8(hf + (size_t)row * D + lane * 8, h);
  float out = 0.f;
#pragma unroll
  for (int o = 0; o < LAT; ++o) {
    float w[8];
    load8(w_out + (size_t)o * D + lane * 8, w);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += h[j] * w[j];
    acc = sm100::warp_sum(acc);
    if ((int)lane == o) out = acc + b_out[o];
  }
  if (lane < LAT) v[(size_t)row * LAT + lane] = out;
}
// dhf[row][c] = sum_o dv[row][o] w_out[o][c]
__global__ void __launch_bounds__(256) final_lin_dh_kernel(const float* __restrict__ dv, const float* __restrict__ w_out, float* __restrict__ dhf) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  __shared__ float s[LAT];
  const int row = blockIdx.x, c = threadIdx.x;
  if (c < LAT) s[c] = dv[(size_t)row * LAT + c];
  __syncthreads();
  float acc = 0.f;
#pragma unroll
  for (int o = 0; o < LAT; ++o) acc += s[o] * w_out[(size_t)o * D + c];
  dhf[(size_t)row * D + c] = acc;
}

// ------------------------------------------------------------------------------------------
// optimizer: global gradient norm, AdamW (torch.optim.AdamW semantics) with norm clipping, refresh of the bf16 packed weights
// ------------------------------------------------------------------------------------------
// deterministic two-stage sum of squares (the clip coefficient must be bit-identical on every data-parallel rank, or the replicas
// drift apart by an ulp per step): partial[block] in a fixed thread order, then one block folds the partials in a fixed order
constexpr int SUMSQ_BLOCKS = 296;
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  __shared__ float red[8];
  float acc = 0.f;
  for (long long i = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * 1024) {
    if (i + 3 < n) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (long long j = i; j < n; ++j) acc += g[j] * g[j];
    }
  }
  acc = sm100::warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[1 + blockIdx.x] = s;
  }
}
__global__ void __launch_bounds__(32) sumsq_final_kernel(float* __restrict__ partial, int n_blocks) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_blocks; i += 32) acc += partial[1 + i];
  acc = sm100::warp_sum(acc);
  if (threadIdx.x == 0) partial[0] = acc;
}
struct AdamParams {
  float* p; const float* g; float* m; float* v; long long n;
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2;   // bc = 1 - beta^step
  const float* sumsq; float max_norm;                    // clip coefficient = min(1, max_norm / (sqrt(sumsq) + 1e-6)); max_norm <= 0: off
  float grad_scale;                                      // applied to g before clipping (1 / world size when the all-reduce summed)
  const int* pk_dst; bf16* pk;                           // packed bf16 position of every parameter (-1: not a GEMM weight)
};
__global__ void __launch_bounds__(256) adamw_kernel(const AdamParams a) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= a.n) return;
  float coef = a.grad_scale;
  if (a.max_norm > 0.f) {
    const float nrm = sqrtf(*a.sumsq) * a.grad_scale;
    coef *= fminf(1.0f, a.max_norm / (nrm + 1e-6f));
  }
  const float g = a.g[i] * coef;
  float p = a.p[i];
  p *= 1.0f - a.lr * a.weight_decay;
  const float m = a.beta1 * a.m[i] + (1.0f - a.beta1) * g;
  const float v = a.beta2 * a.v[i] + (1.0f - a.beta2) * g * g;
  a.m[i] = m; a.v[i] = v;
  const float denom = sqrtf(v) / sqrtf(a.bc2) + a.eps;
  p -= (a.lr / a.bc1) * (m / denom);
  a.p[i] = p;
  const int dst = a.pk_dst ? a.pk_dst[i] : -1;
  if (dst >= 0) a.pk[dst] = __float2bfloat16(p);
}
// packed[dst[i]] = bf16(p[i]) for every GEMM weight (initial pack / after load_state_dict)
__global__ void __launch_bounds__(256) repack_kernel(const float* __restrict__ p, const int* __restrict__ pk_dst, bf16* __restrict__ pk, long long n) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int dst = pk_dst[i];
  if (dst >= 0) pk[dst] = __float2bfloat16(p[i]);
}
// EMA of the parameters (ema_pytorch semantics as configured by ldm_base.yaml:51-53): ema += (1 - decay) * (p - ema)
__global__ void __launch_bounds__(256) ema_kernel(float* __restrict__ ema, const float* __restrict__ p, long long n, float one_minus_decay) {
  sm100::grid_dep_launch();   // programmatic dependent launch: let the next kernel's CTAs become resident ...
  sm100::grid_dep_wait();     // ... and do not touch global memory before the preceding kernel has completed
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n) ema[i] += one_minus_decay * (p[i] - ema[i]);
}

}  // namespace trn
