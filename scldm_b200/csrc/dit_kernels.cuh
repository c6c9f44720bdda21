// DiT denoiser kernels for sm_100a (scLDM generation hot path), everything except the block stack itself.
//
// Work unit: a "slot" = one cell-forward = 16 latent tokens x 256 channels.  A 128-row GEMM tile is exactly 8 slots, so
// LayerNorm rows and attention cells never straddle tiles.
//
//   mod_gemm_kernel            SiLU(t_emb + class_emb) -> the adaLN modulation vectors of every block + the final layer (tcgen05)
//   temb_kernel / cls_kernel   timestep embedding, class-embedding sums
//   inproj_kernel / final_step_tc_kernel   input projection; final layer + CFG combine + ODE stage update + next input projection
//                              (one launch each per evaluation: adaptive solvers, plain forwards, plans the whole-solve kernel
//                              does not cover)
// The block stack (reference layers.py:208-221) and the whole fixed-grid solve live in dit_stack.cuh.
#pragma once

#include "sm100.cuh"

namespace dit {

using bf16 = __nv_bfloat16;

constexpr int D = 256;      // n_embed
constexpr int TOK = 16;     // latent tokens per cell (seq_len)
constexpr int LAT = 16;     // latent channels (n_embed_input)
constexpr int NHEAD = 8;
constexpr int HD = 32;
constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;
constexpr int KSLABS_D = D / BLOCK_K;                  // 4
constexpr int A_SLAB_BYTES = BLOCK_M * BLOCK_K * 2;    // 16384
constexpr int B_SLAB_BYTES = BLOCK_N * BLOCK_K * 2;    // 32768
constexpr int A_SLAB_ELEMS = BLOCK_M * BLOCK_K;
constexpr int B_SLAB_ELEMS = BLOCK_N * BLOCK_K;
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;            // 512
constexpr int NUM_THREADS = 64 + EPI_THREADS;          // warp0 TMA, warp1 MMA, warps 2-17 prologue/epilogue
constexpr int MAX_COMBINE = 8;
constexpr int RESID_WARP_STG = 4096;                   // per-warp epilogue staging of the training GEMMs (train_kernels.cuh)
constexpr int RESID_STG_BYTES = EPI_WARPS * RESID_WARP_STG;   // 64 KB
constexpr int AB_HP = 4;                               // attention: head pairs
constexpr int AB_QN = 192;                             // accumulator columns per head pair: q | k | v, 64 each
constexpr int AB_Q_ITEM_BYTES = AB_QN * BLOCK_K * 2;   // 24 KB: [Wq | Wk | Wv] rows of a head pair, one K slab
constexpr int AB_P_ITEM_BYTES = 128 * BLOCK_K * 2;     // 16 KB: half of a c_proj K slab


// Element (row, col) of the residual stream.  blocked = 0: row-major [rows][256].  blocked = 1: the tile-blocked layout of
// dit_stack_kernel (dit_stack.cuh): X[tile = row / 128][c4 = col / 4][row % 128][4 floats].
__host__ __device__ __forceinline__ size_t x_index(size_t row, int col, int blocked) {
  if (!blocked) return row * (size_t)D + col;
  return (row >> 7) * (size_t)(BLOCK_M * D) + ((((size_t)(col >> 2)) * BLOCK_M + (row & 127)) << 2) + (size_t)(col & 3);
}


// Conditioning row of a slot: table lookup, or computed (no dependent load on the critical path)
//   mode 0: table[slot];  mode 1: identity;  mode 2: CFG layout with shared time (nnets.py:336-378 batched):
//   slots [0,n_u) -> row 0; guided cell j owns n_f slots: pass 0 -> row 0 (unconditional), pass k -> row 1 + j*(n_f-1) + k-1
struct ModIndex {
  const int* table;
  int mode, n_u, n_f, n_slots;
  __device__ __forceinline__ int row(int slot) const {
    if (mode == 0) return table[slot];
    if (slot >= n_slots) return 0;
    if (mode == 1) return slot;
    if (slot < n_u) return 0;
    const int j = (slot - n_u) / n_f, k = (slot - n_u) - j * n_f;
    return k == 0 ? 0 : 1 + j * (n_f - 1) + (k - 1);
  }
};

struct RingState {
  uint32_t stage = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t nstages) {
    if (++stage == nstages) { stage = 0; phase ^= 1; }
  }
};

// Issue the 4 K=16 MMAs of one 64-wide K slab.  a_smem / b_smem are slab base addresses (u32).
__device__ __forceinline__ void issue_slab_mmas(uint32_t tmem_d, uint32_t a_smem, uint32_t b_smem, uint32_t idesc,
                                                bool first_slab) {
  const uint64_t a_desc = sm100::make_kmajor_sw128_desc(a_smem);
  const uint64_t b_desc = sm100::make_kmajor_sw128_desc(b_smem);
#pragma unroll
  for (uint32_t k = 0; k < BLOCK_K / 16; ++k) {
    // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
    sm100::umma_bf16_ss(tmem_d, a_desc + 2ull * k, b_desc + 2ull * k, idesc, (first_slab && k == 0) ? 0u : 1u);
  }
}

// ==========================================================================================
// adaLN modulation table: out[m][n] = sum_k SiLU(temb[.][k] + cls[.][k]) * W[n][k] + bias[n]  (nnets.py:283-287 + every block's and
// the final layer's adaln_modulation Linear, layers.py:193-201, 385-395) - one GEMM with the A tile produced in the kernel and
// resident for the CTA's whole N loop.  18 warps: warp 0 = bulk-TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..17 =
// 16 prologue / epilogue warps (TMEM lane quadrant q = warp % 4, column quarter sub).
// ==========================================================================================
struct ModGemmParams {
  const float* temb;          // [*, 256] timestep embedding rows
  long long temb_row_stride;  // 0 => one shared row, 256 => one row per output row
  const float* cls;           // [n_mod_pad][256] summed class embeddings
  int cond_group;             // > 0 => row m pairs temb row m / cond_group with cls row m % cond_group (the tables of all
  int cond_rows;              //        evaluations of an ODE solve in one GEMM); rows >= cond_rows repeat the last valid row
  const bf16* Wp;             // packed [n_tiles][4 slabs][256 x 64 swizzled]
  int n_tiles_total;
  int tiles_per_cta;
  const float* bias;          // [n_tiles_total * 256]
  float* out;                 // [rows_pad][out_ld] fp32 row-major
  int out_ld;
};

constexpr int MOD_STG_BYTES = 128 * 272;   // epilogue staging: 64 fp32 columns per row, rows padded by 16 B
constexpr size_t mod_gemm_smem_bytes() { return 1024 + KSLABS_D * A_SLAB_BYTES + 3 * B_SLAB_BYTES + MOD_STG_BYTES + 256; }

__global__ void __launch_bounds__(NUM_THREADS, 1) mod_gemm_kernel(const ModGemmParams p) {
  constexpr uint32_t NSTAGE = 3;
  constexpr uint32_t TMEM_COLS = 512;  // two 256-column fp32 accumulators
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (sm100::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smA = smem;                                     // 4 x 16 KB
  uint8_t* smB = smem + KSLABS_D * A_SLAB_BYTES;           // NSTAGE x 32 KB
  uint8_t* smStg = smB + NSTAGE * B_SLAB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smStg + MOD_STG_BYTES);
  uint64_t* full = bars;                     // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;           // [NSTAGE]
  uint64_t* tmem_full = bars + 2 * NSTAGE;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* a_ready = tmem_empty + 2;        // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(a_ready + 1);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_tile = blockIdx.x;
  const int tile0 = blockIdx.y * p.tiles_per_cta;
  const int ntiles = min(p.tiles_per_cta, p.n_tiles_total - tile0);
  const int n_items = ntiles * KSLABS_D;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < NSTAGE; ++i) { sm100::mbar_init(&full[i], 1); sm100::mbar_init(&empty[i], 1); }
    for (uint32_t i = 0; i < 2; ++i) { sm100::mbar_init(&tmem_full[i], 1); sm100::mbar_init(&tmem_empty[i], EPI_WARPS); }
    sm100::mbar_init(a_ready, EPI_WARPS);
    sm100::fence_barrier_init();
    // the first weight slabs go out before the setup barrier (they do not depend on the preceding kernels)
    for (int i = 0; i < (int)NSTAGE && i < n_items; ++i) {
      sm100::mbar_arrive_expect_tx(&full[i], B_SLAB_BYTES);
      sm100::bulk_g2s(smB + i * B_SLAB_BYTES, p.Wp + ((size_t)tile0 * KSLABS_D + i) * B_SLAB_ELEMS, B_SLAB_BYTES, &full[i]);
    }
  }
  if (warp == 1) sm100::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {   // producer: stream the packed weight slabs
      RingState rs;
      for (int i = 0; i < (int)NSTAGE && i < n_items; ++i) rs.advance(NSTAGE);
      for (int i = NSTAGE; i < n_items; ++i) {
        sm100::mbar_wait(&empty[rs.stage], rs.phase ^ 1);
        sm100::mbar_arrive_expect_tx(&full[rs.stage], B_SLAB_BYTES);
        sm100::bulk_g2s(smB + rs.stage * B_SLAB_BYTES, p.Wp + ((size_t)tile0 * KSLABS_D + i) * B_SLAB_ELEMS, B_SLAB_BYTES, &full[rs.stage]);
        rs.advance(NSTAGE);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // MMA issuer
      const uint32_t idesc = sm100::make_idesc_bf16(BLOCK_M, BLOCK_N);
      sm100::mbar_wait(a_ready, 0);
      sm100::tc_fence_after();
      RingState rs;
      for (int t = 0; t < ntiles; ++t) {
        const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
        sm100::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        sm100::tc_fence_after();
        for (int ks = 0; ks < KSLABS_D; ++ks) {
          sm100::mbar_wait(&full[rs.stage], rs.phase);
          sm100::tc_fence_after();
          issue_slab_mmas(tmem_base + acc * BLOCK_N, sm100::smem_u32(smA + ks * A_SLAB_BYTES), sm100::smem_u32(smB + rs.stage * B_SLAB_BYTES), idesc, ks == 0);
          sm100::umma_commit(&empty[rs.stage]);
          rs.advance(NSTAGE);
        }
        sm100::umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    const uint32_t ew = warp - 2, q = warp & 3, sub = ew >> 2, etid = threadIdx.x - 64;
    // ---------- A tile: A[m][k] = SiLU(temb[k] + cls[m][k]), bf16, swizzled; warp ew owns rows [8 ew, 8 ew + 8) ----------
#pragma unroll 4
    for (int i = 0; i < 8; ++i) {
      const int r = ew * 8 + i;
      size_t m = (size_t)row_tile * BLOCK_M + r;
      size_t mt = m, mc = m;
      if (p.cond_group > 0) {
        if (m >= (size_t)p.cond_rows) m = (size_t)p.cond_rows - 1;
        mt = m / (size_t)p.cond_group;
        mc = m % (size_t)p.cond_group;
      }
      const float* tr = p.temb + mt * p.temb_row_stride + lane * 8;
      const float* cr = p.cls + mc * D + lane * 8;
      const float4 t0 = *reinterpret_cast<const float4*>(tr), t1 = *reinterpret_cast<const float4*>(tr + 4);
      const float4 c0 = *reinterpret_cast<const float4*>(cr), c1 = *reinterpret_cast<const float4*>(cr + 4);
      uint4 o;
      o.x = sm100::pack_bf16x2(sm100::silu(t0.x + c0.x), sm100::silu(t0.y + c0.y));
      o.y = sm100::pack_bf16x2(sm100::silu(t0.z + c0.z), sm100::silu(t0.w + c0.w));
      o.z = sm100::pack_bf16x2(sm100::silu(t1.x + c1.x), sm100::silu(t1.y + c1.y));
      o.w = sm100::pack_bf16x2(sm100::silu(t1.z + c1.z), sm100::silu(t1.w + c1.w));
      *reinterpret_cast<uint4*>(smA + (lane >> 3) * A_SLAB_BYTES + sm100::swz_chunk_offset(r, lane & 7)) = o;
    }
    sm100::fence_proxy_async_smem();   // generic-proxy smem writes -> visible to UMMA
    __syncwarp();
    if (lane == 0) sm100::mbar_arrive(a_ready);

    // ---------- epilogue over this CTA's N tiles: + bias, fp32 rows through a padded staging block ----------
    const uint32_t row = q * 32 + lane;                       // accumulator row == TMEM lane
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
      const int tile = tile0 + t;
      sm100::mbar_wait(&tmem_full[acc], acc_phase);
      sm100::tc_fence_after();
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {   // 4 chunks of 64 columns; this warp: 16 of them
        sm100::named_bar_sync(1, EPI_THREADS);
        {
          uint32_t v[16];
          sm100::tmem_ld_32x32b_x16(taddr + ch * 64 + sub * 16, v);
          sm100::tmem_ld_wait();
          const float* bp = p.bias + tile * BLOCK_N + ch * 64 + sub * 16;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 b = *reinterpret_cast<const float4*>(bp + c * 4);
            *reinterpret_cast<float4*>(smStg + row * 272 + (sub * 4 + c) * 16) =
                make_float4(__uint_as_float(v[c * 4 + 0]) + b.x, __uint_as_float(v[c * 4 + 1]) + b.y, __uint_as_float(v[c * 4 + 2]) + b.z,
                            __uint_as_float(v[c * 4 + 3]) + b.w);
          }
        }
        sm100::named_bar_sync(1, EPI_THREADS);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const uint32_t idx = it * EPI_THREADS + etid;
          const uint32_t r = idx >> 4, c = idx & 15;
          *reinterpret_cast<float4*>(p.out + ((size_t)row_tile * BLOCK_M + r) * p.out_ld + tile * BLOCK_N + ch * 64 + c * 4) =
              *reinterpret_cast<const float4*>(smStg + r * 272 + c * 16);
        }
      }
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&tmem_empty[acc]);
    }
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) sm100::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ==========================================================================================
// 16-token self attention of one (slot, head) on one warp: tensor cores via mma.sync m16n8k16 (bf16), fragments from ldmatrix
// ==========================================================================================
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// softmax(Q K^T / sqrt(32)) V for one (slot, head) on one warp.  chunk_addr(token, part, dim) returns the shared-memory
// address (u32) of the 16-byte chunk holding dims [dim, dim+8) (dim a multiple of 8) of q (part 0), k (1) or v (2) for
// that token and this head; all mma fragments come from six ldmatrix.x4 (V transposed on the fly).  The result is the
// C fragment layout of m16n8k16: o[nt] = rows (g, g+8) x dims (8 nt + 2t, +1), g = lane / 4, t = lane % 4.
template <typename ADDR>
__device__ __forceinline__ void attn16_core(ADDR&& chunk_addr, uint32_t lane, float (&o)[4][4]) {
  const uint32_t l7 = lane & 7, m = lane >> 3;     // ldmatrix: this lane supplies row l7 of 8x8 matrix m
  // S = Q K^T : A = Q (16 tokens x 32 dims) in two k-steps; B[k=dim][n=key] = K[key][dim]
  float s[2][4] = {};
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t a[4], kb[4];
    ldmatrix_x4(a, chunk_addr(l7 + 8 * (m & 1), 0, ks * 16 + 8 * (m >> 1)));          // a0..a3
    ldmatrix_x4(kb, chunk_addr((m >> 1) * 8 + l7, 1, ks * 16 + 8 * (m & 1)));         // (nt0: b0, b1), (nt1: b0, b1)
    mma_bf16_16816(s[0], a, kb[0], kb[1]);
    mma_bf16_16816(s[1], a, kb[2], kb[3]);
  }
  // softmax over the 16 keys of rows g (c0,c1) and g+8 (c2,c3); a row lives in the 4 lanes of a quad
  const float scale_log2 = 0.17677669529663687f * 1.4426950408889634f;  // 1/sqrt(32) * log2(e)
  float m0 = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
  float m1 = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  const float ms0 = m0 * scale_log2, ms1 = m1 * scale_log2;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    s[nt][0] = fast_exp2(fmaf(s[nt][0], scale_log2, -ms0));
    s[nt][1] = fast_exp2(fmaf(s[nt][1], scale_log2, -ms0));
    s[nt][2] = fast_exp2(fmaf(s[nt][2], scale_log2, -ms1));
    s[nt][3] = fast_exp2(fmaf(s[nt][3], scale_log2, -ms1));
    l0 += s[nt][0] + s[nt][1];
    l1 += s[nt][2] + s[nt][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float r0 = fast_rcp(l0), r1 = fast_rcp(l1);
  // P (normalised, bf16) as the A operand of O = P V : C-fragment layout == A-fragment layout
  uint32_t pa[4];
  pa[0] = sm100::pack_bf16x2(s[0][0] * r0, s[0][1] * r0);
  pa[1] = sm100::pack_bf16x2(s[0][2] * r1, s[0][3] * r1);
  pa[2] = sm100::pack_bf16x2(s[1][0] * r0, s[1][1] * r0);
  pa[3] = sm100::pack_bf16x2(s[1][2] * r1, s[1][3] * r1);
  // B[k=key][n=dim] = V[key][dim]: 8x8 (key, dim) tiles loaded transposed
#pragma unroll
  for (int np = 0; np < 2; ++np) {
    uint32_t vb[4];
    ldmatrix_x4_trans(vb, chunk_addr((m & 1) * 8 + l7, 2, (np * 2 + (m >> 1)) * 8));   // (nt: keys 0-7, keys 8-15), (nt+1: ...)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float (&oo)[4] = o[np * 2 + j];
      oo[0] = 0.f; oo[1] = 0.f; oo[2] = 0.f; oo[3] = 0.f;
      mma_bf16_16816(oo, pa, vb[2 * j], vb[2 * j + 1]);
    }
  }
}

// ==========================================================================================
// small fp32 kernels: timestep embedding, class-embedding sum, input projection
// ==========================================================================================
// temb[i] = W2 * SiLU(W0 * [cos(t f) | sin(t f)] + b0) + b2  (layers.py:351-364). Weights transposed [in][out].
// 1024 threads: output channel d = tid % 256, K split four ways (64 inputs each, 16 loads in flight per thread) - a single
// 256-thread CTA streaming the two 256 KB matrices was latency-bound at 40 us, which the adaptive solver pays per evaluation.
constexpr int TEMB_THREADS = 1024;
__global__ void __launch_bounds__(TEMB_THREADS) temb_kernel(const float* __restrict__ t, int n_t, const float* __restrict__ w0t,
                                                             const float* __restrict__ b0, const float* __restrict__ w2t,
                                                             const float* __restrict__ b2, float* __restrict__ temb) {
  __shared__ float f[256];
  __shared__ float h[256];
  __shared__ float part[4][256];
  const int i = blockIdx.x;
  if (i >= n_t) return;
  const int d = threadIdx.x & 255, kq = threadIdx.x >> 8;
  const float tv = t[i];
  if (kq == 0) {
    const int k = d & 127;
    const float freq = expf(-9.210340371976184f * (float)k / 128.0f);  // exp(-ln(1e4) k / half)
    const float arg = tv * freq;
    f[d] = (d < 128) ? cosf(arg) : sinf(arg);
  }
  __syncthreads();
  float acc = 0.f;
#pragma unroll 16
  for (int k = kq * 64; k < kq * 64 + 64; ++k) acc += f[k] * w0t[k * D + d];
  part[kq][d] = acc;
  __syncthreads();
  if (kq == 0) h[d] = sm100::silu(b0[d] + ((part[0][d] + part[1][d]) + (part[2][d] + part[3][d])));
  __syncthreads();
  float acc2 = 0.f;
#pragma unroll 16
  for (int k = kq * 64; k < kq * 64 + 64; ++k) acc2 += h[k] * w2t[k * D + d];
  part[kq][d] = acc2;
  __syncthreads();
  if (kq == 0) temb[(size_t)i * D + d] = b2[d] + ((part[0][d] + part[1][d]) + (part[2][d] + part[3][d]));
}

// cls[m] = sum over class tables of emb_c[idx[c][m]]   (nnets.py:403-426, 447-456)
struct ClsParams {
  const float* tables[8];
  const int* idx;   // [n_class][n_mod_pad]
  int n_class;
  int n_mod_pad;
};
__global__ void __launch_bounds__(256) cls_kernel(const ClsParams p, float* __restrict__ cls) {
  const int m = blockIdx.x, d = threadIdx.x;
  float acc = 0.f;
  for (int c = 0; c < p.n_class; ++c) acc += p.tables[c][(size_t)p.idx[c * p.n_mod_pad + m] * D + d];
  cls[(size_t)m * D + d] = acc;
}

// ------------------------------------------------------------------------------------------
// State <-> slot mapping for one model evaluation.
//   states [0, n_u)        : one slot each (slot = state), coefficient 1
//   states [n_u, n_u+n_g)  : n_f consecutive slots starting at n_u + (state-n_u)*n_f, combined as
//                            v = sum_k coef[k] * out[slot_k]   (CFG: coef = [1-sum w, w_1, ...], nnets.py:353-378)
// ------------------------------------------------------------------------------------------
struct StepParams {
  float* X;               // residual stream [slots_pad*16][256]
  const float* mod;       // modulation table
  ModIndex slot_mod;
  int mod_stride;
  int mod_off_final;      // offset of the final layer's (shift | scale) chunks in a mod row
  float eps;
  const float* w_out;     // final_layer.linear.weight [16][256]
  const float* b_out;     // [16]
  const float* w_in;      // input_proj.weight [256][16]
  const float* b_in;      // [256]
  const float* pos;       // pos_embed [16][256]
  int n_u, n_g, n_f;
  float coef[MAX_COMBINE];
  // ODE stage update (fixed-grid explicit RK whose stage s only needs k_{s-1}):
  //   acc (+)= b*dt*v ;  last stage: x_base += acc, x_eval = x_base ; else x_eval = x_base + a*dt*v
  float* x_base;          // [n_states][16][16]
  float* acc;             // [n_states][16][16]
  float* v_out;           // optional: write the combined model output here (plain forward)
  float a_dt, b_dt;
  int first_stage, last_stage;
  int do_update;          // 0: only v_out
  int do_inproj;          // project x_eval for the next evaluation
  int x_blocked;          // layout of X (x_index)
};

__device__ __forceinline__ void state_slots(const StepParams& p, int state, int& slot0, int& nslots) {
  if (state < p.n_u) { slot0 = state; nslots = 1; }
  else { slot0 = p.n_u + (state - p.n_u) * p.n_f; nslots = p.n_f; }
}

// x (state) -> X rows of every slot of that state:  input_proj(x) + pos_embed  (nnets.py:290-291)
__global__ void __launch_bounds__(256) inproj_kernel(const StepParams p) {
  __shared__ float xs[TOK * LAT];
  const int state = blockIdx.x, d = threadIdx.x;
  xs[d] = p.x_base[(size_t)state * TOK * LAT + d];
  __syncthreads();
  float w[LAT];
#pragma unroll
  for (int o = 0; o < LAT; ++o) w[o] = p.w_in[d * LAT + o];
  const float b = p.b_in ? p.b_in[d] : 0.f;
  int slot0, ns;
  state_slots(p, state, slot0, ns);
  for (int tk = 0; tk < TOK; ++tk) {
    float acc = b + p.pos[tk * D + d];
#pragma unroll
    for (int o = 0; o < LAT; ++o) acc += xs[tk * LAT + o] * w[o];
    for (int k = 0; k < ns; ++k) p.X[x_index((size_t)(slot0 + k) * TOK + tk, d, p.x_blocked)] = acc;
  }
}

// final layer (layers.py:397-401: LN(x)*(1+scale)+shift with shift = chunk 0, scale = chunk 1; Linear 256->16)
// + CFG combine + ODE stage update + input projection of the next evaluation point.
// One warp per state, everything in mma.sync fragment layout:
//   per slot: X tile (16 tokens x 256) loaded straight into A-fragment order -> LN (quad shuffles) -> modulate -> bf16
//             -> 32 x m16n8k16 with the final Linear (B fragments in smem) -> + bias -> CFG combine (coef)
//   then the ODE stage update on the 16 x 16 C fragments, and the next evaluation's input projection as
//   2 x 32 MMAs (x split into bf16 hi + lo parts so only the bf16 weight rounding remains) written to every slot.
// ------------------------------------------------------------------------------------------
struct StepTcWeights {
  const uint2* wout_frag;   // final_layer.linear.weight as B fragments [16 ks][2 nt][32 lanes]
  const uint2* win_frag;    // input_proj.weight as B fragments [32 nt][32 lanes]   (K = 16: one k-step)
};

__device__ __forceinline__ void mma_bf16_f(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(128) final_step_tc_kernel(const StepParams p, const StepTcWeights w, int n_states) {
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();   // programmatic dependent launch: X comes from the last MLP kernel
  __shared__ uint2 s_wout[32 * 32];     // 8 KB
  __shared__ uint2 s_win[32 * 32];      // 8 KB
  __shared__ float s_posb[TOK * D];     // 16 KB: pos_embed + input_proj.bias
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 32 * 32; i += 128) { s_wout[i] = w.wout_frag[i]; s_win[i] = w.win_frag[i]; }
  for (int i = tid; i < TOK * D; i += 128) s_posb[i] = p.pos[i] + (p.b_in ? p.b_in[i & (D - 1)] : 0.f);
  __syncthreads();
  const int state = blockIdx.x * 4 + warp;
  if (state >= n_states) return;
  int slot0, ns;
  state_slots(p, state, slot0, ns);
  float coef[MAX_COMBINE];
#pragma unroll
  for (int k = 0; k < MAX_COMBINE; ++k) coef[k] = p.coef[k];
  const float bo[2][2] = {{p.b_out ? p.b_out[2 * t] : 0.f, p.b_out ? p.b_out[2 * t + 1] : 0.f},
                          {p.b_out ? p.b_out[8 + 2 * t] : 0.f, p.b_out ? p.b_out[8 + 2 * t + 1] : 0.f}};
  float vsum[2][4] = {};
  const float inv_d = 1.0f / D;
#pragma unroll 1
  for (int k = 0; k < ns; ++k) {
    const int slot = slot0 + k;
    // element (row g, col 2t) of the slot; a step of 8 rows / 8 columns is a fixed stride in either layout
    const float* x0p = p.X + x_index((size_t)slot * TOK + g, 2 * t, p.x_blocked);
    const int rs8 = p.x_blocked ? 8 * 4 : 8 * D, cs8 = p.x_blocked ? 2 * BLOCK_M * 4 : 8;
    float2 xa[16][4];   // [ks]{(g, 2t), (g+8, 2t), (g, 2t+8), (g+8, 2t+8)}
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      xa[ks][0] = *reinterpret_cast<const float2*>(x0p + 2 * ks * cs8);
      xa[ks][1] = *reinterpret_cast<const float2*>(x0p + 2 * ks * cs8 + rs8);
      xa[ks][2] = *reinterpret_cast<const float2*>(x0p + (2 * ks + 1) * cs8);
      xa[ks][3] = *reinterpret_cast<const float2*>(x0p + (2 * ks + 1) * cs8 + rs8);
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      s0 += (xa[ks][0].x + xa[ks][0].y) + (xa[ks][2].x + xa[ks][2].y);
      s1 += (xa[ks][1].x + xa[ks][1].y) + (xa[ks][3].x + xa[ks][3].y);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float m0 = s0 * inv_d, m1 = s1 * inv_d;
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      xa[ks][0].x -= m0; xa[ks][0].y -= m0; xa[ks][2].x -= m0; xa[ks][2].y -= m0;
      xa[ks][1].x -= m1; xa[ks][1].y -= m1; xa[ks][3].x -= m1; xa[ks][3].y -= m1;
      q0 += xa[ks][0].x * xa[ks][0].x + xa[ks][0].y * xa[ks][0].y + xa[ks][2].x * xa[ks][2].x + xa[ks][2].y * xa[ks][2].y;
      q1 += xa[ks][1].x * xa[ks][1].x + xa[ks][1].y * xa[ks][1].y + xa[ks][3].x * xa[ks][3].x + xa[ks][3].y * xa[ks][3].y;
    }
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    const float r0 = rsqrtf(q0 * inv_d + p.eps), r1 = rsqrtf(q1 * inv_d + p.eps);
    const float* mrow = p.mod + (size_t)p.slot_mod.row(slot) * p.mod_stride + p.mod_off_final + 2 * t;   // shift | scale
    float acc[2][4] = {};
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      const float2 shA = *reinterpret_cast<const float2*>(mrow + 16 * ks), shB = *reinterpret_cast<const float2*>(mrow + 16 * ks + 8);
      const float2 scA = *reinterpret_cast<const float2*>(mrow + D + 16 * ks), scB = *reinterpret_cast<const float2*>(mrow + D + 16 * ks + 8);
      const float gAx = 1.f + scA.x, gAy = 1.f + scA.y, gBx = 1.f + scB.x, gBy = 1.f + scB.y;
      const uint32_t a0 = sm100::pack_bf16x2(xa[ks][0].x * r0 * gAx + shA.x, xa[ks][0].y * r0 * gAy + shA.y);
      const uint32_t a1 = sm100::pack_bf16x2(xa[ks][1].x * r1 * gAx + shA.x, xa[ks][1].y * r1 * gAy + shA.y);
      const uint32_t a2 = sm100::pack_bf16x2(xa[ks][2].x * r0 * gBx + shB.x, xa[ks][2].y * r0 * gBy + shB.y);
      const uint32_t a3 = sm100::pack_bf16x2(xa[ks][3].x * r1 * gBx + shB.x, xa[ks][3].y * r1 * gBy + shB.y);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const uint2 b = s_wout[(ks * 2 + nt) * 32 + lane];
        mma_bf16_f(acc[nt], a0, a1, a2, a3, b.x, b.y);
      }
    }
    const float ck = (ns == 1) ? 1.0f : coef[k];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      vsum[nt][0] += ck * (acc[nt][0] + bo[nt][0]); vsum[nt][1] += ck * (acc[nt][1] + bo[nt][1]);
      vsum[nt][2] += ck * (acc[nt][2] + bo[nt][0]); vsum[nt][3] += ck * (acc[nt][3] + bo[nt][1]);
    }
  }
  // ---- v / ODE stage update in C-fragment layout: (token g | g+8, channels 8nt+2t, +1) ----
  float xe[2][4];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const size_t idx = ((size_t)state * TOK + g + 8 * hr) * LAT + 8 * nt + 2 * t;
      const float2 v = make_float2(vsum[nt][2 * hr], vsum[nt][2 * hr + 1]);
      if (p.v_out) *reinterpret_cast<float2*>(p.v_out + idx) = v;
      float2 xev = make_float2(0.f, 0.f);
      if (p.do_update) {
        float2 ac = p.first_stage ? make_float2(0.f, 0.f) : *reinterpret_cast<const float2*>(p.acc + idx);
        ac.x += p.b_dt * v.x; ac.y += p.b_dt * v.y;
        const float2 xb = *reinterpret_cast<const float2*>(p.x_base + idx);
        if (p.last_stage) {
          xev = make_float2(xb.x + ac.x, xb.y + ac.y);
          *reinterpret_cast<float2*>(p.x_base + idx) = xev;
        } else {
          *reinterpret_cast<float2*>(p.acc + idx) = ac;
          xev = make_float2(xb.x + p.a_dt * v.x, xb.y + p.a_dt * v.y);
        }
      }
      xe[nt][2 * hr] = xev.x; xe[nt][2 * hr + 1] = xev.y;
    }
  }
  if (!p.do_inproj) return;
  // ---- input projection of the next evaluation point: h = x_eval Win^T + b + pos, x_eval = hi + lo (both bf16) ----
  uint32_t ah[4], al[4];
  {
    const float e[4][2] = {{xe[0][0], xe[0][1]}, {xe[0][2], xe[0][3]}, {xe[1][0], xe[1][1]}, {xe[1][2], xe[1][3]}};  // a0..a3
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat16 h0 = __float2bfloat16(e[i][0]), h1 = __float2bfloat16(e[i][1]);
      ah[i] = sm100::pack_bf16x2(e[i][0], e[i][1]);
      al[i] = sm100::pack_bf16x2(e[i][0] - __bfloat162float(h0), e[i][1] - __bfloat162float(h1));
    }
  }
#pragma unroll 4
  for (int nt = 0; nt < 32; ++nt) {
    const float2 p0 = *reinterpret_cast<const float2*>(s_posb + g * D + 8 * nt + 2 * t);
    const float2 p1 = *reinterpret_cast<const float2*>(s_posb + (g + 8) * D + 8 * nt + 2 * t);
    float c[4] = {p0.x, p0.y, p1.x, p1.y};
    const uint2 b = s_win[nt * 32 + lane];
    mma_bf16_f(c, ah[0], ah[1], ah[2], ah[3], b.x, b.y);
    mma_bf16_f(c, al[0], al[1], al[2], al[3], b.x, b.y);
    for (int k = 0; k < ns; ++k) {
      float* dst = p.X + x_index((size_t)(slot0 + k) * TOK + g, 8 * nt + 2 * t, p.x_blocked);
      *reinterpret_cast<float2*>(dst) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(dst + (p.x_blocked ? 8 * 4 : 8 * D)) = make_float2(c[2], c[3]);
    }
  }
}

}  // namespace dit
