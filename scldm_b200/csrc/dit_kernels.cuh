// DiT denoiser kernels for sm_100a (scLDM generation hot path).
//
// Work unit: a "slot" = one cell-forward = 16 latent tokens x 256 channels.  A 128-row GEMM
// tile is exactly 8 slots, so LayerNorm rows and attention cells never straddle tiles.
//
// Per DiT block (reference layers.py:208-221) four launches:
//   gemm_ares<PRO_LN , EPI_QKV   >  LN + adaLN modulate (prologue)  -> QKV GEMM (tcgen05)  -> +bias, bf16
//   attn16_kernel                   16-token self attention per (slot, head) on mma.sync    -> swizzled A tiles
//   gemm_astream<EPI_RESID>         c_proj GEMM (tcgen05, A+B by bulk TMA) -> x += gate*(acc+bias)
//   gemm_ares<PRO_LN , EPI_SWIGLU>  LN + modulate -> [w1|w2] GEMM -> silu(a)*b -> swizzled A tiles
//   gemm_astream<EPI_RESID>         mlp.c_proj GEMM -> x += gate*acc
// plus per model evaluation:
//   gemm_ares<PRO_COND, EPI_MOD>    SiLU(t_emb + class_emb) -> all adaLN modulation vectors of all blocks
//   final_step_kernel               final LN/modulate/linear + CFG combine + ODE stage update + next input_proj
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
constexpr int STG_BYTES = 128 * 272;                   // epilogue staging (fp32 64-col chunk, 16 B row pad)
constexpr int STG_ARES_BYTES = 4 * A_SLAB_BYTES;       // A-resident kernel: 2 x (two 16 KB slabs), double buffered
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;            // 512
constexpr int NUM_THREADS = 64 + EPI_THREADS;          // warp0 TMA, warp1 MMA, warps 2-17 prologue/epilogue
constexpr int MAX_COMBINE = 8;

// Element (row, col) of the residual stream.  blocked = 0: row-major [rows][256].  blocked = 1: the tile-blocked layout of
// dit_stack_kernel (dit_stack.cuh): X[tile = row / 128][c4 = col / 4][row % 128][4 floats].
__host__ __device__ __forceinline__ size_t x_index(size_t row, int col, int blocked) {
  if (!blocked) return row * (size_t)D + col;
  return (row >> 7) * (size_t)(BLOCK_M * D) + ((((size_t)(col >> 2)) * BLOCK_M + (row & 127)) << 2) + (size_t)(col & 3);
}

enum { PRO_LN = 0, PRO_COND = 1 };
enum { EPI_QKV = 0, EPI_SWIGLU = 1, EPI_MOD = 2 };

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

struct AResParams {
  // ---- A operand source -----------------------------------------------------------------
  const float* X;        // PRO_LN: residual stream [rows_pad][256] fp32
  const float* mod;      // PRO_LN: modulation table [n_mod_pad][mod_stride] fp32
  ModIndex slot_mod;     // PRO_LN: slot -> row of mod
  int mod_stride;
  int mod_off_mul;       // column offset of the multiplicative chunk (h = LN(x)*(1+mul)+add)
  int mod_off_add;
  float eps;
  const float* temb;     // PRO_COND: [*, 256] timestep embedding rows
  long long temb_row_stride;  // PRO_COND: 0 => one shared row (sampling), 256 => one row per mod row
  const float* cls;      // PRO_COND: [n_mod_pad][256] summed class embeddings
  int cond_group;        // PRO_COND: > 0 => row m pairs temb row m / cond_group with cls row m % cond_group (the tables of all
  int cond_rows;         //           evaluations of an ODE solve in one GEMM); rows >= cond_rows repeat the last valid row
  // ---- B operand ------------------------------------------------------------------------
  const bf16* Wp;        // packed [n_tiles][4 slabs][256 x 64 swizzled]
  int n_tiles_total;
  int tiles_per_cta;
  const float* bias;     // [n_tiles_total*256] fp32 (EPI_QKV / EPI_MOD) or nullptr
  // ---- output ---------------------------------------------------------------------------
  bf16* out_bf16;        // EPI_QKV : packed [row_tiles][out_ld/64 slabs][128 x 64 swizzled] bf16
  float* out_f32;        // EPI_MOD : [rows_pad][out_ld] fp32 row-major
  int out_ld;
  bf16* out_packed;      // EPI_SWIGLU: [row_tiles][out_slabs][128 x 64 swizzled]
  int out_slabs;
  long long* dbg;        // optional per-CTA phase timestamps (clock64), nullptr in production
};

struct AStreamParams {
  const bf16* Ap;        // packed A [row_tiles][k_slabs][128 x 64 swizzled]
  const bf16* Wp;        // packed B [k_slabs][256 x 64 swizzled]
  int k_slabs;
  const float* bias;     // [256] or nullptr
  float* X;              // residual stream, updated in place
  const float* mod;
  ModIndex slot_mod;
  int mod_stride;
  int mod_off_gate;
  long long* dbg;
};

// ==========================================================================================
// shared pipeline pieces
// ==========================================================================================
// debug timeline: slot i of CTA b lives at dbg[b*32 + i]
__device__ __forceinline__ void dbg_stamp(long long* dbg, int slot) {
  if (dbg != nullptr) {
    long long* d = dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 32;
    d[slot] = clock64();
    if (slot == 0 || slot == 31) {   // wall-clock (ns) twins of the CTA start/end stamps + SM id
      unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
      d[slot == 0 ? 29 : 30] = (long long)g;
      if (slot == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); d[28] = sm; }
    }
  }
}

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


// ------------------------------------------------------------------------------------------
// LN + adaLN-modulate prologue fed by bulk TMA: the producer warp streams the CTA's 128 x 256 fp32 rows of X through
// two 32 KB shared-memory buffers (4 passes of 32 rows); the 16 prologue warps normalise 2 rows each per pass and write
// the bf16 A tile in the swizzled UMMA layout.  No long-latency global loads sit in the warps' dependency chains.
// ------------------------------------------------------------------------------------------
constexpr int XPASS_ROWS = 32;
constexpr int XPASS_BYTES = XPASS_ROWS * D * 4;   // 32 KB = one weight-ring stage
constexpr int XPASS_WARPS = XPASS_ROWS / 8;       // prologue warps reading each pass (8 rows per warp)

// All four X passes are in flight from t=0: passes 0,1 land in a 64 KB scratch region (epilogue staging / H buffers),
// passes 2,3 in the two weight-ring buffers that the first weight slabs do not need yet (the ring starts at physical
// buffer 2; see ring_buf()).  x_full[p] completes when pass p landed; x_empty[p-2] when a borrowed ring buffer is free.
__device__ __forceinline__ uint32_t ring_buf(uint32_t stage) { return (stage + 2u) % 3u; }   // logical stage -> physical buffer

__device__ __forceinline__ uint8_t* x_pass_buffer(int pass, uint8_t* smScratch, uint8_t* smB) {
  return pass < 2 ? smScratch + pass * XPASS_BYTES : smB + (pass - 2) * XPASS_BYTES;
}

__device__ __forceinline__ void producer_issue_x_passes(const float* X, int row_tile, uint8_t* smScratch, uint8_t* smB, uint64_t* x_full) {
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    sm100::mbar_arrive_expect_tx(&x_full[pass], XPASS_BYTES);
    sm100::bulk_g2s(x_pass_buffer(pass, smScratch, smB), X + ((size_t)row_tile * BLOCK_M + pass * XPASS_ROWS) * D, XPASS_BYTES, &x_full[pass]);
  }
}

__device__ __forceinline__ void dbg_stamp(long long* dbg, int slot);
// Row totals of four per-lane partials with a transposing butterfly: 10 shuffles instead of 20 (the prologue is bound by the
// shared-memory / shuffle pipe).  Lanes fold rows pairwise (xor 16, xor 8), finish one row each over 8 lanes, then broadcast.
__device__ __forceinline__ void reduce4_bfly(float (&s)[4], uint32_t lane) {
  const bool b4 = (lane & 16u) != 0, b3 = (lane & 8u) != 0;
  float k0 = b4 ? s[2] : s[0], k1 = b4 ? s[3] : s[1];
  k0 += __shfl_xor_sync(0xffffffffu, b4 ? s[0] : s[2], 16);
  k1 += __shfl_xor_sync(0xffffffffu, b4 ? s[1] : s[3], 16);
  float k = b3 ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, b3 ? k0 : k1, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  s[0] = __shfl_sync(0xffffffffu, k, 0);
  s[1] = __shfl_sync(0xffffffffu, k, 8);
  s[2] = __shfl_sync(0xffffffffu, k, 16);
  s[3] = __shfl_sync(0xffffffffu, k, 24);
}
__device__ __forceinline__ void ln_prologue_tma(uint8_t* smScratch, uint8_t* smB, uint64_t* x_full, uint64_t* x_empty, uint8_t* smA,
                                                const float* mod, const ModIndex& slot_mod, int mod_stride, int off_mul, int off_add,
                                                float eps, int row_tile, uint32_t ew, uint32_t lane, long long* dbg = nullptr,
                                                bool stashed = false, uint32_t x_parity = 0, int exp = 0) {
  // `stashed`: the rows were left in the pass buffers by the previous phase's residual epilogue (resid_epilogue_warp
  // <.., OUT_STASH>) instead of arriving by TMA: nothing to wait for.
  // Warp ew owns rows [8 ew, 8 ew + 8) of the tile: one X pass (ew / 4), one slot (ew / 2) -> one set of modulation
  // vectors.  Four rows are normalised at a time so that four independent reduction chains overlap.  Lane l holds
  // channels [4l, 4l+4) and [128 + 4l, 128 + 4l + 4) of a row (conflict-free 16 B shared-memory reads).
  const float inv_d = 1.0f / D;
  const int pss = ew >> 2;
  if (ew == 0 && lane == 0) dbg_stamp(dbg, 16);
  const float* mrow = mod + (size_t)slot_mod.row(row_tile * 8 + (ew >> 1)) * mod_stride;
  const float4 m0 = *reinterpret_cast<const float4*>(mrow + off_mul + lane * 4), m1 = *reinterpret_cast<const float4*>(mrow + off_mul + 128 + lane * 4);
  const float4 a0 = *reinterpret_cast<const float4*>(mrow + off_add + lane * 4), a1 = *reinterpret_cast<const float4*>(mrow + off_add + 128 + lane * 4);
  if (!stashed) sm100::mbar_wait(&x_full[pss], x_parity);
  if ((ew & 3) == 0 && lane == 0) dbg_stamp(dbg, 22 + pss);
  const uint8_t* xb = x_pass_buffer(pss, smScratch, smB) + (ew & 3) * 8 * (D * 4);
  const uint32_t xoff_even = lane * 16, xoff_odd = lane * 16;
  uint8_t* a_lo = smA + (lane >> 4) * A_SLAB_BYTES + (lane & 1) * 8;
#pragma unroll
  for (int rnd = 0; rnd < 2; ++rnd) {
    float v[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t xo = (i & 1) ? xoff_odd : xoff_even;
      const float4 x0 = *reinterpret_cast<const float4*>(xb + (rnd * 4 + i) * (D * 4) + xo);
      const float4 x1 = *reinterpret_cast<const float4*>(xb + (rnd * 4 + i) * (D * 4) + 512 + xo);
      v[i][0] = x0.x; v[i][1] = x0.y; v[i][2] = x0.z; v[i][3] = x0.w; v[i][4] = x1.x; v[i][5] = x1.y; v[i][6] = x1.z; v[i][7] = x1.w;
    }
    if (rnd == 1 && pss >= 2) {   // hand the borrowed weight-ring buffer back to the producer
      __syncwarp();
      if (lane == 0) sm100::mbar_arrive(&x_empty[pss - 2]);
    }
    float sm_[4], sq[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sm_[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sm_[i] += v[i][j];
    }
    if (exp & 4) reduce4_bfly(sm_, lane);
    else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sm_[i] += __shfl_xor_sync(0xffffffffu, sm_[i], o);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float mean = sm_[i] * inv_d;
      sq[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[i][j] -= mean; sq[i] += v[i][j] * v[i][j]; }
    }
    if (exp & 4) reduce4_bfly(sq, lane);
    else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], o);
      }
    }
    if (ew == 0 && lane == 0) dbg_stamp(dbg, rnd == 0 ? 26 : 18);
    // first use of the modulation vectors: their (two dependent) global loads have been in flight since kernel entry
    const float mul[8] = {1.f + m0.x, 1.f + m0.y, 1.f + m0.z, 1.f + m0.w, 1.f + m1.x, 1.f + m1.y, 1.f + m1.z, 1.f + m1.w};
    const float add[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float rs = rsqrtf(sq[i] * inv_d + eps);
      const int r = ew * 8 + rnd * 4 + i;   // row within the tile
      uint2 lo, hi;
      lo.x = sm100::pack_bf16x2(v[i][0] * rs * mul[0] + add[0], v[i][1] * rs * mul[1] + add[1]);
      lo.y = sm100::pack_bf16x2(v[i][2] * rs * mul[2] + add[2], v[i][3] * rs * mul[3] + add[3]);
      hi.x = sm100::pack_bf16x2(v[i][4] * rs * mul[4] + add[4], v[i][5] * rs * mul[5] + add[5]);
      hi.y = sm100::pack_bf16x2(v[i][6] * rs * mul[6] + add[6], v[i][7] * rs * mul[7] + add[7]);
      const uint32_t off = sm100::swz_chunk_offset(r, (lane & 15) >> 1);
      *reinterpret_cast<uint2*>(a_lo + off) = lo;                       // channels [4l, 4l+4)       -> slab l/16
      *reinterpret_cast<uint2*>(a_lo + 2 * A_SLAB_BYTES + off) = hi;    // channels [128+4l, +4)     -> slab 2 + l/16
    }
    if (ew == 0 && lane == 0) dbg_stamp(dbg, 17 + 2 * rnd);
  }
}

// ==========================================================================================
// GEMM with an A operand produced in-kernel (resident for the CTA's whole N loop)
//
// 18 warps: warp 0 = bulk-TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..17 = 16
// prologue/epilogue warps.  Epilogue warp w may only read TMEM lanes [32*(w%4), +32), so the 16 warps
// form a 4 (lane quadrant q) x 4 (column quarter `sub`) grid over each accumulator chunk.
// ==========================================================================================
template <int PRO, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_ares_kernel(const AResParams p) {
  constexpr uint32_t NSTAGE = 3;
  constexpr int EARLY = PRO == PRO_LN ? 1 : 3;   // weight slabs issued before the setup barrier (PRO_LN lends 2 buffers to X)
  constexpr uint32_t TMEM_COLS = 512;  // two 256-column fp32 accumulators
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;   // keep shared-space provenance (LDS/STS, not generic LD/ST); alignment is checked below
  if (threadIdx.x == 0 && (sm100::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smA = smem;                                     // 4 x 16 KB
  uint8_t* smB = smem + KSLABS_D * A_SLAB_BYTES;           // NSTAGE x 32 KB
  uint8_t* smStg = smB + NSTAGE * B_SLAB_BYTES;            // 64 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smStg + STG_ARES_BYTES);
  uint64_t* full = bars;                 // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;       // [NSTAGE]
  uint64_t* tmem_full = bars + 2 * NSTAGE;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint64_t* a_ready = tmem_empty + 2;        // [1]
  uint64_t* x_full = a_ready + 1;            // [4]  (PRO_LN: X passes landed)
  uint64_t* x_empty = x_full + 4;            // [2]  (borrowed ring buffers handed back)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(x_empty + 2);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const int row_tile = blockIdx.x;
  const int tile0 = blockIdx.y * p.tiles_per_cta;
  const int ntiles = min(p.tiles_per_cta, p.n_tiles_total - tile0);
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 0);

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < NSTAGE; ++i) {
      sm100::mbar_init(&full[i], 1);
      sm100::mbar_init(&empty[i], 1);
    }
    for (uint32_t i = 0; i < 2; ++i) {
      sm100::mbar_init(&tmem_full[i], 1);
      sm100::mbar_init(&tmem_empty[i], EPI_WARPS);
    }
    sm100::mbar_init(a_ready, EPI_WARPS);
    for (uint32_t i = 0; i < 4; ++i) sm100::mbar_init(&x_full[i], 1);
    for (uint32_t i = 0; i < 2; ++i) sm100::mbar_init(&x_empty[i], XPASS_WARPS);
    sm100::fence_barrier_init();
    // first loads go out before the setup barrier: the first weight slab(s), then (once the preceding kernel's output is
    // visible) the X tile (all four passes)
    for (int i = 0; i < EARLY; ++i) {
      sm100::mbar_arrive_expect_tx(&full[i], B_SLAB_BYTES);
      sm100::bulk_g2s(smB + ring_buf(i) * B_SLAB_BYTES, p.Wp + ((size_t)tile0 * KSLABS_D + i) * B_SLAB_ELEMS, B_SLAB_BYTES, &full[i]);
    }
    sm100::grid_dep_wait();
    if constexpr (PRO == PRO_LN) producer_issue_x_passes(p.X, row_tile, smStg, smB, x_full);
  }
  if (warp == 1) sm100::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer: stream packed weight slabs ======================
    if (lane == 0) {
      RingState rs;
      for (int i = 0; i < EARLY; ++i) rs.advance(NSTAGE);
      for (int i = EARLY; i < ntiles * KSLABS_D; ++i) {
        if constexpr (PRO == PRO_LN) {
          if (i == 1 || i == 2) sm100::mbar_wait(&x_empty[i - 1], 0);   // first use of a ring buffer that carried an X pass
        }
        sm100::mbar_wait(&empty[rs.stage], rs.phase ^ 1);
        sm100::mbar_arrive_expect_tx(&full[rs.stage], B_SLAB_BYTES);
        sm100::bulk_g2s(smB + ring_buf(rs.stage) * B_SLAB_BYTES, p.Wp + ((size_t)tile0 * KSLABS_D + i) * B_SLAB_ELEMS, B_SLAB_BYTES,
                        &full[rs.stage]);
        rs.advance(NSTAGE);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) ====================================
    if (lane == 0) {
      const uint32_t idesc = sm100::make_idesc_bf16(BLOCK_M, BLOCK_N);
      dbg_stamp(p.dbg, 1);
      sm100::mbar_wait(a_ready, 0);
      sm100::tc_fence_after();
      dbg_stamp(p.dbg, 2);
      RingState rs;
      for (int t = 0; t < ntiles; ++t) {
        const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
        sm100::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        sm100::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int ks = 0; ks < KSLABS_D; ++ks) {
          sm100::mbar_wait(&full[rs.stage], rs.phase);
          sm100::tc_fence_after();
          issue_slab_mmas(tmem_d, sm100::smem_u32(smA + ks * A_SLAB_BYTES), sm100::smem_u32(smB + ring_buf(rs.stage) * B_SLAB_BYTES),
                          idesc, ks == 0);
          sm100::umma_commit(&empty[rs.stage]);  // frees the B stage when these MMAs retire
          rs.advance(NSTAGE);
        }
        sm100::umma_commit(&tmem_full[acc]);     // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== 16 prologue + epilogue warps (512 threads) ====================
    const uint32_t ew = warp - 2;          // 0..15
    const uint32_t q = warp & 3;           // TMEM lane quadrant this warp may read
    const uint32_t sub = ew >> 2;          // column quarter
    const uint32_t etid = threadIdx.x - 64;

    // ---------- produce the A tile (128 rows x 256 K, bf16, swizzled); warp ew owns rows [8*ew, 8*ew+8) ----------
    if constexpr (PRO == PRO_LN) {
      ln_prologue_tma(smStg, smB, x_full, x_empty, smA, p.mod, p.slot_mod, p.mod_stride, p.mod_off_mul, p.mod_off_add, p.eps, row_tile, ew, lane);
    } else {
      // PRO_COND: A[m][k] = SiLU(temb[k] + cls[m][k])
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
        const float4 t0 = *reinterpret_cast<const float4*>(tr);
        const float4 t1 = *reinterpret_cast<const float4*>(tr + 4);
        const float4 c0 = *reinterpret_cast<const float4*>(cr);
        const float4 c1 = *reinterpret_cast<const float4*>(cr + 4);
        uint4 o;
        o.x = sm100::pack_bf16x2(sm100::silu(t0.x + c0.x), sm100::silu(t0.y + c0.y));
        o.y = sm100::pack_bf16x2(sm100::silu(t0.z + c0.z), sm100::silu(t0.w + c0.w));
        o.z = sm100::pack_bf16x2(sm100::silu(t1.x + c1.x), sm100::silu(t1.y + c1.y));
        o.w = sm100::pack_bf16x2(sm100::silu(t1.z + c1.z), sm100::silu(t1.w + c1.w));
        *reinterpret_cast<uint4*>(smA + (lane >> 3) * A_SLAB_BYTES + sm100::swz_chunk_offset(r, lane & 7)) = o;
      }
    }
    sm100::fence_proxy_async_smem();   // generic-proxy smem writes -> visible to UMMA
    __syncwarp();
    if (lane == 0) sm100::mbar_arrive(a_ready);
    if (etid == 0) dbg_stamp(p.dbg, 3);

    // ---------- epilogue over this CTA's N tiles ----------
    const uint32_t row = q * 32 + lane;                       // accumulator row == TMEM lane
    for (int t = 0; t < ntiles; ++t) {
      const uint32_t acc = t & 1, acc_phase = (t >> 1) & 1;
      const int tile = tile0 + t;
      sm100::mbar_wait(&tmem_full[acc], acc_phase);
      sm100::tc_fence_after();
      if (etid == 0 && t < 12) dbg_stamp(p.dbg, 4 + 2 * t);
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * BLOCK_N;

      if constexpr (EPI == EPI_QKV) {
        // 2 chunks of 128 columns; this warp: 32 columns [32*sub, +32) of the chunk.
        // acc + bias -> bf16 -> two swizzled [128][64] slabs in (double-buffered) staging -> two 16 KB bulk stores into
        // the packed activation layout [row_tile][768/64 slabs][128 x 64 swizzled] that attn16_kernel reads.
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
          const int par = (t * 2 + ch) & 1;
          uint8_t* stg = smStg + par * 2 * A_SLAB_BYTES;
          const float* bp = p.bias + tile * BLOCK_N + ch * 128 + sub * 32;
          float4 bb[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) bb[c] = *reinterpret_cast<const float4*>(bp + c * 4);
          if (etid == 0) sm100::bulk_wait_read<1>();   // the stores issued two chunks ago have drained this buffer
          sm100::named_bar_sync(1, EPI_THREADS);
          {
            uint32_t v[32];
            sm100::tmem_ld_32x32b_x32(taddr + ch * 128 + sub * 32, v);
            sm100::tmem_ld_wait();
            uint8_t* slab = stg + (sub >> 1) * A_SLAB_BYTES;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 b0 = bb[2 * c];
              const float4 b1 = bb[2 * c + 1];
              uint4 o;
              o.x = sm100::pack_bf16x2(__uint_as_float(v[c * 8 + 0]) + b0.x, __uint_as_float(v[c * 8 + 1]) + b0.y);
              o.y = sm100::pack_bf16x2(__uint_as_float(v[c * 8 + 2]) + b0.z, __uint_as_float(v[c * 8 + 3]) + b0.w);
              o.z = sm100::pack_bf16x2(__uint_as_float(v[c * 8 + 4]) + b1.x, __uint_as_float(v[c * 8 + 5]) + b1.y);
              o.w = sm100::pack_bf16x2(__uint_as_float(v[c * 8 + 6]) + b1.z, __uint_as_float(v[c * 8 + 7]) + b1.w);
              *reinterpret_cast<uint4*>(slab + sm100::swz_chunk_offset(row, (sub & 1) * 4 + c)) = o;
            }
          }
          if (ch == 1) {  // both halves of the accumulator have been read
            sm100::tc_fence_before();
            __syncwarp();
            if (lane == 0) sm100::mbar_arrive(&tmem_empty[acc]);
          }
          sm100::fence_proxy_async_smem();
          sm100::named_bar_sync(1, EPI_THREADS);
          if (etid == 0) {
            const size_t slab0 = (size_t)row_tile * (p.out_ld / BLOCK_K) + (size_t)tile * 4 + ch * 2;
            sm100::bulk_s2g(p.out_bf16 + slab0 * A_SLAB_ELEMS, stg, 2 * A_SLAB_BYTES);  // two consecutive slabs
            sm100::bulk_commit();
          }
        }
      } else if constexpr (EPI == EPI_MOD) {
        // 4 chunks of 64 fp32 columns; this warp: 16 columns; staging rows padded to 272 B
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          sm100::named_bar_sync(1, EPI_THREADS);
          {
            uint32_t v[16];
            sm100::tmem_ld_32x32b_x16(taddr + ch * 64 + sub * 16, v);
            sm100::tmem_ld_wait();
            const float* bp = p.bias + tile * BLOCK_N + ch * 64 + sub * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 b = *reinterpret_cast<const float4*>(bp + c * 4);
              float4 o;
              o.x = __uint_as_float(v[c * 4 + 0]) + b.x;
              o.y = __uint_as_float(v[c * 4 + 1]) + b.y;
              o.z = __uint_as_float(v[c * 4 + 2]) + b.z;
              o.w = __uint_as_float(v[c * 4 + 3]) + b.w;
              *reinterpret_cast<float4*>(smStg + row * 272 + (sub * 4 + c) * 16) = o;
            }
          }
          sm100::named_bar_sync(1, EPI_THREADS);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const uint32_t idx = it * EPI_THREADS + etid;
            const uint32_t r = idx >> 4, c = idx & 15;
            const float4 o = *reinterpret_cast<const float4*>(smStg + r * 272 + c * 16);
            float* dst = p.out_f32 + ((size_t)row_tile * BLOCK_M + r) * p.out_ld + tile * BLOCK_N + ch * 64 + c * 4;
            *reinterpret_cast<float4*>(dst) = o;
          }
        }
      } else {
        // EPI_SWIGLU: tile columns [0,128) = w1 rows, [128,256) = w2 rows of hidden [128*tile, +128)
        // -> two 64-wide hidden slabs written as swizzled A slabs and bulk-stored (16 KB contiguous each).
        // this warp: slab hs = sub/2, 32 hidden columns h = sub%2
        const int hs = sub >> 1, h = sub & 1;
        uint8_t* stg = smStg + (t & 1) * 2 * A_SLAB_BYTES;  // double buffered: tile t-2's bulk stores must have drained
        if (etid == 0) sm100::bulk_wait_read<1>();
        sm100::named_bar_sync(1, EPI_THREADS);
        if (tile * 2 + hs < p.out_slabs) {
          uint32_t va[32], vb[32];
          sm100::tmem_ld_32x32b_x32(taddr + hs * 64 + h * 32, va);
          sm100::tmem_ld_32x32b_x32(taddr + 128 + hs * 64 + h * 32, vb);
          sm100::tmem_ld_wait();
          uint8_t* buf = stg + hs * A_SLAB_BYTES;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float hv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              hv[j] = sm100::silu_from_half(__uint_as_float(va[c * 8 + j])) * __uint_as_float(vb[c * 8 + j]);
            uint4 o;
            o.x = sm100::pack_bf16x2(hv[0], hv[1]);
            o.y = sm100::pack_bf16x2(hv[2], hv[3]);
            o.z = sm100::pack_bf16x2(hv[4], hv[5]);
            o.w = sm100::pack_bf16x2(hv[6], hv[7]);
            *reinterpret_cast<uint4*>(buf + sm100::swz_chunk_offset(row, h * 4 + c)) = o;
          }
        }
        // TMEM has been read: release the accumulator to the MMA warp before the store handshake
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&tmem_empty[acc]);
        sm100::fence_proxy_async_smem();
        sm100::named_bar_sync(1, EPI_THREADS);
        if (etid == 0) {
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int slab = tile * 2 + s2;
            if (slab < p.out_slabs)
              sm100::bulk_s2g(p.out_packed + ((size_t)row_tile * p.out_slabs + slab) * A_SLAB_ELEMS, stg + s2 * A_SLAB_BYTES,
                              A_SLAB_BYTES);
          }
          sm100::bulk_commit();
        }
      }
      // accumulator drained -> MMA warp may overwrite it (the SwiGLU path released it right after its TMEM reads)
      if constexpr (EPI == EPI_MOD) {
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(&tmem_empty[acc]);
      }
      if (etid == 0 && t < 12) dbg_stamp(p.dbg, 5 + 2 * t);
    }
    if constexpr (EPI != EPI_MOD) {
      if (etid == 0) sm100::bulk_wait<0>();
    }
  }

  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) sm100::tmem_dealloc(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 31);
}

constexpr size_t ares_smem_bytes() {
  return 1024 + KSLABS_D * A_SLAB_BYTES + 3 * B_SLAB_BYTES + STG_ARES_BYTES + 256;
}
static_assert(2 * XPASS_BYTES <= STG_ARES_BYTES, "X pass buffers alias the epilogue staging");

// ==========================================================================================
// Residual epilogue, one warp at a time and without CTA-wide barriers:
//   X[32q + r][64*sub + c] += gate[slot(r)][c] * (acc[r][c] + bias[c])      r < 32, c < 64
// The accumulator arrives row-per-lane (tcgen05.ld 32x32b); a 4 KB warp-private staging block (128 B rows, the usual
// 16-byte XOR swizzle: conflict free both ways) turns two 32-column halves into row-contiguous order: 4 rows x 128 B per
// warp instruction, i.e. whole cache lines on the global side and conflict-free rows on the shared-memory side.
//   OUT_GLOBAL  st.global of the updated rows (stand-alone kernels)
//   OUT_STASH   the updated rows go to shared memory only (row r at stash + 1024 r, the layout of the TMA X passes that
//               ln_prologue_tma reads): the next phase normalises them from there and the producer thread writes them
//               back with one bulk-TMA store that overlaps that prologue (dit_blocks_kernel)
// ==========================================================================================
constexpr int RESID_WARP_STG = 4096;
constexpr int RESID_STG_BYTES = EPI_WARPS * RESID_WARP_STG;   // 64 KB
enum { OUT_GLOBAL = 0, OUT_STASH = 1 };
#ifndef SCLDM_WB_ITEM
#define SCLDM_WB_ITEM 8
#endif
template <bool HAS_BIAS, int OUT, typename WaitAcc>
__device__ __forceinline__ void resid_epilogue_warp(float* __restrict__ Xtile, uint32_t taddr_q, uint8_t* stg_warp, const float* smGate,
                                                    const float* smBias, uint32_t q, uint32_t sub, uint32_t lane, WaitAcc&& wait_acc,
                                                    uint8_t* stash = nullptr) {
  const uint32_t rg = lane >> 3, cchunk = lane & 7;          // row-contiguous side: row within a 4-row group, 16 B chunk
  const float* xp = Xtile + (size_t)(q * 32 + rg) * D + sub * 64 + cchunk * 4;
  float4 xr[2][8];
  auto ldx = [&](int half) {
#pragma unroll
    for (int it = 0; it < 8; ++it) xr[half][it] = __ldcg(reinterpret_cast<const float4*>(xp + (size_t)it * 4 * D + half * 32));
  };
  ldx(0);
  wait_acc();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    {
      uint32_t v[32];
      sm100::tmem_ld_32x32b_x32(taddr_q + sub * 64 + half * 32, v);
      sm100::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(stg_warp + sm100::swz_chunk_offset(lane, c)) = make_float4(
            __uint_as_float(v[c * 4 + 0]), __uint_as_float(v[c * 4 + 1]), __uint_as_float(v[c * 4 + 2]), __uint_as_float(v[c * 4 + 3]));
    }
    __syncwarp();
    if (half == 0) ldx(1);                                   // the accumulator registers are dead: fetch the second half's rows
    const int col = sub * 64 + half * 32 + cchunk * 4;
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (HAS_BIAS) bb = *reinterpret_cast<const float4*>(smBias + col);
    // rows 4 it + rg of this quadrant belong to slot 2q + it / 4: two gate vectors per half, kept in registers
    const float4 g_lo = *reinterpret_cast<const float4*>(smGate + (q * 2) * D + col);
    const float4 g_hi = *reinterpret_cast<const float4*>(smGate + (q * 2 + 1) * D + col);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const uint32_t r = it * 4 + rg;
      const float4 a = *reinterpret_cast<const float4*>(stg_warp + sm100::swz_chunk_offset(r, cchunk));
      const float4 gg = it < 4 ? g_lo : g_hi;
      float4 o = xr[half][it];
      o.x += gg.x * (a.x + bb.x); o.y += gg.y * (a.y + bb.y);
      o.z += gg.z * (a.z + bb.z); o.w += gg.w * (a.w + bb.w);
      if constexpr (OUT == OUT_GLOBAL) *reinterpret_cast<float4*>(Xtile + (size_t)(q * 32 + r) * D + col) = o;
      else *reinterpret_cast<float4*>(stash + (q * 32 + r) * (D * 4) + col * 4) = o;
    }
    __syncwarp();                                            // staging is rewritten by the second half
  }
  if constexpr (OUT == OUT_STASH) sm100::fence_proxy_async_smem();   // the stash is read by a bulk-TMA store next
}

// ==========================================================================================
// GEMM with both operands streamed by bulk TMA; epilogue x += gate * (acc + bias)
// ==========================================================================================
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_astream_resid_kernel(const AStreamParams p) {
  constexpr uint32_t NSTAGE = 3;
  constexpr uint32_t STAGE_BYTES = A_SLAB_BYTES + B_SLAB_BYTES;  // 48 KB
  constexpr uint32_t TMEM_COLS = 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;   // keep shared-space provenance (LDS/STS, not generic LD/ST); alignment is checked below
  if (threadIdx.x == 0 && (sm100::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smStage = smem;                              // NSTAGE x (A 16 KB | B 32 KB)
  uint8_t* smStg = smem + NSTAGE * STAGE_BYTES;
  float* smGate = reinterpret_cast<float*>(smStg + RESID_STG_BYTES);   // [8 cells][256]
  float* smBias = smGate + 8 * D;                                // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smBias + D);
  uint64_t* full = bars;
  uint64_t* empty = bars + NSTAGE;
  uint64_t* tmem_full = bars + 2 * NSTAGE;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const int row_tile = blockIdx.x;
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 0);

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < NSTAGE; ++i) {
      sm100::mbar_init(&full[i], 1);
      sm100::mbar_init(&empty[i], 1);
    }
    sm100::mbar_init(tmem_full, 1);
    sm100::fence_barrier_init();
    // the first ring of loads goes out before the setup barrier
    const bf16* a_src0 = p.Ap + (size_t)row_tile * p.k_slabs * A_SLAB_ELEMS;
    for (int i = 0; i < (int)NSTAGE && i < p.k_slabs; ++i) {
      sm100::mbar_arrive_expect_tx(&full[i], STAGE_BYTES);
      sm100::bulk_g2s(smStage + i * STAGE_BYTES, a_src0 + (size_t)i * A_SLAB_ELEMS, A_SLAB_BYTES, &full[i]);
      sm100::bulk_g2s(smStage + i * STAGE_BYTES + A_SLAB_BYTES, p.Wp + (size_t)i * B_SLAB_ELEMS, B_SLAB_BYTES, &full[i]);
    }
  }
  if (warp == 1) sm100::tmem_alloc(tmem_ptr_smem, TMEM_COLS);
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      RingState rs;
      const bf16* a_src = p.Ap + (size_t)row_tile * p.k_slabs * A_SLAB_ELEMS;
      for (int i = 0; i < (int)NSTAGE && i < p.k_slabs; ++i) rs.advance(NSTAGE);
      for (int ks = NSTAGE; ks < p.k_slabs; ++ks) {
        sm100::mbar_wait(&empty[rs.stage], rs.phase ^ 1);
        sm100::mbar_arrive_expect_tx(&full[rs.stage], STAGE_BYTES);
        uint8_t* st = smStage + rs.stage * STAGE_BYTES;
        sm100::bulk_g2s(st, a_src + (size_t)ks * A_SLAB_ELEMS, A_SLAB_BYTES, &full[rs.stage]);
        sm100::bulk_g2s(st + A_SLAB_BYTES, p.Wp + (size_t)ks * B_SLAB_ELEMS, B_SLAB_BYTES, &full[rs.stage]);
        rs.advance(NSTAGE);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = sm100::make_idesc_bf16(BLOCK_M, BLOCK_N);
      RingState rs;
      dbg_stamp(p.dbg, 1);
      for (int ks = 0; ks < p.k_slabs; ++ks) {
        sm100::mbar_wait(&full[rs.stage], rs.phase);
        sm100::tc_fence_after();
        if (ks == 0) dbg_stamp(p.dbg, 2);
        const uint32_t st = sm100::smem_u32(smStage + rs.stage * STAGE_BYTES);
        issue_slab_mmas(tmem_base, st, st + A_SLAB_BYTES, idesc, ks == 0);
        sm100::umma_commit(&empty[rs.stage]);
        rs.advance(NSTAGE);
      }
      sm100::umma_commit(tmem_full);
    }
  } else {
    const uint32_t ew = warp - 2;
    const uint32_t q = warp & 3;
    const uint32_t sub = ew >> 2;
    const uint32_t etid = threadIdx.x - 64;
    // while the MMAs run: stage the 8 cells' gate vectors (8 x 256 fp32) and the bias in shared memory
    {
      // 8 cells x 64 float4 = 512 float4: one per thread
      const int cell = etid >> 6, c4 = etid & 63;
      const int mr = p.slot_mod.row(row_tile * 8 + cell);
      reinterpret_cast<float4*>(smGate)[etid] =
          *reinterpret_cast<const float4*>(p.mod + (size_t)mr * p.mod_stride + p.mod_off_gate + c4 * 4);
      if (etid < 64)
        reinterpret_cast<float4*>(smBias)[etid] =
            p.bias != nullptr ? *reinterpret_cast<const float4*>(p.bias + etid * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (etid == 0) dbg_stamp(p.dbg, 3);
    sm100::named_bar_sync(1, EPI_THREADS);   // gates + bias staged (the MMAs are still running)
    resid_epilogue_warp<true, OUT_GLOBAL>(p.X + (size_t)row_tile * BLOCK_M * D, tmem_base + ((q * 32u) << 16), smStg + ew * RESID_WARP_STG, smGate, smBias,
                              q, sub, lane, [&] {
                                sm100::mbar_wait(tmem_full, 0);
                                sm100::tc_fence_after();
                                if (etid == 0) dbg_stamp(p.dbg, 4);
                              });
  }

  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) sm100::tmem_dealloc(tmem_base, TMEM_COLS);
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 31);
}

constexpr size_t astream_smem_bytes() { return 3 * (A_SLAB_BYTES + B_SLAB_BYTES) + RESID_STG_BYTES + 9 * D * 4 + 256; }

// ==========================================================================================
// Fused MLP half of a DiT block (layers.py:219-221): x += gate * c_proj( silu(w1 h) * (w2 h) ),  h = LN(x)*(1+c3)+c4
//
//   prologue   LN + modulate -> A tile (4 swizzled slabs, smem); X rows arrive by bulk TMA (ln_prologue_tma)
//   per hidden chunk j (128 hidden units):
//     M1_j     acc1[128 x 256] = A x [w1_j | w2_j]^T            (TMEM cols 0-255)
//     E1_j     acc1 -> registers -> silu(a)*b -> bf16 -> H_j (2 swizzled slabs in smem = A operand of M2_j)
//     M2_j     acc2[128 x 256] += H_j x w3_j^T                   (TMEM cols 256-511)
//   epilogue   x += gate * acc2   (smem-staged coalesced RMW; staging reuses the A tile)
//
// The MMA issue order M1_0, M1_1, M2_0, M1_2, M2_1, ... keeps the tensor pipe busy while E1_j runs, and the weight
// slabs are packed in exactly that order (pack.py: mlp stream) so the producer is one linear bulk-TMA stream.
// The hidden activations never leave the SM.
// ==========================================================================================
struct MlpFusedParams {
  float* X;               // residual stream [rows_pad][256] fp32, updated in place
  const float* mod;
  ModIndex slot_mod;
  int mod_stride;
  int mod_off_mul, mod_off_add, mod_off_gate;
  float eps;
  const bf16* Wstream;    // [n_slabs][256 x 64 swizzled] in consumption order
  int n_chunks;           // ceil(hidden/128)
  int hid_slabs;          // ceil(hidden/64)
  long long* dbg;
  int exp;                // experiment switches (SCLDM_EXP bit mask), see abi.cu
};

// Shared-memory map common to both fused phases (attention half / MLP half) so that they can alternate inside one
// persistent kernel (dit_blocks_kernel):  [0,64K) A tile | [64K,128K) q/k/v + AO slabs or H buffers | [128K,224K) weight
// ring | 1 KB q bias | mbarriers.  The residual rows of a tile pass through [64K,192K) (TMA passes or the stash).
constexpr int PH_OFF_MID = KSLABS_D * A_SLAB_BYTES;       // 64 KB
constexpr int PH_OFF_RING = PH_OFF_MID + 4 * A_SLAB_BYTES; // 128 KB
constexpr int PH_OFF_BIASQ = PH_OFF_RING + 96 * 1024;      // 224 KB
constexpr int PH_OFF_BARS = PH_OFF_BIASQ + D * 4;
constexpr int PH_BARS_PER_SET = 40;
constexpr int PH_OFF_TMEMPTR = PH_OFF_BARS + 2 * PH_BARS_PER_SET * 8;   // set 0: attention phase, set 1: MLP phase
constexpr size_t phase_smem_bytes() { return PH_OFF_TMEMPTR + 64; }
static_assert(phase_smem_bytes() <= 232448, "exceeds the 227 KB dynamic shared memory limit");
constexpr int PH_OFF_GATE = PH_OFF_RING + 80 * 1024;   // gates + c_proj bias: parked in the (then idle) tail of the weight ring
static_assert(RESID_STG_BYTES <= KSLABS_D * A_SLAB_BYTES, "residual staging lives in the dead A tile");
static_assert(PH_OFF_GATE + 9 * D * 4 <= PH_OFF_BIASQ, "gates + bias fit in the ring tail");

// Write the stashed residual rows of the tile back to global memory (issued by the producer thread at the start of the
// phase that consumes the stash).  Two bulk groups: rows 64-127 first (they sit in the weight ring, which the producer
// needs back soonest: cp.async.bulk.wait_group.read 1), then rows 0-63.
__device__ __forceinline__ void stash_write_back(float* Xtile, const uint8_t* stash) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int half = 1 - g;
    sm100::bulk_s2g(Xtile + (size_t)half * 64 * D, stash + half * 64 * D * 4, 32 * D * 4);
    sm100::bulk_s2g(Xtile + (size_t)(half * 64 + 32) * D, stash + (half * 64 + 32) * D * 4, 32 * D * 4);
    sm100::bulk_commit();
  }
}

// The mbarriers of both phase types are initialised ONCE per kernel (warp 0, one barrier per lane) and never
// re-initialised: a phase that runs for the n-th time waits with parities shifted by how often each barrier has completed
// before.  Per execution every barrier completes an even number of times, except the single-shot ones (a_ready, the final
// accumulator, x_empty, and the MLP's h_ready / h_free, 3 uses each), whose parity therefore alternates with (n & 1);
// x_full completes only in executions whose rows come by TMA.  The MLP weight ring is padded with hand-made completions
// up to an even number of rounds (mlp_phase).
struct PhaseSeq {
  uint32_t odd;     // (previous executions of this phase type) & 1
  uint32_t tma;     // (previous executions with x_tma) & 1
};
// Barrier indices inside a set (40 slots): ring full[8] | empty[8] | 8 unused, then the phase's own barriers.
enum { BI_FULL = 0, BI_EMPTY = 8, BI_PFULL = 16,
       BA_A_READY = 24, BA_ACCQ_FULL = 25, BA_ACCQ_FREE = 26, BA_AO_READY = 27, BA_AO_FREE = 28, BA_ACCP_FULL = 29, BA_X_FULL = 30, BA_X_EMPTY = 34,
       BM_A_READY = 24, BM_ACC1_FULL = 25, BM_ACC1_FREE = 26, BM_H_READY = 27, BM_H_FREE = 29, BM_ACC2_FULL = 31, BM_X_FULL = 32, BM_X_EMPTY = 36 };

// Synchronisation helpers of a phase (one CTA per tile, cta_group::1).  A CTA-pair variant (cta_group::2 MMAs, every weight slab
// split over the two SMs of a cluster) was built and measured in round 1 - 38.6k vs 41.8k cells/s, the accumulator hand-off became
// the critical path - and removed in round 2 (git history: `dit_blocks_kernel<true>`).
struct Cg {
  static constexpr uint32_t WORKER_ARRIVALS = EPI_WARPS;
  __device__ static __forceinline__ void commit(uint64_t* bar) { sm100::umma_commit(bar); }
  __device__ static __forceinline__ void arrive_mma(uint64_t* bar) { sm100::mbar_arrive(bar); }      // a worker warp (one lane) -> the MMA issuer
  __device__ static __forceinline__ void arrive_both(uint64_t* bar) { sm100::mbar_arrive(bar); }
  __device__ static __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { sm100::umma_bf16_ss(d, a, b, idesc, acc); }
  __device__ static __forceinline__ uint32_t rank() { return 0; }
};

__device__ __forceinline__ void issue_slab_mmas_cg(uint32_t tmem_d, uint32_t a_smem, uint32_t b_smem, uint32_t idesc, bool first_slab) {
  const uint64_t a_desc = sm100::make_kmajor_sw128_desc(a_smem);
  const uint64_t b_desc = sm100::make_kmajor_sw128_desc(b_smem);
#pragma unroll
  for (uint32_t k = 0; k < BLOCK_K / 16; ++k) Cg::mma(tmem_d, a_desc + 2ull * k, b_desc + 2ull * k, idesc, (first_slab && k == 0) ? 0u : 1u);
}

__device__ __forceinline__ void phase_barriers_init(uint8_t* smem) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PH_OFF_BARS);
  const uint32_t lane = threadIdx.x & 31;
  if ((threadIdx.x >> 5) == 0) {
    for (uint32_t i = lane; i < PH_BARS_PER_SET; i += 32) {
      const uint32_t W = Cg::WORKER_ARRIVALS;
      const uint32_t ca = (i == BA_A_READY || i == BA_ACCQ_FREE || i == BA_AO_READY) ? W : ((i == BA_X_EMPTY || i == BA_X_EMPTY + 1) ? XPASS_WARPS : 1);
      sm100::mbar_init(&bars[i], ca);
      const uint32_t cm = (i == BM_A_READY || i == BM_ACC1_FREE || i == BM_H_READY || i == BM_H_READY + 1) ? W
                                                                                                         : ((i == BM_X_EMPTY || i == BM_X_EMPTY + 1) ? XPASS_WARPS : 1);
      sm100::mbar_init(&bars[PH_BARS_PER_SET + i], cm);
    }
    sm100::fence_barrier_init();
  }
}

// Weight-ring geometry of a phase (96 KB ring).
struct MlpRing { static constexpr uint32_t NST = 3, STAGE = 32768, FREE0 = 2, EARLY = 1; };
struct AttnRing { static constexpr uint32_t NST = 4, STAGE = 24576, FREE0 = 3, EARLY = 1; };

// One MLP half on the CTA's tile.  Entry: every thread of the CTA, previous phase complete (CTA-wide barrier passed),
// TMEM allocated.  x_tma: the tile's rows are fetched by TMA (otherwise the previous phase stashed them).  STASH: leave
// the updated rows in shared memory for the next phase.  Exit: CTA-wide barrier passed, all async work retired.
template <bool STASH>
__device__ __forceinline__ void mlp_phase(const MlpFusedParams& p, int row_tile, uint8_t* smem, uint32_t tmem_base, bool x_tma, PhaseSeq seq) {
  using R = MlpRing;
  using G = Cg;
  constexpr uint32_t NSTAGE = R::NST;
  uint8_t* smA = smem;                                     // 4 x 16 KB (later: epilogue staging)
  uint8_t* smH = smem + PH_OFF_MID;                        // 2 x (2 x 16 KB); first the X pass buffers / stash
  uint8_t* smB = smem + PH_OFF_RING;                       // NSTAGE stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PH_OFF_BARS) + PH_BARS_PER_SET;
  uint64_t* full = bars + BI_FULL;
  uint64_t* empty = bars + BI_EMPTY;
  uint64_t* a_ready = bars + BM_A_READY;
  uint64_t* acc1_full = bars + BM_ACC1_FULL;
  uint64_t* acc1_free = bars + BM_ACC1_FREE;
  uint64_t* h_ready = bars + BM_H_READY;           // [2]
  uint64_t* h_free = bars + BM_H_FREE;             // [2]
  uint64_t* acc2_full = bars + BM_ACC2_FULL;
  uint64_t* x_full = bars + BM_X_FULL;             // [4]
  uint64_t* x_empty = bars + BM_X_EMPTY;           // [2]
  auto ring_ptr = [&](uint32_t stage) { return smB + ((stage + R::FREE0) % NSTAGE) * R::STAGE; };

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = p.n_chunks;
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 0);
  const uint32_t po = seq.odd, px = seq.tma;
  // parity shifts of the barriers that complete T (acc1) or about T/2 (H buffer b) times per execution
  const uint32_t pt = po & (uint32_t)(T & 1);
  const uint32_t ph0 = po & (uint32_t)(((T + 1) >> 1) & 1), ph1 = po & (uint32_t)((T >> 1) & 1);
  // this CTA's share of weight slab i (pair mode: rows [128 rank, 128 rank + 128) = the rank-th half of its bytes)
  auto slab_src = [&](int i) { return reinterpret_cast<const uint8_t*>(p.Wstream) + (size_t)i * B_SLAB_BYTES + G::rank() * R::STAGE; };
  if (threadIdx.x == 0) {
    // first loads: the first slab(s) into the ring buffers that carry no residual rows, then (x_tma) all four X passes
    for (uint32_t i = 0; i < R::EARLY; ++i) {
      sm100::mbar_arrive_expect_tx(&full[i], R::STAGE);
      sm100::bulk_g2s(ring_ptr(i), slab_src(i), R::STAGE, &full[i]);
    }
    if (x_tma) producer_issue_x_passes(p.X, row_tile, smH, smB, x_full);
    else stash_write_back(p.X + (size_t)row_tile * BLOCK_M * D, smH);
  }
  if (x_tma || !(p.exp & 2)) __syncthreads();   // stash mode: nobody depends on thread 0's issue work (consumers wait on mbarriers)
  auto m2_slabs = [&](int j) { return min(2, p.hid_slabs - 2 * j); };
  const int total = KSLABS_D * T + p.hid_slabs;               // weight slabs of the phase
  int rounds = (total + (int)NSTAGE - 1) / (int)NSTAGE;
  rounds += rounds & 1;
  const int padded = rounds * (int)NSTAGE;                     // ring completions incl. the hand-made ones

  if (warp == 0) {
    // ===================== producer: one linear stream of weight slabs =======================
    if (lane == 0) {
      RingState rs;
      for (uint32_t i = 0; i < R::EARLY; ++i) rs.advance(NSTAGE);   // issued during setup
      for (int i = R::EARLY; i < total; ++i) {
        if (i == (int)R::EARLY) {                                   // first use of ring buffers that carried residual rows
          sm100::mbar_wait(&x_empty[0], po);
          sm100::mbar_wait(&x_empty[1], po);
          if (!x_tma) sm100::bulk_wait_read<1>();                   // ... and the write-back has read rows 64-127 too
        }
        if (i == KSLABS_D - 1 && !x_tma) sm100::bulk_wait_read<0>();   // rows 0-63 (H buffers) read before M1_0 can complete, i.e. before E1_0
        if (i == SCLDM_WB_ITEM && !x_tma) sm100::bulk_wait<0>();       // write-back complete long before the epilogue re-reads X
        sm100::mbar_wait(&empty[rs.stage], rs.phase ^ 1);
        sm100::mbar_arrive_expect_tx(&full[rs.stage], R::STAGE);
        sm100::bulk_g2s(ring_ptr(rs.stage), slab_src(i), R::STAGE, &full[rs.stage]);
        rs.advance(NSTAGE);
      }
      for (int i = total; i < padded; ++i) {   // hand-made completions: every ring barrier ends the phase at parity 0
        sm100::mbar_wait(&empty[rs.stage], rs.phase ^ 1);
        sm100::mbar_arrive(&full[rs.stage]);
        rs.advance(NSTAGE);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && G::rank() == 0) {
      // ===================== MMA issuer (pair mode: the leader CTA, for both CTAs) ==========
      const uint32_t idesc = sm100::make_idesc_bf16(BLOCK_M, BLOCK_N);
      const uint32_t acc1 = tmem_base, acc2 = tmem_base + BLOCK_N;
      dbg_stamp(p.dbg, 1);
      sm100::mbar_wait(a_ready, po);
      sm100::tc_fence_after();
      dbg_stamp(p.dbg, 2);
      RingState rs;
      auto wait_stage = [&]() {
        sm100::mbar_wait(&full[rs.stage], rs.phase);
        sm100::tc_fence_after();
      };
      for (int j = 0; j <= T; ++j) {
        if (j < T) {  // M1_j
          if (j > 0 && !(p.exp & 1)) { sm100::mbar_wait(acc1_free, ((j - 1) & 1) ^ pt); sm100::tc_fence_after(); }
          for (int ks = 0; ks < KSLABS_D; ++ks) {
            wait_stage();
            issue_slab_mmas_cg(acc1, sm100::smem_u32(smA + ks * A_SLAB_BYTES), sm100::smem_u32(ring_ptr(rs.stage)), idesc, ks == 0);
            G::commit(&empty[rs.stage]);
            rs.advance(NSTAGE);
          }
          G::commit(acc1_full);
          // exp bit 0: hand the accumulator over while the tensor pipe is idle (tcgen05.ld is several times faster without
          // MMAs in flight), then queue M2_{j-1} and M1_{j+1} back to back
          if ((p.exp & 1) && j + 1 < T) { sm100::mbar_wait(acc1_free, (j & 1) ^ pt); sm100::tc_fence_after(); }
        }
        if (j >= 1) {  // M2_{j-1}
          const int c = j - 1, b = c & 1;
          sm100::mbar_wait(&h_ready[b], ((c >> 1) & 1) ^ (b ? ph1 : ph0));
          sm100::tc_fence_after();
          const int ns = m2_slabs(c);
          for (int s2 = 0; s2 < ns; ++s2) {
            wait_stage();
            issue_slab_mmas_cg(acc2, sm100::smem_u32(smH + (b * 2 + s2) * A_SLAB_BYTES), sm100::smem_u32(ring_ptr(rs.stage)), idesc,
                                     c == 0 && s2 == 0);
            G::commit(&empty[rs.stage]);
            rs.advance(NSTAGE);
          }
          G::commit(&h_free[b]);
        }
      }
      G::commit(acc2_full);
      for (int i = total; i < padded; ++i) {   // consume the hand-made ring completions
        wait_stage();
        G::arrive_both(&empty[rs.stage]);
        rs.advance(NSTAGE);
      }
    }
  } else {
    // ===================== 16 prologue / epilogue warps ====================================
    const uint32_t ew = warp - 2, q = warp & 3, sub = ew >> 2, etid = threadIdx.x - 64;
    const uint32_t row = q * 32 + lane;
    ln_prologue_tma(smH, smB, x_full, x_empty, smA, p.mod, p.slot_mod, p.mod_stride, p.mod_off_mul, p.mod_off_add, p.eps, row_tile, ew, lane, p.dbg, !x_tma, px, p.exp);
    sm100::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) G::arrive_mma(a_ready);
    if (etid == 0) dbg_stamp(p.dbg, 3);

    // ---------- E1_j: SwiGLU of hidden chunk j into the A slabs of M2_j ----------
    const int hs = sub >> 1, h = sub & 1;   // this warp: slab hs of the chunk, 32-column half h
    for (int j = 0; j < T; ++j) {
      sm100::mbar_wait(acc1_full, (j & 1) ^ pt);
      sm100::tc_fence_after();
      if (etid == 0 && j < 6) dbg_stamp(p.dbg, 4 + 2 * j);
      const uint32_t taddr = tmem_base + ((q * 32u) << 16);
      uint32_t va[32], vb[32];
      sm100::tmem_ld_32x32b_x32(taddr + hs * 64 + h * 32, va);
      sm100::tmem_ld_32x32b_x32(taddr + 128 + hs * 64 + h * 32, vb);
      sm100::tmem_ld_wait();
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) G::arrive_mma(acc1_free);               // M1_{j+1} may overwrite acc1
      const int b = j & 1;
      if (j >= 2) sm100::mbar_wait(&h_free[b], (((j >> 1) - 1) & 1) ^ (b ? ph1 : ph0));   // M2_{j-2} finished reading this buffer
      uint8_t* buf = smH + (b * 2 + hs) * A_SLAB_BYTES;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float hv[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) hv[jj] = sm100::silu_from_half(__uint_as_float(va[c * 8 + jj])) * __uint_as_float(vb[c * 8 + jj]);
        uint4 o;
        o.x = sm100::pack_bf16x2(hv[0], hv[1]);
        o.y = sm100::pack_bf16x2(hv[2], hv[3]);
        o.z = sm100::pack_bf16x2(hv[4], hv[5]);
        o.w = sm100::pack_bf16x2(hv[6], hv[7]);
        *reinterpret_cast<uint4*>(buf + sm100::swz_chunk_offset(row, h * 4 + c)) = o;
      }
      sm100::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) G::arrive_mma(&h_ready[b]);
      if (etid == 0 && j < 6) dbg_stamp(p.dbg, 5 + 2 * j);
    }

    // ---------- final epilogue: x += gate * acc2 ----------
    float* smGate = reinterpret_cast<float*>(smem + PH_OFF_GATE);   // every weight slab has been consumed: the ring is idle
    const int gcell = etid >> 6, gc4 = etid & 63;
    const float4 gate_v = *reinterpret_cast<const float4*>(p.mod + (size_t)p.slot_mod.row(row_tile * 8 + gcell) * p.mod_stride + p.mod_off_gate + gc4 * 4);
    resid_epilogue_warp<false, STASH ? OUT_STASH : OUT_GLOBAL>(p.X + (size_t)row_tile * BLOCK_M * D, tmem_base + ((q * 32u) << 16) + BLOCK_N, smA + ew * RESID_WARP_STG, smGate,
                                      nullptr, q, sub, lane, [&] {
                                        sm100::mbar_wait(acc2_full, po);
                                        sm100::tc_fence_after();
                                        if (etid == 0) dbg_stamp(p.dbg, 20);
                                        reinterpret_cast<float4*>(smGate)[etid] = gate_v;
                                        sm100::named_bar_sync(1, EPI_THREADS);
                                      }, smH);
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 31);
}

__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_fused_kernel(const MlpFusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;   // keep shared-space provenance (LDS/STS, not generic LD/ST); alignment is checked below
  if (threadIdx.x == 0 && (sm100::smem_u32(smem) & 1023u) != 0) __trap();
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + PH_OFF_TMEMPTR);
  if ((threadIdx.x >> 5) == 1) sm100::tmem_alloc(tmem_ptr_smem, 512);
  phase_barriers_init(smem);
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();   // X and the modulation table come from the preceding kernels
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  mlp_phase<false>(p, blockIdx.x, smem, tmem_base, true, PhaseSeq{0, 0});
  if ((threadIdx.x >> 5) == 1) sm100::tmem_dealloc(tmem_base, 512);
}

constexpr size_t mlp_fused_smem_bytes() { return phase_smem_bytes(); }
static_assert(2 * XPASS_BYTES <= 4 * A_SLAB_BYTES, "X pass buffers alias the H buffers");

// ==========================================================================================
// 16-token self attention, one warp per (slot, head), tensor cores via mma.sync m16n8k16 (bf16)
//   qkv: packed [row_tiles][12 slabs][128 x 64 swizzled] bf16; logical columns q | k | v, heads = contiguous
//        32-channel groups (layers.py:147-151)
//   out: swizzled A tiles [row_tiles][4 slabs][128 x 64] bf16 (A operand of the c_proj GEMM)
// ==========================================================================================
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
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

__global__ void __launch_bounds__(256) attn16_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out_packed, int n_slots) {
  // one CTA per slot (cell-forward), one warp per head.  The slot's q|k|v rows (16 tokens x 768 channels) are 12
  // contiguous 2 KB segments of the packed activation slabs: stage them with 12 bulk-TMA copies, then read the mma
  // fragments from shared memory (the swizzle is resolved per access).
  __shared__ __align__(128) uint8_t s_qkv[12 * 2048];
  __shared__ uint64_t s_bar;
  const int slot = blockIdx.x;
  const int head = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t g = lane >> 2, t = lane & 3;
  if (slot >= n_slots) return;
  const size_t row0 = (size_t)slot * TOK;                 // first token row of this slot
  const uint8_t* tile_base = reinterpret_cast<const uint8_t*>(qkv + (row0 >> 7) * (3 * D / BLOCK_K) * A_SLAB_ELEMS);
  const uint32_t rbase = row0 & 127;                      // multiple of 16: the 16 rows share (r & 7) patterns with r - rbase
  if (threadIdx.x == 0) {
    sm100::mbar_init(&s_bar, 1);
    sm100::fence_barrier_init();
    sm100::mbar_arrive_expect_tx(&s_bar, 12 * 2048);
#pragma unroll
    for (int sl = 0; sl < 12; ++sl)
      sm100::bulk_g2s(s_qkv + sl * 2048, tile_base + (size_t)sl * A_SLAB_BYTES + rbase * 128, 2048, &s_bar);
  }
  __syncthreads();
  sm100::mbar_wait(&s_bar, 0);
  const uint32_t s_base = sm100::smem_u32(s_qkv);
  auto chunk_addr = [&](uint32_t token, uint32_t part, uint32_t dim) -> uint32_t {
    const uint32_t col = part * D + head * HD + dim;
    // (rbase + token) & 7 == token & 7 because rbase is a multiple of 16
    return s_base + (col >> 6) * 2048 + sm100::swz_chunk_offset(token, (col & 63) >> 3);
  };
  float o[4][4];
  attn16_core(chunk_addr, lane, o);
  // O rows (g, g+8) x dims (nt*8 + 2t, +1): assemble the slot's 16 x 256 output as 4 swizzled 2 KB slab segments in
  // shared memory (reusing the q slabs, which every warp has finished reading after the barrier) and bulk-store them
  __syncthreads();
  uint8_t* s_out = s_qkv;   // q columns [0,256) = slabs 0..3 are dead now
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int col = head * HD + nt * 8 + 2 * t;
    uint8_t* seg = s_out + (col >> 6) * 2048 + (col & 7) * 2;
    *reinterpret_cast<uint32_t*>(seg + sm100::swz_chunk_offset(g, (col & 63) >> 3)) = sm100::pack_bf16x2(o[nt][0], o[nt][1]);
    *reinterpret_cast<uint32_t*>(seg + sm100::swz_chunk_offset(g + 8, (col & 63) >> 3)) = sm100::pack_bf16x2(o[nt][2], o[nt][3]);
  }
  sm100::fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint8_t* out_base = reinterpret_cast<uint8_t*>(out_packed + (row0 >> 7) * KSLABS_D * A_SLAB_ELEMS) + rbase * 128;
#pragma unroll
    for (int sl = 0; sl < KSLABS_D; ++sl) sm100::bulk_s2g(out_base + (size_t)sl * A_SLAB_BYTES, s_out + sl * 2048, 2048);
    sm100::bulk_commit();
    sm100::bulk_wait_read<0>();
  }
}

// ==========================================================================================
// Fused attention half of a DiT block (layers.py:213-218):
//     x += gate * c_proj( attention( c_attn( LN(x) * (1 + c0) + c1 ) ) )
// One CTA per 128-row tile (8 slots).  q/k/v and the attention output never leave the SM:
//
//   prologue    LN + modulate -> A tile (4 swizzled slabs, smem); X rows arrive by bulk TMA (ln_prologue_tma)
//   per head pair hp (heads 2hp, 2hp+1; 4 of them):
//     Q_hp      accq[128 x 192] = A x [Wq_hp | Wk_hp | Wv_hp]^T              (TMEM cols 0-191)
//     E_hp      accq + bias -> bf16 q|k|v slabs in smem -> 16 (slot, head) attention jobs, one per warp (mma.sync)
//               -> AO_hp slab (smem) = K slab hp of the c_proj A operand
//     P_hp      accp[128 x 256] += AO_hp x Wproj[:, 64hp:64hp+64]^T          (TMEM cols 256-511, two N=128 halves)
//   epilogue    x += gate * (accp + bias)   (resid_epilogue_warp)
//
// MMA issue order Q_0, Q_1, P_0, Q_2, P_1, Q_3, P_2, P_3 (the tensor pipe works on Q_{hp+1} while the warps run E_hp);
// the weights are packed in exactly that order (pack.py: attn stream), 24 KB per Q item (192 x 64 slab, one per K
// slab) and 16 KB per P item (128 x 64).
// ==========================================================================================
struct AttnBlockParams {
  float* X;               // residual stream [rows_pad][256] fp32, updated in place
  const float* mod;
  ModIndex slot_mod;
  int mod_stride;
  int mod_off_mul, mod_off_add, mod_off_gate;
  float eps;
  const bf16* Wstream;    // one layer of the attention weight stream (512 KB)
  const float* bias_q;    // [256] c_attn.bias[0:256].  The k bias cancels in the softmax (a per-query constant shift of the
                          // scores) and the v bias passes through the attention (rows of P sum to 1), so it is folded into
  const float* bias_proj; // [256] = c_proj.bias + c_proj.weight @ c_attn.bias[512:768]   (pack.py: b_proj_fused)
  long long* dbg;
  int exp;                // experiment switches (SCLDM_EXP bit mask), see abi.cu
};

constexpr int AB_HP = 4;                        // head pairs
constexpr int AB_QN = 192;                      // accumulator columns per head pair: q | k | v, 64 each
constexpr int AB_Q_ITEM_BYTES = AB_QN * BLOCK_K * 2;    // 24 KB
constexpr int AB_P_ITEM_BYTES = 128 * BLOCK_K * 2;      // 16 KB
constexpr size_t attn_block_smem_bytes() { return phase_smem_bytes(); }

// One attention half on the CTA's tile; same entry / exit contract as mlp_phase.
template <bool STASH>
__device__ __forceinline__ void attn_phase(const AttnBlockParams& p, int row_tile, uint8_t* smem, uint32_t tmem_base, bool x_tma, PhaseSeq seq) {
  using R = AttnRing;
  using G = Cg;
  constexpr uint32_t NSTAGE = R::NST;
  constexpr uint32_t Q_BYTES = AB_Q_ITEM_BYTES, P_BYTES = AB_P_ITEM_BYTES;   // this CTA's share of an item
  constexpr int N_ITEMS = AB_HP * (KSLABS_D + 2);
  static_assert(N_ITEMS % NSTAGE == 0 && ((N_ITEMS / NSTAGE) & 1) == 0, "every ring barrier must complete an even number of times per phase");
  uint8_t* smA = smem;
  uint8_t* smQKV = smem + PH_OFF_MID;     // q | k | v slabs of the current head pair; first residual rows 0-63
  uint8_t* smAO = smQKV + 3 * A_SLAB_BYTES;
  uint8_t* smW = smem + PH_OFF_RING;      // weight ring; first residual rows 64-127
  float* smGate = reinterpret_cast<float*>(smem + PH_OFF_GATE);       // parked once every weight item has been consumed
  float* smBiasP = smGate + 8 * D;
  float* smBiasQ = reinterpret_cast<float*>(smem + PH_OFF_BIASQ);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PH_OFF_BARS);
  uint64_t* full = bars + BI_FULL;
  uint64_t* empty = bars + BI_EMPTY;
  uint64_t* a_ready = bars + BA_A_READY;
  uint64_t* accq_full = bars + BA_ACCQ_FULL;
  uint64_t* accq_free = bars + BA_ACCQ_FREE;
  uint64_t* ao_ready = bars + BA_AO_READY;
  uint64_t* ao_free = bars + BA_AO_FREE;
  uint64_t* accp_full = bars + BA_ACCP_FULL;
  uint64_t* x_full = bars + BA_X_FULL;             // [4]
  uint64_t* x_empty = bars + BA_X_EMPTY;           // [2]
  auto ring_ptr = [&](uint32_t stage) { return smW + ((stage + R::FREE0) % NSTAGE) * R::STAGE; };

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 0);
  const uint32_t po = seq.odd, px = seq.tma;
  // item i of the stream: Q items (4 per head pair) and P items (2 per head pair) in the order Q0 Q1 P0 Q2 P1 Q3 P2 P3.
  // The producer walks them with a running source pointer; in pair mode it takes the rank-th half of every item.
  if (threadIdx.x == 0) {
    // the first item(s) go into the ring buffers that carry no residual rows; then (x_tma) all four X passes
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.Wstream);
    for (uint32_t i = 0; i < R::EARLY; ++i) {   // R::EARLY <= 4: these are Q items of head pair 0
      sm100::mbar_arrive_expect_tx(&full[i], Q_BYTES);
      sm100::bulk_g2s(ring_ptr(i), src + (size_t)i * AB_Q_ITEM_BYTES + G::rank() * Q_BYTES, Q_BYTES, &full[i]);
    }
    if (x_tma) producer_issue_x_passes(p.X, row_tile, smQKV, smW, x_full);
    else stash_write_back(p.X + (size_t)row_tile * BLOCK_M * D, smQKV);
  }
  if (x_tma || !(p.exp & 2)) __syncthreads();
  if (warp == 0) {
    // ===================== producer: one linear stream of weight items ======================
    if (lane == 0) {
      RingState rs;
      const uint32_t rank = G::rank();
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.Wstream);
      int item = 0;
      auto issue = [&](uint32_t item_bytes, uint32_t my_bytes) {
        if (item == (int)R::EARLY) {          // first use of ring buffers that carried residual rows 64-127
          sm100::mbar_wait(&x_empty[0], po);
          sm100::mbar_wait(&x_empty[1], po);
          if (!x_tma) sm100::bulk_wait_read<1>();   // ... and the write-back has read them too
        }
        if (item == KSLABS_D - 1 && !x_tma) sm100::bulk_wait_read<0>();   // rows 0-63 (q/k/v staging) read before Q_0 can complete
        if (item == SCLDM_WB_ITEM && !x_tma) sm100::bulk_wait<0>();       // write-back complete long before this phase's epilogue re-reads X
        if (item >= (int)R::EARLY) {          // the first items were issued during setup
          sm100::mbar_wait(&empty[rs.stage], rs.phase ^ 1);
          sm100::mbar_arrive_expect_tx(&full[rs.stage], my_bytes);
          sm100::bulk_g2s(ring_ptr(rs.stage), src + rank * my_bytes, my_bytes, &full[rs.stage]);
        }
        src += item_bytes;
        ++item;
        rs.advance(NSTAGE);
      };
      for (int step = 0; step <= AB_HP; ++step) {
        if (step < AB_HP)
          for (int ks = 0; ks < KSLABS_D; ++ks) issue(AB_Q_ITEM_BYTES, Q_BYTES);
        if (step >= 1)
          for (int half = 0; half < 2; ++half) issue(AB_P_ITEM_BYTES, P_BYTES);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && G::rank() == 0) {
      // ===================== MMA issuer (pair mode: the leader CTA, for both CTAs) ==========
      const uint32_t idesc_q = sm100::make_idesc_bf16(BLOCK_M, AB_QN);
      const uint32_t idesc_p = sm100::make_idesc_bf16(BLOCK_M, 128);
      const uint32_t accq = tmem_base, accp = tmem_base + 256;
      dbg_stamp(p.dbg, 1);
      sm100::mbar_wait(a_ready, po);
      sm100::tc_fence_after();
      dbg_stamp(p.dbg, 2);
      RingState rs;
      auto wait_stage = [&]() {
        sm100::mbar_wait(&full[rs.stage], rs.phase);
        sm100::tc_fence_after();
      };
      for (int step = 0; step <= AB_HP; ++step) {
        if (step < AB_HP) {  // Q_step
          if (step > 0) { sm100::mbar_wait(accq_free, (step - 1) & 1); sm100::tc_fence_after(); }
          for (int ks = 0; ks < KSLABS_D; ++ks) {
            wait_stage();
            issue_slab_mmas_cg(accq, sm100::smem_u32(smA + ks * A_SLAB_BYTES), sm100::smem_u32(ring_ptr(rs.stage)), idesc_q, ks == 0);
            G::commit(&empty[rs.stage]);
            rs.advance(NSTAGE);
          }
          G::commit(accq_full);
        }
        if (step >= 1) {  // P_{step-1}
          const int hp = step - 1;
          sm100::mbar_wait(ao_ready, hp & 1);
          sm100::tc_fence_after();
          for (int half = 0; half < 2; ++half) {
            wait_stage();
            issue_slab_mmas_cg(accp + half * 128, sm100::smem_u32(smAO), sm100::smem_u32(ring_ptr(rs.stage)), idesc_p, hp == 0);
            G::commit(&empty[rs.stage]);
            rs.advance(NSTAGE);
          }
          G::commit(ao_free);
        }
      }
      G::commit(accp_full);
    }
  } else {
    // ===================== 16 worker warps ==================================================
    const uint32_t ew = warp - 2, q = warp & 3, sub = ew >> 2, etid = threadIdx.x - 64;
    const uint32_t row = q * 32 + lane;
    ln_prologue_tma(smQKV, smW, x_full, x_empty, smA, p.mod, p.slot_mod, p.mod_stride, p.mod_off_mul, p.mod_off_add, p.eps, row_tile, ew, lane, p.dbg, !x_tma, px, p.exp);
    sm100::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) G::arrive_mma(a_ready);
    // gate / bias vectors: needed by the final epilogue only; the loads are issued now and parked in shared memory then
    const int gcell = etid >> 6, gc4 = etid & 63;
    const float4 gate_v = *reinterpret_cast<const float4*>(p.mod + (size_t)p.slot_mod.row(row_tile * 8 + gcell) * p.mod_stride + p.mod_off_gate + gc4 * 4);
    float4 bias_v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (etid < 64) bias_v = *reinterpret_cast<const float4*>(p.bias_proj + etid * 4);
    else if (etid < 128) reinterpret_cast<float4*>(smBiasQ)[etid - 64] = *reinterpret_cast<const float4*>(p.bias_q + (etid - 64) * 4);
    if (etid == 0) dbg_stamp(p.dbg, 3);

    const uint32_t taddr_q = tmem_base + ((q * 32u) << 16);
    const uint32_t slot = ew >> 1, h = ew & 1;        // this warp's attention job within a head pair
    const uint32_t g = lane >> 2, t = lane & 3;
    const uint32_t qkv_base = sm100::smem_u32(smQKV);
    for (int hp = 0; hp < AB_HP; ++hp) {
      sm100::mbar_wait(accq_full, hp & 1);
      sm100::tc_fence_after();
      if (etid == 0) dbg_stamp(p.dbg, 4 + 3 * hp);
      uint32_t v0[32], v1[16];
      sm100::tmem_ld_32x32b_x32(taddr_q + sub * 48, v0);
      sm100::tmem_ld_32x32b_x16(taddr_q + sub * 48 + 32, v1);
      sm100::tmem_ld_wait();
      sm100::tc_fence_before();
      __syncwarp();
      if (lane == 0) G::arrive_mma(accq_free);               // Q_{hp+1} may overwrite the accumulator
      sm100::named_bar_sync(1, EPI_THREADS);
#pragma unroll
      for (int c8 = 0; c8 < 6; ++c8) {
        const uint32_t gcol = sub * 48 + c8 * 8;             // accumulator column: [0,64) q, [64,128) k, [128,192) v
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (gcol < 64) {                                      // q columns carry a bias (see AttnBlockParams)
          b0 = *reinterpret_cast<const float4*>(smBiasQ + hp * 64 + gcol);
          b1 = *reinterpret_cast<const float4*>(smBiasQ + hp * 64 + gcol + 4);
        }
        const uint32_t* vv = c8 < 4 ? &v0[c8 * 8] : &v1[(c8 - 4) * 8];
        uint4 o;
        o.x = sm100::pack_bf16x2(__uint_as_float(vv[0]) + b0.x, __uint_as_float(vv[1]) + b0.y);
        o.y = sm100::pack_bf16x2(__uint_as_float(vv[2]) + b0.z, __uint_as_float(vv[3]) + b0.w);
        o.z = sm100::pack_bf16x2(__uint_as_float(vv[4]) + b1.x, __uint_as_float(vv[5]) + b1.y);
        o.w = sm100::pack_bf16x2(__uint_as_float(vv[6]) + b1.z, __uint_as_float(vv[7]) + b1.w);
        *reinterpret_cast<uint4*>(smQKV + (gcol >> 6) * A_SLAB_BYTES + sm100::swz_chunk_offset(row, (gcol & 63) >> 3)) = o;
      }
      sm100::named_bar_sync(1, EPI_THREADS);                 // q/k/v of this head pair staged
      if (etid == 0) dbg_stamp(p.dbg, 5 + 3 * hp);
      auto chunk_addr = [&](uint32_t token, uint32_t part, uint32_t dim) -> uint32_t {
        return qkv_base + part * A_SLAB_BYTES + sm100::swz_chunk_offset(slot * TOK + token, (h * HD + dim) >> 3);
      };
      float o[4][4];
      attn16_core(chunk_addr, lane, o);
      if (hp > 0) sm100::mbar_wait(ao_free, (hp - 1) & 1);   // P_{hp-1} finished reading the AO slab
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint32_t col = h * HD + nt * 8 + 2 * t;
        uint8_t* dst = smAO + (col & 7) * 2;
        *reinterpret_cast<uint32_t*>(dst + sm100::swz_chunk_offset(slot * TOK + g, col >> 3)) = sm100::pack_bf16x2(o[nt][0], o[nt][1]);
        *reinterpret_cast<uint32_t*>(dst + sm100::swz_chunk_offset(slot * TOK + g + 8, col >> 3)) = sm100::pack_bf16x2(o[nt][2], o[nt][3]);
      }
      sm100::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) G::arrive_mma(ao_ready);
      if (etid == 0) dbg_stamp(p.dbg, 6 + 3 * hp);
    }

    // ---------- final epilogue: x += gate * (accp + bias) ----------
    resid_epilogue_warp<true, STASH ? OUT_STASH : OUT_GLOBAL>(p.X + (size_t)row_tile * BLOCK_M * D, tmem_base + ((q * 32u) << 16) + 256, smA + ew * RESID_WARP_STG, smGate, smBiasP,
                                     q, sub, lane, [&] {
                                       sm100::mbar_wait(accp_full, po);
                                       sm100::tc_fence_after();
                                       if (etid == 0) dbg_stamp(p.dbg, 20);
                                       reinterpret_cast<float4*>(smGate)[etid] = gate_v;   // every MMA has retired: the A tile is dead
                                       if (etid < 64) reinterpret_cast<float4*>(smBiasP)[etid] = bias_v;
                                       sm100::named_bar_sync(1, EPI_THREADS);
                                     }, smQKV);
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) dbg_stamp(p.dbg, 31);
}

__global__ void __launch_bounds__(NUM_THREADS, 1) attn_block_kernel(const AttnBlockParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (sm100::smem_u32(smem) & 1023u) != 0) __trap();
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + PH_OFF_TMEMPTR);
  if ((threadIdx.x >> 5) == 1) sm100::tmem_alloc(tmem_ptr_smem, 512);
  phase_barriers_init(smem);
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();   // X and the modulation table come from the preceding kernels
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  attn_phase<false>(p, blockIdx.x, smem, tmem_base, true, PhaseSeq{0, 0});
  if ((threadIdx.x >> 5) == 1) sm100::tmem_dealloc(tmem_base, 512);
}

// ==========================================================================================
// The whole block stack of one DiT evaluation as ONE persistent kernel: tiles are independent through all layers, so
// each CTA walks its tiles through attention half / MLP half of every layer without ever synchronising with other CTAs.
// Between phases the residual rows stay in shared memory (stash), TMEM stays allocated, and CTAs drift out of lock-step,
// which spreads the L2-bound residual read-modify-write over time.
// ==========================================================================================
struct BlocksParams {
  AttnBlockParams attn;   // layer 0; layer l adds the strides below
  MlpFusedParams mlp;
  int n_layer, n_tiles;
  long long attn_w_stride, mlp_w_stride;   // elements per layer in the two weight streams
  long long* dbg;                          // optional phase timeline of (second tile, layer dbg_layer) per CTA
  int dbg_layer;
  int stagger_cycles;                      // CTA b starts (b % 8) * stagger_cycles late: see dit_blocks_kernel
};

__global__ void __launch_bounds__(NUM_THREADS, 1) dit_blocks_kernel(const BlocksParams bp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (sm100::smem_u32(smem) & 1023u) != 0) __trap();
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + PH_OFF_TMEMPTR);
  if ((threadIdx.x >> 5) == 1) {
    sm100::tmem_alloc(tmem_ptr_smem, 512);
  }
  phase_barriers_init(smem);
  sm100::grid_dep_launch();
  sm100::grid_dep_wait();
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Every tile costs the same, so CTAs launched together would stay in lock-step and hit the L2-bound stretches at the same
  // moment on all SMs.  An optional one-off start offset per CTA (pair) keeps them out of phase for the rest of the kernel.
  const int group = (int)blockIdx.x, n_groups = (int)gridDim.x;
  constexpr int per_group = 1, rank = 0;
  if (bp.stagger_cycles > 0) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64(), d = (long long)(group & 7) * bp.stagger_cycles;
      while (clock64() - t0 < d) {}
    }
    __syncthreads();
  }
  int* smRows = reinterpret_cast<int*>(smem + PH_OFF_TMEMPTR + 16);   // conditioning row of each of the tile's 8 slots
  uint32_t n_attn = 0, n_mlp = 0, n_tma = 0;                          // executions so far (mbarrier parities, see PhaseSeq)
  for (int t0 = group * per_group; t0 < bp.n_tiles; t0 += n_groups * per_group) {
    const int tile = t0 + rank;
    // resolve the tile's slot -> conditioning-row indices once, so that the per-phase loads of the modulation vectors are
    // not behind a second dependent global load (visible to all warps after the first phase's setup barrier; the previous
    // tile's last phase ended with a CTA-wide barrier)
    if (threadIdx.x < 8) smRows[threadIdx.x] = bp.attn.slot_mod.row(tile * 8 + threadIdx.x);
    ModIndex tile_rows{};
    tile_rows.table = smRows - tile * 8;
    tile_rows.mode = 0;
    for (int l = 0; l < bp.n_layer; ++l) {
      AttnBlockParams ap = bp.attn;
      ap.slot_mod = tile_rows;
      ap.Wstream += (size_t)l * bp.attn_w_stride;
      ap.bias_q += (size_t)l * 3 * D;
      ap.bias_proj += (size_t)l * D;
      ap.mod_off_mul += l * 6 * D; ap.mod_off_add += l * 6 * D; ap.mod_off_gate += l * 6 * D;
      const bool dbg_on = bp.dbg != nullptr && l == bp.dbg_layer && t0 == (group + n_groups) * per_group;
      ap.dbg = dbg_on ? bp.dbg : nullptr;
      attn_phase<true>(ap, tile, smem, tmem_base, l == 0, PhaseSeq{n_attn & 1, n_tma & 1});
      ++n_attn;
      if (l == 0) ++n_tma;
      MlpFusedParams mp = bp.mlp;
      mp.slot_mod = tile_rows;
      mp.Wstream += (size_t)l * bp.mlp_w_stride;
      mp.mod_off_mul += l * 6 * D; mp.mod_off_add += l * 6 * D; mp.mod_off_gate += l * 6 * D;
      mp.dbg = dbg_on ? bp.dbg + 2 * (1 << 17) : nullptr;
      if (l + 1 < bp.n_layer) mlp_phase<true>(mp, tile, smem, tmem_base, false, PhaseSeq{n_mlp & 1, 0});
      else mlp_phase<false>(mp, tile, smem, tmem_base, false, PhaseSeq{n_mlp & 1, 0});
      ++n_mlp;
    }
  }
  if ((threadIdx.x >> 5) == 1) {
    sm100::tmem_dealloc(tmem_base, 512);
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
// One block per state, 8 warps, each warp handles 2 tokens.
__global__ void __launch_bounds__(512) final_step_kernel(const StepParams p, int n_states) {
  __shared__ float s_wout[LAT * D];   // 16 KB  [o][d]
  __shared__ float s_win[LAT * D];    // 16 KB  transposed to [o][d] (w_in is [d][o]): conflict-free float4 reads
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < LAT * D; i += 512) {
    s_wout[i] = p.w_out[i];
    s_win[(i & (LAT - 1)) * D + (i >> 4)] = p.w_in[i];
  }
  __syncthreads();
  const float inv_d = 1.0f / D;
  const int tk = warp;  // one warp per (state, token); the weight tiles are loaded into shared memory once per CTA
#pragma unroll 1
  for (int state = blockIdx.x; state < n_states; state += gridDim.x) {
    int slot0, ns;
    state_slots(p, state, slot0, ns);
    float vsum = 0.f;  // lane o (<16) accumulates combined output channel o
#pragma unroll 1
    for (int k = 0; k < ns; ++k) {
      const int slot = slot0 + k;
      const float4 x0 = *reinterpret_cast<const float4*>(p.X + x_index((size_t)slot * TOK + tk, lane * 8, p.x_blocked));
      const float4 x1 = *reinterpret_cast<const float4*>(p.X + x_index((size_t)slot * TOK + tk, lane * 8 + 4, p.x_blocked));
      const float* mrow = p.mod + (size_t)p.slot_mod.row(slot) * p.mod_stride + p.mod_off_final + lane * 8;
      const float4 sh0 = *reinterpret_cast<const float4*>(mrow), sh1 = *reinterpret_cast<const float4*>(mrow + 4);
      const float4 sc0 = *reinterpret_cast<const float4*>(mrow + D), sc1 = *reinterpret_cast<const float4*>(mrow + D + 4);
      float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const float shift[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
      const float scale[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[j];
      const float mean = sm100::warp_sum(s) * inv_d;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { v[j] -= mean; ss += v[j] * v[j]; }
      const float rstd = rsqrtf(sm100::warp_sum(ss) * inv_d + p.eps);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = v[j] * rstd * (1.f + scale[j]) + shift[j];
      // 16 partial dot products per lane, then a reduce-scatter butterfly (16 shuffles instead of 16 x 5)
      float part[LAT];
#pragma unroll
      for (int o = 0; o < LAT; ++o) {
        const float4 w0 = *reinterpret_cast<const float4*>(s_wout + o * D + lane * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(s_wout + o * D + lane * 8 + 4);
        part[o] = v[0] * w0.x + v[1] * w0.y + v[2] * w0.z + v[3] * w0.w + v[4] * w1.x + v[5] * w1.y + v[6] * w1.z + v[7] * w1.w;
      }
#pragma unroll
      for (int width = 8, off = 16; width >= 1; width >>= 1, off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < width; ++i) {
          const float send = hi ? part[i] : part[i + width];
          const float keep = hi ? part[i + width] : part[i];
          part[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      // lanes 2o and 2o+1 now hold the two halves of output o
      float full = part[0] + __shfl_xor_sync(0xffffffffu, part[0], 1);
      float mine = __shfl_sync(0xffffffffu, full, (lane & 15) * 2);
      if (lane < LAT) mine += (p.b_out ? p.b_out[lane] : 0.f);
      const float coef = (ns == 1) ? 1.0f : p.coef[k];
      vsum += coef * mine;
    }
    // lane o < 16 holds v[state][tk][o]
    float x_eval = 0.f;
    if (lane < LAT) {
      const size_t idx = ((size_t)state * TOK + tk) * LAT + lane;
      if (p.v_out) p.v_out[idx] = vsum;
      if (p.do_update) {
        float acc = p.first_stage ? 0.f : p.acc[idx];
        acc += p.b_dt * vsum;
        const float xb = p.x_base[idx];
        if (p.last_stage) {
          x_eval = xb + acc;
          p.x_base[idx] = x_eval;
        } else {
          p.acc[idx] = acc;
          x_eval = xb + p.a_dt * vsum;
        }
      }
    }
    if (p.do_inproj) {
      float h[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = (p.b_in ? p.b_in[lane * 8 + j] : 0.f) + p.pos[tk * D + lane * 8 + j];
#pragma unroll
      for (int o = 0; o < LAT; ++o) {
        const float xo = __shfl_sync(0xffffffffu, x_eval, o);
        const float4 w0 = *reinterpret_cast<const float4*>(s_win + o * D + lane * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(s_win + o * D + lane * 8 + 4);
        h[0] += xo * w0.x; h[1] += xo * w0.y; h[2] += xo * w0.z; h[3] += xo * w0.w;
        h[4] += xo * w1.x; h[5] += xo * w1.y; h[6] += xo * w1.z; h[7] += xo * w1.w;
      }
      for (int k = 0; k < ns; ++k) {
        *reinterpret_cast<float4*>(p.X + x_index((size_t)(slot0 + k) * TOK + tk, lane * 8, p.x_blocked)) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(p.X + x_index((size_t)(slot0 + k) * TOK + tk, lane * 8 + 4, p.x_blocked)) = make_float4(h[4], h[5], h[6], h[7]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Tensor-core version of final_step_kernel: one warp per state, everything in mma.sync fragment layout.
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
