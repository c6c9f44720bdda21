// dit_stack_kernel: the adaLN block stack of the DiT denoiser (reference layers.py:208-221, nnets.py:273-297) as one persistent
// kernel - and, in whole-solve mode (template argument SOLVE, see StackParams), every evaluation of a fixed-grid ODE solve with its
// input projection, final layer, CFG combine and Runge-Kutta stage updates in ONE launch.
//
// Design points (what changed against the first-generation block kernel of round 1, and why):
//   * The residual stream is kept in a TILE-BLOCKED layout in global memory / L2:  X[tile][c4 = col / 4][row (128)][4 floats].
//     A TMEM accumulator arrives row-per-lane (tcgen05.ld 32x32b); with this layout the lane that owns a row reads and writes
//     that row's residual values as 16-byte accesses that are contiguous across the warp (512 B per instruction): the
//     transposing shared-memory staging of the old residual epilogue and the shuffle reductions of the LayerNorm prologue
//     disappear.
//   * The residual epilogue of phase P and the LayerNorm + adaLN-modulate prologue of phase P + 1 are ONE step ("boundary"):
//     x_new = x_old + gate * (acc + bias) is formed in registers, written to global memory and parked in the dead
//     accumulator's own TMEM columns (tcgen05.st); the LayerNorm row statistics are per-lane partial sums combined over the
//     four column-quarter warps through 8 KB of shared memory; the second pass re-reads x_new from TMEM, normalises, modulates
//     and writes the bf16 A tile of the next phase.  No 128 KB shared-memory stash: the weight ring belongs to the producer
//     for the whole kernel, so the next phase's first weight slabs are in flight while the boundary runs.
//   * One continuous weight pipeline: producer and MMA issuer walk a single item sequence over tiles, layers and phases with a
//     3 x 32 KB ring; mbarrier parities are running use counts (no per-phase parity bookkeeping).
//   * c_proj of the attention half is one N = 256 MMA group per head pair (the two 128-row halves of the packed item are
//     contiguous = one 256 x 64 slab).
//
//   * The write-back of x_new (128 KB per boundary, SM -> L2 stores move 32 B/clk/SM = 4k cycles) is done by a dedicated
//     DRAIN warpgroup (one warp per TMEM lane quadrant) straight from the TMEM park while the workers and the tensor pipe are
//     already in the next phase; the old rows come back through the dead chunk accumulator's TMEM columns (loaded by the
//     workers while the last MMAs of the phase retire).  setmaxnreg hands the registers the producer / issuer / drain warps do
//     not need to the 16 worker warps.
//
// Roles (24 warps = 6 warpgroups): warp 0 lane 0 = bulk-TMA producer, warp 1 lane 0 = tcgen05.mma issuer (warps 2, 3 idle),
// warps 4..7 = drain warps (TMEM lane quadrant q = warp % 4), warps 8..23 = 16 worker warps (quadrant q = warp % 4, column
// quarter sub = (warp - 8) / 4).
#pragma once

#include "dit_kernels.cuh"

namespace dit {

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct StackParams {
  float* X;                 // residual stream, updated in place: tile-blocked (x_index) when io_blocked, else row-major [rows][256]
  int io_blocked;
  float* scratch;           // io_blocked == 0: gridDim.x x 128 KB of CTA-private blocked storage for the rows between the first and
                            // the last boundary of a tile (stays in L2); unused otherwise (the tile's own rows of X serve)
  const float* mod;         // adaLN table [cond rows][mod_stride]: per layer (mul1 add1 gate1 mul2 add2 gate2) x 256
  ModIndex slot_mod;
  int mod_stride;
  float eps;
  const bf16* w_attn;       // attention weight stream, per layer attn_w_stride elements (pack.py)
  const bf16* w_mlp;        // MLP weight stream
  long long attn_w_stride, mlp_w_stride;
  const float* bias_q;      // [L][768] c_attn.bias (only the q part is used: the k bias cancels in the softmax, the v bias is folded into bias_proj)
  const float* bias_proj;   // [L][256] fused c_proj bias
  int n_layer, n_tiles;
  int n_chunks, hid_slabs;  // MLP: ceil(hidden / 128) hidden chunks, ceil(hidden / 64) K slabs of c_proj
  int hid_last;             // hidden units of the last chunk, padded to a multiple of 32 (32 .. 128): its [w1|w2] tile is 2 hid_last wide
  long long* dbg;           // optional timeline of (second tile, layer dbg_layer): 128 stamps per CTA
  int dbg_layer;
  // ---- whole-solve mode (n_evals > 0): every evaluation of a fixed-grid ODE solve in this one launch.  A tile's rows never meet
  // another tile's, in time as little as across layers, so each CTA carries its tiles through ALL evaluations: input projection
  // (tcgen05, the state as bf16 hi | lo parts) -> block stack -> final LayerNorm / modulate / Linear (tcgen05) -> CFG combine ->
  // explicit Runge-Kutta stage update of the state (transport/integrators.py:100-112, nnets.py:336-378).  X is not used: the
  // residual rows live in `scratch` (one tile per CTA, L2 resident).
  int n_evals;
  long long mod_eval_stride;   // floats between the modulation tables of consecutive evaluations
  const float4* stage;         // [n_evals] {a_dt, b_dt, first_stage, last_stage}: acc (+)= b_dt v; last: x += acc; else x_eval = x + a_dt v
  const bf16* w_solve;         // input projection + pos_embed + bias slab (32 KB), final-linear slabs (8 KB), pack.py
  const float* b_out;          // [16] final_layer.linear.bias
  float* x_base;               // [n_states][16][16] ODE state, updated in place
  float* acc;                  // [n_states][16][16] stage accumulator (multi-stage methods)
  int n_u, n_g, n_f;           // state <-> slot mapping (StepParams); n_f <= 2
  int slot_shift;              // 1: one unused slot sits between the unguided and the guided slots (odd n_u), so that the two slots
                               // of a guided state are always lanes l and l ^ 16 of one warp.  Kernel-internal: slot s >= n_u of the
                               // kernel is slot s - slot_shift of the plan.
  float coef[2];
  int mod_off_final;           // column of the final layer's (shift | scale) chunks in a modulation row
};

constexpr int S2_NST = 3;
constexpr int S2_STAGE = 32768;
constexpr int S2_OFF_MID = KSLABS_D * A_SLAB_BYTES;             // 64 KB: q|k|v + AO slabs / H buffers / boundary park
constexpr int S2_OFF_RING = S2_OFF_MID + 4 * A_SLAB_BYTES;      // 128 KB
constexpr int S2_OFF_BIASQ = S2_OFF_RING + S2_NST * S2_STAGE;   // 224 KB
constexpr int S2_OFF_BIASP = S2_OFF_BIASQ + D * 4;             // c_proj bias of the attention half (boundary pass 1)
constexpr int S2_OFF_BARS = S2_OFF_BIASP + D * 4;
enum { SB_FULL = 0, SB_EMPTY = 3, SB_ACCA_FULL = 7, SB_ACCA_FREE = 8, SB_AO_READY = 9, SB_AO_FREE = 10,
       SB_H_READY = 11, SB_H_FREE = 13, SB_ACCB_FULL = 15, SB_PARK_READY = 16, SB_PARK_DRAINED = 17, SB_XFULL = 18, SB_Z_READY = 19,
       SB_A_READY = 20 /* [4]: one per K slab of the A tile */, SB_COUNT = 24 };
constexpr int S2_WARPS = 24, S2_THREADS = S2_WARPS * 32, S2_WORKER_WARP0 = 8, S2_DRAIN_WARP0 = 4;
constexpr int S2_REGS_LAUNCH = 80, S2_REGS_IDLE = 56, S2_REGS_DRAIN = 40, S2_REGS_WORKER = 96;
static_assert(S2_THREADS * S2_REGS_LAUNCH >= 128 * S2_REGS_IDLE + 128 * S2_REGS_DRAIN + 512 * S2_REGS_WORKER, "setmaxnreg budget");
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr int S2_OFF_TMEMPTR = S2_OFF_BARS + SB_COUNT * 8;
constexpr int S2_OFF_ROWS = S2_OFF_TMEMPTR + 16;
constexpr size_t stack_smem_bytes() { return S2_OFF_ROWS + 8 * 4 + 16; }
static_assert(stack_smem_bytes() <= 232448, "exceeds the 227 KB dynamic shared memory limit");
// boundary park (inside the mid region): per-slot vectors of the tile's 8 slots + the LayerNorm partials
constexpr int PK_GATE = 0, PK_MUL = 8192, PK_ADD = 16384, PK_EXCH = 24576;   // 32 KB = one H buffer

enum { B_FIRST = 0, B_MID = 1, B_LAST = 2 };

struct BoundaryArgs {
  const float* xin;      // column 0 of this lane's row of x_old; consecutive 4-column groups are in_cs floats apart
  int in_cs;
  const float* mod;
  const int* rows;       // shared memory: conditioning row of each of the tile's 8 slots
  int mod_stride;
  int off_gate;          // column offset of the finishing phase's gate chunk          (B_MID, B_LAST)
  const float* bias;     // c_proj bias of the finishing phase (attention half) or unused
  int off_mul, off_add;  // LayerNorm modulation of the starting phase                  (B_FIRST, B_MID)
  float eps;
  const float* next_bias_q;   // q bias of the attention half that starts after this boundary (nullptr: an MLP half starts)
  float* sm_bias_q;
};

struct BoundarySync {
  uint64_t* accB_full; uint32_t accB_parity;
  uint64_t* a_ready;
  uint64_t* park_ready;
  uint64_t* park_drained; uint32_t n_drained;   // drained boundaries so far (all of them must be complete before x_old is read)
  bool drain;            // hand the parked rows to the drain warps (false: nobody needs them in global memory)
};

// One boundary on the calling worker warp.  `region_free()` returns once every worker warp has finished the phase's last chunk
// (the chunk accumulator's TMEM columns and the park region are dead); it is called after the global loads have been issued.
template <int KIND, bool HAS_BIAS, bool HAS_X, typename RegionFree>
__device__ __forceinline__ void boundary_step(const BoundaryArgs& b, const BoundarySync& sy, uint8_t* smem, uint8_t* park, uint32_t tmem_base,
                                              uint32_t q, uint32_t sub, uint32_t lane, uint32_t etid, RegionFree&& region_free, long long* dbg) {
  const uint32_t row = q * 32 + lane;
  // ---- issue every global load first: this thread's share of the per-slot vectors, then the first half of the old rows ----
  const int pslot = etid >> 6, pc4 = etid & 63;
  const float* mrow = b.mod + (size_t)b.rows[pslot] * b.mod_stride + pc4 * 4;
  float4 gate_v = make_float4(0.f, 0.f, 0.f, 0.f), mul_v = gate_v, add_v = gate_v, bias_v = gate_v;
  if constexpr (KIND != B_FIRST) gate_v = b.off_gate >= 0 ? *reinterpret_cast<const float4*>(mrow + b.off_gate) : make_float4(1.f, 1.f, 1.f, 1.f);
  if constexpr (KIND != B_LAST) {
    mul_v = *reinterpret_cast<const float4*>(mrow + b.off_mul);
    add_v = *reinterpret_cast<const float4*>(mrow + b.off_add);
  }
  if constexpr (HAS_BIAS) { if (etid < 64) bias_v = *reinterpret_cast<const float4*>(b.bias + etid * 4); }
  // the drain warps have written every earlier x_new (the rows read below) and are done with the TMEM park
  if (sy.n_drained > 0) sm100::mbar_wait(sy.park_drained, (sy.n_drained - 1) & 1);
  // Column ownership of this warp: in every K slab j of the tile the 16 columns [64 j + 16 sub, + 16).  The passes walk the slabs in
  // order ("step" st = slab st), so that the A tile of the next phase completes slab by slab and its first MMA group can start on
  // slab 0 while the later slabs are still being normalised (per-slab a_ready barriers).
  uint32_t xr[32], xr2[32];   // x_old of steps 0, 1 and 2, 3
  auto load_x = [&](uint32_t (&dst)[32], int pair) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c4 = (2 * pair + (i >> 2)) * 16 + sub * 4 + (i & 3);   // 4-column group of the row
      const float4 t = __ldcg(reinterpret_cast<const float4*>(b.xin + c4 * b.in_cs));
      dst[4 * i + 0] = __float_as_uint(t.x); dst[4 * i + 1] = __float_as_uint(t.y);
      dst[4 * i + 2] = __float_as_uint(t.z); dst[4 * i + 3] = __float_as_uint(t.w);
    }
  };
  constexpr bool has_x = HAS_X;   // false: the accumulator IS the new row (input projection incl. pos_embed + bias)
  if constexpr (has_x) {
    load_x(xr, 0);
    if constexpr (KIND != B_FIRST) load_x(xr2, 1);   // nothing else is live here: everything in flight at once
  }
  region_free();
  if (dbg && etid == 0) dbg[4] = clock64();
  if constexpr (KIND != B_FIRST) *reinterpret_cast<float4*>(park + PK_GATE + etid * 16) = gate_v;
  if constexpr (KIND != B_LAST) {
    *reinterpret_cast<float4*>(park + PK_MUL + etid * 16) = make_float4(1.f + mul_v.x, 1.f + mul_v.y, 1.f + mul_v.z, 1.f + mul_v.w);
    *reinterpret_cast<float4*>(park + PK_ADD + etid * 16) = add_v;
  }
  if constexpr (HAS_BIAS) { if (etid < 64) *reinterpret_cast<float4*>(smem + S2_OFF_BIASP + etid * 16) = bias_v; }
  if constexpr (KIND != B_LAST) {
    if (b.next_bias_q != nullptr && etid < 64) reinterpret_cast<float4*>(b.sm_bias_q)[etid] = *reinterpret_cast<const float4*>(b.next_bias_q + etid * 4);
  }
  const uint32_t taddrA = tmem_base + ((q * 32u) << 16) + sub * 64;   // dead chunk accumulator: warp-private staging of x_old (step st at + 16 st)
  const uint32_t taddr = tmem_base + ((q * 32u) << 16) + 256 + sub * 16;   // c_proj accumulator, then the park of x_new: step st at + 64 st
  if constexpr (KIND != B_FIRST) {
    // the old rows wait in TMEM while the phase's last MMAs retire (this warp writes and reads the same lanes / columns)
    if constexpr (has_x) {
      tmem_st_32x32b_x32(taddrA, xr);
      tmem_st_32x32b_x32(taddrA + 32, xr2);
      tmem_st_wait();
    }
    if (dbg && etid == 0) dbg[5] = clock64();
    sm100::mbar_wait(sy.accB_full, sy.accB_parity);
    sm100::tc_fence_after();
  }
  if (dbg && etid == 0) dbg[0] = clock64();
  sm100::named_bar_sync(1, EPI_THREADS);   // park complete
  const uint32_t slot = row >> 4;          // slot of this lane's row within the tile
  float2* exch = reinterpret_cast<float2*>(park + PK_EXCH);
  // ---- pass 1: x_new = x_old + gate * (acc + bias) -> TMEM park, row-statistics partials ----
  // 16 columns at a time, the TMEM loads of the next step in flight while this one is processed (tcgen05.wait::ld covers
  // every outstanding load, so the next loads are issued right after the wait)
  auto stats16 = [&](const uint32_t (&w)[16], float& m, float& qq) {
    sm100::f32x2 s2 = sm100::pack2u(w[0], w[1]);
#pragma unroll
    for (int i = 1; i < 8; ++i) s2 = sm100::add2(s2, sm100::pack2u(w[2 * i], w[2 * i + 1]));
    float sa, sb;
    sm100::unpack2(s2, sa, sb);
    m = (sa + sb) * (1.0f / 16.0f);
    const sm100::f32x2 nm = sm100::pack2(-m, -m);
    sm100::f32x2 q2 = sm100::pack2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const sm100::f32x2 d = sm100::add2(sm100::pack2u(w[2 * i], w[2 * i + 1]), nm);
      q2 = sm100::fma2(d, d, q2);
    }
    sm100::unpack2(q2, sa, sb);
    qq = sa + sb;
  };
  if constexpr (KIND != B_FIRST) {
    uint32_t av[2][16], xv[2][16];
    sm100::tmem_ld_32x32b_x16(taddr, av[0]);
    if constexpr (has_x) sm100::tmem_ld_32x32b_x16(taddrA, xv[0]);
    float m_prev = 0.f, q_prev = 0.f;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      sm100::tmem_ld_wait();
      if (st < 3) {
        sm100::tmem_ld_32x32b_x16(taddr + (st + 1) * 64, av[(st + 1) & 1]);
        if constexpr (has_x) sm100::tmem_ld_32x32b_x16(taddrA + (st + 1) * 16, xv[(st + 1) & 1]);
      }
      uint32_t (&v)[16] = av[st & 1];
      const uint32_t (&xo)[16] = xv[st & 1];
      const int col0 = st * 64 + sub * 16;
      if constexpr (has_x) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 g = *reinterpret_cast<const float4*>(park + PK_GATE + (slot * D + col0 + i * 4) * 4);
        sm100::f32x2 a0 = sm100::pack2u(v[4 * i + 0], v[4 * i + 1]), a1 = sm100::pack2u(v[4 * i + 2], v[4 * i + 3]);
        if constexpr (HAS_BIAS) {
          const float4 bb = *reinterpret_cast<const float4*>(smem + S2_OFF_BIASP + (col0 + i * 4) * 4);
          a0 = sm100::add2(a0, sm100::pack2(bb.x, bb.y));
          a1 = sm100::add2(a1, sm100::pack2(bb.z, bb.w));
        }
        a0 = sm100::fma2(sm100::pack2(g.x, g.y), a0, sm100::pack2u(xo[4 * i + 0], xo[4 * i + 1]));
        a1 = sm100::fma2(sm100::pack2(g.z, g.w), a1, sm100::pack2u(xo[4 * i + 2], xo[4 * i + 3]));
        sm100::unpack2u(a0, v[4 * i + 0], v[4 * i + 1]);
        sm100::unpack2u(a1, v[4 * i + 2], v[4 * i + 3]);
      }
      tmem_st_32x32b_x16(taddr + st * 64, v);
      }
      if constexpr (KIND != B_LAST) {
        float m, qq;
        stats16(v, m, qq);
        if (st & 1) {   // two 16-column groups -> one 32-column partial (Chan et al., equal counts)
          const float dm = m - m_prev;
          exch[(sub * 2 + (st >> 1)) * BLOCK_M + row] = make_float2(0.5f * (m + m_prev), fmaf(8.0f * dm, dm, qq + q_prev));
        }
        m_prev = m; q_prev = qq;
      }
    }
  } else {
#pragma unroll
    for (int pair = 0; pair < 2; ++pair) {
      if (pair == 1) load_x(xr, 1);
      uint32_t lo[16], hi[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { lo[i] = xr[i]; hi[i] = xr[16 + i]; }
      tmem_st_32x32b_x16(taddr + (2 * pair) * 64, lo);
      tmem_st_32x32b_x16(taddr + (2 * pair + 1) * 64, hi);
      float ma, qa, mb, qb;
      stats16(lo, ma, qa);
      stats16(hi, mb, qb);
      const float dm = mb - ma;
      exch[(sub * 2 + pair) * BLOCK_M + row] = make_float2(0.5f * (ma + mb), fmaf(8.0f * dm, dm, qa + qb));
    }
  }
  tmem_st_wait();
  if constexpr (KIND != B_FIRST) {   // hand the parked rows to the drain warps (B_FIRST: they are in global memory already)
    sm100::tc_fence_before();
    __syncwarp();
    if (lane == 0 && sy.drain) sm100::mbar_arrive(sy.park_ready);
  }
  if (dbg && etid == 0) dbg[1] = clock64();
  if constexpr (KIND == B_LAST) return;
  sm100::named_bar_sync(2 + q, 128);       // the four column-quarter warps of this lane quadrant
  // ---- combine the 8 partials of this row (equal counts: Chan et al.) ----
  float mean, rstd;
  {
    float2 pr[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pr[k] = exch[k * BLOCK_M + row];
    float sm_ = 0.f, m2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sm_ += pr[k].x; m2 += pr[k].y; }
    mean = sm_ * 0.125f;
    float dev = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float d = pr[k].x - mean; dev = fmaf(d, d, dev); }
    m2 = fmaf(32.0f, dev, m2);
    rstd = rsqrtf(m2 * (1.0f / D) + b.eps);
  }
  if (dbg && etid == 0) dbg[2] = clock64();
  // ---- pass 2: A tile of the next phase = bf16( (x_new - mean) * rstd * (1 + mul) + add ), slab by slab ----
  const sm100::f32x2 rstd2 = sm100::pack2(rstd, rstd), nmr2 = sm100::pack2(-mean * rstd, -mean * rstd);   // (x - mean) * rstd = fma(x, rstd, nmr)
  uint32_t pv[4][16];
#pragma unroll
  for (int st = 0; st < 4; ++st) sm100::tmem_ld_32x32b_x16(taddr + st * 64, pv[st]);
  sm100::tmem_ld_wait();
  sm100::tc_fence_before();
#pragma unroll
  for (int st = 0; st < 4; ++st) {
    const uint32_t (&v)[16] = pv[st];
    const int col0 = st * 64 + sub * 16;
    uint8_t* a_slab = smem + st * A_SLAB_BYTES;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float h[8];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4 mu = *reinterpret_cast<const float4*>(park + PK_MUL + (slot * D + col0 + c * 8 + k * 4) * 4);
        const float4 ad = *reinterpret_cast<const float4*>(park + PK_ADD + (slot * D + col0 + c * 8 + k * 4) * 4);
        const sm100::f32x2 t0 = sm100::fma2(sm100::pack2u(v[c * 8 + 4 * k + 0], v[c * 8 + 4 * k + 1]), rstd2, nmr2);
        const sm100::f32x2 t1 = sm100::fma2(sm100::pack2u(v[c * 8 + 4 * k + 2], v[c * 8 + 4 * k + 3]), rstd2, nmr2);
        sm100::unpack2(sm100::fma2(t0, sm100::pack2(mu.x, mu.y), sm100::pack2(ad.x, ad.y)), h[4 * k + 0], h[4 * k + 1]);
        sm100::unpack2(sm100::fma2(t1, sm100::pack2(mu.z, mu.w), sm100::pack2(ad.z, ad.w)), h[4 * k + 2], h[4 * k + 3]);
      }
      uint4 o;
      o.x = sm100::pack_bf16x2(h[0], h[1]);
      o.y = sm100::pack_bf16x2(h[2], h[3]);
      o.z = sm100::pack_bf16x2(h[4], h[5]);
      o.w = sm100::pack_bf16x2(h[6], h[7]);
      *reinterpret_cast<uint4*>(a_slab + sm100::swz_chunk_offset(row, sub * 2 + c)) = o;
    }
    if (st & 1) {   // one proxy fence per two slabs: this warp's share of K slabs st - 1 and st is in place
      sm100::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) { sm100::mbar_arrive(sy.a_ready + st - 1); sm100::mbar_arrive(sy.a_ready + st); }
    }
  }
  if (dbg && etid == 0) dbg[3] = clock64();
}

template <bool SOLVE>
__global__ void __launch_bounds__(S2_THREADS, 1) dit_stack_kernel(const StackParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (sm100::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smA = smem;
  uint8_t* smMid = smem + S2_OFF_MID;
  uint8_t* smRing = smem + S2_OFF_RING;
  float* smBiasQ = reinterpret_cast<float*>(smem + S2_OFF_BIASQ);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S2_OFF_BARS);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + S2_OFF_TMEMPTR);
  int* smRows = reinterpret_cast<int*>(smem + S2_OFF_ROWS);
  uint64_t* full = bars + SB_FULL;
  uint64_t* empty = bars + SB_EMPTY;
  uint64_t* a_ready = bars + SB_A_READY;
  uint64_t* accA_full = bars + SB_ACCA_FULL;
  uint64_t* accA_free = bars + SB_ACCA_FREE;
  uint64_t* ao_ready = bars + SB_AO_READY;
  uint64_t* ao_free = bars + SB_AO_FREE;
  uint64_t* h_ready = bars + SB_H_READY;    // [2]
  uint64_t* h_free = bars + SB_H_FREE;      // [2]
  uint64_t* accB_full = bars + SB_ACCB_FULL;
  uint64_t* park_ready = bars + SB_PARK_READY;
  uint64_t* park_drained = bars + SB_PARK_DRAINED;
  uint64_t* xfull = bars + SB_XFULL;
  uint64_t* z_ready = bars + SB_Z_READY;

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) sm100::tmem_alloc(tmem_ptr_smem, 512);
  if (threadIdx.x == 0) {
    for (int i = 0; i < SB_COUNT; ++i) {
      const bool workers = (i >= SB_A_READY && i < SB_A_READY + 4) || i == SB_ACCA_FREE || i == SB_AO_READY || i == SB_H_READY || i == SB_H_READY + 1 || i == SB_PARK_READY;
      sm100::mbar_init(&bars[i], workers ? EPI_WARPS : ((i == SB_PARK_DRAINED || i == SB_Z_READY) ? 4 : 1));
    }
    sm100::fence_barrier_init();
  }
  sm100::grid_dep_launch();
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int T = p.n_chunks;
  const int hl = p.hid_last;   // width of the last hidden chunk
  // Before an attention half the boundary park lives in H buffer (T & 1) (the one the last MLP chunk does not use), before an
  // MLP half in the (dead) q/k/v staging.  The OTHER 32 KB of the mid region is free from the finishing phase's last MMA until
  // the starting phase's first chunk epilogue: the 4th weight item of the starting phase is prefetched there (the ring holds
  // three), so the first MMA group of a phase never waits for a weight slab.
  uint8_t* const park_pre_attn = smMid + (T & 1) * 2 * A_SLAB_BYTES;
  uint8_t* const x_attn = smMid + ((T & 1) ^ 1) * 2 * A_SLAB_BYTES;
  uint8_t* const park_pre_mlp = smMid;
  uint8_t* const x_mlp = smMid + 2 * A_SLAB_BYTES;
  // blocked interior storage of a tile: its own rows of X, or this CTA's scratch tile when X is row-major
  auto interior = [&](int tile) { return p.io_blocked ? p.X + (size_t)tile * BLOCK_M * D : p.scratch + (size_t)blockIdx.x * BLOCK_M * D; };
  constexpr bool solve = SOLVE;
  const int n_ev = solve ? p.n_evals : 1;
  constexpr uint32_t SOLVE_I_BYTES = B_SLAB_BYTES, SOLVE_F_BYTES = 4 * 16 * BLOCK_K * 2;

  if (warp < S2_DRAIN_WARP0) {
    reg_dec<S2_REGS_IDLE>();
    if (warp == 0 && lane == 0) {
      // ===================== producer: one continuous stream of weight items (independent of preceding kernels) ==============
      RingState rs;
      auto put = [&](const uint8_t* src, uint32_t bytes) {
        sm100::mbar_wait(&empty[rs.stage], rs.phase ^ 1);
        sm100::mbar_arrive_expect_tx(&full[rs.stage], bytes);
        sm100::bulk_g2s(smRing + rs.stage * S2_STAGE, src, bytes, &full[rs.stage]);
        rs.advance(S2_NST);
      };
      uint32_t n_phase = 0;
      auto put_x = [&](const uint8_t* src, uint32_t bytes, uint8_t* dst) {   // the extra slot is free once the previous phase's MMAs have retired
        if (n_phase > 0) sm100::mbar_wait(accB_full, (n_phase - 1) & 1);
        sm100::mbar_arrive_expect_tx(xfull, bytes);
        sm100::bulk_g2s(dst, src, bytes, xfull);
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int e = 0; e < n_ev; ++e) {
          if (solve) { put(reinterpret_cast<const uint8_t*>(p.w_solve), SOLVE_I_BYTES); ++n_phase; }
          for (int l = 0; l < p.n_layer; ++l) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.w_attn + (size_t)l * p.attn_w_stride);
            for (int step = 0; step <= AB_HP; ++step) {
              if (step < AB_HP)
                for (int ks = 0; ks < KSLABS_D; ++ks) {
                  if (step == 0 && ks == KSLABS_D - 1) put_x(src, AB_Q_ITEM_BYTES, x_attn);
                  else put(src, AB_Q_ITEM_BYTES);
                  src += AB_Q_ITEM_BYTES;
                }
              if (step >= 1) { put(src, 2 * AB_P_ITEM_BYTES); src += 2 * AB_P_ITEM_BYTES; }
            }
            ++n_phase;
            src = reinterpret_cast<const uint8_t*>(p.w_mlp + (size_t)l * p.mlp_w_stride);
            for (int j = 0; j <= T; ++j) {   // M1_0 M1_1 M2_0 M1_2 M2_1 ...
              if (j < T) {
                const uint32_t bytes = j + 1 < T ? (uint32_t)B_SLAB_BYTES : (uint32_t)(2 * hl * BLOCK_K * 2);   // [w1 | w2] rows of the chunk, one K slab
                for (int ks = 0; ks < KSLABS_D; ++ks) {
                  if (j == 0 && ks == KSLABS_D - 1) put_x(src, bytes, x_mlp);
                  else put(src, bytes);
                  src += bytes;
                }
              }
              if (j >= 1)
                for (int s2 = 0, ns = min(2, p.hid_slabs - 2 * (j - 1)); s2 < ns; ++s2) { put(src, B_SLAB_BYTES); src += B_SLAB_BYTES; }
            }
            ++n_phase;
          }
          if (solve) put(reinterpret_cast<const uint8_t*>(p.w_solve) + SOLVE_I_BYTES, SOLVE_F_BYTES);
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer ========================================================================================
      const uint32_t idesc_q = sm100::make_idesc_bf16(BLOCK_M, AB_QN);
      const uint32_t idesc_n = sm100::make_idesc_bf16(BLOCK_M, BLOCK_N);
      const uint32_t idesc_l = sm100::make_idesc_bf16(BLOCK_M, 2 * hl);   // last hidden chunk
      const uint32_t accA = tmem_base, accB = tmem_base + 256;
      const uint32_t a_base = sm100::smem_u32(smA), ao_base = sm100::smem_u32(smMid + 3 * A_SLAB_BYTES), h_base = sm100::smem_u32(smMid);
      RingState rs;
      uint32_t n_ar = 0, u_accA = 0, n_aor = 0, n_hr0 = 0, n_hr1 = 0, n_dr = 0, n_xu = 0, n_z = 0;
      const uint32_t idesc_f = sm100::make_idesc_bf16(BLOCK_M, 16);
      auto wait_stage = [&]() -> uint32_t {
        sm100::mbar_wait(&full[rs.stage], rs.phase);
        sm100::tc_fence_after();
        return sm100::smem_u32(smRing + rs.stage * S2_STAGE);
      };
      auto done_stage = [&]() {
        sm100::umma_commit(&empty[rs.stage]);
        rs.advance(S2_NST);
      };
      // the c_proj accumulator's columns hold the parked x_new of the preceding boundary until the drain warps are done with it
      auto wait_drained = [&]() {
        if (n_dr > 0) { sm100::mbar_wait(park_drained, (n_dr - 1) & 1); sm100::tc_fence_after(); }
      };
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
       for (int e = 0; e < n_ev; ++e) {
        if (solve) {
          // ---- input projection + pos_embed + bias: accB[128 x 256] = [z_hi | z_lo | 1hot(tok) | 1hot(tok)] x [Win | Win | posb_hi | posb_lo]^T ----
          sm100::mbar_wait(z_ready, n_z & 1); ++n_z;
          sm100::tc_fence_after();
          wait_drained();
          const uint32_t bs = wait_stage();
          issue_slab_mmas(accB, a_base, bs, idesc_n, true);
          done_stage();
          sm100::umma_commit(accB_full);
          ++n_dr;   // the boundary that follows is drained
        }
        for (int l = 0; l < p.n_layer; ++l) {
          // ---- attention half: Q0 Q1 P0 Q2 P1 Q3 P2 P3 (Q0 follows the A tile slab by slab) ----
          for (int step = 0; step <= AB_HP; ++step) {
            if (step < AB_HP) {
              if (step > 0) { sm100::mbar_wait(accA_free, (u_accA - 1) & 1); sm100::tc_fence_after(); }
              for (int ks = 0; ks < KSLABS_D; ++ks) {
                if (step == 0) { sm100::mbar_wait(a_ready + ks, n_ar & 1); sm100::tc_fence_after(); }
                if (step == 0 && ks == KSLABS_D - 1) {   // prefetched into the extra slot
                  sm100::mbar_wait(xfull, n_xu & 1); ++n_xu;
                  sm100::tc_fence_after();
                  issue_slab_mmas(accA, a_base + ks * A_SLAB_BYTES, sm100::smem_u32(x_attn), idesc_q, false);
                  continue;
                }
                const uint32_t bs = wait_stage();
                issue_slab_mmas(accA, a_base + ks * A_SLAB_BYTES, bs, idesc_q, ks == 0);
                done_stage();
              }
              sm100::umma_commit(accA_full); ++u_accA;
            }
            if (step >= 1) {
              sm100::mbar_wait(ao_ready, n_aor & 1); ++n_aor;
              sm100::tc_fence_after();
              if (step == 1) wait_drained();
              const uint32_t bs = wait_stage();
              issue_slab_mmas(accB, ao_base, bs, idesc_n, step == 1);
              done_stage();
              sm100::umma_commit(ao_free);
            }
          }
          sm100::umma_commit(accB_full);
          ++n_dr; ++n_ar;   // the attention -> MLP boundary
          // ---- MLP half: M1_0 M1_1 M2_0 M1_2 M2_1 ... ----
          for (int j = 0; j <= T; ++j) {
            if (j < T) {
              if (j > 0) { sm100::mbar_wait(accA_free, (u_accA - 1) & 1); sm100::tc_fence_after(); }
              for (int ks = 0; ks < KSLABS_D; ++ks) {
                if (j == 0) { sm100::mbar_wait(a_ready + ks, n_ar & 1); sm100::tc_fence_after(); }
                if (j == 0 && ks == KSLABS_D - 1) {      // prefetched into the extra slot
                  sm100::mbar_wait(xfull, n_xu & 1); ++n_xu;
                  sm100::tc_fence_after();
                  issue_slab_mmas(accA, a_base + ks * A_SLAB_BYTES, sm100::smem_u32(x_mlp), T == 1 ? idesc_l : idesc_n, false);
                  continue;
                }
                const uint32_t bs = wait_stage();
                issue_slab_mmas(accA, a_base + ks * A_SLAB_BYTES, bs, j + 1 < T ? idesc_n : idesc_l, ks == 0);
                done_stage();
              }
              sm100::umma_commit(accA_full); ++u_accA;
            }
            if (j >= 1) {
              const int c = j - 1, hb = c & 1;
              if (hb) { sm100::mbar_wait(&h_ready[1], n_hr1 & 1); ++n_hr1; }
              else { sm100::mbar_wait(&h_ready[0], n_hr0 & 1); ++n_hr0; }
              sm100::tc_fence_after();
              if (c == 0) wait_drained();
              const int ns = min(2, p.hid_slabs - 2 * c);
              for (int s2 = 0; s2 < ns; ++s2) {
                const uint32_t bs = wait_stage();
                issue_slab_mmas(accB, h_base + (hb * 2 + s2) * A_SLAB_BYTES, bs, idesc_n, c == 0 && s2 == 0);
                done_stage();
              }
              sm100::umma_commit(&h_free[hb]);
            }
          }
          sm100::umma_commit(accB_full);
          if (!(solve && l + 1 == p.n_layer)) ++n_dr;   // the MLP -> next attention (or end of tile) boundary; not drained before the final layer
          ++n_ar;
        }
        if (solve) {
          // ---- final linear: accA[128 x 16] = LN_mod(x) x Wout^T ----
          const uint32_t bs = wait_stage();
          for (int ks = 0; ks < KSLABS_D; ++ks) {
            sm100::mbar_wait(a_ready + ks, n_ar & 1);
            sm100::tc_fence_after();
            issue_slab_mmas(accA, a_base + ks * A_SLAB_BYTES, bs + ks * (16 * BLOCK_K * 2), idesc_f, ks == 0);
          }
          ++n_ar;
          done_stage();
          sm100::umma_commit(accA_full); ++u_accA;
        }
       }
      }
    }
  } else if (warp < S2_WORKER_WARP0) {
    // ===================== drain warps: parked x_new (TMEM) -> global memory ==================================================
    reg_dec<S2_REGS_DRAIN>();
    const uint32_t dq = warp & 3, row = dq * 32 + lane;
    const uint32_t taddr = tmem_base + 256 + ((dq * 32u) << 16);
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
      float* const x_in = interior(tile) + row * 4;
      float* const x_io = p.X + (size_t)tile * BLOCK_M * D + (p.io_blocked ? row * 4 : row * D);
      const int io_cs = p.io_blocked ? BLOCK_M * 4 : 4;
      for (int e = 0; e < n_ev; ++e) {
      for (int bnd = 0; bnd < 2 * p.n_layer; ++bnd) {
        // plain mode: boundary bnd follows phase bnd (the last one writes the io buffer); solve mode: boundary 0 follows the input
        // projection, boundary bnd >= 1 phase bnd - 1 (the one before the final layer is not drained): everything stays interior
        const bool last = !solve && bnd == 2 * p.n_layer - 1;
        float* dst = last ? x_io : x_in;
        const int cs = last ? io_cs : BLOCK_M * 4;
        sm100::mbar_wait(park_ready, n & 1); ++n;
        sm100::tc_fence_after();
#pragma unroll 1
        for (int c16 = 0; c16 < 16; ++c16) {
          uint32_t v[16];
          sm100::tmem_ld_32x32b_x16(taddr + c16 * 16, v);
          sm100::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(dst + (c16 * 4 + i) * cs) =
                make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
        sm100::tc_fence_before();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(park_drained);
        if (p.dbg != nullptr && warp == S2_DRAIN_WARP0 && lane == 0 && !solve && (bnd >> 1) == p.dbg_layer && tile == (int)(blockIdx.x + gridDim.x))
          p.dbg[(size_t)blockIdx.x * 128 + ((bnd & 1) ? 47 : 27)] = clock64();
      }
      }
    }
  } else {
    // ===================== 16 worker warps =====================================================================================
    reg_inc<S2_REGS_WORKER>();
    const uint32_t ew = warp - S2_WORKER_WARP0, q = warp & 3, sub = ew >> 2, etid = threadIdx.x - S2_WORKER_WARP0 * 32;
    const uint32_t row = q * 32 + lane;
    const uint32_t taddr_q = tmem_base + ((q * 32u) << 16);
    uint8_t* smQKV = smMid;
    uint8_t* smAO = smMid + 3 * A_SLAB_BYTES;
    const uint32_t qkv_base = sm100::smem_u32(smQKV);
    // attention job of this warp within a head pair: a slot of its own TMEM lane quadrant and one head, so that the q/k/v rows
    // it stages (bf16, warp-private 3 KB block: [part][token][64 B], 16-byte chunks XOR-swizzled with (token / 2) % 4) are
    // exactly the ones its own mma.sync job reads: no barrier between staging and the attention core
    const uint32_t job_slot = 2 * q + (sub >> 1), job_h = sub & 1;
    const bool job_lane = (lane >> 4) == (sub >> 1);        // lanes whose TMEM rows belong to the job's slot
    const uint32_t job_tok = lane & 15;
    uint8_t* const stg = smQKV + ew * 3072;
    const uint32_t stg_base = sm100::smem_u32(stg);
    const uint32_t g = lane >> 2, t4 = lane & 3;
    const int hs = sub >> 1, hh = sub & 1;                  // SwiGLU: slab hs of the chunk, 32-column half hh
    uint32_t n_accA = 0, n_accB = 0, n_ao = 0, n_h0 = 0, n_h1 = 0, n_dr = 0;
    BoundarySync sy{};
    sy.accB_full = accB_full; sy.a_ready = a_ready; sy.park_ready = park_ready; sy.park_drained = park_drained;
    sm100::grid_dep_wait();   // X and the modulation table come from the preceding kernels
    if (p.dbg != nullptr && etid == 0) {
      unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      p.dbg[(size_t)blockIdx.x * 128 + 60] = clock64(); p.dbg[(size_t)blockIdx.x * 128 + 62] = (long long)gt;
    }
    int tile_no = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_no) {
      if (etid < 8) {
        const int s = tile * 8 + (int)etid;
        smRows[etid] = (solve && s >= p.n_u) ? (s < p.n_u + p.slot_shift ? 0 : p.slot_mod.row(s - p.slot_shift)) : p.slot_mod.row(s);
      }
      sm100::named_bar_sync(1, EPI_THREADS);   // rows resolved; every warp has left the previous tile
      if (p.dbg != nullptr && etid == 0 && tile_no < 4) p.dbg[(size_t)blockIdx.x * 128 + 52 + tile_no] = clock64();
      BoundaryArgs ba{};
      // lane-resolved addresses of (row, column 64 sub) in the io buffer and in the blocked interior storage
      const float* const x_io = p.X + (size_t)tile * BLOCK_M * D + (p.io_blocked ? (size_t)row * 4 : (size_t)row * D);   // column 0 of this lane's row
      const int io_cs = p.io_blocked ? BLOCK_M * 4 : 4;
      const float* const x_in = interior(tile) + (size_t)row * 4;
      ba.mod = p.mod; ba.rows = smRows; ba.mod_stride = p.mod_stride; ba.eps = p.eps;
      ba.sm_bias_q = smBiasQ;
      // ---- whole-solve mode: state <-> slot mapping of this lane's row (StepParams), recomputed where needed ----
      struct RowState { int state, k; bool valid, pair; };
      auto row_state = [&]() {
        RowState r{0, 0, false, false};
        const int s_g = tile * 8 + (int)(row >> 4);
        const int g0 = p.n_u + p.slot_shift;   // first guided slot
        r.valid = s_g < g0 + p.n_g * p.n_f && !(s_g >= p.n_u && s_g < g0);
        if (s_g < g0) r.state = s_g < p.n_u ? s_g : 0;
        else { const int j = (s_g - g0) / p.n_f; r.k = (s_g - g0) - j * p.n_f; r.state = p.n_u + j; r.pair = p.n_f == 2; }
        return r;
      };
      // the state of this row as the A operand of the input projection: 128-byte row = [hi 16 bf16 | lo 16 bf16 | unused]
      auto write_z = [&](const float (&z)[16]) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat16 h0 = __float2bfloat16(z[2 * i]), h1 = __float2bfloat16(z[2 * i + 1]);
          hi[i] = sm100::pack_bf16x2(z[2 * i], z[2 * i + 1]);
          lo[i] = sm100::pack_bf16x2(z[2 * i] - __bfloat162float(h0), z[2 * i + 1] - __bfloat162float(h1));
        }
        *reinterpret_cast<uint4*>(smA + sm100::swz_chunk_offset(row, 0)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(smA + sm100::swz_chunk_offset(row, 1)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        *reinterpret_cast<uint4*>(smA + sm100::swz_chunk_offset(row, 2)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        *reinterpret_cast<uint4*>(smA + sm100::swz_chunk_offset(row, 3)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
        // one-hot(token) twice (bf16 1.0 = 0x3F80): picks pos_embed + bias (hi and lo parts) out of the weight slab
        const int tok = (int)(row & 15);
        const uint32_t one = 0x3F80u << (16 * (tok & 1));
        const int w = (tok & 7) >> 1;
        const uint4 oh = make_uint4(w == 0 ? one : 0u, w == 1 ? one : 0u, w == 2 ? one : 0u, w == 3 ? one : 0u), zero4 = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(smA + sm100::swz_chunk_offset(row, 4)) = tok < 8 ? oh : zero4;
        *reinterpret_cast<uint4*>(smA + sm100::swz_chunk_offset(row, 5)) = tok < 8 ? zero4 : oh;
        *reinterpret_cast<uint4*>(smA + sm100::swz_chunk_offset(row, 6)) = tok < 8 ? oh : zero4;
        *reinterpret_cast<uint4*>(smA + sm100::swz_chunk_offset(row, 7)) = tok < 8 ? zero4 : oh;
        sm100::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) sm100::mbar_arrive(z_ready);
      };
      if constexpr (solve) {
        if (sub == 0) {
          const RowState rs0 = row_state();
          float z[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 t = rs0.valid ? *reinterpret_cast<const float4*>(p.x_base + ((size_t)rs0.state * TOK + (row & 15)) * LAT + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            z[4 * i] = t.x; z[4 * i + 1] = t.y; z[4 * i + 2] = t.z; z[4 * i + 3] = t.w;
          }
          write_z(z);
        }
      } else {
        ba.xin = x_io; ba.in_cs = io_cs;
        ba.off_mul = 0; ba.off_add = D; ba.next_bias_q = p.bias_q;
        sy.n_drained = n_dr; sy.drain = true;
        boundary_step<B_FIRST, false, true>(ba, sy, smem, park_pre_attn, tmem_base, q, sub, lane, etid, [] {}, nullptr);
      }
     for (int e = 0; e < n_ev; ++e) {
      if constexpr (solve) {
        // ---- boundary after the input projection: x = (pos_embed + bias) + acc, LN1 / modulate of layer 0 ----
        ba.mod = p.mod + (size_t)e * p.mod_eval_stride;
        ba.xin = nullptr; ba.in_cs = 0;
        ba.off_gate = -1; ba.bias = nullptr; ba.off_mul = 0; ba.off_add = D; ba.next_bias_q = p.bias_q;
        sy.accB_parity = n_accB & 1; sy.n_drained = n_dr; sy.drain = true;
        boundary_step<B_MID, false, false>(ba, sy, smem, park_pre_attn, tmem_base, q, sub, lane, etid,
                                    [] { sm100::named_bar_sync(1, EPI_THREADS); /* every warp is past the previous evaluation's final step */ },
                                    (p.dbg != nullptr && tile_no == 1 && e > 0) ? p.dbg + (size_t)blockIdx.x * 128 + 64 : nullptr);
        ++n_accB; ++n_dr;
        ba.xin = x_in; ba.in_cs = BLOCK_M * 4;
        if (p.dbg != nullptr && tile_no == 1 && etid == 0) p.dbg[(size_t)blockIdx.x * 128 + 51] = clock64();
      }
      for (int l = 0; l < p.n_layer; ++l) {
        long long* dbg = (p.dbg != nullptr && l == p.dbg_layer && tile_no == 1) ? p.dbg + (size_t)blockIdx.x * 128 : nullptr;
        const int mo = l * 6 * D;
        // ================= attention half =================
        if (dbg && etid == 0) dbg[4] = clock64();
        for (int hp = 0; hp < AB_HP; ++hp) {
          sm100::mbar_wait(accA_full, n_accA & 1); ++n_accA;
          sm100::tc_fence_after();
          if (dbg && etid == 0) dbg[8 + 3 * hp] = clock64();
          __syncwarp();                                     // the previous job's ldmatrix reads of the staging block are done
          // q | k | v of this head: 32 accumulator columns each (q at 32 h, k at 64 + 32 h, v at 128 + 32 h); the next part's
          // TMEM load is in flight while this one is converted
          uint32_t pv[2][32];
          sm100::tmem_ld_32x32b_x32(taddr_q + job_h * 32, pv[0]);
#pragma unroll
          for (int part = 0; part < 3; ++part) {
            sm100::tmem_ld_wait();
            if (part < 2) sm100::tmem_ld_32x32b_x32(taddr_q + (part + 1) * 64 + job_h * 32, pv[(part + 1) & 1]);
            const uint32_t (&vv)[32] = pv[part & 1];
            if (job_lane) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                if (part == 0) {                             // q columns carry a bias
                  b0 = *reinterpret_cast<const float4*>(smBiasQ + hp * 64 + job_h * 32 + c * 8);
                  b1 = *reinterpret_cast<const float4*>(smBiasQ + hp * 64 + job_h * 32 + c * 8 + 4);
                }
                uint4 o;
                o.x = sm100::pack_bf16x2(__uint_as_float(vv[c * 8 + 0]) + b0.x, __uint_as_float(vv[c * 8 + 1]) + b0.y);
                o.y = sm100::pack_bf16x2(__uint_as_float(vv[c * 8 + 2]) + b0.z, __uint_as_float(vv[c * 8 + 3]) + b0.w);
                o.z = sm100::pack_bf16x2(__uint_as_float(vv[c * 8 + 4]) + b1.x, __uint_as_float(vv[c * 8 + 5]) + b1.y);
                o.w = sm100::pack_bf16x2(__uint_as_float(vv[c * 8 + 6]) + b1.z, __uint_as_float(vv[c * 8 + 7]) + b1.w);
                *reinterpret_cast<uint4*>(stg + part * 1024 + job_tok * 64 + ((c ^ ((job_tok >> 1) & 3)) << 4)) = o;
              }
            }
            if (part == 1) {                                 // every load of the accumulator has completed
              sm100::tmem_ld_wait();
              sm100::tc_fence_before();
              __syncwarp();
              if (lane == 0) sm100::mbar_arrive(accA_free);
            }
          }
          __syncwarp();
          if (dbg && etid == 0) dbg[9 + 3 * hp] = clock64();
          auto chunk_addr = [&](uint32_t token, uint32_t part, uint32_t dim) -> uint32_t {
            return stg_base + part * 1024 + token * 64 + (((dim >> 3) ^ ((token >> 1) & 3)) << 4);
          };
          float o[4][4];
          attn16_core(chunk_addr, lane, o);
          if (n_ao > 0) sm100::mbar_wait(ao_free, (n_ao - 1) & 1);   // the previous c_proj MMAs finished reading the AO slab
          ++n_ao;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const uint32_t col = job_h * HD + nt * 8 + 2 * t4;
            uint8_t* dst = smAO + (col & 7) * 2;
            *reinterpret_cast<uint32_t*>(dst + sm100::swz_chunk_offset(job_slot * TOK + g, col >> 3)) = sm100::pack_bf16x2(o[nt][0], o[nt][1]);
            *reinterpret_cast<uint32_t*>(dst + sm100::swz_chunk_offset(job_slot * TOK + g + 8, col >> 3)) = sm100::pack_bf16x2(o[nt][2], o[nt][3]);
          }
          sm100::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) sm100::mbar_arrive(ao_ready);
          if (dbg && etid == 0) dbg[10 + 3 * hp] = clock64();
        }
        // ---- boundary: attention residual + LN2 / modulate ----
        ba.off_gate = mo + 2 * D; ba.bias = p.bias_proj + (size_t)l * D;
        ba.off_mul = mo + 3 * D; ba.off_add = mo + 4 * D; ba.next_bias_q = nullptr;
        sy.accB_parity = n_accB & 1; sy.n_drained = n_dr; sy.drain = true;
        boundary_step<B_MID, true, true>(ba, sy, smem, park_pre_mlp, tmem_base, q, sub, lane, etid,
                                   [] { sm100::named_bar_sync(1, EPI_THREADS); /* every attention job has read its q/k/v */ },
                                   dbg ? dbg + 20 : nullptr);
        ++n_accB; ++n_dr;
        ba.xin = x_in; ba.in_cs = BLOCK_M * 4;
        // ================= MLP half =================
        for (int j = 0; j < T; ++j) {
          sm100::mbar_wait(accA_full, n_accA & 1); ++n_accA;
          sm100::tc_fence_after();
          if (dbg && etid == 0 && j < 6) dbg[28 + 2 * j] = clock64();
          // the chunk's accumulator: columns [0, w) = w1 part, [w, 2 w) = w2 part (w = 128, the last chunk hid_last); this warp
          // handles hidden units [32 sub, 32 sub + 32) of the chunk when the chunk has them
          const int cw = j + 1 < T ? 128 : hl;
          const bool active = (int)(sub * 32) < cw;
          uint32_t va[32], vb[32];
          sm100::tmem_ld_32x32b_x32(taddr_q + sub * 32, va);        // (an inactive warp reads columns it does not use)
          sm100::tmem_ld_32x32b_x32(taddr_q + cw + sub * 32, vb);
          sm100::tmem_ld_wait();
          sm100::tc_fence_before();
          __syncwarp();
          if (lane == 0) sm100::mbar_arrive(accA_free);
          const int hb = j & 1;
          const uint32_t cnt = hb ? n_h1 : n_h0;
          if (cnt > 0) sm100::mbar_wait(&h_free[hb], (cnt - 1) & 1);   // the MMAs that read this H buffer last have retired
          if (hb) ++n_h1; else ++n_h0;
          uint8_t* buf = smMid + (hb * 2 + hs) * A_SLAB_BYTES;
          if (active) {
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
              *reinterpret_cast<uint4*>(buf + sm100::swz_chunk_offset(row, hh * 4 + c)) = o;
            }
          } else if (hs < min(2, p.hid_slabs - 2 * j)) {   // beyond the chunk's width but inside a slab the c_proj MMAs read (zero weights there)
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(buf + sm100::swz_chunk_offset(row, hh * 4 + c)) = make_uint4(0u, 0u, 0u, 0u);
          }
          sm100::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) sm100::mbar_arrive(&h_ready[hb]);
          if (dbg && etid == 0 && j < 6) dbg[29 + 2 * j] = clock64();
        }
        // ---- boundary: MLP residual (+ LN1 / modulate of the next layer).  The park goes into the H buffer that the last
        //      chunk does not use; it is free once the MMAs of chunk T - 2 have retired. ----
        const int pb = T & 1;
        uint8_t* park = park_pre_attn;
        const uint32_t pcnt = pb ? n_h1 : n_h0;
        auto h_region_free = [&] {
          if (pcnt > 0) sm100::mbar_wait(&h_free[pb], (pcnt - 1) & 1);
          sm100::named_bar_sync(1, EPI_THREADS);   // every warp has read the last chunk's accumulator
        };
        ba.off_gate = mo + 5 * D; ba.bias = nullptr;
        ba.off_mul = mo + 6 * D; ba.off_add = mo + 7 * D; ba.next_bias_q = p.bias_q + (size_t)(l + 1) * 3 * D;
        sy.accB_parity = n_accB & 1; sy.n_drained = n_dr; sy.drain = true;
        if (l + 1 < p.n_layer) {
          boundary_step<B_MID, false, true>(ba, sy, smem, park, tmem_base, q, sub, lane, etid, h_region_free, dbg ? dbg + 40 : nullptr);
          ++n_dr;
        } else if constexpr (!solve) {
          boundary_step<B_LAST, false, true>(ba, sy, smem, park, tmem_base, q, sub, lane, etid, h_region_free, dbg ? dbg + 40 : nullptr);
          ++n_dr;
        } else {
          // final layer (layers.py:397-401): LN(x) * (1 + scale) + shift -> A tile of the final Linear; the rows themselves are dead
          ba.off_mul = p.mod_off_final + D; ba.off_add = p.mod_off_final; ba.next_bias_q = nullptr;
          sy.drain = false;
          boundary_step<B_MID, false, true>(ba, sy, smem, park, tmem_base, q, sub, lane, etid, h_region_free, dbg ? dbg + 40 : nullptr);
        }
        ++n_accB;
        if (dbg && etid == 0) dbg[5] = clock64();
      }
      if constexpr (solve) {
        // ---- v = final Linear (+ bias), CFG combine, Runge-Kutta stage update, next evaluation point (column quarter 0 warps) ----
        long long* dbs = (p.dbg != nullptr && tile_no == 1 && etid == 0) ? p.dbg + (size_t)blockIdx.x * 128 : nullptr;
        if (dbs) dbs[48] = clock64();
        sm100::mbar_wait(accA_full, n_accA & 1); ++n_accA;
        sm100::tc_fence_after();
        if (dbs) dbs[49] = clock64();
        if (sub == 0) {
          uint32_t o16[16];
          sm100::tmem_ld_32x32b_x16(taddr_q, o16);
          sm100::tmem_ld_wait();
          sm100::tc_fence_before();
          __syncwarp();
          if (lane == 0) sm100::mbar_arrive(accA_free);
          const RowState rs1 = row_state();
          const int st_state = rs1.state, st_k = rs1.k;
          const bool st_valid = rs1.valid, st_pair = rs1.pair;
          float* xb = p.x_base + ((size_t)st_state * TOK + (row & 15)) * LAT;
          float* ac = p.acc + ((size_t)st_state * TOK + (row & 15)) * LAT;
          const float4 stg4 = p.stage[e];   // {a_dt, b_dt, first_stage, last_stage}
          float4 xb4[4], ac4[4];
          float v[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bo = *reinterpret_cast<const float4*>(p.b_out + 4 * i);
            xb4[i] = st_valid ? *reinterpret_cast<const float4*>(xb + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            ac4[i] = (st_valid && stg4.z == 0.f) ? *reinterpret_cast<const float4*>(ac + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[4 * i + 0] = __uint_as_float(o16[4 * i + 0]) + bo.x; v[4 * i + 1] = __uint_as_float(o16[4 * i + 1]) + bo.y;
            v[4 * i + 2] = __uint_as_float(o16[4 * i + 2]) + bo.z; v[4 * i + 3] = __uint_as_float(o16[4 * i + 3]) + bo.w;
          }
          // a guided state owns two consecutive slots = lanes l and l ^ 16 of this warp: v = c0 v_0 + c1 v_1
          const float c_mine = st_pair ? p.coef[st_k] : 1.f, c_other = st_pair ? p.coef[st_k ^ 1] : 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float other = __shfl_xor_sync(0xffffffffu, v[i], 16);
            v[i] = fmaf(c_other, other, c_mine * v[i]);
          }
          float z[16];
          __syncwarp();   // both lanes of a pair have read the state before one of them updates it
          const bool writer = st_valid && st_k == 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 a = ac4[i];
            a.x = fmaf(stg4.y, v[4 * i + 0], a.x); a.y = fmaf(stg4.y, v[4 * i + 1], a.y);
            a.z = fmaf(stg4.y, v[4 * i + 2], a.z); a.w = fmaf(stg4.y, v[4 * i + 3], a.w);
            float4 xe;
            if (stg4.w != 0.f) {
              xe = make_float4(xb4[i].x + a.x, xb4[i].y + a.y, xb4[i].z + a.z, xb4[i].w + a.w);
              if (writer) *reinterpret_cast<float4*>(xb + 4 * i) = xe;
            } else {
              if (writer) *reinterpret_cast<float4*>(ac + 4 * i) = a;
              xe = make_float4(fmaf(stg4.x, v[4 * i + 0], xb4[i].x), fmaf(stg4.x, v[4 * i + 1], xb4[i].y),
                               fmaf(stg4.x, v[4 * i + 2], xb4[i].z), fmaf(stg4.x, v[4 * i + 3], xb4[i].w));
            }
            z[4 * i] = xe.x; z[4 * i + 1] = xe.y; z[4 * i + 2] = xe.z; z[4 * i + 3] = xe.w;
          }
          if (e + 1 < n_ev) write_z(z);   // the final Linear's MMAs have retired: the A tile is dead
          if (dbs) dbs[50] = clock64();
        } else {
          __syncwarp();
          if (lane == 0) sm100::mbar_arrive(accA_free);
        }
      }
     }
    }
    if (p.dbg != nullptr && etid == 0) {
      unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      p.dbg[(size_t)blockIdx.x * 128 + 61] = clock64(); p.dbg[(size_t)blockIdx.x * 128 + 63] = (long long)gt;
    }
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) sm100::tmem_dealloc(tmem_base, 512);
}

}  // namespace dit
