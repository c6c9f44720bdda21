// dit_stack_kernel: the adaLN block stack of the DiT denoiser (reference layers.py:208-221, nnets.py:273-297) as one
// persistent kernel, second generation.
//
// What changed against dit_blocks_kernel (dit_kernels.cuh), and why:
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
  const float* bias_q;      // [L][768] c_attn.bias (only the q part is used, see AttnBlockParams)
  const float* bias_proj;   // [L][256] fused c_proj bias
  int n_layer, n_tiles;
  int n_chunks, hid_slabs;  // MLP: ceil(hidden / 128), ceil(hidden / 64)
  long long* dbg;           // optional timeline of (second tile, layer dbg_layer): 64 stamps per CTA
  int dbg_layer;
};

constexpr int S2_NST = 3;
constexpr int S2_STAGE = 32768;
constexpr int S2_OFF_MID = KSLABS_D * A_SLAB_BYTES;             // 64 KB: q|k|v + AO slabs / H buffers / boundary park
constexpr int S2_OFF_RING = S2_OFF_MID + 4 * A_SLAB_BYTES;      // 128 KB
constexpr int S2_OFF_BIASQ = S2_OFF_RING + S2_NST * S2_STAGE;   // 224 KB
constexpr int S2_OFF_BIASP = S2_OFF_BIASQ + D * 4;             // c_proj bias of the attention half (boundary pass 1)
constexpr int S2_OFF_BARS = S2_OFF_BIASP + D * 4;
enum { SB_FULL = 0, SB_EMPTY = 3, SB_A_READY = 6, SB_ACCA_FULL = 7, SB_ACCA_FREE = 8, SB_AO_READY = 9, SB_AO_FREE = 10,
       SB_H_READY = 11, SB_H_FREE = 13, SB_ACCB_FULL = 15, SB_PARK_READY = 16, SB_PARK_DRAINED = 17, SB_XFULL = 18, SB_COUNT = 19 };
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
  const float* xin;      // this lane's row of x_old at column 64 sub; consecutive 4-column groups are in_cs floats apart
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
};

// One boundary on the calling worker warp.  `region_free()` returns once every worker warp has finished the phase's last chunk
// (the chunk accumulator's TMEM columns and the park region are dead); it is called after the global loads have been issued.
template <int KIND, bool HAS_BIAS, typename RegionFree>
__device__ __forceinline__ void boundary_step(const BoundaryArgs& b, const BoundarySync& sy, uint8_t* smem, uint8_t* park, uint32_t tmem_base,
                                              uint32_t q, uint32_t sub, uint32_t lane, uint32_t etid, RegionFree&& region_free, long long* dbg) {
  const uint32_t row = q * 32 + lane;
  // ---- issue every global load first: this thread's share of the per-slot vectors, then the first half of the old rows ----
  const int pslot = etid >> 6, pc4 = etid & 63;
  const float* mrow = b.mod + (size_t)b.rows[pslot] * b.mod_stride + pc4 * 4;
  float4 gate_v = make_float4(0.f, 0.f, 0.f, 0.f), mul_v = gate_v, add_v = gate_v, bias_v = gate_v;
  if constexpr (KIND != B_FIRST) gate_v = *reinterpret_cast<const float4*>(mrow + b.off_gate);
  if constexpr (KIND != B_LAST) {
    mul_v = *reinterpret_cast<const float4*>(mrow + b.off_mul);
    add_v = *reinterpret_cast<const float4*>(mrow + b.off_add);
  }
  if constexpr (HAS_BIAS) { if (etid < 64) bias_v = *reinterpret_cast<const float4*>(b.bias + etid * 4); }
  // the drain warps have written every earlier x_new (the rows read below) and are done with the TMEM park
  if (sy.n_drained > 0) sm100::mbar_wait(sy.park_drained, (sy.n_drained - 1) & 1);
  uint32_t xr[32], xr2[32];
  auto load_x = [&](uint32_t (&dst)[32], int half) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 t = __ldcg(reinterpret_cast<const float4*>(b.xin + (half * 8 + i) * b.in_cs));
      dst[4 * i + 0] = __float_as_uint(t.x); dst[4 * i + 1] = __float_as_uint(t.y);
      dst[4 * i + 2] = __float_as_uint(t.z); dst[4 * i + 3] = __float_as_uint(t.w);
    }
  };
  load_x(xr, 0);
  if constexpr (KIND != B_FIRST) load_x(xr2, 1);   // nothing else is live here: both halves in flight at once
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
  const uint32_t taddrA = tmem_base + ((q * 32u) << 16) + sub * 64;   // dead chunk accumulator: staging of x_old
  const uint32_t taddr = taddrA + 256;                                // c_proj accumulator, then the park of x_new
  if constexpr (KIND != B_FIRST) {
    // the old rows wait in TMEM while the phase's last MMAs retire (this warp writes and reads the same lanes / columns)
    tmem_st_32x32b_x32(taddrA, xr);
    tmem_st_32x32b_x32(taddrA + 32, xr2);
    tmem_st_wait();
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
    sm100::tmem_ld_32x32b_x16(taddrA, xv[0]);
    float m_prev = 0.f, q_prev = 0.f;
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      sm100::tmem_ld_wait();
      if (st < 3) {
        sm100::tmem_ld_32x32b_x16(taddr + (st + 1) * 16, av[(st + 1) & 1]);
        sm100::tmem_ld_32x32b_x16(taddrA + (st + 1) * 16, xv[(st + 1) & 1]);
      }
      uint32_t (&v)[16] = av[st & 1];
      const uint32_t (&xo)[16] = xv[st & 1];
      const int col0 = sub * 64 + st * 16;
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
      tmem_st_32x32b_x16(taddr + st * 16, v);
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
    for (int half = 0; half < 2; ++half) {
      if (half == 1) load_x(xr, 1);
      tmem_st_32x32b_x32(taddr + half * 32, xr);
      float ma, qa, mb, qb;
      uint32_t lo[16], hi[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { lo[i] = xr[i]; hi[i] = xr[16 + i]; }
      stats16(lo, ma, qa);
      stats16(hi, mb, qb);
      const float dm = mb - ma;
      exch[(sub * 2 + half) * BLOCK_M + row] = make_float2(0.5f * (ma + mb), fmaf(8.0f * dm, dm, qa + qb));
    }
  }
  tmem_st_wait();
  if constexpr (KIND != B_FIRST) {   // hand the parked rows to the drain warps (B_FIRST: they are in global memory already)
    sm100::tc_fence_before();
    __syncwarp();
    if (lane == 0) sm100::mbar_arrive(sy.park_ready);
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
  // ---- pass 2: A tile of the next phase = bf16( (x_new - mean) * rstd * (1 + mul) + add ), K slab `sub` ----
  uint8_t* a_slab = smem + sub * A_SLAB_BYTES;
  const sm100::f32x2 rstd2 = sm100::pack2(rstd, rstd), nmr2 = sm100::pack2(-mean * rstd, -mean * rstd);   // (x - mean) * rstd = fma(x, rstd, nmr)
  uint32_t pv[2][32];
  sm100::tmem_ld_32x32b_x32(taddr, pv[0]);
  sm100::tmem_ld_32x32b_x32(taddr + 32, pv[1]);
  sm100::tmem_ld_wait();
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const uint32_t (&v)[32] = pv[half];
    const int col0 = sub * 64 + half * 32;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
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
      *reinterpret_cast<uint4*>(a_slab + sm100::swz_chunk_offset(row, half * 4 + c)) = o;
    }
  }
  sm100::tc_fence_before();
  sm100::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) sm100::mbar_arrive(sy.a_ready);
  if (dbg && etid == 0) dbg[3] = clock64();
}

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

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) sm100::tmem_alloc(tmem_ptr_smem, 512);
  if (threadIdx.x == 0) {
    for (int i = 0; i < SB_COUNT; ++i) {
      const bool workers = i == SB_A_READY || i == SB_ACCA_FREE || i == SB_AO_READY || i == SB_H_READY || i == SB_H_READY + 1 || i == SB_PARK_READY;
      sm100::mbar_init(&bars[i], workers ? EPI_WARPS : (i == SB_PARK_DRAINED ? 4 : 1));
    }
    sm100::fence_barrier_init();
  }
  sm100::grid_dep_launch();
  sm100::tc_fence_before();
  __syncthreads();
  sm100::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int T = p.n_chunks;
  const int mlp_items = KSLABS_D * T + p.hid_slabs;
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
          for (int i = 0; i < mlp_items; ++i) {
            if (i == KSLABS_D - 1) put_x(src, B_SLAB_BYTES, x_mlp);
            else put(src, B_SLAB_BYTES);
            src += B_SLAB_BYTES;
          }
          ++n_phase;
        }
      }
    } else if (warp == 1 && lane == 0) {
      // ===================== MMA issuer ========================================================================================
      const uint32_t idesc_q = sm100::make_idesc_bf16(BLOCK_M, AB_QN);
      const uint32_t idesc_n = sm100::make_idesc_bf16(BLOCK_M, BLOCK_N);
      const uint32_t accA = tmem_base, accB = tmem_base + 256;
      const uint32_t a_base = sm100::smem_u32(smA), ao_base = sm100::smem_u32(smMid + 3 * A_SLAB_BYTES), h_base = sm100::smem_u32(smMid);
      RingState rs;
      uint32_t n_ar = 0, u_accA = 0, n_aor = 0, n_hr0 = 0, n_hr1 = 0, n_dr = 0, n_ph = 0;
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
        for (int l = 0; l < p.n_layer; ++l) {
          // ---- attention half: Q0 Q1 P0 Q2 P1 Q3 P2 P3 ----
          sm100::mbar_wait(a_ready, n_ar & 1); ++n_ar;
          sm100::tc_fence_after();
          for (int step = 0; step <= AB_HP; ++step) {
            if (step < AB_HP) {
              if (step > 0) { sm100::mbar_wait(accA_free, (u_accA - 1) & 1); sm100::tc_fence_after(); }
              for (int ks = 0; ks < KSLABS_D; ++ks) {
                if (step == 0 && ks == KSLABS_D - 1) {   // prefetched into the extra slot
                  sm100::mbar_wait(xfull, n_ph & 1);
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
          ++n_dr; ++n_ph;   // the attention -> MLP boundary
          // ---- MLP half: M1_0 M1_1 M2_0 M1_2 M2_1 ... ----
          sm100::mbar_wait(a_ready, n_ar & 1); ++n_ar;
          sm100::tc_fence_after();
          for (int j = 0; j <= T; ++j) {
            if (j < T) {
              if (j > 0) { sm100::mbar_wait(accA_free, (u_accA - 1) & 1); sm100::tc_fence_after(); }
              for (int ks = 0; ks < KSLABS_D; ++ks) {
                if (j == 0 && ks == KSLABS_D - 1) {      // prefetched into the extra slot
                  sm100::mbar_wait(xfull, n_ph & 1);
                  sm100::tc_fence_after();
                  issue_slab_mmas(accA, a_base + ks * A_SLAB_BYTES, sm100::smem_u32(x_mlp), idesc_n, false);
                  continue;
                }
                const uint32_t bs = wait_stage();
                issue_slab_mmas(accA, a_base + ks * A_SLAB_BYTES, bs, idesc_n, ks == 0);
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
          ++n_dr; ++n_ph;   // the MLP -> next attention (or end of tile) boundary
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
      for (int bnd = 0; bnd < 2 * p.n_layer; ++bnd) {
        const bool last = bnd == 2 * p.n_layer - 1;
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
        if (p.dbg != nullptr && warp == S2_DRAIN_WARP0 && lane == 0 && (bnd >> 1) == p.dbg_layer && tile == (int)(blockIdx.x + gridDim.x))
          p.dbg[(size_t)blockIdx.x * 64 + ((bnd & 1) ? 47 : 27)] = clock64();
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
    int tile_no = 0;
    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++tile_no) {
      if (etid < 8) smRows[etid] = p.slot_mod.row(tile * 8 + etid);
      sm100::named_bar_sync(1, EPI_THREADS);   // rows resolved; every warp has left the previous tile
      BoundaryArgs ba{};
      // lane-resolved addresses of (row, column 64 sub) in the io buffer and in the blocked interior storage
      const float* const x_io = p.X + (size_t)tile * BLOCK_M * D + (p.io_blocked ? ((size_t)(sub * 16) * BLOCK_M + row) * 4 : (size_t)row * D + sub * 64);
      const int io_cs = p.io_blocked ? BLOCK_M * 4 : 4;
      const float* const x_in = interior(tile) + ((size_t)(sub * 16) * BLOCK_M + row) * 4;
      ba.xin = x_io; ba.in_cs = io_cs;
      ba.mod = p.mod; ba.rows = smRows; ba.mod_stride = p.mod_stride; ba.eps = p.eps;
      ba.off_mul = 0; ba.off_add = D;
      ba.sm_bias_q = smBiasQ; ba.next_bias_q = p.bias_q;
      sy.n_drained = n_dr;
      boundary_step<B_FIRST, false>(ba, sy, smem, park_pre_attn, tmem_base, q, sub, lane, etid, [] {}, nullptr);
      for (int l = 0; l < p.n_layer; ++l) {
        long long* dbg = (p.dbg != nullptr && l == p.dbg_layer && tile_no == 1) ? p.dbg + (size_t)blockIdx.x * 64 : nullptr;
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
                if (part == 0) {                             // q columns carry a bias (see AttnBlockParams)
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
        sy.accB_parity = n_accB & 1; sy.n_drained = n_dr;
        boundary_step<B_MID, true>(ba, sy, smem, park_pre_mlp, tmem_base, q, sub, lane, etid,
                                   [] { sm100::named_bar_sync(1, EPI_THREADS); /* every attention job has read its q/k/v */ },
                                   dbg ? dbg + 20 : nullptr);
        ++n_accB; ++n_dr;
        ba.xin = x_in; ba.in_cs = BLOCK_M * 4;
        // ================= MLP half =================
        for (int j = 0; j < T; ++j) {
          sm100::mbar_wait(accA_full, n_accA & 1); ++n_accA;
          sm100::tc_fence_after();
          if (dbg && etid == 0 && j < 6) dbg[28 + 2 * j] = clock64();
          uint32_t va[32], vb[32];
          sm100::tmem_ld_32x32b_x32(taddr_q + hs * 64 + hh * 32, va);
          sm100::tmem_ld_32x32b_x32(taddr_q + 128 + hs * 64 + hh * 32, vb);
          sm100::tmem_ld_wait();
          sm100::tc_fence_before();
          __syncwarp();
          if (lane == 0) sm100::mbar_arrive(accA_free);
          const int hb = j & 1;
          const uint32_t cnt = hb ? n_h1 : n_h0;
          if (cnt > 0) sm100::mbar_wait(&h_free[hb], (cnt - 1) & 1);   // the MMAs that read this H buffer last have retired
          if (hb) ++n_h1; else ++n_h0;
          uint8_t* buf = smMid + (hb * 2 + hs) * A_SLAB_BYTES;
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
        sy.accB_parity = n_accB & 1; sy.n_drained = n_dr;
        if (l + 1 < p.n_layer)
          boundary_step<B_MID, false>(ba, sy, smem, park, tmem_base, q, sub, lane, etid, h_region_free, dbg ? dbg + 40 : nullptr);
        else
          boundary_step<B_LAST, false>(ba, sy, smem, park, tmem_base, q, sub, lane, etid, h_region_free, dbg ? dbg + 40 : nullptr);
        ++n_accB; ++n_dr;
        if (dbg && etid == 0) dbg[5] = clock64();
      }
    }
  }
  sm100::tc_fence_before();
  __syncthreads();
  if (warp == 1) sm100::tmem_dealloc(tmem_base, 512);
}

}  // namespace dit
