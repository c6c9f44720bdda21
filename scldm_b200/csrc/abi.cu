// C-ABI entry points (include/scldm_b200.h) and kernel launch sequences.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/scldm_b200.h"
#include "dit_kernels.cuh"
#include "dit_stack.cuh"
#include "vae_kernels.cuh"
#include "csr_kernels.cuh"
#include "train_kernels.cuh"
#include "vae256_kernels.cuh"
#include "eval_kernels.cuh"
#include "vae_train_kernels.cuh"

namespace {

thread_local char g_err[512] = "";
long long* g_dbg_clk = nullptr;   // optional device buffer for kernel phase timestamps (tools/kernel_timeline.py)
int g_dbg_layer = -1;             // which layer's kernels write to it
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_OK(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) return fail(SCLDM_ECUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// ---- optional live per-kernel timing (CUDA events on the launching stream; not graph-capturable) ----
struct ProfRec { const char* name; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_prof_pool;
bool g_prof_on = false;
// Runtime options (scldm_set_option / scldm_get_option).  The SCLDM_* environment variables only provide the initial values.
int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
int g_num_sms = 148;
int g_opt_dec_cpb = env_int("SCLDM_DEC_CPB", 0);      // cells per MCAB decode CTA (0: heuristic)
int g_opt_dec_occ = env_int("SCLDM_DEC_OCC", 2);      // resident CTAs per SM the MCAB decode kernel is compiled for (2: 128 registers, 3: 80 registers + spills)
int g_opt_pdl = env_int("SCLDM_PDL", 1);              // programmatic dependent launch between the kernels of a call
int g_opt_mod_batch = env_int("SCLDM_MOD_BATCH", 1);  // adaLN vectors of all evaluations of a fixed-grid solve from ONE GEMM
int g_opt_solve = env_int("SCLDM_SOLVE", 1);          // fixed-grid ODE solves as ONE launch of dit_stack_kernel (every evaluation in the kernel)
#define g_dec_cpb g_opt_dec_cpb
#define g_dec_occ g_opt_dec_occ
#define g_use_pdl (g_opt_pdl != 0)
#define g_mod_batch (g_opt_mod_batch != 0)
struct OptionEntry { const char* name; int* value; };
const OptionEntry g_options[] = {{"dec_cpb", &g_opt_dec_cpb}, {"dec_occ", &g_opt_dec_occ},
                                 {"pdl", &g_opt_pdl}, {"mod_batch", &g_opt_mod_batch}, {"solve", &g_opt_solve}};
cudaStream_t g_prof_stream = nullptr;

cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
inline void prof_begin(const char* name) {
  if (!g_prof_on) return;
  ProfRec r{name, prof_event(), prof_event()};
  cudaEventRecord(r.a, g_prof_stream);
  g_prof.push_back(r);
}
inline void prof_end() {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().b, g_prof_stream);
}

#define LAUNCH(name, ...)                                                                      \
  do {                                                                                         \
    prof_begin(name);                                                                          \
    __VA_ARGS__;                                                                               \
    prof_end();                                                                                \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                        \
    cudaError_t _e = cudaPeekAtLastError();                                                    \
    if (_e != cudaSuccess) return fail(SCLDM_ECUDA, "launch %s: %s", name, cudaGetErrorString(_e)); \
  } while (0)

// Launch with programmatic stream serialization: the kernel's CTAs may become resident (barrier init, TMEM allocation,
// first weight loads) while the preceding kernel drains; the kernel itself orders its dependent accesses with
// griddepcontrol.wait.  Plain launch while per-kernel event timing is on.
template <typename... KArgs, typename... Args>
inline void launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster > 1) {   // thread-block clusters (CTA pairs for cta_group::2)
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (!g_prof_on && g_use_pdl) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  launch_ex(kern, grid, block, smem, st, 1, std::forward<Args>(args)...);
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// CTA-private blocked storage of dit_stack_kernel when its input / output is row-major (one 128 x 256 fp32 tile per CTA)
inline size_t stack_scratch_bytes(size_t rows16) {
  const size_t tiles = (rows16 + 127) / 128;
  return (tiles < 160 ? tiles : 160) * (size_t)(128 * 256 * 4);
}
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

struct DitWs {
  float* X;
  float* mod;
  float* cls;
  float* temb;
  float* acc;
  float* tvals;
  float4* stage;
  size_t total;
};

// ODE solves on a fixed grid know every evaluation time in advance: when the table fits MOD_BATCH_BYTES, the adaLN vectors of
// ALL evaluations come from one GEMM before the loop (rows = evaluation x conditioning row) instead of one 20 us launch per
// evaluation.  Returns the padded row count of that table, or 0 when the per-evaluation path is used.
constexpr size_t MOD_BATCH_BYTES = 192u << 20;
size_t mod_batch_rows(const scldm_dit_weights* w, const scldm_dit_plan* plan, int n_evals) {
  if (!g_mod_batch || n_evals < 2) return 0;
  const size_t rows = align_up((size_t)n_evals * plan->n_mod, dit::BLOCK_M);
  return rows * (size_t)w->mod_stride * 4 <= MOD_BATCH_BYTES ? rows : 0;
}

DitWs carve_dit(void* base, const scldm_dit_weights* w, const scldm_dit_plan* plan, int n_evals) {
  const size_t slots_pad = scldm_dit_slots_pad(plan);
  const size_t batch_rows = mod_batch_rows(w, plan, n_evals);
  const size_t mod_pad = scldm_dit_mod_pad(plan), mod_rows = batch_rows > mod_pad ? batch_rows : mod_pad;
  const size_t rows = slots_pad * dit::TOK;
  const size_t n_states = (size_t)plan->n_u + plan->n_g;
  const size_t temb_rows = (size_t)n_evals > mod_pad ? (size_t)n_evals : mod_pad;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  char* b = static_cast<char*>(base);
  DitWs ws;
  ws.X = reinterpret_cast<float*>(b + take((rows + dit::BLOCK_M) * dit::D * 4));   // + one tile: the whole-solve kernel may shift the guided slots by one
  ws.mod = reinterpret_cast<float*>(b + take(mod_rows * (size_t)w->mod_stride * 4));
  ws.cls = reinterpret_cast<float*>(b + take(mod_pad * dit::D * 4));
  ws.temb = reinterpret_cast<float*>(b + take(temb_rows * dit::D * 4));
  ws.acc = reinterpret_cast<float*>(b + take(n_states * dit::TOK * dit::LAT * 4));
  ws.tvals = reinterpret_cast<float*>(b + take(align_up((size_t)(n_evals > 0 ? n_evals : 1) * 4, 1024)));
  ws.stage = reinterpret_cast<float4*>(b + take(align_up((size_t)(n_evals > 0 ? n_evals : 1) * 16, 1024)));
  ws.total = off;
  return ws;
}

int check_dit(const scldm_dit_weights* w, const scldm_dit_plan* plan) {
  if (!w || !plan) return fail(SCLDM_EINVAL, "null weights/plan");
  if (w->n_layer < 1 || w->hidden < 1 || w->hidden > 768) return fail(SCLDM_EINVAL, "unsupported DiT dims: n_layer=%d hidden=%d", w->n_layer, w->hidden);
  if (w->hid_slabs != ceil_div(w->hidden, 64) || w->mlp1_tiles != ceil_div(w->hidden, 128))
    return fail(SCLDM_EINVAL, "inconsistent hid_slabs/mlp1_tiles");
  if (w->mod_stride != w->n_layer * 6 * dit::D + 2 * dit::D) return fail(SCLDM_EINVAL, "bad mod_stride %d", w->mod_stride);
  if (w->n_class < 0 || w->n_class > SCLDM_MAX_CLASSES) return fail(SCLDM_EINVAL, "n_class %d out of range", w->n_class);
  if (plan->n_u < 0 || plan->n_g < 0 || plan->n_u + plan->n_g < 1) return fail(SCLDM_EINVAL, "empty plan");
  if (plan->n_g > 0 && (plan->n_f < 1 || plan->n_f > SCLDM_MAX_COMBINE)) return fail(SCLDM_EINVAL, "n_f %d out of range", plan->n_f);
  if (plan->n_mod < 1) return fail(SCLDM_EINVAL, "n_mod must be >= 1");
  if (!plan->slot_mod || (w->n_class > 0 && !plan->cls_idx)) return fail(SCLDM_EINVAL, "null index arrays");
  if (plan->slot_mode < 0 || plan->slot_mode > 2) return fail(SCLDM_EINVAL, "bad slot_mode %d", plan->slot_mode);
  if (!w->w_attn_stream || !w->w_mlp_stream || !w->b_proj_fused || !w->b_qkv || !w->wout_frag || !w->win_frag) return fail(SCLDM_EINVAL, "null packed weights");
  return SCLDM_OK;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SCLDM_OK;
}

struct TimeArgs { float t[64]; };
__global__ void set_times_kernel(float* dst, TimeArgs a, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = a.t[threadIdx.x];
}
struct StageArgs { float4 s[64]; };   // {a_dt, b_dt, first_stage, last_stage} per evaluation (dit_stack_kernel, whole-solve mode)
__global__ void set_stage_kernel(float4* dst, StageArgs a, int n) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = a.s[threadIdx.x];
}

// adaLN modulation vectors of every block + final layer for all conditioning rows
int launch_mod(const scldm_dit_weights* w, const scldm_dit_plan* plan, const DitWs& ws, const float* temb, long long temb_stride,
               cudaStream_t st, int batch_rows = 0, int n_evals = 0) {
  const int mod_pad = batch_rows > 0 ? batch_rows : scldm_dit_mod_pad(plan);
  dit::ModGemmParams p{};
  p.temb = temb;
  p.temb_row_stride = temb_stride;
  p.cls = ws.cls;
  if (batch_rows > 0) { p.cond_group = plan->n_mod; p.cond_rows = n_evals * plan->n_mod; p.temb_row_stride = dit::D; }
  p.Wp = static_cast<const dit::bf16*>(w->w_mod);
  p.n_tiles_total = w->mod_stride / dit::BLOCK_N;
  const int row_tiles = mod_pad / dit::BLOCK_M;
  int groups = ceil_div(2 * 148, row_tiles);              // aim for ~2 waves of CTAs
  if (groups > p.n_tiles_total) groups = p.n_tiles_total;
  if (groups < 1) groups = 1;
  p.tiles_per_cta = ceil_div(p.n_tiles_total, groups);
  p.bias = w->b_mod;
  p.out = ws.mod;
  p.out_ld = w->mod_stride;
  dim3 grid(row_tiles, ceil_div(p.n_tiles_total, p.tiles_per_cta));
  LAUNCH("mod_gemm", launch_pdl(dit::mod_gemm_kernel, grid, dim3(dit::NUM_THREADS), dit::mod_gemm_smem_bytes(), st, p));
  return SCLDM_OK;
}

dit::ModIndex mod_index(const scldm_dit_plan* plan) {
  dit::ModIndex mi{};
  mi.table = plan->slot_mod;
  mi.mode = plan->slot_mode;
  mi.n_u = plan->n_u;
  mi.n_f = plan->n_g > 0 ? plan->n_f : 1;
  mi.n_slots = plan->n_u + plan->n_g * mi.n_f;
  return mi;
}

dit::StackParams stack_params(const scldm_dit_weights* w, const scldm_dit_plan* plan, const DitWs& ws, int row_tiles) {
  dit::StackParams sp{};
  sp.X = ws.X; sp.mod = ws.mod; sp.slot_mod = mod_index(plan); sp.mod_stride = w->mod_stride; sp.eps = w->eps;
  sp.w_attn = static_cast<const dit::bf16*>(w->w_attn_stream); sp.attn_w_stride = 4LL * dit::D * dit::D;
  sp.w_mlp = static_cast<const dit::bf16*>(w->w_mlp_stream);
  sp.hid_last = (w->hidden - 128 * (w->mlp1_tiles - 1) + 31) / 32 * 32;
  sp.mlp_w_stride = ((long long)(w->mlp1_tiles - 1) * dit::KSLABS_D + w->hid_slabs) * dit::B_SLAB_ELEMS + (long long)dit::KSLABS_D * 2 * sp.hid_last * dit::BLOCK_K;
  sp.bias_q = w->b_qkv; sp.bias_proj = w->b_proj_fused;
  sp.n_layer = w->n_layer; sp.n_tiles = row_tiles; sp.n_chunks = w->mlp1_tiles; sp.hid_slabs = w->hid_slabs;
  sp.dbg = g_dbg_clk; sp.dbg_layer = g_dbg_layer;
  return sp;
}

// Fixed-grid solves whose modulation tables were all precomputed run as ONE launch of dit_stack_kernel when a state has at most
// two slots (plain forwards, CFG of a 1-class or joint model): the kernel keeps the two slots of a guided state in one warp.
bool use_solve(const scldm_dit_weights* w, const scldm_dit_plan* plan, size_t batch_rows) {
  if (!g_opt_solve || !w->w_solve || batch_rows == 0) return false;
  return plan->n_g == 0 || plan->n_f <= 2;
}

// the n_layer adaLN blocks on the residual stream ws.X (reference layers.py:208-221)
// `rowmajor_scratch` (dit_stack_kernel only): ws.X is row-major and this buffer (scldm_stack_scratch_bytes) holds the blocked interior
int launch_blocks(const scldm_dit_weights* w, const scldm_dit_plan* plan, const DitWs& ws, cudaStream_t st, float* rowmajor_scratch = nullptr) {
  const int row_tiles = scldm_dit_slots_pad(plan) * dit::TOK / dit::BLOCK_M;
  dit::StackParams sp = stack_params(w, plan, ws, row_tiles);
  sp.io_blocked = rowmajor_scratch == nullptr; sp.scratch = rowmajor_scratch;
  const int grid = row_tiles < g_num_sms ? row_tiles : g_num_sms;
  LAUNCH("dit_stack", launch_pdl(dit::dit_stack_kernel<false>, dim3(grid), dim3(dit::S2_THREADS), dit::stack_smem_bytes(), st, sp));
  return SCLDM_OK;
}

int launch_final(const scldm_dit_weights* w, const dit::StepParams& s, int n_states, cudaStream_t st) {
  dit::StepTcWeights tw{static_cast<const uint2*>(w->wout_frag), static_cast<const uint2*>(w->win_frag)};
  LAUNCH("final_step_tc", launch_pdl(dit::final_step_tc_kernel, dim3(ceil_div(n_states, 4)), dim3(128), 0, st, s, tw, n_states));
  return SCLDM_OK;
}

dit::StepParams make_step(const scldm_dit_weights* w, const scldm_dit_plan* plan, const DitWs& ws) {
  dit::StepParams s{};
  s.X = ws.X; s.mod = ws.mod; s.slot_mod = mod_index(plan); s.mod_stride = w->mod_stride;
  s.mod_off_final = w->n_layer * 6 * dit::D; s.eps = w->eps;
  s.w_out = w->w_out; s.b_out = w->b_out; s.w_in = w->w_in; s.b_in = w->b_in; s.pos = w->pos;
  s.n_u = plan->n_u; s.n_g = plan->n_g; s.n_f = plan->n_g > 0 ? plan->n_f : 1;
  for (int i = 0; i < SCLDM_MAX_COMBINE; ++i) s.coef[i] = plan->coef[i];
  s.acc = ws.acc;
  s.x_blocked = 1;   // dit_stack_kernel keeps the residual stream tile-blocked (x_index)
  return s;
}

int launch_cls(const scldm_dit_weights* w, const scldm_dit_plan* plan, const DitWs& ws, cudaStream_t st) {
  const int mod_pad = scldm_dit_mod_pad(plan);
  dit::ClsParams c{};
  for (int i = 0; i < w->n_class; ++i) c.tables[i] = w->class_tables[i];
  c.idx = plan->cls_idx; c.n_class = w->n_class; c.n_mod_pad = mod_pad;
  LAUNCH("cls", dit::cls_kernel<<<mod_pad, 256, 0, st>>>(c, ws.cls));
  return SCLDM_OK;
}

// cudaFuncSetAttribute and the SM count are per DEVICE: prepared once for every device a call is made on (the caller makes the
// tensors' device current, ops.py::_on_arg_device)
constexpr int MAX_DEVICES = 64;
int g_sms_of[MAX_DEVICES];
int prepare_kernels() {
  static std::atomic<int> done[MAX_DEVICES];
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEVICES) return fail(SCLDM_EINVAL, "device index %d out of range", dev);
  if (done[dev].load()) { g_num_sms = g_sms_of[dev]; return SCLDM_OK; }
  int rc;
  if ((rc = set_smem(dit::mod_gemm_kernel, dit::mod_gemm_smem_bytes()))) return rc;
  if ((rc = set_smem(dit::dit_stack_kernel<false>, dit::stack_smem_bytes()))) return rc;
  if ((rc = set_smem(dit::dit_stack_kernel<true>, dit::stack_smem_bytes()))) return rc;
  {
    int n = 0;
    CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    g_sms_of[dev] = n > 0 ? n : 148;
    g_num_sms = g_sms_of[dev];
  }
  if ((rc = set_smem(vae::dec_latent_multi_kernel, vae::dec_latent_multi_smem_bytes()))) return rc;
  if ((rc = set_smem(vae::mcab_decode_kernel, (vae::MW_TOTAL + vae::TOK * vae::KV) * sizeof(float)))) return rc;
  done[dev].store(1);
  return SCLDM_OK;
}

}  // namespace

extern "C" {

int32_t scldm_dit_slots_pad(const scldm_dit_plan* plan) {
  const int n_f = plan->n_g > 0 ? plan->n_f : 1;
  const int slots = plan->n_u + plan->n_g * n_f;
  return ceil_div(slots, 8) * 8;
}
int32_t scldm_dit_mod_pad(const scldm_dit_plan* plan) { return ceil_div(plan->n_mod, dit::BLOCK_M) * dit::BLOCK_M; }

size_t scldm_dit_workspace_bytes(const scldm_dit_weights* w, const scldm_dit_plan* plan, int32_t n_evals) {
  if (!w || !plan) return 0;
  return carve_dit(nullptr, w, plan, n_evals).total + 1024;
}

int32_t scldm_dit_workspace_layout(const scldm_dit_weights* w, const scldm_dit_plan* plan, int32_t n_evals, size_t* offsets,
                                   int32_t max_entries) {
  if (!w || !plan || !offsets) return 0;
  const DitWs ws = carve_dit(nullptr, w, plan, n_evals);
  const size_t v[6] = {(size_t)ws.X, (size_t)ws.mod, (size_t)ws.cls, (size_t)ws.temb, (size_t)ws.acc, (size_t)ws.tvals};
  int n = 0;
  for (; n < 6 && n < max_entries; ++n) offsets[n] = v[n];
  return n;
}

namespace {
// one evaluation; t_mod != nullptr: one time per conditioning row, else every row shares the time `t_shared`
int dit_forward_impl(const scldm_dit_weights* w, const scldm_dit_plan* plan, const float* x, const float* t_mod, float t_shared, float* v_out,
                     void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = check_dit(w, plan))) return rc;
  if (!x || !v_out || !workspace) return fail(SCLDM_EINVAL, "null buffer");
  if (workspace_bytes < scldm_dit_workspace_bytes(w, plan, 0)) return fail(SCLDM_ENOMEM, "workspace too small");
  if ((rc = prepare_kernels())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* base = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  const DitWs ws = carve_dit(base, w, plan, 0);
  const int mod_pad = scldm_dit_mod_pad(plan);
  const int n_states = plan->n_u + plan->n_g;

  if ((rc = launch_cls(w, plan, ws, st))) return rc;
  if (t_mod) {
    LAUNCH("temb", dit::temb_kernel<<<mod_pad, dit::TEMB_THREADS, 0, st>>>(t_mod, mod_pad, w->temb_w0t, w->temb_b0, w->temb_w2t, w->temb_b2, ws.temb));
    if ((rc = launch_mod(w, plan, ws, ws.temb, dit::D, st))) return rc;
  } else {   // shared time: one embedding row, broadcast by the adaLN GEMM's prologue (row stride 0)
    TimeArgs ta{};
    ta.t[0] = t_shared;
    LAUNCH("set_times", set_times_kernel<<<1, 64, 0, st>>>(ws.tvals, ta, 1));
    LAUNCH("temb", dit::temb_kernel<<<1, dit::TEMB_THREADS, 0, st>>>(ws.tvals, 1, w->temb_w0t, w->temb_b0, w->temb_w2t, w->temb_b2, ws.temb));
    if ((rc = launch_mod(w, plan, ws, ws.temb, 0, st))) return rc;
  }

  dit::StepParams s = make_step(w, plan, ws);
  s.x_base = const_cast<float*>(x);  // read only in inproj
  LAUNCH("inproj", dit::inproj_kernel<<<n_states, 256, 0, st>>>(s));
  if ((rc = launch_blocks(w, plan, ws, st))) return rc;
  s.v_out = v_out; s.do_update = 0; s.do_inproj = 0;
  if ((rc = launch_final(w, s, n_states, st))) return rc;
  return SCLDM_OK;
}
}  // namespace

int scldm_dit_forward(const scldm_dit_weights* w, const scldm_dit_plan* plan, const float* x, const float* t_mod, float* v_out,
                      void* workspace, size_t workspace_bytes, void* stream) {
  if (!t_mod) return fail(SCLDM_EINVAL, "null buffer");
  return dit_forward_impl(w, plan, x, t_mod, 0.f, v_out, workspace, workspace_bytes, stream);
}

int scldm_dit_forward_shared_t(const scldm_dit_weights* w, const scldm_dit_plan* plan, const float* x, float t, float* v_out,
                               void* workspace, size_t workspace_bytes, void* stream) {
  return dit_forward_impl(w, plan, x, nullptr, t, v_out, workspace, workspace_bytes, stream);
}

int scldm_dit_sample_ode(const scldm_dit_weights* w, const scldm_dit_plan* plan, float* x, const float* t_grid_host, int32_t n_grid,
                         int32_t method, void* workspace, size_t workspace_bytes, void* stream) {
  int rc;
  if ((rc = check_dit(w, plan))) return rc;
  if (!x || !t_grid_host || !workspace) return fail(SCLDM_EINVAL, "null buffer");
  if (n_grid < 2) return fail(SCLDM_EINVAL, "need at least 2 grid points");
  if (method != SCLDM_ODE_EULER && method != SCLDM_ODE_HEUN2 && method != SCLDM_ODE_MIDPOINT)
    return fail(SCLDM_EINVAL, "unsupported ODE method %d (fixed-grid euler/heun2/midpoint only)", method);
  const int n_steps = n_grid - 1;
  const int stages = method == SCLDM_ODE_EULER ? 1 : 2;
  const int n_evals = n_steps * stages;
  if (workspace_bytes < scldm_dit_workspace_bytes(w, plan, n_evals)) return fail(SCLDM_ENOMEM, "workspace too small");
  if ((rc = prepare_kernels())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* base = reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  const DitWs ws = carve_dit(base, w, plan, n_evals);
  const int n_states = plan->n_u + plan->n_g;

  // evaluation times (fp32 arithmetic as torchdiffeq's fixed-grid solvers: t0, t0+dt/2 or t1)
  {
    int e = 0;
    while (e < n_evals) {
      TimeArgs ta{};
      int n = 0;
      for (; n < 64 && e + n < n_evals; ++n) {
        const int idx = e + n, k = idx / stages, sg = idx % stages;
        const float t0 = t_grid_host[k], t1 = t_grid_host[k + 1];
        const float dt = t1 - t0;
        float tv = t0;
        if (sg == 1) tv = (method == SCLDM_ODE_HEUN2) ? t1 : t0 + 0.5f * dt;
        ta.t[n] = tv;
      }
      LAUNCH("set_times", set_times_kernel<<<1, 64, 0, st>>>(ws.tvals + e, ta, n));
      e += n;
    }
  }
  LAUNCH("temb", dit::temb_kernel<<<n_evals, dit::TEMB_THREADS, 0, st>>>(ws.tvals, n_evals, w->temb_w0t, w->temb_b0, w->temb_w2t, w->temb_b2, ws.temb));
  if ((rc = launch_cls(w, plan, ws, st))) return rc;

  const int batch_rows = (int)mod_batch_rows(w, plan, n_evals);
  if (use_solve(w, plan, batch_rows)) {
    // ---- the whole solve in one launch: every evaluation, the CFG combine and the stage updates inside dit_stack_kernel ----
    for (int e0 = 0; e0 < n_evals; e0 += 64) {
      StageArgs sa{};
      int n = 0;
      for (; n < 64 && e0 + n < n_evals; ++n) {
        const int idx = e0 + n, k = idx / stages, sg = idx % stages;
        const float dt = t_grid_host[k + 1] - t_grid_host[k];
        float a_dt, b_dt;
        if (method == SCLDM_ODE_EULER) { a_dt = 0.f; b_dt = dt; }
        else if (method == SCLDM_ODE_HEUN2) { a_dt = dt; b_dt = 0.5f * dt; }
        else { a_dt = 0.5f * dt; b_dt = sg == 0 ? 0.f : dt; }
        sa.s[n] = make_float4(a_dt, b_dt, sg == 0 ? 1.f : 0.f, sg == stages - 1 ? 1.f : 0.f);
      }
      LAUNCH("set_stage", set_stage_kernel<<<1, 64, 0, st>>>(ws.stage + e0, sa, n));
    }
    if ((rc = launch_mod(w, plan, ws, ws.temb, dit::D, st, batch_rows, n_evals))) return rc;
    const int n_f = plan->n_g > 0 ? plan->n_f : 1;
    const int slot_shift = (n_f == 2 && (plan->n_u & 1)) ? 1 : 0;   // see StackParams::slot_shift
    const int row_tiles = ceil_div(plan->n_u + slot_shift + plan->n_g * n_f, 8);
    dit::StackParams sp = stack_params(w, plan, ws, row_tiles);
    sp.slot_shift = slot_shift;
    sp.X = nullptr; sp.io_blocked = 0; sp.scratch = ws.X;   // residual rows: one CTA-private tile each, L2 resident
    sp.n_evals = n_evals; sp.mod_eval_stride = (long long)plan->n_mod * w->mod_stride; sp.stage = ws.stage;
    sp.w_solve = static_cast<const dit::bf16*>(w->w_solve); sp.b_out = w->b_out;
    sp.x_base = x; sp.acc = ws.acc;
    sp.n_u = plan->n_u; sp.n_g = plan->n_g; sp.n_f = n_f;
    sp.coef[0] = plan->coef[0]; sp.coef[1] = plan->coef[1];
    sp.mod_off_final = w->n_layer * 6 * dit::D;
    const int grid = row_tiles < g_num_sms ? row_tiles : g_num_sms;
    LAUNCH("dit_solve", launch_pdl(dit::dit_stack_kernel<true>, dim3(grid), dim3(dit::S2_THREADS), dit::stack_smem_bytes(), st, sp));
    return SCLDM_OK;
  }
  dit::StepParams s = make_step(w, plan, ws);
  s.x_base = x;
  LAUNCH("inproj", dit::inproj_kernel<<<n_states, 256, 0, st>>>(s));
  s.do_update = 1;
  if (batch_rows > 0 && (rc = launch_mod(w, plan, ws, ws.temb, dit::D, st, batch_rows, n_evals))) return rc;
  for (int k = 0; k < n_steps; ++k) {
    const float dt = t_grid_host[k + 1] - t_grid_host[k];
    for (int sg = 0; sg < stages; ++sg) {
      const int e = k * stages + sg;
      DitWs wse = ws;
      if (batch_rows > 0) {   // this evaluation's rows of the precomputed table
        wse.mod = ws.mod + (size_t)e * plan->n_mod * w->mod_stride;
        s.mod = wse.mod;
      } else if ((rc = launch_mod(w, plan, ws, ws.temb + (size_t)e * dit::D, 0, st))) return rc;
      if ((rc = launch_blocks(w, plan, wse, st))) return rc;
      s.first_stage = sg == 0;
      s.last_stage = sg == stages - 1;
      if (method == SCLDM_ODE_EULER) { s.a_dt = 0.f; s.b_dt = dt; }
      else if (method == SCLDM_ODE_HEUN2) { s.a_dt = dt; s.b_dt = 0.5f * dt; }
      else { s.a_dt = 0.5f * dt; s.b_dt = sg == 0 ? 0.f : dt; }
      s.do_inproj = !(k == n_steps - 1 && s.last_stage);
      if ((rc = launch_final(w, s, n_states, st))) return rc;
    }
  }
  return SCLDM_OK;
}

int scldm_vae_qside(const scldm_vae_dec_weights* w, float* qp, void* qp_bf16, void* stream) {
  if (!w || (!qp && !qp_bf16)) return fail(SCLDM_EINVAL, "null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LAUNCH("qside", vae::qside_kernel<<<ceil_div(w->n_ids, 128), 128, 0, st>>>(w->emb, w->ca_ln1q_w, w->ca_ln1q_b, w->ca_wq, w->eps, w->n_ids, qp, static_cast<__nv_bfloat16*>(qp_bf16)));
  return SCLDM_OK;
}

size_t scldm_vae_decode_workspace_bytes(int32_t n_cells, int32_t n_genes) {
  const size_t tiles = ceil_div(n_genes, 128);
  return align_up((size_t)n_cells * vae::TOK * vae::KV * 4, 1024) + align_up((size_t)n_cells * n_genes * 4, 1024) +
         align_up((size_t)n_cells * tiles * 8 * 8, 1024) + align_up((size_t)n_cells * 1024 * 2, 1024) + 1024;   // 8 softmax partials per gene tile
}

int scldm_vae_decode(const scldm_vae_dec_weights* w, const float* qp, const void* qp_bf16, const float* z, int32_t n_cells,
                     const int64_t* genes, int32_t n_genes, const float* lib, float* mu, float* theta, float* counts, uint64_t seed,
                     int64_t cell_offset, int32_t precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (!w || !z || !genes || !lib || !workspace) return fail(SCLDM_EINVAL, "null argument");
  if (precision != SCLDM_DECODE_TC && precision != SCLDM_DECODE_FP32) return fail(SCLDM_EINVAL, "bad precision %d", precision);
  if (precision == SCLDM_DECODE_TC && (!qp_bf16 || !w->mcab_wfrag || !w->mcab_small)) return fail(SCLDM_EINVAL, "tensor-core decode needs qp_bf16 + fragment weights");
  if (precision == SCLDM_DECODE_FP32 && !qp) return fail(SCLDM_EINVAL, "fp32 decode needs the fp32 Q-side table");
  if (n_cells < 1 || n_genes < 1) return fail(SCLDM_EINVAL, "empty decode: n_cells=%d n_genes=%d", n_cells, n_genes);
  if (workspace_bytes < scldm_vae_decode_workspace_bytes(n_cells, n_genes)) return fail(SCLDM_ENOMEM, "workspace too small");
  const bool unshared = w->theta_tbl == nullptr;   // unshared-theta head: theta is [n_cells][n_genes] and always needed (the NB draw reads it)
  if (unshared && !theta) return fail(SCLDM_EINVAL, "unshared-theta head: a theta buffer [n_cells][n_genes] is required");
  int rc;
  if ((rc = prepare_kernels())) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tiles = ceil_div(n_genes, 128);
  char* b = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
  float* kv = reinterpret_cast<float*>(b);
  b += align_up((size_t)n_cells * vae::TOK * vae::KV * 4, 1024);
  float* logits_ws = reinterpret_cast<float*>(b);
  b += align_up((size_t)n_cells * n_genes * 4, 1024);
  float2* partials = reinterpret_cast<float2*>(b);
  b += align_up((size_t)n_cells * tiles * 8 * 8, 1024);
  __nv_bfloat16* kvb = reinterpret_cast<__nv_bfloat16*>(b);
  float* logits = mu ? mu : logits_ws;  // finalised in place when mu is requested
  const bool tc = precision == SCLDM_DECODE_TC;

  vae::DecLatentParams dp{};
  dp.z = z; dp.win_t = w->win_t; dp.blocks = w->blocks; dp.n_layer = w->n_layer;
  dp.ca_ln1_w = w->ca_ln1_w; dp.ca_ln1_b = w->ca_ln1_b; dp.ca_wkv_t = w->ca_wkv_t; dp.eps = w->eps;
  dp.kv = tc ? nullptr : kv; dp.kvb = tc ? kvb : nullptr;
  LAUNCH("dec_latent", vae::dec_latent_multi_kernel<<<ceil_div(n_cells, vae::DL_CELLS), 128 * vae::DL_CELLS, vae::dec_latent_multi_smem_bytes(), st>>>(dp, n_cells));

  // enough blocks for ~4 waves, but amortise the gene-side loads over several cells
  int cpb = (int)(((long long)tiles * n_cells) / (148LL * 2 * 4));
  if (cpb < 1) cpb = 1;
  if (cpb > 32) cpb = 32;
  if (g_dec_cpb > 0) cpb = g_dec_cpb;
  if (tc) {
    vae::McabTcParams mp{};
    mp.emb = w->emb; mp.qp = static_cast<const __nv_bfloat16*>(qp_bf16); mp.genes = reinterpret_cast<const long long*>(genes);
    mp.G = n_genes; mp.kvb = kvb; mp.n_cells = n_cells; mp.cells_per_block = cpb;
    mp.wfrag = static_cast<const uint32_t*>(w->mcab_wfrag); mp.small = w->mcab_small; mp.eps = w->eps;
    mp.logits = logits; mp.partials = partials; mp.gene_tiles = tiles; mp.log_theta = unshared ? theta : nullptr;
    if (g_dec_occ == 3) LAUNCH("mcab_decode_tc", vae::mcab_decode_tc_kernel<3><<<dim3(tiles, ceil_div(n_cells, cpb)), 256, 0, st>>>(mp));
    else LAUNCH("mcab_decode_tc", vae::mcab_decode_tc_kernel<2><<<dim3(tiles, ceil_div(n_cells, cpb)), 256, 0, st>>>(mp));
  } else {
    vae::McabParams mp{};
    mp.emb = w->emb; mp.qp = qp; mp.genes = reinterpret_cast<const long long*>(genes); mp.G = n_genes; mp.kv = kv; mp.n_cells = n_cells;
    mp.cells_per_block = cpb;
    mp.wblob = w->mcab_blob; mp.eps = w->eps; mp.logits = logits; mp.partials = partials; mp.gene_tiles = tiles;
    mp.log_theta = unshared ? theta : nullptr;
    LAUNCH("mcab_decode", vae::mcab_decode_kernel<<<dim3(tiles, ceil_div(n_cells, cpb)), 128, (vae::MW_TOTAL + vae::TOK * vae::KV) * sizeof(float), st>>>(mp));
  }

  vae::NbParams np{};
  np.logits = logits; np.partials = partials; np.gene_tiles = tc ? tiles * 8 : tiles; np.G = n_genes; np.n_cells = n_cells; np.lib = lib;
  np.theta_tbl = w->theta_tbl; np.genes = reinterpret_cast<const long long*>(genes); np.mu = mu; np.theta = theta; np.counts = counts;
  np.seed = seed; np.cell_offset = cell_offset;
  int gx = ceil_div(n_genes, 256 * 4);
  if (gx < 1) gx = 1;
  LAUNCH("nb_finalize", vae::nb_finalize_kernel<<<dim3(n_cells, gx), 256, 0, st>>>(np));
  return SCLDM_OK;
}

int scldm_vae_encode(const scldm_vae_enc_weights* w, const int64_t* genes_subset, const float* counts_subset, int32_t n_cells,
                     int32_t seq_len, float* z, void* stream) {
  if (!w || !genes_subset || !counts_subset || !z) return fail(SCLDM_EINVAL, "null argument");
  if (n_cells < 1 || seq_len < 1) return fail(SCLDM_EINVAL, "empty encode: n_cells=%d seq_len=%d", n_cells, seq_len);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  vae::EncParams p{};
  p.agg = w->agg_func;
  p.emb = w->emb; p.genes = reinterpret_cast<const long long*>(genes_subset); p.counts = counts_subset; p.S = seq_len; p.n_cells = n_cells;
  p.wkv_frag = static_cast<const uint32_t*>(w->wkv_frag); p.q_tbl = static_cast<const __nv_bfloat16*>(w->q_tbl);
  p.ln1_w = w->ln1_w; p.ln1_b = w->ln1_b; p.inducing = w->inducing; p.wproj_t = w->wproj_t; p.ln2_w = w->ln2_w; p.ln2_b = w->ln2_b;
  p.w1_t = w->w1_t; p.w2_t = w->w2_t; p.w3_t = w->w3_t; p.pos = w->has_pos ? w->pos : nullptr; p.blocks = w->blocks; p.n_layer = w->n_layer;
  p.wlat_t = w->wlat_t; p.eps = w->eps; p.z = z;
  LAUNCH("mcab_encode", vae::mcab_encode_kernel<<<n_cells, 256, 0, st>>>(p));
  return SCLDM_OK;
}

int scldm_randn_cells(float* out, int32_t n_cells, int32_t per_cell, uint64_t seed, int64_t cell_offset, uint32_t stream_id,
                      void* stream) {
  if (!out || n_cells < 1 || per_cell < 1) return fail(SCLDM_EINVAL, "bad randn arguments");
  const long long n = (long long)n_cells * per_cell;
  LAUNCH("randn_cells", vae::randn_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n_cells, per_cell, seed,
                                                                                                      cell_offset, stream_id));
  return SCLDM_OK;
}

int scldm_csr_count(const float* dense, int32_t rows, int32_t G, int32_t* row_nnz, int64_t* indptr, void* stream) {
  if (!dense || !row_nnz || !indptr || rows < 0 || G < 1) return fail(SCLDM_EINVAL, "bad csr_count arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows > 0) LAUNCH("csr_count", csr::count_kernel<<<rows, csr::THREADS, 0, st>>>(dense, G, row_nnz));
  LAUNCH("csr_scan", csr::scan_kernel<<<1, 1024, 0, st>>>(row_nnz, rows, reinterpret_cast<long long*>(indptr)));
  return SCLDM_OK;
}

int scldm_csr_fill(const float* dense, int32_t rows, int32_t G, const int64_t* indptr, int32_t* indices, float* data, void* stream) {
  if (!dense || !indptr || rows < 0 || G < 1) return fail(SCLDM_EINVAL, "bad csr_fill arguments");
  if (rows == 0) return SCLDM_OK;
  // indices / data may be NULL only when the matrix has no non-zero at all (nothing is written then)
  LAUNCH("csr_fill", csr::fill_kernel<<<rows, csr::THREADS, 0, static_cast<cudaStream_t>(stream)>>>(dense, G, reinterpret_cast<const long long*>(indptr),
                                                                                           indices, data));
  return SCLDM_OK;
}

int scldm_nb_nll(const float* x, const float* mu, const float* theta, int64_t theta_row_stride, int32_t rows, int32_t G, float* nll,
                 void* stream) {
  if (!x || !mu || !theta || !nll || rows < 0 || G < 1 || (theta_row_stride != 0 && theta_row_stride < G))
    return fail(SCLDM_EINVAL, "bad nb_nll arguments");
  if (rows == 0) return SCLDM_OK;
  LAUNCH("nb_nll", vae::nb_nll_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, mu, theta, (long long)theta_row_stride, G, nll));
  return SCLDM_OK;
}

int scldm_tokenize_expressed(const float* dense, int32_t rows, int32_t G, const int64_t* gene_ids, int32_t S, int64_t mask_idx,
                             int64_t* genes_subset, float* counts_subset, float* library, int32_t* overflow, void* stream) {
  if (!dense || !gene_ids || !genes_subset || !counts_subset || !library || !overflow || rows < 0 || G < 1 || S < 1)
    return fail(SCLDM_EINVAL, "bad tokenize_expressed arguments");
  if (rows == 0) return SCLDM_OK;
  LAUNCH("tokenize_expressed", csr::tokenize_expressed_kernel<<<rows, csr::THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
                                   dense, G, reinterpret_cast<const long long*>(gene_ids), S, (long long)mask_idx,
                                   reinterpret_cast<long long*>(genes_subset), counts_subset, library, overflow));
  return SCLDM_OK;
}

uint64_t scldm_launch_count(void) { return g_launches.load(); }

int scldm_set_option(const char* name, int32_t value) {
  if (!name) return fail(SCLDM_EINVAL, "null option name");
  for (const OptionEntry& o : g_options)
    if (strcmp(o.name, name) == 0) { *o.value = value; return SCLDM_OK; }
  return fail(SCLDM_EINVAL, "unknown option '%s' (dec_cpb, dec_occ, pdl, mod_batch, solve)", name);
}
int32_t scldm_get_option(const char* name) {
  if (name)
    for (const OptionEntry& o : g_options)
      if (strcmp(o.name, name) == 0) return *o.value;
  return -1;
}

void scldm_debug_timeline(long long* device_buf, int32_t layer) {
  g_dbg_clk = device_buf;
  g_dbg_layer = layer;
}

void scldm_prof_enable(int32_t on, void* stream) {
  g_prof_on = on != 0;
  g_prof_stream = static_cast<cudaStream_t>(stream);
}

// "name count total_ms\n" per kernel class since the last call; returns the number of bytes written
int32_t scldm_prof_summary(char* buf, int32_t cap) {
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
    g_prof_pool.push_back(r.a);
    g_prof_pool.push_back(r.b);
  }
  g_prof.clear();
  int off = 0;
  for (auto& kv : agg) {
    int n = snprintf(buf + off, cap > off ? cap - off : 0, "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    if (n < 0 || off + n >= cap) break;
    off += n;
  }
  return off;
}
const char* scldm_last_error(void) { return g_err; }
const char* scldm_version(void) { return "scldm_b200 0.1 (sm_100a)"; }

}  // extern "C"

#include "train_abi.inc"
#include "vae256_abi.inc"
#include "vae_train_abi.inc"

extern "C" {

int scldm_pair_stats(const float* x, int32_t nx, const float* y, int32_t ny, int32_t D, float* out, void* stream) {
  if (!x || !y || !out || nx < 1 || ny < 1 || D < 1) return fail(SCLDM_EINVAL, "bad pair_stats arguments");
  LAUNCH("pair_stats", evk::pair_stats_kernel<<<dim3(ceil_div(ny, evk::PT), ceil_div(nx, evk::PT)), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, nx, y, ny, D, out));
  return SCLDM_OK;
}

int scldm_sinkhorn(const float* K, const float* M, const float* a, const float* b, int32_t n, int32_t m, float* u, float* v, int32_t n_iter, float* res,
                   void* stream) {
  if (!K || !M || !a || !b || !u || !v || !res || n < 1 || m < 1 || n_iter < 0) return fail(SCLDM_EINVAL, "bad sinkhorn arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int it = 0; it < n_iter; ++it) {
    LAUNCH("sinkhorn_ktu", evk::sinkhorn_ktu_kernel<<<ceil_div(m, 256), 256, 0, st>>>(K, u, b, n, m, v));
    LAUNCH("sinkhorn_kv", evk::sinkhorn_kv_kernel<<<ceil_div(n, 8), 256, 0, st>>>(K, v, a, n, m, u));
  }
  CUDA_OK(cudaMemsetAsync(res, 0, 8, st));
  LAUNCH("sinkhorn_eval", evk::sinkhorn_eval_kernel<<<ceil_div(m, 256), 256, 0, st>>>(K, M, u, v, b, n, m, res));
  return SCLDM_OK;
}

int scldm_sde_drift(const float* v, const float* x, float t, int32_t form, float norm, float* drift, int64_t n, void* stream) {
  if (!v || !x || !drift || n < 1 || form < 0 || form > 5) return fail(SCLDM_EINVAL, "bad sde_drift arguments");
  LAUNCH("sde_drift", evk::sde_drift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(v, x, t, form, norm, drift, (long long)n));
  return SCLDM_OK;
}

int scldm_sde_kick(const float* x, const float* noise, float t, float dt, int32_t form, float norm, uint64_t seed, int64_t cell_offset, int32_t per_cell,
                   uint32_t step, float* out, int64_t n, void* stream) {
  if (!x || !out || n < 1 || per_cell < 1 || form < 0 || form > 5) return fail(SCLDM_EINVAL, "bad sde_kick arguments");
  LAUNCH("sde_kick", evk::sde_kick_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, noise, t, dt, form, norm, seed, (long long)cell_offset,
                                                                                                         per_cell, step, out, (long long)n));
  return SCLDM_OK;
}

int scldm_axpy2(const float* a, float c1, const float* d1, float c2, const float* d2, float* out, int64_t n, void* stream) {
  if (!a || !d1 || !out || n < 1) return fail(SCLDM_EINVAL, "bad axpy2 arguments");
  LAUNCH("axpy2", evk::axpy2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, c1, d1, c2, d2, out, (long long)n));
  return SCLDM_OK;
}

}  // extern "C"
