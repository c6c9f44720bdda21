/* scldm_b200 C-ABI: B200 (sm_100a) kernels for scLDM's generation hot path.
 *
 * The reference (czi-ai/scldm) has NO FFI / plugin interface for this path -- it is plain
 * nn.Module composition.  This header is therefore the *new* seam a
 * maintainer would bind underneath the reference's Python classes; every entry point names the
 * reference code it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host
 *   - the caller owns every buffer incl. the workspace; the library never allocates or frees
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no host
 *     synchronisation and no H2D/D2H copies => capturable in a CUDA graph
 *   - return value: 0 on success, negative SCLDM_E* on error; scldm_last_error() gives the message
 *   - there is no CPU fallback: without a CUDA device every compute call returns SCLDM_ECUDA
 */
#ifndef SCLDM_B200_H
#define SCLDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCLDM_OK 0
#define SCLDM_EINVAL -1  /* bad argument / unsupported shape */
#define SCLDM_ECUDA -2   /* CUDA runtime error              */
#define SCLDM_ENOMEM -3  /* workspace too small             */

#define SCLDM_MAX_CLASSES 8
#define SCLDM_MAX_COMBINE 8

#define SCLDM_ODE_EULER 0
#define SCLDM_ODE_HEUN2 1
#define SCLDM_ODE_MIDPOINT 2

/* Packed DiT weights (device).  Produced from the reference state_dict by scldm_b200/pack.py.
 * bf16 matrices are stored as 128B-swizzled K-major UMMA tiles [n_tile][k_slab][256 x 64]. */
typedef struct scldm_dit_weights {
  int32_t n_layer;    /* DiT blocks (nnets.py:247-262)                                        */
  int32_t hidden;     /* SwiGLU hidden width, 684 for n_embed=256 (layers.py:165-167)          */
  int32_t hid_slabs;  /* ceil(hidden/64): K slabs of mlp.c_proj                                */
  int32_t mlp1_tiles; /* ceil(hidden/128): N tiles of the fused [w1|w2] GEMM                   */
  int32_t mod_stride; /* n_layer*1536 + 512: floats per row of the modulation table            */
  int32_t n_class;    /* number of class-embedding tables (sorted by class name)               */
  float eps;          /* LayerNorm eps (1e-8)                                                  */
  const void* w_mod;    /* bf16 [mod_stride/256][4][256x64]: all adaLN Linear weights, stacked   */
  const float* b_mod;   /* [mod_stride]                                                          */
  const float* b_qkv;   /* [n_layer][768] attn.c_attn.bias (the kernels add the q part; k cancels in the softmax, v is folded below) */
  const void* w_mlp_stream; /* bf16, per layer the SwiGLU MLP in the consumption order of dit_stack_kernel: N tile j of the fused
                               [w1|w2] GEMM = four K slabs [w1 rows of hidden chunk j (halved) | w2 rows of chunk j] x 64, c_proj
                               slab c = [256 x 64] for hidden columns 64c..64c+63; order M1_0, M1_1, M2_0, M1_2, M2_1, ...
                               (M2_j = the one or two c_proj slabs of chunk j).  Chunks are 128 hidden units wide; the last one
                               is padded to a multiple of 16 only.                                                         */
  const void* w_attn_stream; /* bf16 [n_layer][262144]: attn.c_attn + attn.c_proj in consumption order.  Per head pair hp (heads
                                2hp, 2hp+1) a "Q item" is the 192 x 64 K-major swizzled slab [Wq rows 64hp.. | Wk rows 64hp.. |
                                Wv rows 64hp..] for one 64-wide K slab, a "P item" the 256 x 64 slab c_proj.weight[:, 64hp..64hp+64];
                                order Q_0 (4 items), Q_1, P_0, Q_2, P_1, Q_3, P_2, P_3                                      */
  const float* b_proj_fused; /* [n_layer][256] = c_proj.bias + c_proj.weight @ c_attn.bias[512:768]: the v bias passes through
                                the softmax-weighted averaging unchanged                                                    */
  const float* temb_w0t; /* t_embedder.mlp.0.weight^T [256][256] */
  const float* temb_b0;
  const float* temb_w2t; /* t_embedder.mlp.2.weight^T [256][256] */
  const float* temb_b2;
  const float* w_in;   /* input_proj.weight [256][16]        */
  const float* b_in;   /* [256]                               */
  const float* pos;    /* pos_embed [16][256]                 */
  const float* w_out;  /* final_layer.linear.weight [16][256] */
  const float* b_out;  /* [16]                                */
  const void* wout_frag; /* bf16 final_layer.linear.weight in mma.sync B-fragment order [16][2][32][4]                         */
  const void* win_frag;  /* bf16 input_proj.weight in mma.sync B-fragment order [1][32][32][4]                                   */
  const float* class_tables[SCLDM_MAX_CLASSES]; /* class_embeddings.<name>.weight [(V+1)][256] */
  const void* w_solve;  /* bf16, whole-solve kernel (scldm_dit_sample_ode): one 256 x 64 K-major swizzled slab, row n =
                           [input_proj.weight[n, 0:16] | the same again | bf16 hi part of (pos_embed + input_proj.bias)[0:16, n] |
                           its bf16 lo part] (the state enters as bf16 hi | lo parts followed by one-hot(token) twice), then
                           final_layer.linear.weight as four 16 x 64 slabs.  NULL: one block-stack launch + one step kernel per
                           evaluation                                                                                         */
} scldm_dit_weights;

/* Which model evaluations one call performs and how they are combined.
 *   states [0,n_u)       : evaluated once (slot = state)
 *   states [n_u,n_u+n_g) : evaluated n_f times (consecutive slots); v = sum_k coef[k]*out_k
 * This expresses DiT.forward (n_g=0) and DiT.forward_with_cfg (nnets.py:336-378: first half
 * unconditional, second half guided with coef = [1-sum(w), w_1, ...]).                        */
typedef struct scldm_dit_plan {
  int32_t n_u, n_g, n_f;
  float coef[SCLDM_MAX_COMBINE];
  int32_t n_mod;          /* distinct conditioning rows                                   */
  const int32_t* cls_idx; /* [n_class][n_mod_pad] embedding row per class (null = vocab)   */
  const int32_t* slot_mod;/* [slots_pad] conditioning row of every slot                    */
  int32_t slot_mode;      /* 0: use the slot_mod table; 1: slot_mod[s] == s; 2: the shared-time CFG layout
                             (slots [0,n_u) -> row 0; guided cell j, pass k: k==0 -> row 0, else 1 + j*(n_f-1) + k-1).
                             Modes 1/2 let the kernels compute the row instead of loading it (no dependent load).  */
} scldm_dit_plan;

/* padded sizes the index arrays / workspace must honour (multiples of 8 slots / 128 mod rows) */
int32_t scldm_dit_slots_pad(const scldm_dit_plan* plan);
int32_t scldm_dit_mod_pad(const scldm_dit_plan* plan);
size_t scldm_dit_workspace_bytes(const scldm_dit_weights* w, const scldm_dit_plan* plan, int32_t n_evals);

/* Introspection for tests: byte offsets (from the 1024-aligned workspace base) of
 * {X, qkv, attn_out, hidden, mod, cls, temb, acc, tvals}; returns the number of entries written. */
int32_t scldm_dit_workspace_layout(const scldm_dit_weights* w, const scldm_dit_plan* plan, int32_t n_evals, size_t* offsets,
                                   int32_t max_entries);

/* Replaces DiT.forward / DiT.forward_with_cfg (nnets.py:273-297, 336-378).
 *   x      [n_u+n_g][16][16] fp32 states;  t_mod [n_mod_pad] fp32 time of every conditioning row
 *   v_out  [n_u+n_g][16][16] fp32 combined model output                                        */
int scldm_dit_forward(const scldm_dit_weights* w, const scldm_dit_plan* plan, const float* x, const float* t_mod,
                      float* v_out, void* workspace, size_t workspace_bytes, void* stream);

/* Same evaluation when every conditioning row shares one time t (what an ODE solver's drift call passes:
 * `th.ones(x.size(0)) * t`, integrators.py:105): one timestep-embedding row instead of one per conditioning row.
 * Used by the adaptive dopri5 sampler, whose evaluation times are not known in advance.                          */
int scldm_dit_forward_shared_t(const scldm_dit_weights* w, const scldm_dit_plan* plan, const float* x, float t, float* v_out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Replaces Sampler.sample_ode(...)(x, model) -> [-1] for fixed-grid solvers
 * (transport.py:324-369, integrators.py:100-112 + torchdiffeq fixed-grid step).  x is advanced in
 * place over t_grid_host[0..n_grid) (n_grid points => n_grid-1 steps); all rows share t.         */
int scldm_dit_sample_ode(const scldm_dit_weights* w, const scldm_dit_plan* plan, float* x, const float* t_grid_host,
                         int32_t n_grid, int32_t method, void* workspace, size_t workspace_bytes, void* stream);

/* Packed VAE decoder weights (device, fp32).  scldm_b200/pack.py documents every layout. */
typedef struct scldm_vae_dec_weights {
  int32_t n_layer;
  int32_t n_ids;           /* n_genes + 1 rows in the embedding / theta tables */
  float eps;
  const float* win_t;      /* decoder_latent_input.1.weight^T [16][32]                       */
  const float* blocks;     /* n_layer packed Blocks (decoder.decoder_layers.i)                */
  const float* ca_ln1_w;   /* decoder_cross_attention.ln_1                                    */
  const float* ca_ln1_b;
  const float* ca_wkv_t;   /* decoder_cross_attention.attn.c_attn.weight^T [32][64]           */
  const float* ca_ln1q_w;  /* decoder_cross_attention.ln_1q                                   */
  const float* ca_ln1q_b;
  const float* ca_wq;      /* decoder_cross_attention.attn.c_attn_q.weight [32][32]           */
  const float* mcab_blob;  /* c_proj | ln_2 | mlp.w1 | mlp.w2 | mlp.c_proj^T | head w | head b | theta-head w | theta-head b */
  const float* emb;        /* input_layer.gene_embedding.weight [n_ids][32]                   */
  const float* theta_tbl;  /* decoder_head.theta.weight [n_ids]; NULL = unshared-theta head   */
  const void* mcab_wfrag;  /* bf16 MCAB weights in mma.sync B-fragment order (80 x 64 u32)    */
  const float* mcab_small; /* ln_2.weight[32] | ln_2.bias[32] | head w[32] | head b | pad[3] | theta-head w[32] | b | pad[3] */
} scldm_vae_dec_weights;

#define SCLDM_DECODE_TC 0   /* MCAB on tensor cores (bf16 operands, fp32 accumulate): default   */
#define SCLDM_DECODE_FP32 1 /* MCAB in fp32 on CUDA cores (exact variant)                       */

/* Cell-invariant query side of the decoder MCAB: qp[g] = c_attn_q(ln_1q(emb[g])) (layers.py:253,326).
 * qp [n_ids][32] fp32 and qp_bf16 [n_ids][32] bf16 (either may be NULL).                              */
int scldm_vae_qside(const scldm_vae_dec_weights* w, float* qp, void* qp_bf16, void* stream);

size_t scldm_vae_decode_workspace_bytes(int32_t n_cells, int32_t n_genes);

/* Replaces TransformerVAE.decode (vae.py:71-87) [+ NegativeBinomial.sample, models.py:819].
 *   z [n_cells][16][16]; genes [n_genes] int64 vocabulary ids shared by all cells; lib [n_cells]
 *   mu [n_cells][n_genes] / theta [n_genes] / counts [n_cells][n_genes]: any may be NULL.
 *   Unshared-theta head (theta_tbl == NULL; stochastic_layers.py:111-113): theta is [n_cells][n_genes] and required. */
int scldm_vae_decode(const scldm_vae_dec_weights* w, const float* qp, const void* qp_bf16, const float* z, int32_t n_cells,
                     const int64_t* genes, int32_t n_genes, const float* lib, float* mu, float* theta, float* counts,
                     uint64_t seed, int64_t cell_offset, int32_t precision, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Packed VAE encoder weights (device).  scldm_b200/pack.py::PackedVAEEncoder. */
typedef struct scldm_vae_enc_weights {
  int32_t n_layer;
  int32_t has_pos;
  int32_t agg_func;       /* count transform of InputTransformerVAE (layers.py:28-44): token = emb[gene] * f(count) with
                             0: log1p(c)   1: c == 0 ? -1 : log1p(c) ("log1pzero")   2: asinh(sqrt(c + 1)) ("anscombe")   3: sqrt(c + 1) */
  float eps;
  const float* emb;       /* input_layer.gene_embedding.weight [n_ids][32]                                  */
  const void* wkv_frag;   /* encoder.ca_layer.attn.c_attn.weight (k|v) in mma.sync B-fragment order          */
  const void* q_tbl;      /* bf16 [16][32]: c_attn_q(ln_1q(inducing_points)) -- cell invariant (layers.py:312-313) */
  const float* ln1_w;     /* encoder.ca_layer.ln_1                                                          */
  const float* ln1_b;
  const float* inducing;  /* encoder.ca_layer.inducing_points [16][32]                                      */
  const float* wproj_t;   /* encoder.ca_layer.attn.c_proj.weight^T [32][32]                                 */
  const float* ln2_w;
  const float* ln2_b;
  const float* w1_t;      /* encoder.ca_layer.mlp.w1.weight^T [32][88]                                      */
  const float* w2_t;
  const float* w3_t;      /* encoder.ca_layer.mlp.c_proj.weight^T [88][32]                                  */
  const float* pos;       /* encoder.pos_embed [16][32]                                                     */
  const float* blocks;    /* n_layer packed Blocks (encoder.encoder_layers.i)                               */
  const float* wlat_t;    /* encoder.encoder_latent_input.0.weight^T [32][16]                               */
} scldm_vae_enc_weights;

/* Replaces TransformerVAE.encode (vae.py:58-69): genes_subset [n_cells][S] int64, counts_subset [n_cells][S] fp32
 * -> z [n_cells][16][16] fp32.  Padding tokens (id 0, count 0) are NOT masked, as in the reference.          */
int scldm_vae_encode(const scldm_vae_enc_weights* w, const int64_t* genes_subset, const float* counts_subset, int32_t n_cells,
                     int32_t seq_len, float* z, void* stream);

/* N(0,1) draws keyed by (seed, global cell index, element): latent noise (models.py:788) and the
 * size-factor normals (models.py:585-596).                                                        */
int scldm_randn_cells(float* out, int32_t n_cells, int32_t per_cell, uint64_t seed, int64_t cell_offset,
                      uint32_t stream_id, void* stream);

/* Device-side CSR of a dense (rows, G) fp32 count matrix: the arrays scipy.sparse.csr_matrix(dense) would hold
 * (reference: process_generation_output builds them on the host from the dense D2H copy, src/scldm/_utils.py:186-200).
 *   scldm_csr_count: row_nnz [rows] int32 scratch, indptr [rows+1] int64 -> the caller reads indptr[rows] (= nnz) to size
 *                    the outputs of
 *   scldm_csr_fill : indices [nnz] int32 (ascending within a row), data [nnz] fp32.                              */
int scldm_csr_count(const float* dense, int32_t rows, int32_t G, int32_t* row_nnz, int64_t* indptr, void* stream);
int scldm_csr_fill(const float* dense, int32_t rows, int32_t G, const int64_t* indptr, int32_t* indices, float* data, void* stream);

/* "expressed" tokenizer on the device (reference: tokenize_cells(sample_genes="expressed"), src/scldm/datamodule.py:708-731):
 * dense (rows, G) counts -> genes_subset / counts_subset (rows, S): per cell the expressed genes packed left in gene order
 * (token = gene_ids[g]), padded with mask_idx / 0; library [rows] = row sums.  *overflow (device int32, zeroed by the
 * caller) receives the largest expressed-gene count that exceeded S (the reference raises ValueError in that case).  */
int scldm_tokenize_expressed(const float* dense, int32_t rows, int32_t G, const int64_t* gene_ids, int32_t S, int64_t mask_idx,
                             int64_t* genes_subset, float* counts_subset, float* library, int32_t* overflow, void* stream);

/* NB reconstruction loss per cell: nll[r] = -sum_g log_nb_positive(x[r][g], mu[r][g], theta[..][g]) with eps = 1e-8
 * (reference: src/scldm/distributions.py:6-42, summed over genes as VAE.loss does, src/scldm/models.py:233-247).
 * theta_row_stride = G for a (rows, G) theta, 0 for one shared (G,) row.                                          */
int scldm_nb_nll(const float* x, const float* mu, const float* theta, int64_t theta_row_stride, int32_t rows, int32_t G, float* nll,
                 void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * DiT training step (reference: LatentDiffusion.training_step, src/scldm/models.py:634-666; Transport.training_losses,
 * src/scldm/transport/transport.py:110-150; DiT.forward in training mode, src/scldm/nnets.py:273-297; torch autograd;
 * torch.optim.AdamW, experiments/configs/model/ldm_base.yaml:36-40; gradient_clip_val, configs/training/default.yaml:15).
 *
 * All trainable DiT parameters live in ONE flat fp32 buffer (`params`), their gradients in a second one with the same
 * layout (`grads`) - the buffer DDP all-reduces.  The layout is ordered by when a gradient is complete during the backward
 * pass (final layer, blocks n_layer-1 .. 0, then the conditioning / input tensors), so contiguous ranges can be all-reduced
 * while earlier blocks are still being differentiated.  bf16 copies of the GEMM weights in UMMA tile order (`pk_*`) are
 * refreshed by the optimizer step.
 * --------------------------------------------------------------------------------------------------------------- */
#define SCLDM_MAX_LAYERS 32

typedef struct scldm_dit_train_layout {   /* element offsets into params / grads */
  int64_t w_qkv[SCLDM_MAX_LAYERS];   /* blocks.i.attn.c_attn.weight [768][256]           */
  int64_t b_qkv[SCLDM_MAX_LAYERS];   /* blocks.i.attn.c_attn.bias   [768]                */
  int64_t w_proj[SCLDM_MAX_LAYERS];  /* blocks.i.attn.c_proj.weight [256][256]           */
  int64_t b_proj[SCLDM_MAX_LAYERS];  /* blocks.i.attn.c_proj.bias   [256]                */
  int64_t w1[SCLDM_MAX_LAYERS];      /* blocks.i.mlp.w1.weight      [hidden][256]        */
  int64_t w2[SCLDM_MAX_LAYERS];      /* blocks.i.mlp.w2.weight      [hidden][256]        */
  int64_t w3[SCLDM_MAX_LAYERS];      /* blocks.i.mlp.c_proj.weight  [256][hidden]        */
  int64_t w_mod[SCLDM_MAX_LAYERS];   /* blocks.i.adaln_modulation.1.weight [1536][256]   */
  int64_t b_mod;                     /* all adaLN biases, contiguous: blocks 0..n_layer-1 [1536] each, then the final layer's [512] */
  int64_t w_mod_final;               /* final_layer.adaln_modulation.1.weight [512][256] */
  int64_t w_out, b_out;              /* final_layer.linear [16][256], [16]               */
  int64_t temb_w0, temb_b0, temb_w2, temb_b2;   /* t_embedder.mlp.{0,2} [256][256], [256] */
  int64_t w_in, b_in;                /* input_proj [256][16], [256]                      */
  int64_t class_tab[SCLDM_MAX_CLASSES];          /* class_embeddings.<name>.weight [(V+1)][256], sorted by class name */
  int64_t n_params;
} scldm_dit_train_layout;

typedef struct scldm_dit_train {
  int32_t n_layer, hidden, n_class;
  float eps;
  scldm_dit_train_layout off;
  float* params;       /* flat fp32 master weights                                            */
  float* grads;        /* flat fp32 gradients (same layout)                                   */
  const float* pos;    /* pos_embed [16][256] (requires_grad=False in the reference)           */
  /* bf16 GEMM weights as 128B-swizzled K-major tiles [n_tile][k_slab][256 x 64] (T = ceil(hidden/128)):               */
  const void* pk_qkv;  /* [n_layer][3][4]                                                      */
  const void* pk_proj; /* [n_layer][1][4]                                                      */
  const void* pk_w12;  /* [n_layer][T][4]   tile j rows 0-127 = w1[128j..], rows 128-255 = w2[128j..] (zero padded)     */
  const void* pk_w3;   /* [n_layer][1][2T]  K (= hidden) zero padded to 128 T                   */
  const void* pk_mod;  /* [6 n_layer + 2][4]                                                   */
} scldm_dit_train;

size_t scldm_dit_train_workspace_bytes(const scldm_dit_train* tr, int32_t n_cells);

/* DiT.forward in training mode with every activation the backward pass needs kept in `workspace`.
 *   x [n_cells][16][16] fp32, t [n_cells] fp32, cls_idx [n_class][n_cells] int32 embedding rows (label dropout already
 *   applied by the caller: dropped labels = the class's null row), v_out [n_cells][16][16].  n_cells % 8 == 0.          */
int scldm_dit_train_forward(const scldm_dit_train* tr, const float* x, const float* t, const int32_t* cls_idx, int32_t n_cells,
                            float* v_out, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the call above: dv [n_cells][16][16] = dLoss/dv_out.  Gradients are ACCUMULATED into tr->grads unless
 * zero_grads != 0 (then the buffer is cleared first); dx (nullable) receives dLoss/dx.
 * events: n_events cudaEvent_t handles; events[k] is recorded on `stream` once the backward of block ev_after_layer[k]
 * (and of everything after it in the network) has been enqueued - the gradients of the flat-buffer prefix ending with
 * that block are then final, so the caller can start their all-reduce on another stream.                               */
int scldm_dit_train_backward(const scldm_dit_train* tr, const float* dv, float* dx, int32_t n_cells, int32_t zero_grads,
                             void* const* events, const int32_t* ev_after_layer, int32_t n_events, void* workspace,
                             size_t workspace_bytes, void* stream);

/* Fused optimizer step over flat buffers: g *= grad_scale; clip by global norm (max_norm <= 0: off; Lightning /
 * torch.nn.utils.clip_grad_norm_ semantics); AdamW (torch.optim.AdamW: decoupled weight decay, bias correction with
 * `step` >= 1); then the bf16 packed copy pk[pk_dst[i]] of every parameter with pk_dst[i] >= 0 is refreshed.
 * scratch: 512 device floats (deterministic two-stage norm: identical clip coefficient on every data-parallel rank).  */
int scldm_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int32_t step, float max_norm, float grad_scale, float* scratch,
                     const int32_t* pk_dst, void* pk, void* stream);
/* pk[pk_dst[i]] = bf16(params[i]) (initial pack, after load_state_dict) */
int scldm_repack(const float* params, const int32_t* pk_dst, void* pk, int64_t n, void* stream);
/* ema += (1 - decay) * (params - ema)   (ema_pytorch update with the decay the caller scheduled) */
int scldm_ema_update(float* ema, const float* params, int64_t n, float decay, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * VAE training step, n_embed = 32 (reference: VAE.training_step, src/scldm/models.py:249-287; TransformerVAE.forward,
 * src/scldm/vae.py:29-56; VAE.loss, models.py:233-247; log_nb_positive, distributions.py:6-42; torch autograd through
 * Encoder / Decoder / CrossAttentionBlock / Block, nnets.py:82-208, layers.py:177-330; optimizer: AdamWLegacy,
 * optimizers.py:72-141 = scldm_adamw_step above with pk_dst = NULL).
 *
 * All trainable VAE parameters live in ONE flat fp32 buffer, gradients in a second one with the same layout.  Parameter
 * groups (element offsets, every tensor padded to a multiple of 4 floats):
 *   Block [12672]:               ln_1.weight 32 | ln_1.bias 32 | attn.c_attn.weight [96][32] | attn.c_proj.weight [32][32] |
 *                                ln_2.weight | ln_2.bias | mlp.w1.weight [88][32] | mlp.w2.weight [88][32] | mlp.c_proj.weight [32][88]
 *   CrossAttentionBlock [12736]: ln_1.w | ln_1.b | ln_1q.w | ln_1q.b | ln_2.w | ln_2.b | attn.c_attn.weight [64][32] (k | v) |
 *                                attn.c_attn_q.weight [32][32] | attn.c_proj.weight [32][32] | mlp.w1 | mlp.w2 | mlp.c_proj
 * Covered: bias = False, use_adaln = False, shared_theta = True, shared_embedding = True, the multiplicative agg_func
 * variants (vae_base.yaml).                                                                                        */
typedef struct scldm_vae_train {
  int32_t n_layer, n_ids, agg_func, has_pos;
  float eps;
  float* params;        /* flat fp32 parameters                                                            */
  float* grads;         /* flat fp32 gradients (same layout); may be NULL for forward-only calls             */
  int64_t emb;          /* input_layer.gene_embedding.weight [n_ids][32]                                    */
  int64_t theta;        /* decoder_head.theta.weight [n_ids]                                                */
  int64_t head_w;       /* decoder_head.params.weight [32]                                                  */
  int64_t head_b;       /* decoder_head.params.bias [1]                                                     */
  int64_t enc_ca;       /* encoder.ca_layer (CrossAttentionBlock group)                                     */
  int64_t dec_ca;       /* decoder.decoder_cross_attention (CrossAttentionBlock group)                      */
  int64_t inducing;     /* encoder.ca_layer.inducing_points [16][32]                                        */
  int64_t enc_blocks;   /* encoder.encoder_layers.0 .. n_layer-1, contiguous Block groups                   */
  int64_t dec_blocks;   /* decoder.decoder_layers.0 .. n_layer-1                                            */
  int64_t enc_lat;      /* encoder.encoder_latent_input.0.weight [16][32]                                   */
  int64_t dec_lat;      /* decoder.decoder_latent_input.1.weight [32][16]                                   */
  int64_t n_params;
  const float* pos;     /* encoder.pos_embed [16][32] (requires_grad = False, nnets.py:103-106) or NULL     */
} scldm_vae_train;

size_t scldm_vae_train_workspace_bytes(const scldm_vae_train* tr, int32_t n_cells, int32_t S, int32_t G);

/* One forward (+ backward) of the VAE on n_cells cells:
 *   genes_subset / counts_subset [n_cells][S]  encoder tokens ("expressed" packing, datamodule.py:708-731)
 *   genes [G] the shared gene-id row, counts [n_cells][G], library [n_cells]
 *   nll [n_cells] = -sum_g log_nb_positive (the per-cell term of VAE.loss); z_out [n_cells][16][16], mu_out [n_cells][G] nullable.
 * backward != 0: d(loss_scale * sum_cells nll) / d params is ACCUMULATED into tr->grads (cleared first if zero_grads != 0);
 * loss_scale = 1 / global batch gives the gradient of `recon_loss.sum(dim=1).mean()`.
 * exact != 0: the decoder MCAB GEMMs use 3 x TF32 (fp32-grade products) instead of TF32.
 * workspace: 256-byte aligned, scldm_vae_train_workspace_bytes.  Stream-ordered, no host sync.                      */
int scldm_vae_train_step(const scldm_vae_train* tr, const int64_t* genes_subset, const float* counts_subset, int32_t S, const int64_t* genes,
                         const float* counts, const float* library, int32_t n_cells, int32_t G, float loss_scale, int32_t backward,
                         int32_t zero_grads, int32_t exact, float* nll, float* z_out, float* mu_out, void* workspace, size_t workspace_bytes,
                         void* stream);

/* Backward of the LAST scldm_vae_train_step(backward = 0) on the same workspace / inputs, from caller-provided gradients of the NB
 * parameters (the torch.autograd bridge, scldm_b200/vae_training.py::differentiable_forward): dmu [n_cells][G] = dLoss/dmu,
 * dtheta [n_cells][G] = dLoss/dtheta of the theta row expanded over the cells (nullable).  Parameter gradients are accumulated into
 * tr->grads (cleared first if zero_grads != 0).                                                                     */
int scldm_vae_train_backward(const scldm_vae_train* tr, const int64_t* genes_subset, const float* counts_subset, int32_t S, const int64_t* genes,
                             const float* library, const float* dmu, const float* dtheta, int32_t n_cells, int32_t G, int32_t zero_grads,
                             int32_t exact, void* workspace, size_t workspace_bytes, void* stream);

/* Stand-alone access to the slab GEMM for unit tests: mode 0 forward  out[M][N]  = A[M][K] W[N][K]^T (+bias[N]),
 * mode 1 dgrad out[M][K] = dY[M][N] W[N][K], mode 2 wgrad out[N][K] += dY[M][N]^T A[M][K].  a_f32 / dy_f32 are fp32
 * row-major inputs (converted to bf16 slab tensors in `workspace`), w_packed = bf16 [N/256][K/64][256 x 64] tiles.
 * M % 128 == 0, N % 256 == 0, K % 256 == 0.                                                                            */
int scldm_test_gemm(int32_t mode, const float* a_f32, const float* dy_f32, const void* w_packed, const float* bias, int32_t M,
                    int32_t N, int32_t K, int32_t split_k, float* out, void* workspace, size_t workspace_bytes, void* stream);
/* unit-test probe of the small-mean NB sampler: k[i] = inverse CDF of NB(mu[i], theta[i]) at u[i] */
int scldm_test_nb_invert(const float* u, const float* mu, const float* theta, float* k, int32_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Census-scale VAE, n_embed = 256 (8 self-attention heads x 32, 4 cross heads x 64, SwiGLU hidden 684; bias = False,
 * use_adaln = False, shared theta): the same reference entry points as scldm_vae_decode / scldm_vae_encode
 * (TransformerVAE.decode, src/scldm/vae.py:71-87; .encode, vae.py:58-69; CrossAttentionBlock, layers.py:267-330) with every
 * MCAB contraction on tcgen05 (rows = gene tokens / count tokens) and the latent Blocks on the DiT block-stack kernel.
 * One struct describes either side (decoder: decoder.* / decoder_head.*, encoder: encoder.*); unused members are NULL.
 * --------------------------------------------------------------------------------------------------------------- */
typedef struct scldm_vae256_weights {
  int32_t n_layer;            /* latent Blocks                                                                       */
  int32_t n_ids;              /* rows of the gene-embedding table (n_genes + 1)                                        */
  int32_t mlp_tiles;          /* T = ceil(hidden / 128)                                                                */
  int32_t has_pos;            /* encoder: add pos_embed after the MCAB                                                 */
  float eps;
  float head_b;               /* decoder_head.params.bias                                                              */
  const float* emb;           /* input_layer.gene_embedding.weight [n_ids][256]                                        */
  scldm_dit_weights blocks;   /* the latent Blocks packed like DiT blocks (w_attn_stream, w_mlp_stream, zero biases)    */
  const float* blocks_mod;    /* [n_layer*1536 + 512]: per Block (ln_1.weight - 1 | ln_1.bias | 1 | ln_2.weight - 1 | ln_2.bias | 1) */
  const float* ln1_mod;       /* MCAB ln_1 as a modulation row [512]: (weight - 1 | bias)                               */
  const void* w_kv;           /* MCAB attn.c_attn.weight  (k | v)  bf16 tiles [2][4][256 x 64]                          */
  const float* ln1q_w;        /* decoder: ln_1q.weight / bias [256] (Q side, per vocabulary)                            */
  const float* ln1q_b;
  const void* w_q;            /* decoder: attn.c_attn_q.weight tiles [1][4]                                             */
  const void* w_proj;         /* MCAB attn.c_proj.weight tiles [1][4]                                                   */
  const float* ln2_w;         /* MCAB ln_2 [256]                                                                        */
  const float* ln2_b;
  const float* ln2_mod;       /* encoder: ln_2 as a modulation row [512]                                                */
  const void* w_12;           /* MCAB mlp [w1 | w2] tiles [T][4] (tile j rows 0-127 = w1[128j..], 128-255 = w2[128j..])  */
  const void* w_3;            /* encoder: mlp.c_proj tiles [1][2T]                                                      */
  const float* lat_w;         /* decoder_latent_input.1.weight [256][16]                                                */
  const float* head_w;        /* decoder_head.params.weight [256]                                                       */
  const float* head_v;        /* [128 T] = mlp.c_proj.weight^T head_w: the NB-head Linear folded through the last MLP projection */
  const float* theta_tbl;     /* decoder_head.theta.weight [n_ids]                                                      */
  const float* q_tbl;         /* encoder: c_attn_q(ln_1q(inducing_points)) [16][256]                                    */
  const float* inducing;      /* encoder: ca_layer.inducing_points [16][256]                                            */
  const float* pos;           /* encoder: pos_embed [16][256]                                                           */
  const float* ones;          /* [256] ones (unit gate of the MLP residual)                                             */
  const float* out_w;         /* encoder_latent_input.0.weight [16][256]                                                */
} scldm_vae256_weights;

/* qp_bf16 [n_ids][256] = c_attn_q(ln_1q(emb)) for the whole vocabulary (cell invariant; cache it per vocabulary) */
size_t scldm_vae256_qside_workspace_bytes(int32_t n_ids);
int scldm_vae256_qside(const scldm_vae256_weights* w, void* qp_bf16, void* workspace, size_t workspace_bytes, void* stream);
/* same contract as scldm_vae_decode (genes shared by all cells; mu / theta / counts nullable) */
size_t scldm_vae256_decode_workspace_bytes(int32_t n_cells, int32_t n_genes);
int scldm_vae256_decode(const scldm_vae256_weights* w, const void* qp_bf16, const float* z, int32_t n_cells, const int64_t* genes,
                        int32_t n_genes, const float* lib, float* mu, float* theta, float* counts, uint64_t seed, int64_t cell_offset,
                        void* workspace, size_t workspace_bytes, void* stream);
/* same contract as scldm_vae_encode */
size_t scldm_vae256_encode_workspace_bytes(int32_t n_cells, int32_t seq_len);
int scldm_vae256_encode(const scldm_vae256_weights* w, const int64_t* genes_subset, const float* counts_subset, int32_t n_cells,
                        int32_t seq_len, float* z, void* workspace, size_t workspace_bytes, void* stream);

/* ---- generation evaluation (src/scldm/evaluations.py) and the SDE sampler (transport/integrators.py:7-75) ------------------- */
/* out [4][nx][ny] fp32: per pair (i, j) of rows of x [nx][D] and y [ny][D]: sum x*y, sum |x - y|, sum |x + y|, sum min(x, y) -
 * the sufficient statistics of RBFKernel / BrayCurtisKernel / TanimotoKernel / RuzickaKernel (evaluations.py:10-69) and of the
 * torch.cdist cost matrix of `wasserstein` (evaluations.py:100). */
int scldm_pair_stats(const float* x, int32_t nx, const float* y, int32_t ny, int32_t D, float* out, void* stream);
/* n_iter Sinkhorn-Knopp iterations on the Gibbs kernel K [n][m] (POT sinkhorn_knopp: v = b / K^T u; u = a / K v), then
 * res[0] = <u K v, M> and res[1] = || v * K^T u - b ||_1 (res: 2 device floats, zeroed here). */
int scldm_sinkhorn(const float* K, const float* M, const float* a, const float* b, int32_t n, int32_t m, float* u, float* v, int32_t n_iter,
                   float* res, void* stream);
#define SCLDM_DIFFUSION_CONSTANT 0
#define SCLDM_DIFFUSION_SBDM 1
#define SCLDM_DIFFUSION_SIGMA 2
#define SCLDM_DIFFUSION_LINEAR 3
#define SCLDM_DIFFUSION_DECREASING 4
#define SCLDM_DIFFUSION_INCDEC 5
/* SDE drift of the Linear path with a velocity model: drift = v + D(t) (t v - x) / (1 - t)   (transport.py:231-233, path.py:52-95) */
int scldm_sde_drift(const float* v, const float* x, float t, int32_t diffusion_form, float diffusion_norm, float* drift, int64_t n, void* stream);
/* out = x + sqrt(2 D(t)) sqrt(dt) w; w = noise[n] when non-NULL, else Philox N(0,1) keyed by (seed, cell_offset + i / per_cell, i % per_cell, step) */
int scldm_sde_kick(const float* x, const float* noise, float t, float dt, int32_t diffusion_form, float diffusion_norm, uint64_t seed, int64_t cell_offset,
                   int32_t per_cell, uint32_t step, float* out, int64_t n, void* stream);
/* out = a + c1 d1 (+ c2 d2 when d2 != NULL) */
int scldm_axpy2(const float* a, float c1, const float* d1, float c2, const float* d2, float* out, int64_t n, void* stream);

/* Live per-kernel timing for bench.py: when enabled every launch is bracketed by CUDA events recorded on
 * `stream` (must be the stream the calls run on; disables CUDA-graph capturability while on).
 * scldm_prof_summary synchronises the device and writes "name count total_ms\n" lines.            */
void scldm_prof_enable(int32_t on, void* stream);
int32_t scldm_prof_summary(char* buf, int32_t cap);

/* Debug: when device_buf != NULL the four GEMM kernels of DiT block `layer` write clock64() phase stamps into
 * device_buf (4 regions of 2^17 int64: qkv, proj, mlp1, mlp2; 32 stamps per CTA).  NULL disables.      */
void scldm_debug_timeline(long long* device_buf, int32_t layer);

/* Runtime options (cross-checks / profiling; the defaults are the product path).  name: "mega" (1: one persistent block-stack kernel per
 * evaluation, 0: one kernel per block half), "pdl" (programmatic dependent launch), "mod_batch" (adaLN table of a whole solve from one
 * GEMM), "exp" (bit mask of dit_blocks_kernel micro-variants), "dec_cpb" / "dec_occ" (MCAB decode launch shape).  The SCLDM_<NAME>
 * environment variables give the initial values only.  scldm_get_option returns -1 for an unknown name. */
int scldm_set_option(const char* name, int32_t value);
int32_t scldm_get_option(const char* name);

/* kernels launched by this library since load (bench.py reports it as gpu_launches) */
uint64_t scldm_launch_count(void);
const char* scldm_last_error(void);
const char* scldm_version(void);

#ifdef __cplusplus
}
#endif
#endif
