"""ctypes binding of the C-ABI in `include/scldm_b200.h` (libscldm_b200.so, built in-tree by
`scldm_b200/build.py`).  There is no fallback: if the library is missing the import of any
compute path raises."""

from __future__ import annotations

import ctypes as C
import os

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libscldm_b200.so")

MAX_CLASSES = 8
MAX_COMBINE = 8
ODE_METHODS = {"euler": 0, "heun2": 1, "midpoint": 2}
DECODE_PRECISION = {"bf16": 0, "fp32": 1}

EXPORTS = [
    "scldm_dit_slots_pad", "scldm_dit_mod_pad", "scldm_dit_workspace_bytes", "scldm_dit_workspace_layout",
    "scldm_dit_forward", "scldm_dit_forward_shared_t", "scldm_dit_sample_ode", "scldm_vae_qside", "scldm_vae_decode_workspace_bytes",
    "scldm_vae_decode", "scldm_vae_encode", "scldm_randn_cells", "scldm_csr_count", "scldm_csr_fill", "scldm_tokenize_expressed", "scldm_nb_nll", "scldm_dit_train_workspace_bytes", "scldm_dit_train_forward", "scldm_dit_train_backward", "scldm_adamw_step", "scldm_repack",
    "scldm_ema_update", "scldm_vae_train_workspace_bytes", "scldm_vae_train_step", "scldm_vae_train_backward", "scldm_vae256_qside_workspace_bytes", "scldm_vae256_qside", "scldm_vae256_decode_workspace_bytes", "scldm_vae256_decode",
    "scldm_vae256_encode_workspace_bytes", "scldm_vae256_encode", "scldm_pair_stats", "scldm_sinkhorn", "scldm_sde_drift", "scldm_sde_kick", "scldm_axpy2", "scldm_test_gemm", "scldm_test_nb_invert", "scldm_prof_enable", "scldm_prof_summary", "scldm_debug_timeline", "scldm_set_option", "scldm_get_option", "scldm_launch_count", "scldm_last_error", "scldm_version",
]


class DitWeights(C.Structure):
    _fields_ = [
        ("n_layer", C.c_int32), ("hidden", C.c_int32), ("hid_slabs", C.c_int32), ("mlp1_tiles", C.c_int32),
        ("mod_stride", C.c_int32), ("n_class", C.c_int32), ("eps", C.c_float),
        ("w_mod", C.c_void_p), ("b_mod", C.c_void_p), ("b_qkv", C.c_void_p),
        ("w_mlp_stream", C.c_void_p), ("w_attn_stream", C.c_void_p), ("b_proj_fused", C.c_void_p),
        ("temb_w0t", C.c_void_p), ("temb_b0", C.c_void_p), ("temb_w2t", C.c_void_p), ("temb_b2", C.c_void_p),
        ("w_in", C.c_void_p), ("b_in", C.c_void_p), ("pos", C.c_void_p), ("w_out", C.c_void_p), ("b_out", C.c_void_p),
        ("wout_frag", C.c_void_p), ("win_frag", C.c_void_p),
        ("class_tables", C.c_void_p * MAX_CLASSES),
        ("w_solve", C.c_void_p),
    ]


class DitPlan(C.Structure):
    _fields_ = [
        ("n_u", C.c_int32), ("n_g", C.c_int32), ("n_f", C.c_int32), ("coef", C.c_float * MAX_COMBINE),
        ("n_mod", C.c_int32), ("cls_idx", C.c_void_p), ("slot_mod", C.c_void_p), ("slot_mode", C.c_int32),
    ]


class VaeDecWeights(C.Structure):
    _fields_ = [
        ("n_layer", C.c_int32), ("n_ids", C.c_int32), ("eps", C.c_float),
        ("win_t", C.c_void_p), ("blocks", C.c_void_p), ("ca_ln1_w", C.c_void_p), ("ca_ln1_b", C.c_void_p),
        ("ca_wkv_t", C.c_void_p), ("ca_ln1q_w", C.c_void_p), ("ca_ln1q_b", C.c_void_p), ("ca_wq", C.c_void_p),
        ("mcab_blob", C.c_void_p), ("emb", C.c_void_p), ("theta_tbl", C.c_void_p),
        ("mcab_wfrag", C.c_void_p), ("mcab_small", C.c_void_p),
    ]


class VaeEncWeights(C.Structure):
    _fields_ = [("n_layer", C.c_int32), ("has_pos", C.c_int32), ("agg_func", C.c_int32), ("eps", C.c_float)] + [
        (n, C.c_void_p) for n in ("emb", "wkv_frag", "q_tbl", "ln1_w", "ln1_b", "inducing", "wproj_t", "ln2_w", "ln2_b", "w1_t", "w2_t",
                                  "w3_t", "pos", "blocks", "wlat_t")]


MAX_LAYERS = 32


class DitTrainLayout(C.Structure):
    _fields_ = [(n, C.c_int64 * MAX_LAYERS) for n in ("w_qkv", "b_qkv", "w_proj", "b_proj", "w1", "w2", "w3", "w_mod")] + [
        (n, C.c_int64) for n in ("b_mod", "w_mod_final", "w_out", "b_out", "temb_w0", "temb_b0", "temb_w2", "temb_b2", "w_in", "b_in")] + [
        ("class_tab", C.c_int64 * MAX_CLASSES), ("n_params", C.c_int64)]


class DitTrain(C.Structure):
    _fields_ = [("n_layer", C.c_int32), ("hidden", C.c_int32), ("n_class", C.c_int32), ("eps", C.c_float), ("off", DitTrainLayout),
                ("params", C.c_void_p), ("grads", C.c_void_p), ("pos", C.c_void_p),
                ("pk_qkv", C.c_void_p), ("pk_proj", C.c_void_p), ("pk_w12", C.c_void_p), ("pk_w3", C.c_void_p), ("pk_mod", C.c_void_p)]


class Vae256Weights(C.Structure):
    _fields_ = [("n_layer", C.c_int32), ("n_ids", C.c_int32), ("mlp_tiles", C.c_int32), ("has_pos", C.c_int32), ("eps", C.c_float), ("head_b", C.c_float),
                ("emb", C.c_void_p), ("blocks", DitWeights), ("blocks_mod", C.c_void_p), ("ln1_mod", C.c_void_p), ("w_kv", C.c_void_p),
                ("ln1q_w", C.c_void_p), ("ln1q_b", C.c_void_p), ("w_q", C.c_void_p), ("w_proj", C.c_void_p), ("ln2_w", C.c_void_p), ("ln2_b", C.c_void_p),
                ("ln2_mod", C.c_void_p), ("w_12", C.c_void_p), ("w_3", C.c_void_p), ("lat_w", C.c_void_p), ("head_w", C.c_void_p), ("head_v", C.c_void_p),
                ("theta_tbl", C.c_void_p), ("q_tbl", C.c_void_p), ("inducing", C.c_void_p), ("pos", C.c_void_p), ("ones", C.c_void_p), ("out_w", C.c_void_p)]


class VaeTrain(C.Structure):
    _fields_ = [("n_layer", C.c_int32), ("n_ids", C.c_int32), ("agg_func", C.c_int32), ("has_pos", C.c_int32), ("eps", C.c_float),
                ("params", C.c_void_p), ("grads", C.c_void_p)] + [
        (n, C.c_int64) for n in ("emb", "theta", "head_w", "head_b", "enc_ca", "dec_ca", "inducing", "enc_blocks", "dec_blocks", "enc_lat", "dec_lat",
                                 "n_params")] + [("pos", C.c_void_p)]


_lib = None


def load() -> C.CDLL:
    """Load libscldm_b200.so (once) and declare the prototypes of include/scldm_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "scldm_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    lib.scldm_dit_slots_pad.argtypes = [P(DitPlan)]
    lib.scldm_dit_slots_pad.restype = C.c_int32
    lib.scldm_dit_mod_pad.argtypes = [P(DitPlan)]
    lib.scldm_dit_mod_pad.restype = C.c_int32
    lib.scldm_dit_workspace_bytes.argtypes = [P(DitWeights), P(DitPlan), C.c_int32]
    lib.scldm_dit_workspace_bytes.restype = C.c_size_t
    lib.scldm_dit_workspace_layout.argtypes = [P(DitWeights), P(DitPlan), C.c_int32, P(C.c_size_t), C.c_int32]
    lib.scldm_dit_workspace_layout.restype = C.c_int32
    lib.scldm_dit_forward.argtypes = [P(DitWeights), P(DitPlan), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.scldm_dit_forward.restype = C.c_int
    lib.scldm_dit_forward_shared_t.argtypes = [P(DitWeights), P(DitPlan), C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.scldm_dit_forward_shared_t.restype = C.c_int
    lib.scldm_dit_sample_ode.argtypes = [P(DitWeights), P(DitPlan), C.c_void_p, P(C.c_float), C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_size_t, C.c_void_p]
    lib.scldm_dit_sample_ode.restype = C.c_int
    lib.scldm_vae_qside.argtypes = [P(VaeDecWeights), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scldm_vae_qside.restype = C.c_int
    lib.scldm_vae_decode_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.scldm_vae_decode_workspace_bytes.restype = C.c_size_t
    lib.scldm_vae_decode.argtypes = [P(VaeDecWeights), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_int32, C.c_void_p,
                                     C.c_size_t, C.c_void_p]
    lib.scldm_vae_decode.restype = C.c_int
    lib.scldm_vae_encode.argtypes = [P(VaeEncWeights), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.scldm_vae_encode.restype = C.c_int
    lib.scldm_randn_cells.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_int64, C.c_uint32, C.c_void_p]
    lib.scldm_randn_cells.restype = C.c_int
    lib.scldm_csr_count.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scldm_csr_count.restype = C.c_int
    lib.scldm_csr_fill.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scldm_csr_fill.restype = C.c_int
    lib.scldm_nb_nll.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.scldm_nb_nll.restype = C.c_int
    lib.scldm_tokenize_expressed.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p]
    lib.scldm_tokenize_expressed.restype = C.c_int
    lib.scldm_dit_train_workspace_bytes.argtypes = [P(DitTrain), C.c_int32]
    lib.scldm_dit_train_workspace_bytes.restype = C.c_size_t
    lib.scldm_dit_train_forward.argtypes = [P(DitTrain), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.scldm_dit_train_forward.restype = C.c_int
    lib.scldm_dit_train_backward.argtypes = [P(DitTrain), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, P(C.c_void_p), P(C.c_int32), C.c_int32,
                                             C.c_void_p, C.c_size_t, C.c_void_p]
    lib.scldm_dit_train_backward.restype = C.c_int
    lib.scldm_adamw_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_int32, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.scldm_adamw_step.restype = C.c_int
    lib.scldm_repack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.scldm_repack.restype = C.c_int
    lib.scldm_ema_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]
    lib.scldm_ema_update.restype = C.c_int
    lib.scldm_vae_train_workspace_bytes.argtypes = [P(VaeTrain), C.c_int32, C.c_int32, C.c_int32]
    lib.scldm_vae_train_workspace_bytes.restype = C.c_size_t
    lib.scldm_vae_train_step.argtypes = [P(VaeTrain), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p]
    lib.scldm_vae_train_step.restype = C.c_int
    lib.scldm_vae_train_backward.argtypes = [P(VaeTrain), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                             C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.scldm_vae_train_backward.restype = C.c_int
    lib.scldm_vae256_qside_workspace_bytes.argtypes = [C.c_int32]
    lib.scldm_vae256_qside_workspace_bytes.restype = C.c_size_t
    lib.scldm_vae256_qside.argtypes = [P(Vae256Weights), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.scldm_vae256_qside.restype = C.c_int
    lib.scldm_vae256_decode_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.scldm_vae256_decode_workspace_bytes.restype = C.c_size_t
    lib.scldm_vae256_decode.argtypes = [P(Vae256Weights), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint64, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.scldm_vae256_decode.restype = C.c_int
    lib.scldm_vae256_encode_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.scldm_vae256_encode_workspace_bytes.restype = C.c_size_t
    lib.scldm_vae256_encode.argtypes = [P(Vae256Weights), C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.scldm_vae256_encode.restype = C.c_int
    lib.scldm_pair_stats.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.scldm_pair_stats.restype = C.c_int
    lib.scldm_sinkhorn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.scldm_sinkhorn.restype = C.c_int
    lib.scldm_sde_drift.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]
    lib.scldm_sde_drift.restype = C.c_int
    lib.scldm_sde_kick.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_uint64, C.c_int64, C.c_int32, C.c_uint32, C.c_void_p,
                                   C.c_int64, C.c_void_p]
    lib.scldm_sde_kick.restype = C.c_int
    lib.scldm_axpy2.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.scldm_axpy2.restype = C.c_int
    lib.scldm_test_gemm.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]
    lib.scldm_test_gemm.restype = C.c_int
    lib.scldm_test_nb_invert.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.scldm_test_nb_invert.restype = C.c_int
    lib.scldm_prof_enable.argtypes = [C.c_int32, C.c_void_p]
    lib.scldm_prof_enable.restype = None
    lib.scldm_prof_summary.argtypes = [C.c_char_p, C.c_int32]
    lib.scldm_prof_summary.restype = C.c_int32
    lib.scldm_debug_timeline.argtypes = [C.c_void_p, C.c_int32]
    lib.scldm_debug_timeline.restype = None
    lib.scldm_set_option.argtypes = [C.c_char_p, C.c_int32]
    lib.scldm_set_option.restype = C.c_int
    lib.scldm_get_option.argtypes = [C.c_char_p]
    lib.scldm_get_option.restype = C.c_int32
    lib.scldm_launch_count.argtypes = []
    lib.scldm_launch_count.restype = C.c_uint64
    lib.scldm_last_error.argtypes = []
    lib.scldm_last_error.restype = C.c_char_p
    lib.scldm_version.argtypes = []
    lib.scldm_version.restype = C.c_char_p
    _lib = lib
    return lib


class ScldmError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().scldm_last_error().decode(errors="replace")
        raise ScldmError(f"{what} failed (code {rc}): {msg}")
