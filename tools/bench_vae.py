"""Census-vocabulary VAE throughput (BASELINE configs[4], inference part): MCAB encode and MCAB decode + NB head on one GPU.
G = 36 130 genes, S = 8 000 tokens per cell (`datamodule/default.yaml:130`).  Prints one JSON line.
Usage (GPU box): python tools/bench_vae.py [dataset] [cells] [--embed 32|256]
--embed 256 runs the census-scale VAE width (tensor-core MCAB, SURVEY.md 8d: 48.2 GFLOP per decoded cell at G = 36 130)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scldm_b200 import ops, synthetic
from scldm_b200.config import DATASETS, dataset_configs
from scldm_b200.vae import TransformerVAE

argv = list(sys.argv[1:])
EMBED = 32
if "--embed" in argv:
    i = argv.index("--embed")
    EMBED = int(argv[i + 1])
    del argv[i:i + 2]
dataset = argv[0] if len(argv) > 0 else "census"
B = int(argv[1]) if len(argv) > 1 else (1024 if EMBED == 32 else 128)
dev = torch.device("cuda:0")
_, vcfg = dataset_configs(dataset)
if EMBED != 32:
    from dataclasses import replace

    vcfg = replace(vcfg, n_embed=EMBED)
G, S = vcfg.n_genes, DATASETS[dataset]["genes_seq_len"]
vae = TransformerVAE.from_config(vcfg)
vae.load_state_dict(synthetic.vae_state_dict(vcfg, 1234))
vae = vae.to(dev).eval()
gen = torch.Generator().manual_seed(7)
# "expressed"-mode tokens: n_expr ~ U(0.05 G, 0.3 G) clipped to S, distinct gene ids packed left, counts 1 + Poisson(2)
counts_dense = torch.zeros(B, G)
for i in range(B):
    n = int(min(S, torch.randint(int(0.05 * G), int(0.3 * G), (1,), generator=gen)))
    idx = torch.randperm(G, generator=gen)[:n]
    counts_dense[i, idx] = 1.0 + torch.poisson(torch.full((n,), 2.0), generator=gen)
counts_dense = counts_dense.to(dev)
gene_row = torch.arange(1, G + 1, device=dev)
tok = ops.tokenize_expressed(counts_dense, gene_row, S)
gs, cs, lib = tok["genes_subset"], tok["counts_subset"], tok["library_size"]
z = torch.randn(B, 16, 16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    tot = 0.0
    for _ in range(steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / steps


ms_tok = timed(lambda: ops.tokenize_expressed(counts_dense, gene_row, S))
ms_enc = timed(lambda: vae.encode(None, None, cs, gs))
ms_dec = timed(lambda: vae.decode(z, gene_row.unsqueeze(0).expand(B, -1), lib))
counts_out = torch.empty(B, G, device=dev)
ms_decs = timed(lambda: vae.decode_counts(z, gene_row, lib.reshape(-1), seed=1, out_counts=counts_out))
mu_d = vae.decode(z, gene_row.unsqueeze(0).expand(B, -1), lib)
ms_nll = timed(lambda: ops.nb_nll(counts_dense, mu_d.mu, mu_d.theta))
ms_csr = timed(lambda: ops.counts_to_csr(counts_out))
E, H, M, L, NV = EMBED, vcfg.hidden, 16, 16, vcfg.n_layer
# decode per cell = M 2LE + M N_v (8E^2 + 6 E H + 4 M E) + M 4E^2 + G (4E^2 + 4ME + 6 E H + 2E)   (23 104 FLOP per gene token at E = 32)
dec_flop = M * 2 * L * E + M * NV * (8 * E * E + 6 * E * H + 4 * M * E) + M * 4 * E * E + G * (4 * E * E + 4 * M * E + 6 * E * H + 2 * E)
enc_flop = S * (4 * E * E + 4 * M * E) + M * (2 * E * E + 6 * E * H) + M * NV * (8 * E * E + 6 * E * H + 4 * M * E) + M * 2 * E * L
line = {
    "workload": f"{dataset}-vocabulary VAE: G={G}, S={S}, {B} cells, E={EMBED}, 8+8 layers (synthetic weights / tokens)",
    "decode_gflop_per_cell": round(dec_flop / 1e9, 2), "encode_gflop_per_cell": round(enc_flop / 1e9, 3),
    "tokenize_expressed_cells_per_s": round(B / ms_tok * 1e3), "encode_cells_per_s": round(B / ms_enc * 1e3),
    "decode_mu_theta_cells_per_s": round(B / ms_dec * 1e3), "decode_sample_counts_cells_per_s": round(B / ms_decs * 1e3),
    "encode_tflops": round(B * enc_flop / ms_enc / 1e9, 1), "decode_tflops": round(B * dec_flop / ms_dec / 1e9, 1),
    "decode_hbm_gbs_algorithmic": round(B * G * 8 / ms_decs / 1e6, 1),
    "nb_nll_cells_per_s": round(B / ms_nll * 1e3), "nb_nll_hbm_gbs": round(B * G * 8 / ms_nll / 1e6, 1),      # reads counts + mu (theta row is shared)
    "csr_cells_per_s": round(B / ms_csr * 1e3), "csr_hbm_gbs": round(B * G * 8 / ms_csr / 1e6, 1),            # two reads of the dense matrix (+ nnz writes)
    "decode_frac_of_bf16_peak": round(B * dec_flop / ms_dec / 1e9 / 1393.8, 4), "encode_frac_of_bf16_peak": round(B * enc_flop / ms_enc / 1e9 / 1393.8, 4),
    "note": "fractions vs the sustained cuBLAS bf16 rate of MEASURED_PEAKS.json (1393.8 TFLOP/s), algorithmic FLOPs of SURVEY 8(d) incl. the work the kernels skip (cached Q side, head folded through mlp.c_proj); CUDA events, L2 flushed between steps; decode_mu_theta returns the NB distribution (mu, theta), decode_sample_counts the Gamma-Poisson draw only",
}
print(json.dumps(line))
