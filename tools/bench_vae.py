"""Census-vocabulary VAE throughput (BASELINE configs[4], inference part): MCAB encode and MCAB decode + NB head on one GPU.
G = 36 130 genes, S = 8 000 tokens per cell (`datamodule/default.yaml:130`).  Prints one JSON line.
Usage (GPU box): python tools/bench_vae.py [dataset] [cells]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scldm_b200 import ops, synthetic
from scldm_b200.config import DATASETS, dataset_configs
from scldm_b200.vae import TransformerVAE

dataset = sys.argv[1] if len(sys.argv) > 1 else "census"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
_, vcfg = dataset_configs(dataset)
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
E, H, M = 32, 88, 16
dec_flop = 2.2e6 + G * 23104
enc_flop = S * (4 * E * E + 4 * M * E) + 2.0e6        # pooling over S tokens + tail blocks
line = {
    "workload": f"{dataset}-vocabulary VAE: G={G}, S={S}, {B} cells, E=32, 8+8 layers (synthetic weights / tokens)",
    "tokenize_expressed_cells_per_s": round(B / ms_tok * 1e3), "encode_cells_per_s": round(B / ms_enc * 1e3),
    "decode_mu_theta_cells_per_s": round(B / ms_dec * 1e3), "decode_sample_counts_cells_per_s": round(B / ms_decs * 1e3),
    "encode_tflops": round(B * enc_flop / ms_enc / 1e9, 1), "decode_tflops": round(B * dec_flop / ms_dec / 1e9, 1),
    "decode_hbm_gbs_algorithmic": round(B * G * 8 / ms_decs / 1e6, 1),
    "nb_nll_cells_per_s": round(B / ms_nll * 1e3), "nb_nll_hbm_gbs": round(B * G * 8 / ms_nll / 1e6, 1),      # reads counts + mu (theta row is shared)
    "csr_cells_per_s": round(B / ms_csr * 1e3), "csr_hbm_gbs": round(B * G * 8 / ms_csr / 1e6, 1),            # two reads of the dense matrix (+ nnz writes)
    "note": "CUDA events, L2 flushed between steps; decode_mu_theta returns the NB distribution (mu, theta), decode_sample_counts the Gamma-Poisson draw only",
}
print(json.dumps(line))
