"""One census-vocabulary decode + encode of the n_embed = 256 VAE (for ncu): python tools/profile_v256.py [cells]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scldm_b200 import synthetic  # noqa: E402
from scldm_b200.config import VAEConfig  # noqa: E402
from scldm_b200.vae import TransformerVAE  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cfg = VAEConfig(n_genes=36130, n_embed=256)
vae = TransformerVAE.from_config(cfg)
vae.load_state_dict(synthetic.vae_state_dict(cfg, 1234))
vae = vae.cuda().eval()
z = torch.randn(B, 16, 16, device="cuda")
genes = torch.arange(1, 36131, device="cuda")
lib = torch.full((B,), 5000.0, device="cuda")
gs = torch.randint(1, 36131, (B, 8000), device="cuda")
cs = torch.ones(B, 8000, device="cuda")
for _ in range(2):
    vae.decode_counts(z, genes, lib, seed=1)
    vae.encode(None, None, cs, gs)
torch.cuda.synchronize()
