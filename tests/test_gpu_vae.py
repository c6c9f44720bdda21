"""GPU parity: fused VAE decoder (latent blocks -> MCAB -> NB head -> Gamma-Poisson) vs oracle / golden.

Two MCAB variants share every other kernel:
  precision="fp32": CUDA-core fp32 throughout      -> mu rel-L2 <= 1e-4
  precision="bf16": tensor cores (mma.sync), bf16 operands / fp32 accumulate / fp32 residual (default)
                                                    -> mu rel-L2 <= 1e-2 (measured 2.8e-3)
Both: |sum_g mu - library| / library <= 1e-4; theta (fp32 exp of a table) rel <= 1e-5.
Sampled counts: distributional agreement only (moments within Monte-Carlo error)."""

import os

import numpy as np
import pytest
import torch

from oracle import scldm_oracle as O
from oracle.make_golden import WEIGHT_SEED, vae_inputs
from scldm_b200 import synthetic
from scldm_b200.config import DiTConfig, VAEConfig

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


MU_TOL = {"fp32": 1e-4, "bf16": 1e-2}   # measured on B200: 8e-7 / 2.8e-3


def make_vae(cfg, precision="bf16"):
    from scldm_b200.vae import TransformerVAE

    vae = TransformerVAE.from_config(cfg)
    sd = synthetic.vae_state_dict(cfg, WEIGHT_SEED)
    vae.load_state_dict(sd, strict=True)
    vae.decode_precision = precision
    return vae.cuda().eval(), sd


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,G,B,S", [("vae_small", 1500, 3, 400), ("vae_dentate", 17002, 2, 600)])
def test_decode_vs_golden(golden_dir, name, G, B, S, precision):
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    cfg = VAEConfig(n_genes=G)
    vae, sd = make_vae(cfg, precision)
    z, genes, lib, _, _ = vae_inputs(name, cfg, B, S)
    nb = vae.decode(z.cuda(), genes.cuda(), lib.cuda())
    e_mu, e_th = rel_l2(nb.mu, g["mu"]), rel_l2(nb.theta[0], g["theta"])
    tot = (nb.mu.sum(1).cpu() / lib[:, 0] - 1).abs().max().item()
    big = torch.from_numpy(g["mu"]) >= 1e-3 * float(g["mu"].max())
    e_rel = float(((nb.mu.cpu() - torch.from_numpy(g["mu"])).abs() / torch.from_numpy(g["mu"]))[big].max())
    print(name, precision, f"mu rel-L2 {e_mu:.2e} max-rel(big) {e_rel:.2e} theta {e_th:.2e} |sum/lib-1| {tot:.2e}")
    assert e_mu < MU_TOL[precision] and e_th < 1e-5 and tot < 1e-4, (e_mu, e_th, tot)
    assert e_rel < 5 * MU_TOL[precision]
    assert nb.mu.shape == (B, G) and nb.theta.shape == (B, G)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decode_many_cells_vs_oracle(precision):
    """cells_per_block > 1 path, ragged last gene tile (G=1000 is not a multiple of 128)."""
    cfg = VAEConfig(n_genes=1000, n_layer=3)
    vae, sd = make_vae(cfg, precision)
    B = 700
    z = synthetic.randn("dm.z", (B, 16, 16))
    lib = torch.exp(8.0 + 0.3 * synthetic.randn("dm.lib", (B, 1)))
    genes = torch.arange(1, 1001).unsqueeze(0).expand(B, -1)
    nb = vae.decode(z.cuda(), genes.cuda(), lib.cuda())
    with torch.no_grad():
        mu, theta = O.vae_decode(z, genes, lib, sd, cfg)
    e = rel_l2(nb.mu, mu)
    print("many cells", precision, f"{e:.2e}")
    assert e < MU_TOL[precision]
    assert rel_l2(nb.theta, theta) < 1e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decode_gene_subset_and_order(precision):
    """gene ids are looked up, not assumed to be arange: a shuffled subset gives the matching columns' logits."""
    cfg = VAEConfig(n_genes=600, n_layer=2)
    vae, sd = make_vae(cfg, precision)
    B = 5
    z = synthetic.randn("gs.z", (B, 16, 16))
    lib = torch.full((B, 1), 1000.0)
    perm = torch.randperm(600, generator=torch.Generator().manual_seed(5))[:333] + 1
    genes = perm.unsqueeze(0).repeat(B, 1)
    nb = vae.decode(z.cuda(), genes.cuda(), lib.cuda())
    with torch.no_grad():
        mu, theta = O.vae_decode(z, genes, lib, sd, cfg)
    assert rel_l2(nb.mu, mu) < MU_TOL[precision] and rel_l2(nb.theta, theta) < 1e-5


def test_nb_sampling_moments():
    """counts ~ NB(mu, theta): per-gene mean and variance over many cells sharing one latent."""
    cfg = VAEConfig(n_genes=256, n_layer=1)
    vae, sd = make_vae(cfg)
    B = 20000
    z = synthetic.randn("nb.z", (1, 16, 16)).expand(B, -1, -1).contiguous()
    lib = torch.full((B,), 3000.0)
    genes = torch.arange(1, 257)
    counts, mu, theta = vae.decode_counts(z.cuda(), genes.cuda(), lib.cuda(), seed=11, cell_offset=0, want_mu=True)
    counts, mu, theta = counts.cpu().double(), mu[0].cpu().double(), theta.cpu().double()
    assert bool((counts >= 0).all()) and bool((counts == counts.round()).all())
    m, v = counts.mean(0), counts.var(0)
    exp_v = mu + mu**2 / theta
    # standard error of the mean ~ sqrt(var/B); allow 5 sigma, plus 8% on the variance
    assert bool(((m - mu).abs() <= 5 * (exp_v / B).sqrt() + 1e-3).all()), float(((m - mu).abs() / (exp_v / B).sqrt()).max())
    big = mu > 0.5
    assert float(((v[big] - exp_v[big]).abs() / exp_v[big]).median()) < 0.05
    assert float(((v[big] - exp_v[big]).abs() / exp_v[big]).max()) < 0.35
    # zero fraction for the NB: (theta/(theta+mu))^theta
    p0 = (theta / (theta + mu)) ** theta
    z0 = (counts == 0).double().mean(0)
    assert float((z0 - p0).abs().max()) < 0.02
    # determinism + offset sensitivity
    c2, _, _ = vae.decode_counts(z.cuda(), genes.cuda(), lib.cuda(), seed=11, cell_offset=0)
    c3, _, _ = vae.decode_counts(z.cuda(), genes.cuda(), lib.cuda(), seed=11, cell_offset=B)
    assert torch.equal(c2.cpu().double(), counts) and not torch.equal(c3.cpu().double(), counts)


def test_randn_cells_statistics_and_sharding_invariance():
    from scldm_b200 import ops

    a = ops.randn_cells(4096, 256, seed=5, cell_offset=0, stream_id=1, device="cuda")
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1) < 5e-3
    b = ops.randn_cells(1024, 256, seed=5, cell_offset=1024, stream_id=1, device="cuda")
    assert torch.equal(a[1024:2048], b)  # keyed by global cell index


def test_full_sample_vs_golden(golden_dir):
    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport

    g = dict(np.load(os.path.join(golden_dir, "sample_me1.npz")))
    dcfg = DiTConfig(class_vocab_sizes={"clusters": 14})
    vcfg = VAEConfig(n_genes=1500)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(dcfg, WEIGHT_SEED))
    vae, _ = make_vae(vcfg)
    ldm = LatentDiffusion(vae, dit.cuda().eval(), create_transport("Linear", "velocity"), sampling_method="euler", num_steps=50)
    B = 2
    genes = torch.arange(1, vcfg.n_genes + 1).unsqueeze(0).repeat(B, 1).cuda()
    counts, z, mu = ldm.sample({"clusters": torch.from_numpy(g["label"]).cuda()}, {"clusters": 2.0}, B, genes,
                               z0=torch.from_numpy(g["z0"]).cuda(), log_size_factors=torch.from_numpy(g["log_size_factors"]).cuda(),
                               return_mu=True)
    e_z, e_mu = rel_l2(z, g["z_final"]), rel_l2(mu, g["mu"])
    print(f"sample: z {e_z:.2e} mu {e_mu:.2e}")
    assert e_z < 5e-3 and e_mu < 1e-2, (e_z, e_mu)
    assert counts.shape == (2 * B, vcfg.n_genes) and z.shape == (2 * B, 16, 16)
    lib = torch.exp(torch.from_numpy(g["log_size_factors"]))
    assert torch.allclose(mu.sum(1).cpu(), torch.cat([lib, lib]), rtol=1e-4)


def test_sample_rng_path_and_size_factors():
    """device-drawn noise / size factors: shapes, determinism under a fixed seed, class-dependent library size."""
    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport

    dcfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=1)
    vcfg = VAEConfig(n_genes=500, n_layer=1)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(dcfg, WEIGHT_SEED))
    vae, _ = make_vae(vcfg)
    mu_t, sd_t = synthetic.size_factor_tables(dcfg.class_vocab_sizes)
    B = 64
    lab = {"clusters": synthetic.randint("rp.lab", 14, (B,)).cuda()}
    genes = torch.arange(1, 501).unsqueeze(0).repeat(B, 1).cuda()
    outs = []
    for _ in range(2):
        ldm = LatentDiffusion(vae, dit.cuda().eval(), create_transport("Linear", "velocity"), sampling_method="euler", mu_size_factor=mu_t,
                              sd_size_factor=sd_t, num_steps=5, seed=3, cell_chunk=24)
        outs.append(ldm.sample(lab, {"clusters": 1.0}, B, genes, return_mu=True))
    (c1, z1, m1), (c2, z2, m2) = outs
    assert torch.equal(c1, c2) and torch.equal(z1, z2)
    lib = m1.sum(1)[:B].log().cpu()
    want = torch.tensor([mu_t["clusters"][int(c)] for c in lab["clusters"].cpu()])
    assert float((lib - want).abs().max()) < 6 * 0.5  # within 6 sd of the class mean
    # chunking invariance: one big chunk gives the same cells
    ldm = LatentDiffusion(vae, dit, create_transport("Linear", "velocity"), sampling_method="euler", mu_size_factor=mu_t, sd_size_factor=sd_t,
                          num_steps=5, seed=3, cell_chunk=4096)
    c3, z3, _ = ldm.sample(lab, {"clusters": 1.0}, B, genes, return_mu=True)
    assert torch.equal(z3, z1) and torch.equal(c3, c1)


def test_shard_invariance_single_gpu():
    """generating cells [0,10) in one call == generating [0,6) and [6,10) as two 'ranks' with global offsets
    (what scldm_b200.dist.sample_sharded does per rank): RNG is keyed by the global cell index."""
    from scldm_b200.dist import shard_range
    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport

    dcfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=1)
    vcfg = VAEConfig(n_genes=300, n_layer=1)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(dcfg, WEIGHT_SEED))
    vae, _ = make_vae(vcfg)
    mu_t, sd_t = synthetic.size_factor_tables(dcfg.class_vocab_sizes)
    mk = lambda: LatentDiffusion(vae, dit.cuda().eval(), create_transport("Linear", "velocity"), sampling_method="euler", mu_size_factor=mu_t,  # noqa: E731
                                 sd_size_factor=sd_t, num_steps=4, seed=9)
    B = 10
    lab = {"clusters": synthetic.randint("si.lab", 14, (B,)).cuda()}
    genes = torch.arange(1, 301).unsqueeze(0).repeat(B, 1).cuda()
    c_all, z_all = mk().sample(lab, {"clusters": 1.5}, B, genes)
    parts = []
    for r in range(2):
        a, b = shard_range(B, r, 2)
        parts.append(mk().sample({"clusters": lab["clusters"][a:b]}, {"clusters": 1.5}, b - a, genes[a:b], cell_offset=a))
    n0 = parts[0][0].shape[0] // 2
    n1 = parts[1][0].shape[0] // 2
    c_cat = torch.cat([parts[0][0][:n0], parts[1][0][:n1], parts[0][0][n0:], parts[1][0][n1:]])
    z_cat = torch.cat([parts[0][1][:n0], parts[1][1][:n1], parts[0][1][n0:], parts[1][1][n1:]])
    assert torch.equal(z_cat, z_all) and torch.equal(c_cat, c_all)


@pytest.mark.parametrize("name,G,B,S", [("vae_small", 1500, 3, 400), ("vae_dentate", 17002, 2, 600)])
def test_encode_vs_golden(golden_dir, name, G, B, S):
    """MCAB encode (tensor-core flash pooling, bf16 operands) vs the reference: z is LayerNorm-ed (unit variance), so
    rel-L2 == RMS error; tolerance 2e-2.  Padding tokens (id 0, count 0) are part of the fixture (unmasked, quirk 3)."""
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    cfg = VAEConfig(n_genes=G)
    vae, sd = make_vae(cfg)
    _, _, _, cs, gs = vae_inputs(name, cfg, B, S)
    assert int((gs == 0).sum()) > 0  # the fixture really contains padding
    z = vae.encode(None, None, cs.cuda(), gs.cuda())
    e = rel_l2(z, g["z_enc"])
    print(name, f"encode z rel-L2 {e:.2e}")
    assert z.shape == (B, 16, 16) and e < 1e-2, e


@pytest.mark.parametrize("agg", ["log1pzero", "anscombe", "sqrt"])
def test_encode_count_transforms_vs_golden(golden_dir, agg):
    """agg_func variants of InputTransformerVAE (reference layers.py:28-44) in the encoder kernel vs reference-minted vectors; with
    'anscombe' / 'sqrt' the zero-count padding tokens carry f(0) != 0 and, unmasked as in the reference, do contribute."""
    g = dict(np.load(os.path.join(golden_dir, "vae_agg.npz")))
    cfg = VAEConfig(n_genes=1500, agg_func=agg)
    vae, sd = make_vae(cfg)
    z = vae.encode(None, None, torch.from_numpy(g["counts_subset"]).cuda(), torch.from_numpy(g["genes_subset"]).cuda())
    e = rel_l2(z, g["z_enc_" + agg])
    print(agg, f"encode z rel-L2 {e:.2e}")
    assert e < 1e-2, e


def test_encode_ragged_lengths_vs_oracle():
    """S not a multiple of 16 (ragged last token block), S < 128 (some warps idle), many cells."""
    cfg = VAEConfig(n_genes=700, n_layer=2)
    vae, sd = make_vae(cfg)
    for S, B in ((37, 5), (250, 33), (1000, 4)):
        gen = torch.Generator().manual_seed(S)
        gs = torch.randint(0, 701, (B, S), generator=gen)
        cs = torch.poisson(torch.full((B, S), 2.0), generator=gen) * (gs > 0)
        z = vae.encode(None, None, cs.cuda(), gs.cuda())
        with torch.no_grad():
            zo = O.vae_encode(cs, gs, sd, cfg)
        e = rel_l2(z, zo)
        print("encode", S, B, f"{e:.2e}")
        assert e < 1e-2, e


def test_joint_size_factors_match_oracle():
    """joint-key size-factor statistics (`models.py:498-550`, Replogle / Parse1M configs): the device lookup uses the same
    standard normals as the oracle's per-cell loop; missing joint keys / statistics give exactly 0."""
    from scldm_b200 import ops
    from scldm_b200.models import STREAM_SIZE_FACTOR, LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport

    dcfg = DiTConfig(class_vocab_sizes={"cell_line": 4, "gene": 7}, n_layer=1, condition_strategy="joint")
    vcfg = VAEConfig(n_genes=200, n_layer=1)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(dcfg, WEIGHT_SEED))
    vae, _ = make_vae(vcfg)
    j2c, mu_vec, sd_vec, c = {}, {}, {}, 0
    for i in range(4):
        for j in range(7):
            if (i + j) % 5 == 0:
                continue  # combination never seen in training: no joint key
            j2c[f"{i}_{j}"] = c
            if c % 6 != 1:  # a few classes lack statistics
                mu_vec[c], sd_vec[c] = 7.0 + 0.1 * c, 0.2 + 0.01 * c
            c += 1
    ldm = LatentDiffusion(vae, dit.cuda().eval(), create_transport("Linear", "velocity"), sampling_method="euler", mu_size_factor={"joint": mu_vec},
                          sd_size_factor={"joint": sd_vec}, joint_idx_2_classes=j2c, joint_key="joint",
                          joint_components=["cell_line", "gene"], num_steps=3, seed=5)
    B = 64
    cond = {"cell_line": synthetic.randint("jl.a", 4, (B,)), "gene": synthetic.randint("jl.b", 7, (B,))}
    lsf = ldm._sample_log_size_factors({k: v.cuda() for k, v in cond.items()}, B, cell_offset=100).cpu()
    eps = ops.randn_cells(B, 1, 5, 100, STREAM_SIZE_FACTOR, "cuda").reshape(-1).cpu()
    ref = O.sample_joint_log_size_factors(cond, ["cell_line", "gene"], j2c, mu_vec, sd_vec, B, eps)
    assert torch.allclose(lsf, ref, atol=1e-5) and int((ref == 0).sum()) > 0 and int((ref != 0).sum()) > 0
    counts, z = ldm.sample({k: v.cuda() for k, v in cond.items()}, {"cell_line": 1.0, "gene": 2.0}, B,
                           torch.arange(1, 201).unsqueeze(0).repeat(B, 1).cuda())
    assert counts.shape == (2 * B, 200) and bool(torch.isfinite(counts).all())


@pytest.mark.gpu
@pytest.mark.parametrize("rows,G,density", [(37, 17002, 0.08), (5, 1500, 0.5), (3, 129, 0.0), (64, 2000, 1.0), (1, 7, 0.3)])
def test_device_csr_matches_scipy(rows, G, density):
    """counts_to_csr builds exactly what scipy.sparse.csr_matrix(dense) holds (the reference's host-side step,
    _utils.py:186-200): bit-exact indptr / indices / data, including empty rows, ragged G and a full matrix."""
    from scipy import sparse

    from scldm_b200 import ops

    g = torch.Generator().manual_seed(rows * 1000 + G)
    dense = torch.poisson(torch.full((rows, G), 3.0), generator=g) + 1.0
    dense = dense * (torch.rand(rows, G, generator=g) < density)
    if rows > 2:
        dense[1] = 0.0                      # an all-zero row
    ref = sparse.csr_matrix(dense.numpy())
    indptr, indices, data = ops.counts_to_csr(dense.cuda())
    assert indptr.dtype == torch.int64 and indices.dtype == torch.int32 and data.dtype == torch.float32
    assert np.array_equal(indptr.cpu().numpy(), ref.indptr.astype(np.int64))
    assert np.array_equal(indices.cpu().numpy(), ref.indices.astype(np.int32))
    assert np.array_equal(data.cpu().numpy(), ref.data)


@pytest.mark.gpu
def test_sample_csr_equals_dense_sample():
    from scldm_b200.models import csr_to_scipy

    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport

    dcfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=1)
    vcfg = VAEConfig(n_genes=300, n_layer=1)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(dcfg, WEIGHT_SEED))
    vae, _ = make_vae(vcfg)
    mu_t, sd_t = synthetic.size_factor_tables(dcfg.class_vocab_sizes)
    ldm = LatentDiffusion(vae, dit.cuda().eval(), create_transport("Linear", "velocity"), sampling_method="euler", mu_size_factor=mu_t, sd_size_factor=sd_t,
                          num_steps=5, seed=3)
    B = 6
    lab = {"clusters": synthetic.randint("csr.lab", 14, (B,)).cuda()}
    genes = torch.arange(1, vcfg.n_genes + 1, device="cuda").unsqueeze(0).expand(B, -1)
    counts, z = ldm.sample(lab, {"clusters": 2.0}, B, genes, cell_offset=0)
    csr, z2 = ldm.sample_csr(lab, {"clusters": 2.0}, B, genes, cell_offset=0)
    m = csr_to_scipy(csr, vcfg.n_genes)
    assert torch.equal(z, z2)
    assert np.array_equal(m.toarray(), counts.cpu().numpy())
    assert m.nnz == int((counts != 0).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("rows,G,S,density", [(33, 17002, 6147, 0.2), (4, 300, 300, 1.0), (3, 129, 16, 0.0), (7, 2000, 512, 0.1)])
def test_device_tokenizer_matches_reference_expressed_mode(rows, G, S, density):
    """ops.tokenize_expressed == the numpy restatement of tokenize_cells(sample_genes="expressed") (datamodule.py:708-731):
    bit-exact token / count arrays (padding included) and row sums; too many expressed genes raises like the reference."""
    from scldm_b200 import ops

    g = torch.Generator().manual_seed(rows + G + S)
    dense = (torch.poisson(torch.full((rows, G), 2.0), generator=g) + 1.0) * (torch.rand(rows, G, generator=g) < density)
    gene_ids = torch.randperm(G, generator=g) + 1
    ref = O.tokenize_cells_expressed(dense.numpy(), gene_ids.numpy(), S, mask_idx=0)
    out = ops.tokenize_expressed(dense.cuda(), gene_ids.cuda(), S, mask_idx=0)
    assert np.array_equal(out["genes_subset"].cpu().numpy(), ref["genes_subset"])
    assert np.array_equal(out["counts_subset"].cpu().numpy(), ref["counts_subset"])
    assert np.allclose(out["library_size"].cpu().numpy(), ref["library_size"], rtol=1e-6)
    if density > 0:
        with pytest.raises(ValueError):
            ops.tokenize_expressed(dense.cuda(), gene_ids.cuda(), max(1, int((dense > 0).sum(1).max()) - 1))
        with pytest.raises(ValueError):
            O.tokenize_cells_expressed(dense.numpy(), gene_ids.numpy(), max(1, int((dense > 0).sum(1).max()) - 1))


def test_sample_streams_to_pinned_host_buffers():
    """`sample(host_out=...)`: rows are copied out piece by piece on a side stream behind their decode; the host buffers must
    hold exactly what the call returns on the device (several ODE chunks and ragged decode pieces), and bad buffers raise."""
    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport

    dcfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=1)
    vcfg = VAEConfig(n_genes=700, n_layer=1)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(dcfg, WEIGHT_SEED))
    vae, _ = make_vae(vcfg)
    mu_t, sd_t = synthetic.size_factor_tables(dcfg.class_vocab_sizes)
    B = 53
    lab = {"clusters": synthetic.randint("ho.lab", 14, (B,)).cuda()}
    genes = torch.arange(1, 701).unsqueeze(0).repeat(B, 1).cuda()
    mk = lambda: LatentDiffusion(vae, dit.cuda().eval(), create_transport("Linear", "velocity"), sampling_method="euler", mu_size_factor=mu_t,  # noqa: E731
                                 sd_size_factor=sd_t, num_steps=4, seed=11, cell_chunk=24)
    c_ref, z_ref = mk().sample(lab, {"clusters": 2.0}, B, genes)
    ldm = mk()
    ldm.decode_piece = 7
    counts_h = torch.full((2 * B, 700), -1.0).pin_memory()
    z_h = torch.full((2 * B, 16, 16), -1.0).pin_memory()
    c_dev, z_dev = ldm.sample(lab, {"clusters": 2.0}, B, genes, host_out=(counts_h, z_h))
    torch.cuda.current_stream().synchronize()
    assert torch.equal(c_dev, c_ref) and torch.equal(z_dev, z_ref)
    assert torch.equal(counts_h, c_ref.cpu()) and torch.equal(z_h, z_ref.cpu())
    with pytest.raises(ValueError):
        mk().sample(lab, {"clusters": 2.0}, B, genes, host_out=(torch.empty(2 * B, 700), z_h))          # not pinned
    with pytest.raises(ValueError):
        mk().sample(lab, {"clusters": 2.0}, B, genes, host_out=(counts_h[:B], z_h))                     # wrong shape


def test_nb_nll_and_vae_forward_vs_golden(golden_dir):
    """fused NB reconstruction loss and the inference `TransformerVAE.forward` against the reference's outputs
    (tests/golden/vae_loss_small.npz).  The loss kernel is fp32 throughout: per-cell sums agree to 1e-5 relative on the
    reference's own (mu, theta); through the bf16 tensor-core encode/decode the loss moves by < 1e-3 relative."""
    from scldm_b200 import ops

    g = dict(np.load(os.path.join(golden_dir, "vae_loss_small.npz")))
    counts, mu, theta = (torch.from_numpy(g[k]).cuda() for k in ("counts", "mu", "theta"))
    per_cell = ops.nb_nll(counts, mu, theta)                       # (B, G) theta
    assert rel_l2(per_cell, g["per_cell"]) < 1e-5
    per_cell_shared = ops.nb_nll(counts, mu, theta[0])             # theta is one row for shared_theta
    assert torch.allclose(per_cell_shared, per_cell, rtol=1e-6)
    odd = ops.nb_nll(counts[:, :1499].contiguous(), mu[:, :1499].contiguous(), theta[:, :1499].contiguous())   # G % 4 != 0: scalar path
    ref_odd = (-torch.from_numpy(g["log_nb"])[:, :1499]).sum(1)
    assert rel_l2(odd, ref_odd) < 1e-5
    cfg = VAEConfig(n_genes=1500)
    vae, _ = make_vae(cfg)
    B = counts.shape[0]
    genes = torch.arange(1, cfg.n_genes + 1).unsqueeze(0).repeat(B, 1).cuda()
    params, h_z = vae(counts, genes, torch.from_numpy(g["lib"]).cuda(), torch.from_numpy(g["counts_subset"]).cuda(),
                      torch.from_numpy(g["genes_subset"]).cuda())
    loss = vae.reconstruction_loss(counts, params)
    e_mu, e_z = rel_l2(params["mu"], g["mu"]), rel_l2(h_z, g["h_z"])
    e_l = abs(float(loss["llh"]) - float(g["llh"])) / abs(float(g["llh"]))
    print(f"vae forward: mu {e_mu:.2e} h_z {e_z:.2e} llh rel {e_l:.2e}")
    assert e_mu < 1.5e-2 and e_z < 1e-2 and e_l < 1e-3, (e_mu, e_z, e_l)
    with pytest.raises(RuntimeError):
        ops.nb_nll(counts.cpu(), mu.cpu(), theta.cpu())            # no CPU fallback


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decode_unshared_theta_vs_golden(golden_dir, precision):
    """NB head with `shared_theta=False` (`params` = Linear(E->2), theta = exp(second channel), stochastic_layers.py:106-113):
    mu and the per-cell theta (B, G) against the reference; the NB draw uses that per-cell theta."""
    g = dict(np.load(os.path.join(golden_dir, "vae_unshared_theta.npz")))
    cfg = VAEConfig(n_genes=1500, shared_theta=False)
    vae, _ = make_vae(cfg, precision)
    B = g["z"].shape[0]
    genes = torch.arange(1, cfg.n_genes + 1).unsqueeze(0).repeat(B, 1).cuda()
    lib = torch.from_numpy(g["lib"]).cuda()
    d = vae.decode(torch.from_numpy(g["z"]).cuda(), genes, lib)
    e_mu, e_th = rel_l2(d.mu, g["mu"]), rel_l2(d.theta, g["theta"])
    print(f"unshared theta [{precision}]: mu {e_mu:.2e} theta {e_th:.2e}")
    assert d.theta.shape == (B, cfg.n_genes)
    assert e_mu < MU_TOL[precision] and e_th < MU_TOL[precision]
    assert torch.allclose(d.mu.sum(1).cpu(), torch.from_numpy(g["lib"]).reshape(-1), rtol=1e-4)
    counts = d.sample()
    assert counts.shape == (B, cfg.n_genes) and bool((counts >= 0).all()) and bool((counts == counts.round()).all())
    # total count vs total mean within 5 standard deviations of the NB sum (var = mu + mu^2 / theta; theta spans 0.006 ... 112 here)
    sd_total = float((d.mu + d.mu**2 / d.theta).double().sum().sqrt())
    assert abs(float(counts.double().sum()) - float(d.mu.double().sum())) < 5 * sd_total
    # and the over-dispersed genes really are drawn with their own theta: zero fraction matches (theta/(theta+mu))^theta on average
    p0 = (d.theta / (d.theta + d.mu)).pow(d.theta).double().mean()
    assert abs(float((counts == 0).double().mean()) - float(p0)) < 0.03
    c2, _, _ = vae.decode_counts(torch.from_numpy(g["z"]).cuda(), genes, lib.reshape(-1), seed=3)
    assert c2.shape == (B, cfg.n_genes)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_decode_is_invariant_to_latent_token_order(precision):
    """The decoder treats the 16 latents as a SET (no positional encoding in `Decoder.forward`, nnets.py:200-208; self-attention and
    the MCAB softmax over keys are permutation-equivariant / -invariant): decoding a permuted latent set gives the same NB means.
    Checked at the full dentate_gyrus vocabulary (G = 17 002)."""
    cfg = VAEConfig(n_genes=17002)
    vae, _ = make_vae(cfg, precision)
    B = 24
    z = synthetic.randn("perm.z", (B, 16, 16)).cuda()
    lib = torch.exp(8.0 + 0.3 * synthetic.randn("perm.lsf", (B, 1))).cuda()
    genes = torch.arange(1, cfg.n_genes + 1).cuda().unsqueeze(0).expand(B, -1)
    perm = torch.from_numpy(np.random.default_rng(3).permutation(16)).cuda()
    mu_a = vae.decode(z, genes, lib).mu
    mu_b = vae.decode(z[:, perm].contiguous(), genes, lib).mu
    e = rel_l2(mu_b, mu_a)
    print(f"latent permutation [{precision}]: mu rel-L2 {e:.2e}")
    assert e < (1e-5 if precision == "fp32" else 5e-3)


def test_full_size_generation_properties():
    """BASELINE configs[1] at the bench's size (2368 cells -> 4736 rows, G = 17 002, 49-eval Euler + CFG): properties that do not need
    the CPU oracle - shapes, finiteness, integer counts, sum_g mu = library size per row, unconditional and guided halves share the
    library sizes, and chunking invariance of the latents (cells are independent, RNG keyed by the global cell index)."""
    import bench

    ldm, dcfg, vcfg = bench.build_models(torch.device("cuda"))
    B, G = 2368, vcfg.n_genes
    lab = {"clusters": torch.randint(0, 14, (B,), generator=torch.Generator().manual_seed(5)).cuda()}
    genes = torch.arange(1, G + 1).cuda().unsqueeze(0).expand(B, -1)
    ldm.cells_generated = 0
    counts, z, mu = ldm.sample(lab, {"clusters": 2.0}, B, genes, return_mu=True)
    assert counts.shape == (2 * B, G) and z.shape == (2 * B, 16, 16) and mu.shape == (2 * B, G)
    assert bool(torch.isfinite(z).all()) and bool(torch.isfinite(mu).all()) and bool((mu > 0).all())
    assert bool((counts >= 0).all()) and bool((counts == counts.round()).all())
    lib = mu.sum(1)
    assert torch.allclose(lib[:B], lib[B:], rtol=1e-4)                      # same library size for the unconditional / guided twin
    assert 6.0 < float(lib.log().mean()) < 10.0                             # synthetic tables: mu_c ~ U(7,9) on the log scale
    theta = torch.exp(ldm.vae_model.decoder_head.theta.weight.detach()[1:, 0]).double()
    sd_total = float((mu.double() + mu.double() ** 2 / theta[None, :]).sum().sqrt())
    assert abs(float(counts.double().sum()) - float(mu.double().sum())) < 5 * sd_total      # 80 M NB draws: total within 5 sd
    assert float((z[:B] - z[B:]).abs().max()) > 1e-3                        # guidance 2.0 moves the guided half
    assert 0.5 < float(z.std()) < 2.0
    ldm.cells_generated = 0
    ldm.cell_chunk = 592
    _, z2 = ldm.sample(lab, {"clusters": 2.0}, B, genes)
    assert torch.equal(z2, z)
    # 16 random cells (32 of the 4736 rows: unconditional + guided twins) against the fp32 oracle on the CPU, from the same noise and
    # size factors (re-derived from the Philox streams the call used)
    from scldm_b200 import ops
    from scldm_b200.models import STREAM_NOISE

    idx = torch.randperm(B, generator=torch.Generator().manual_seed(9))[:16]
    z0 = ops.randn_cells(B, 256, ldm.seed, 0, STREAM_NOISE, torch.device("cuda")).view(B, 16, 16)[idx.cuda()].cpu()
    lsf = ldm._sample_log_size_factors(lab, B, 0)[idx.cuda()].cpu()
    dsd, vsd = synthetic.dit_state_dict(dcfg, 1234), synthetic.vae_state_dict(vcfg, 1234)
    with torch.no_grad():
        mu_o, _, z_o = O.latent_diffusion_sample(z0, {"clusters": lab["clusters"][idx.cuda()].cpu()}, {"clusters": 2.0}, genes[:16].cpu(), lsf, dsd, dcfg, vsd, vcfg,
                                                 num_steps=50)
    rows = torch.cat([idx, B + idx]).cuda()
    e_z, e_mu = rel_l2(z[rows], z_o), rel_l2(mu[rows], mu_o)
    print(f"full size, 32 sampled rows vs oracle: z {e_z:.2e}, mu {e_mu:.2e}")
    assert e_z < 5e-3 and e_mu < 1e-2, (e_z, e_mu)


def test_default_solver_is_the_references_dopri5():
    """`LatentDiffusion.sample` of the reference always integrates with `sample_ode()`'s defaults (models.py:793: adaptive dopri5,
    50 points, atol = rtol = 1e-5).  A default-constructed drop-in does the same: its latents agree with a very fine fixed-grid
    Heun solve of the oracle far better than the 49-step Euler configuration does."""
    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport

    dcfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=2)
    vcfg = VAEConfig(n_genes=300, n_layer=1)
    dsd = synthetic.dit_state_dict(dcfg, WEIGHT_SEED)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(dsd)
    vae, _ = make_vae(vcfg)
    B = 3
    z0 = synthetic.randn("dd.z0", (B, 16, 16))
    lab = {"clusters": synthetic.randint("dd.lab", 14, (B,))}
    w = {"clusters": 1.5}
    genes = torch.arange(1, 301).unsqueeze(0).repeat(B, 1).cuda()
    lsf = torch.full((B,), 7.0)
    ldm = LatentDiffusion(vae, dit.cuda().eval(), create_transport("Linear", "velocity"))
    assert ldm.sampling_method == "dopri5" and ldm.num_steps == 50 and ldm.atol == 1e-5 and ldm.rtol == 1e-5
    _, z_def = ldm.sample({k: v.cuda() for k, v in lab.items()}, w, B, genes, z0=z0.cuda(), log_size_factors=lsf.cuda())
    nfe = ldm.transport_sampler.last_nfe
    ldm_e = LatentDiffusion(vae, dit, create_transport("Linear", "velocity"), sampling_method="euler")
    _, z_eul = ldm_e.sample({k: v.cuda() for k, v in lab.items()}, w, B, genes, z0=z0.cuda(), log_size_factors=lsf.cuda())
    lab2 = {"clusters": torch.cat([lab["clusters"], lab["clusters"]])}
    with torch.no_grad():
        ref = O.sample_ode(torch.cat([z0, z0]), lambda x, t: O.dit_forward_with_cfg(x, t, lab2, w, dsd, dcfg), num_steps=401, method="heun2")[-1]
    e_def, e_eul = rel_l2(z_def, ref), rel_l2(z_eul, ref)
    print(f"default (dopri5, nfe {nfe}) vs fine Heun: {e_def:.2e}; 49-step Euler vs fine Heun: {e_eul:.2e}")
    assert e_def < 5e-3 and nfe > 49


def test_unconditional_sample_without_labels():
    """`sample(condition=None, guidance_weight=None, ...)` (models.py:776-819 with no labels): size factors are zeros
    (library size 1), every forward is unconditional, so the two halves of the output carry the same latents."""
    from scldm_b200.models import LatentDiffusion
    from scldm_b200.nnets import DiT
    from scldm_b200.transport import create_transport

    dcfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=1)
    vcfg = VAEConfig(n_genes=200, n_layer=1)
    dsd = synthetic.dit_state_dict(dcfg, WEIGHT_SEED)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(dsd)
    vae, _ = make_vae(vcfg)
    B = 5
    genes = torch.arange(1, 201).unsqueeze(0).repeat(B, 1).cuda()
    z0 = synthetic.randn("un.z0", (B, 16, 16))
    ldm = LatentDiffusion(vae, dit.cuda().eval(), create_transport("Linear", "velocity"), sampling_method="euler", num_steps=6)
    counts, z, mu = ldm.sample(None, None, B, genes, z0=z0.cuda(), return_mu=True)
    assert counts.shape == (2 * B, 200) and torch.equal(z[:B], z[B:])
    assert torch.allclose(mu.sum(1), torch.ones(2 * B, device="cuda"), rtol=1e-4)      # exp(0) library size
    with torch.no_grad():
        z_o = O.sample_ode(torch.cat([z0, z0]), lambda x, t: O.dit_forward_with_cfg(x, t, None, None, dsd, dcfg), num_steps=6, method="euler")[-1]
    assert rel_l2(z, z_o) < 5e-3, rel_l2(z, z_o)
    with pytest.raises(ValueError):
        ldm.sample(None, None, B, genes[:2])                                           # genes batch dimension must match (models.py:777)


def test_torch_library_ops_run_the_same_kernels():
    """`torch.ops.scldm_b200.vae_decode / vae_encode / dit_sample_ode` (scldm_b200/torch_ops.py) return exactly what the Python API does."""
    from scldm_b200 import ops, torch_ops
    from scldm_b200.nnets import DiT

    cfg = VAEConfig(n_genes=500, n_layer=2)
    vae, _ = make_vae(cfg)
    B = 6
    z = synthetic.randn("to.z", (B, 16, 16)).cuda()
    lib = torch.full((B,), 2000.0).cuda()
    genes = torch.arange(1, 501).cuda()
    hd = torch_ops.register(vae.packed_decoder())
    mu, theta, counts = torch.ops.scldm_b200.vae_decode(z, genes, lib, 7, 0, hd)
    mu2, theta2, counts2 = ops.vae_decode(vae.packed_decoder(), z, genes, lib, want_mu=True, want_counts=True, seed=7, cell_offset=0)
    assert torch.equal(mu, mu2) and torch.equal(theta, theta2) and torch.equal(counts, counts2)
    gs = torch.randint(1, 501, (B, 64), generator=torch.Generator().manual_seed(1)).cuda()
    cs = torch.ones(B, 64).cuda()
    he = torch_ops.register(vae.packed_encoder())
    assert torch.equal(torch.ops.scldm_b200.vae_encode(gs, cs, he), vae.encode(None, None, cs, gs))
    dcfg = DiTConfig(class_vocab_sizes={"clusters": 14}, n_layer=2)
    dit = DiT(**dcfg.kwargs())
    dit.load_state_dict(synthetic.dit_state_dict(dcfg, WEIGHT_SEED))
    dit = dit.cuda().eval()
    lab = {"clusters": synthetic.randint("to.lab", 14, (2 * B,)).cuda()}
    plan, _ = dit.cfg_plan(lab, {"clusters": 2.0}, B, torch.device("cuda"), shared_time=True)
    hp = torch_ops.register(plan)
    x = torch.cat([z, z])
    grid = torch.linspace(0, 1, 6)
    assert torch.equal(torch.ops.scldm_b200.dit_sample_ode(x, grid, "euler", hp), ops.dit_sample_ode(plan, x.clone(), grid, "euler"))
    assert ops.get_option("solve") == 1 and ops.get_option("nonexistent") == -1
    torch_ops.release(hd), torch_ops.release(he), torch_ops.release(hp)
