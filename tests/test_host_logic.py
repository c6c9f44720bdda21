"""CPU tests of the host side: weight packing layout, CFG slot layout semantics, state_dict contract of the
drop-in modules, the C-ABI library (loads, exports every declared symbol), error behaviour without CUDA."""

import ctypes
import json
import os
import re

import pytest
import torch

from oracle import scldm_oracle as O
from oracle.make_golden import WEIGHT_SEED, dit_inputs, golden_cases
from scldm_b200 import _lib, pack, synthetic
from scldm_b200.config import DiTConfig, VAEConfig, swiglu_hidden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_swiglu_hidden():
    assert swiglu_hidden(256, 4) == 684 and swiglu_hidden(32, 4) == 88  # layers.py:165-167


def test_pack_roundtrip_and_swizzle_positions():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(700, 300, generator=g)
    p = pack.pack_kmajor_tiles(w, 256)
    assert p.shape == (3, 5, 256 * 64) and p.dtype == torch.bfloat16
    back = pack.unpack_kmajor_tiles(p, 256)
    assert torch.equal(back[:700, :300], w.to(torch.bfloat16))
    assert float(back[700:].abs().max()) == 0 and float(back[:, 300:].abs().max()) == 0
    # explicit address check of the 128B swizzle: element (r, c) of tile (nt, ks) sits at
    # r*64 + ((c//8) ^ (r&7))*8 + c%8   (elements; csrc/sm100.cuh swz_chunk_offset)
    for (r, c) in [(0, 0), (1, 0), (7, 63), (9, 17), (255, 8), (130, 40)]:
        nt, ks = 1, 2
        off = r * 64 + (((c // 8) ^ (r & 7)) * 8) + c % 8
        assert p[nt, ks, off] == w[nt * 256 + r, ks * 64 + c].to(torch.bfloat16)


def test_dit_pack_layout_cpu():
    """PackedDiT arranges [w1|w2] tiles and the stacked adaLN matrix as the kernels index them."""
    cfg = DiTConfig(class_vocab_sizes={"a": 3}, n_layer=2)
    sd = synthetic.dit_state_dict(cfg, 1)
    pk = pack.PackedDiT(sd, cfg, "cpu")
    assert pk.mod_stride == 2 * 1536 + 512 and pk.hid_slabs == 11 and pk.mlp1_tiles == 6
    mod = pack.unpack_kmajor_tiles(pk.w_mod, 256)
    assert torch.equal(mod[1536:3072], sd["blocks.1.adaln_modulation.1.weight"].to(torch.bfloat16))
    assert torch.equal(mod[3072:3584], sd["final_layer.adaln_modulation.1.weight"].to(torch.bfloat16))
    # MLP stream of layer 1: M1_0 M1_1 M2_0 M1_2 M2_1 ... ; full [w1|w2] tiles are 4 x (256 x 64), the last one 4 x (2 hid_last x 64)
    assert pk.hid_last == 64
    st = pk.w_mlp_stream[1]
    tile, slab = 4 * 256 * 64, 256 * 64
    m1_1 = pack.unpack_kmajor_tiles(st[tile: 2 * tile].view(1, 4, slab), 256)                 # M1_1 follows M1_0
    assert torch.equal(m1_1[:128], (0.5 * sd["blocks.1.mlp.w1.weight"][128:256]).to(torch.bfloat16))   # stored halved (silu_from_half)
    assert torch.equal(m1_1[128:], sd["blocks.1.mlp.w2.weight"][128:256].to(torch.bfloat16))
    m2_0 = pack.unpack_kmajor_tiles(st[2 * tile: 2 * tile + 2 * slab].view(1, 2, slab), 256)  # then the two c_proj slabs of chunk 0
    assert torch.equal(m2_0, sd["blocks.1.mlp.c_proj.weight"][:, :128].to(torch.bfloat16))
    off = 5 * tile + 8 * slab                                                                  # M1_0..M1_4 and M2_0..M2_3 precede M1_5
    last = pack.unpack_kmajor_tiles(st[off: off + 4 * 128 * 64].view(1, 4, 128 * 64), 128)    # [w1 64 | w2 64] rows
    assert torch.equal(last[:44], (0.5 * sd["blocks.1.mlp.w1.weight"][640:684]).to(torch.bfloat16))
    assert float(last[44:64].abs().max()) == 0 and float(last[64 + 44:].abs().max()) == 0
    assert torch.equal(last[64:64 + 44], sd["blocks.1.mlp.w2.weight"][640:684].to(torch.bfloat16))
    assert st.numel() == off + 4 * 128 * 64 + 3 * slab                                          # + M2_4 (2 slabs) + M2_5 (1 slab)


def test_attention_stream_pack_layout_cpu():
    """The fused attention-block kernel walks one linear weight stream: per head pair hp four 192x64 [Wq|Wk|Wv] slabs
    (Q items) and two 128x64 c_proj slabs (P items), in the MMA issue order Q0 Q1 P0 Q2 P1 Q3 P2 P3; the v bias is
    folded into the c_proj bias (rows of softmax sum to 1) and the k bias is dropped (cancels in the softmax)."""
    cfg = DiTConfig(class_vocab_sizes={"a": 3}, n_layer=2)
    sd = synthetic.dit_state_dict(cfg, 1)
    pk = pack.PackedDiT(sd, cfg, "cpu")
    stream = pk.w_attn_stream[1]                       # layer 1
    assert stream.numel() == 4 * 256 * 256
    q_item, p_item = 192 * 64, 128 * 64
    order = [("q", 0), ("q", 1), ("p", 0), ("q", 2), ("p", 1), ("q", 3), ("p", 2), ("p", 3)]
    wqkv = sd["blocks.1.attn.c_attn.weight"].to(torch.bfloat16)
    wp = sd["blocks.1.attn.c_proj.weight"].to(torch.bfloat16)
    off = 0
    for kind, hp in order:
        if kind == "q":
            tile = pack.unpack_kmajor_tiles(stream[off: off + 4 * q_item].view(1, 4, q_item), 192)   # [192, 256]
            rows = torch.cat([torch.arange(part * 256 + 64 * hp, part * 256 + 64 * hp + 64) for part in range(3)])
            assert torch.equal(tile, wqkv[rows])
            off += 4 * q_item
        else:
            tile = pack.unpack_kmajor_tiles(stream[off: off + 2 * p_item].view(2, 1, p_item), 128)   # [256, 64]
            assert torch.equal(tile, wp[:, 64 * hp: 64 * hp + 64])
            off += 2 * p_item
    assert off == stream.numel()
    bqkv = sd["blocks.1.attn.c_attn.bias"].double()
    want = sd["blocks.1.attn.c_proj.bias"].double() + sd["blocks.1.attn.c_proj.weight"].double() @ bqkv[512:]
    assert torch.allclose(pk.b_proj_fused[1].double(), want, atol=1e-6)
    # attention is invariant to the k bias and shifts by the v bias: softmax(q(k+bk)^T)(v+bv) = softmax(q k^T) v + bv
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(16, 32, generator=g, dtype=torch.float64) for _ in range(3))
    bk, bv = torch.randn(32, generator=g, dtype=torch.float64), torch.randn(32, generator=g, dtype=torch.float64)
    att = lambda q, k, v: torch.softmax(q @ k.T / 32 ** 0.5, -1) @ v  # noqa: E731
    assert torch.allclose(att(q, k + bk, v + bv), att(q, k, v) + bv, atol=1e-12)


def test_vae_pack_layout_cpu():
    cfg = VAEConfig(n_genes=50, n_layer=2)
    sd = synthetic.vae_state_dict(cfg, 1)
    pk = pack.PackedVAEDecoder(sd, cfg, "cpu")
    assert pk.blocks.shape == (2, pack.VAE_BLOCK_STRIDE) and pk.mcab_blob.numel() == pack.MCAB_TOTAL
    b = pk.blocks[1]
    assert torch.equal(b[128:128 + 32 * 96].view(32, 96), sd["decoder.decoder_layers.1.attn.c_attn.weight"].T)
    assert torch.equal(pk.mcab_blob[1088:1088 + 88 * 32].view(88, 32), sd["decoder.decoder_cross_attention.mlp.w1.weight"])
    # constants duplicated in csrc/vae_kernels.cuh
    src = open(os.path.join(ROOT, "scldm_b200", "csrc", "vae_kernels.cuh")).read()
    assert "MW_LN2W = 1024" in src and "BLK_WQKV = 128" in src and "HID = 88" in src


def test_mma_b_fragment_layout():
    """pack.mma_b_frags: lane (g,t) of fragment (ks,nt) holds w[8nt+g][16ks+2t,+1] and w[8nt+g][16ks+2t+8,+9]."""
    g0 = torch.Generator().manual_seed(1)
    w = torch.randn(88, 32, generator=g0)
    f = pack.mma_b_frags(w)
    assert f.shape == (2, 11, 32, 4)
    wb = w.to(torch.bfloat16)
    for ks, nt, lane in [(0, 0, 0), (1, 10, 31), (0, 5, 13), (1, 3, 6)]:
        g, t = lane // 4, lane % 4
        r = 8 * nt + g
        exp = torch.stack([wb[r, 16 * ks + 2 * t], wb[r, 16 * ks + 2 * t + 1], wb[r, 16 * ks + 2 * t + 8], wb[r, 16 * ks + 2 * t + 9]])
        assert torch.equal(f[ks, nt, lane], exp)
    f3 = pack.mma_b_frags(torch.randn(32, 88, generator=g0))  # K = 88 -> padded to 96
    assert f3.shape == (6, 4, 32, 4) and float(f3[5, :, :, :].float().abs().sum()) > 0
    assert float(f3[5, 0, 2, 2:].float().abs().sum()) == 0  # lane t=2: columns 80+2t+8 >= 88 are padding


def test_drop_in_state_dict_contract(golden_dir):
    """the drop-in modules expose exactly the reference's state_dict keys/shapes and load them strictly."""
    from scldm_b200.nnets import DiT
    from scldm_b200.vae import TransformerVAE

    keys = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    for name, case in golden_cases().items():
        m = DiT(**case["cfg"].kwargs())
        assert {k: list(v.shape) for k, v in m.state_dict().items()} == keys[name]
        m.load_state_dict(synthetic.dit_state_dict(case["cfg"], WEIGHT_SEED), strict=True)
    for name, G in (("vae_small", 1500), ("vae_dentate", 17002)):
        cfg = VAEConfig(n_genes=G)
        v = TransformerVAE.from_config(cfg)
        assert {k: list(t.shape) for k, t in v.state_dict().items()} == keys[name]
        v.load_state_dict(synthetic.vae_state_dict(cfg, WEIGHT_SEED), strict=True)
        assert v.config() == cfg


def test_fresh_dit_init_matches_reference_scheme():
    """adaLN-zero + zero final layer (nnets.py:480-492), sincos pos_embed (sin first)."""
    from scldm_b200.nnets import DiT

    m = DiT(**DiTConfig(class_vocab_sizes={"c": 4}, n_layer=2).kwargs())
    assert float(m.blocks[1].adaln_modulation[1].weight.abs().max()) == 0
    assert float(m.final_layer.linear.weight.abs().max()) == 0
    assert torch.allclose(m.pos_embed[0], torch.from_numpy(synthetic.sincos_pos_embed(256, 16)))
    assert m.class_embeddings["c"].weight.shape == (5, 256)


@pytest.mark.parametrize("name", list(golden_cases().keys()))
@pytest.mark.parametrize("shared_time", [False, True])
def test_cfg_layout_reproduces_forward_with_cfg(golden_dir, name, shared_time):
    """Evaluate every slot of the batched CFG layout with the ORACLE's plain forward and combine with `coef`:
    must equal the oracle's (= the reference's) forward_with_cfg.  Pins the host-side batching logic."""
    from scldm_b200.nnets import DiT

    case = golden_cases()[name]
    cfg, B = case["cfg"], case["B"]
    sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
    x, t, labels = dit_inputs(name, cfg, B)
    if shared_time:
        t = torch.full_like(t, 0.3)
    m = DiT(**cfg.kwargs())
    lay = m.cfg_layout(labels, case["scales"], B, "cpu", shared_time)
    names = sorted(cfg.class_vocab_sizes)
    n_f, coef = lay["n_f"], lay["coef"]
    assert lay["slot_mod"].numel() == B + B * n_f

    def run_slot(state_row, mod_row):
        # labels of this conditioning row; a null index (== vocab) means "class absent" for the oracle
        lab = {}
        for ci, cname in enumerate(names):
            idx = int(lay["cls_idx"][ci, mod_row])
            if idx != cfg.class_vocab_sizes[cname]:
                lab[cname] = torch.tensor([idx])
        tt = t[state_row:state_row + 1] if lay["t_index"] is None else t[lay["t_index"][mod_row]].reshape(1)
        return O.dit_forward(x[state_row:state_row + 1], tt, lab, sd, cfg)[0]

    with torch.no_grad():
        ref = O.dit_forward_with_cfg(x, t, labels, case["scales"], sd, cfg)
        for s in (0, B - 1):  # unguided states
            out = run_slot(s, int(lay["slot_mod"][s]))
            assert torch.allclose(out, ref[s], atol=2e-5, rtol=1e-4)
        for j in (0, B - 1):  # guided states
            acc = 0
            for k in range(n_f):
                acc = acc + coef[k] * run_slot(B + j, int(lay["slot_mod"][B + j * n_f + k]))
            assert torch.allclose(acc, ref[B + j], atol=5e-5, rtol=1e-3)


def _mod_row(mode, slot, n_u, n_f, n_slots, table):
    """python mirror of csrc/dit_kernels.cuh::ModIndex::row"""
    if mode == 0:
        return int(table[slot])
    if slot >= n_slots:
        return 0
    if mode == 1:
        return slot
    if slot < n_u:
        return 0
    j, k = divmod(slot - n_u, n_f)
    return 0 if k == 0 else 1 + j * (n_f - 1) + (k - 1)


@pytest.mark.parametrize("name", list(golden_cases().keys()))
def test_slot_mode_closed_forms_match_tables(name):
    """the kernels may compute the conditioning row instead of loading it: the closed forms must equal the tables."""
    from scldm_b200.nnets import DiT
    from scldm_b200.ops import DitPlan

    case = golden_cases()[name]
    cfg, B = case["cfg"], case["B"]
    _, _, labels = dit_inputs(name, cfg, B)
    m = DiT(**cfg.kwargs())
    for shared in (False, True):
        lay = m.cfg_layout(labels, case["scales"], B, "cpu", shared)
        mode = DitPlan.SLOT_MODES[lay["slot_mode"]]
        n_slots = lay["slot_mod"].numel()
        for slot in range(n_slots):
            assert _mod_row(mode, slot, lay["n_u"], lay["n_f"], n_slots, lay["slot_mod"]) == int(lay["slot_mod"][slot])
        # every slot points at an existing row (the enumerated label table of a small label space may hold unused combinations)
        assert int(lay["slot_mod"].max()) + 1 <= lay["cls_idx"].shape[1]


@pytest.mark.parametrize("name", list(golden_cases().keys()))
def test_deduplicated_conditioning_rows_are_equivalent(name):
    """shared-time plans keep one adaLN row per distinct label combination: every slot must still see its own labels."""
    from scldm_b200.nnets import DiT

    case = golden_cases()[name]
    cfg, B = case["cfg"], case["B"]
    _, _, labels = dit_inputs(name, cfg, B)
    m = DiT(**cfg.kwargs())
    lay_d = m.cfg_layout(labels, case["scales"], B, "cpu", True)
    m.dedup_conditions = False
    lay_f = m.cfg_layout(labels, case["scales"], B, "cpu", True)
    assert lay_d["slot_mode"] == "table" and lay_f["slot_mode"] == "cfg_shared"
    per_slot_d = lay_d["cls_idx"][:, lay_d["slot_mod"].long()]
    per_slot_f = lay_f["cls_idx"][:, lay_f["slot_mod"].long()]
    assert torch.equal(per_slot_d, per_slot_f)
    # one row per label combination, never a duplicate: either the combinations that occur (torch.unique) or, for small label
    # spaces, the whole enumerated space (no device synchronisation while the plan is built)
    total = 1
    for v in cfg.class_vocab_sizes.values():
        total *= v + 1
    assert lay_d["cls_idx"].shape[1] <= max(lay_f["cls_idx"].shape[1], 1 + total)
    assert torch.unique(lay_d["cls_idx"][:, 1:], dim=1).shape[1] == lay_d["cls_idx"].shape[1] - 1


def test_abi_library_loads_and_exports_declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "scldm_b200.h")).read()
    declared = set(re.findall(r"\b(scldm_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"scldm_b200"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in lib.scldm_version()
    assert lib.scldm_vae_decode_workspace_bytes(4, 1000) > 4 * 1000 * 4
    # ctypes struct sizes match the C structs (8-byte pointers, natural alignment)
    assert ctypes.sizeof(_lib.DitWeights) == 7 * 4 + 4 + 18 * 8 + 8 * 8   # 7 ints + pad, 18 pointers, 8 class tables
    assert ctypes.sizeof(_lib.DitPlan) == 3 * 4 + 8 * 4 + 4 + 2 * 8 + 8


def test_no_cpu_fallback():
    """the product path fails loudly without CUDA instead of computing on the CPU."""
    from scldm_b200.nnets import DiT

    cfg = DiTConfig(class_vocab_sizes={"c": 4}, n_layer=1)
    m = DiT(**cfg.kwargs()).eval()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            m(torch.zeros(2, 16, 16), torch.zeros(2), {"c": torch.zeros(2, dtype=torch.long)})
    with pytest.raises(RuntimeError, match="weight container"):
        m.blocks[0](torch.zeros(1, 16, 256))


def test_product_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "scldm_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("with the oracle", ""), f


def test_abi_argument_errors_are_reported_not_undefined():
    """Error behaviour of the C-ABI (include/scldm_b200.h): bad arguments return SCLDM_EINVAL (-1) with a message in
    scldm_last_error(), before any CUDA work - so this runs without a GPU.  No exceptions cross the boundary."""
    lib = _lib.load()
    none = None
    calls = {
        "scldm_csr_count": lambda: lib.scldm_csr_count(none, 1, 1, none, none, none),
        "scldm_csr_fill": lambda: lib.scldm_csr_fill(none, 1, 1, none, none, none, none),
        "scldm_nb_nll": lambda: lib.scldm_nb_nll(none, none, none, 0, 1, 1, none, none),
        "scldm_tokenize_expressed": lambda: lib.scldm_tokenize_expressed(none, 1, 1, none, 1, 0, none, none, none, none, none),
    }
    for name, call in calls.items():
        rc = call()
        msg = lib.scldm_last_error().decode()
        assert rc == -1 and len(msg) > 0, (name, rc, msg)
    # theta stride shorter than a row is rejected (would read across rows)
    buf = (ctypes.c_float * 8)()
    rc = lib.scldm_nb_nll(buf, buf, buf, 2, 1, 4, buf, none)
    assert rc == -1 and "nb_nll" in lib.scldm_last_error().decode()
    # _lib.check turns a negative code into a RuntimeError that carries the library's message
    with pytest.raises(RuntimeError, match="nb_nll"):
        _lib.check(rc, "scldm_nb_nll")


@pytest.mark.parametrize("name", ["dit_me2", "dit_joint"])
def test_training_mode_label_dropout_matches_reference(golden_dir, name):
    """Training-mode CFG label dropout (`nnets.py:389-456`) is host logic here: under the same CPU seed the label rows the
    kernels would receive reproduce the summed class embedding the reference DiT builds in train mode (same sequence of
    `torch.randint` / `torch.rand` draws; the golden holds three seeds so both the class pick and the drop mask vary)."""
    import numpy as np

    from scldm_b200.nnets import DiT

    g = dict(np.load(os.path.join(golden_dir, "dit_label_dropout.npz")))
    cfg = golden_cases()[name]["cfg"]
    sd = synthetic.dit_state_dict(cfg, WEIGHT_SEED)
    m = DiT(**cfg.kwargs())
    m.load_state_dict(sd)
    m.train()
    labels = {k: torch.from_numpy(g[f"{name}.label.{k}"]) for k in cfg.class_vocab_sizes}
    n = 64
    seen = set()
    for rep in range(3):
        torch.manual_seed(1000 + rep)
        rows = m._cls_rows(m._active_labels(labels, force_drop_ids=True), n, "cpu")        # [n_class, n] table rows
        emb = sum(sd[f"class_embeddings.{cname}.weight"][rows[i].long()] for i, cname in enumerate(sorted(cfg.class_vocab_sizes)))
        assert torch.allclose(emb, torch.from_numpy(g[f"{name}.emb{rep}"]), atol=1e-6)
        seen.add(tuple(rows.reshape(-1).tolist()))
    assert len(seen) == 3                                     # the three seeds really gave different masks
    m.eval()
    if cfg.condition_strategy != "joint":
        with pytest.raises(AssertionError):
            m._active_labels(labels, force_drop_ids=True)     # the reference asserts this too (nnets.py:399-400)
    rows = m._cls_rows(m._active_labels(labels, force_drop_ids=False), n, "cpu")
    nulls = torch.tensor([cfg.class_vocab_sizes[c] for c in sorted(cfg.class_vocab_sizes)])
    assert int((rows != nulls[:, None]).any(1).sum()) == (len(labels) if cfg.condition_strategy == "joint" else 1)


def test_training_flat_layout_and_pack_map():
    """Host logic of `scldm_b200.training`: the flat buffer is ordered by gradient completion (final layer, blocks L-1..0, tail), the
    adaLN biases are contiguous in block order, and the bf16 tile arena is a pure gather of the flat buffer that reproduces
    `pack.pack_kmajor_tiles` of every GEMM weight."""
    import torch

    from scldm_b200 import synthetic
    from scldm_b200.config import DiTConfig
    from scldm_b200.pack import pack_kmajor_tiles
    from scldm_b200.training import flat_layout, pack_sources, wsd_schedule

    cfg = DiTConfig(class_vocab_sizes={"b": 3, "a": 5}, n_layer=3, condition_strategy="joint")
    sd = synthetic.dit_state_dict(cfg)
    shapes = {k: tuple(v.shape) for k, v in sd.items() if k != "pos_embed"}
    names, off, layer_end, n = flat_layout(cfg, shapes)
    assert set(names) == set(shapes) and len(names) == len(shapes)
    assert names[0].startswith("final_layer") and names.index("blocks.2.attn.c_attn.weight") < names.index("blocks.0.attn.c_attn.weight")
    assert layer_end[2] < layer_end[1] < layer_end[0] < n
    assert all(o % 4 == 0 for o in off.values())
    for l in range(3):
        assert off[f"blocks.{l}.adaln_modulation.1.bias"] == off["blocks.0.adaln_modulation.1.bias"] + l * 1536
    assert names.index("class_embeddings.a.weight") < names.index("class_embeddings.b.weight")   # sorted class names
    flat = torch.zeros(n)
    for k in names:
        flat[off[k]: off[k] + sd[k].numel()] = sd[k].reshape(-1)
    src, pko = pack_sources(cfg, off, shapes)
    pk = torch.where(src >= 0, flat[src.clamp_min(0)], torch.zeros(())).to(torch.bfloat16)
    H, T = cfg.hidden, -(-cfg.hidden // 128)
    ref_qkv = torch.stack([pack_kmajor_tiles(sd[f"blocks.{l}.attn.c_attn.weight"], 256) for l in range(3)]).reshape(-1)
    assert torch.equal(pk[pko["qkv"]: pko["qkv"] + ref_qkv.numel()], ref_qkv)
    w3 = torch.zeros(256, T * 128)
    w3[:, :H] = sd["blocks.1.mlp.c_proj.weight"]
    ref_w3 = pack_kmajor_tiles(w3, 256).reshape(-1)
    o = pko["w3"] + ref_w3.numel()
    assert torch.equal(pk[o: o + ref_w3.numel()], ref_w3)
    mod = torch.cat([sd[f"blocks.{l}.adaln_modulation.1.weight"] for l in range(3)] + [sd["final_layer.adaln_modulation.1.weight"]], 0)
    assert torch.equal(pk[pko["mod"]:], pack_kmajor_tiles(mod, 256).reshape(-1))
    # every GEMM weight element appears exactly once in the arena
    used = src[src >= 0]
    assert used.numel() == used.unique().numel()
    # the reference's LR schedule (scldm/_utils.py:19-60)
    f = wsd_schedule(1000, final_lr_factor=0.1, num_warmup_steps=100, init_div_factor=100, fract_decay=0.1)
    assert abs(f(0) - 0.01) < 1e-12 and f(100) == 1.0 and f(899) == 1.0 and 0.1 < f(950) < 1.0 and f(1000) == 0.1


def test_torch_library_ops_are_registered_with_fake_kernels():
    """`torch.ops.scldm_b200.*`: the hot-path entry points are real PyTorch operators with shape-propagating fake kernels (so
    `torch.compile` / `torch.export` can trace through them); the device kernels themselves need a GPU."""
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode

    from scldm_b200 import torch_ops  # noqa: F401

    for name in ("dit_forward", "dit_sample_ode", "vae_encode", "vae_decode"):
        assert hasattr(torch.ops.scldm_b200, name)
    with FakeTensorMode():
        mu, theta, counts = torch.ops.scldm_b200.vae_decode(torch.empty(5, 16, 16), torch.empty(700, dtype=torch.int64), torch.empty(5), 0, 0, 1)
        assert mu.shape == (5, 700) and theta.shape == (700,) and counts.shape == (5, 700)
        assert torch.ops.scldm_b200.vae_encode(torch.empty(3, 40, dtype=torch.int64), torch.empty(3, 40), 1).shape == (3, 16, 16)
        assert torch.ops.scldm_b200.dit_sample_ode(torch.empty(4, 16, 16), torch.empty(50), "euler", 1).shape == (4, 16, 16)
    with pytest.raises(RuntimeError, match="unknown handle"):
        torch.ops.scldm_b200.vae_encode(torch.zeros(1, 4, dtype=torch.int64), torch.zeros(1, 4), 987654)
