"""Model / dataset shape descriptions for the scLDM generation hot path.

The dims mirror the only model configuration the reference ships
(`experiments/configs/model/ldm_base.yaml:15-28`, `vae_base.yaml:6-41`) and the
dataset table in `experiments/configs/datamodule/default.yaml:41-137`.
"""

from __future__ import annotations

from dataclasses import dataclass, field


def swiglu_hidden(n_embed: int, multiple_of: int) -> int:
    """Hidden width of the SwiGLU MLP (reference `layers.py:165-167`)."""
    hidden = int(2 * (n_embed * 4) / 3)
    return multiple_of * ((hidden + multiple_of - 1) // multiple_of)


@dataclass(frozen=True)
class DiTConfig:
    n_embed: int = 256
    n_embed_input: int = 16
    n_layer: int = 8
    n_head: int = 8
    seq_len: int = 16
    dropout: float = 0.0
    bias: bool = True
    norm_layer: str = "layernorm"
    multiple_of: int = 4
    layernorm_eps: float = 1e-8
    class_vocab_sizes: dict = field(default_factory=dict)
    cfg_dropout_prob: float = 0.8
    condition_strategy: str = "mutually_exclusive"

    @property
    def hidden(self) -> int:
        return swiglu_hidden(self.n_embed, self.multiple_of)

    def kwargs(self) -> dict:
        return dict(
            n_embed=self.n_embed,
            n_embed_input=self.n_embed_input,
            n_layer=self.n_layer,
            n_head=self.n_head,
            seq_len=self.seq_len,
            dropout=self.dropout,
            bias=self.bias,
            norm_layer=self.norm_layer,
            multiple_of=self.multiple_of,
            layernorm_eps=self.layernorm_eps,
            class_vocab_sizes=dict(self.class_vocab_sizes),
            cfg_dropout_prob=self.cfg_dropout_prob,
            condition_strategy=self.condition_strategy,
        )


@dataclass(frozen=True)
class VAEConfig:
    n_genes: int = 17002
    n_embed: int = 32
    n_embed_latent: int = 16
    n_inducing_points: int = 16
    n_layer: int = 8
    n_head: int = 8
    n_head_cross: int = 4
    dropout: float = 0.0
    bias: bool = False
    multiple_of: int = 4
    layernorm_eps: float = 1e-8
    norm_layer: str = "layernorm"
    positional_encoding: bool = True
    shared_embedding: bool = True
    use_adaln: bool = False
    shared_theta: bool = True
    agg_func: str = "log1p"

    @property
    def hidden(self) -> int:
        return swiglu_hidden(self.n_embed, self.multiple_of)

    def encoder_kwargs(self) -> dict:
        return dict(
            n_layer=self.n_layer,
            n_inducing_points=self.n_inducing_points,
            n_embed=self.n_embed,
            n_embed_latent=self.n_embed_latent,
            n_head=self.n_head,
            n_head_cross=self.n_head_cross,
            dropout=self.dropout,
            bias=self.bias,
            multiple_of=self.multiple_of,
            layernorm_eps=self.layernorm_eps,
            norm_layer=self.norm_layer,
            positional_encoding=self.positional_encoding,
        )

    def decoder_kwargs(self) -> dict:
        return dict(
            n_genes=self.n_genes,
            n_embed=self.n_embed,
            n_embed_latent=self.n_embed_latent,
            n_head=self.n_head,
            n_head_cross=self.n_head_cross,
            n_layer=self.n_layer,
            n_inducing_points=self.n_inducing_points,
            dropout=self.dropout,
            bias=self.bias,
            multiple_of=self.multiple_of,
            layernorm_eps=self.layernorm_eps,
            norm_layer=self.norm_layer,
            shared_embedding=self.shared_embedding,
            use_adaln=self.use_adaln,
        )

    def head_kwargs(self) -> dict:
        return dict(
            n_genes=self.n_genes,
            shared_theta=self.shared_theta,
            n_embed=self.n_embed,
            norm_layer=self.norm_layer,
            layernorm_eps=self.layernorm_eps,
        )

    def input_kwargs(self) -> dict:
        return dict(n_genes=self.n_genes, n_embed=self.n_embed, agg_func=self.agg_func)


# Dataset shapes: `datamodule/default.yaml:41-137` + `metadata/*.json` sizes.
DATASETS: dict[str, dict] = {
    "dentate_gyrus": dict(
        n_genes=17002, genes_seq_len=6147, class_vocab_sizes={"clusters": 14},
        guidance_weight={"clusters": 1.0}, condition_strategy="mutually_exclusive",
    ),
    "hlca": dict(
        n_genes=27997, genes_seq_len=10186, class_vocab_sizes={"cell_type": 50},
        guidance_weight={"cell_type": 1.0}, condition_strategy="mutually_exclusive",
    ),
    "tabula_muris": dict(
        n_genes=19734, genes_seq_len=9059, class_vocab_sizes={"tissue": 16},
        guidance_weight={"tissue": 1.0}, condition_strategy="mutually_exclusive",
    ),
    "parse1m": dict(
        n_genes=2000, genes_seq_len=2000, class_vocab_sizes={"cell_type": 18, "cytokine": 91},
        guidance_weight={"cell_type": 1.0, "cytokine": 1.0}, condition_strategy="joint",
    ),
    "replogle": dict(
        n_genes=2000, genes_seq_len=2000, class_vocab_sizes={"cell_line": 4, "gene": 2024},
        guidance_weight={"cell_line": 1.0, "gene": 1.0}, condition_strategy="joint",
    ),
    "census": dict(
        n_genes=36130, genes_seq_len=8000, class_vocab_sizes={},
        guidance_weight=None, condition_strategy="mutually_exclusive",
    ),
}


def dataset_configs(name: str) -> tuple[DiTConfig, VAEConfig]:
    d = DATASETS[name]
    dit = DiTConfig(class_vocab_sizes=dict(d["class_vocab_sizes"]), condition_strategy=d["condition_strategy"])
    vae = VAEConfig(n_genes=d["n_genes"])
    return dit, vae
