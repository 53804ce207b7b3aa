"""DualAR model configuration.

Mirrors the field names, defaults and derivation rules of the reference's
``RQTransformerModelArgs`` / ``BaseModelArgs`` dataclasses
(reference: modeling/model/rq_transformer.py:25-114) so that a reference
``config.json`` loads unchanged.  Only the fields the decode path reads are
interpreted; unknown keys are preserved in ``extra`` and written back by
``save``.
"""
from __future__ import annotations

import dataclasses
import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, Optional


def _find_multiple(n: int, k: int) -> int:
    return n if n % k == 0 else n + k - (n % k)


@dataclass
class RQTransformerModelArgs:
    model_type: str = "dual_ar"

    vocab_size: int = 32000
    n_layer: int = 32
    n_head: int = 32
    dim: int = 4096
    intermediate_size: Optional[int] = 16384
    n_local_heads: int = -1
    head_dim: int = 64
    rope_base: float = 10000
    norm_eps: float = 1e-5
    max_seq_len: int = 2048
    dropout: float = 0.0
    tie_word_embeddings: bool = True
    attention_qkv_bias: bool = False

    codebook_size: int = 160
    num_codebooks: int = 4

    use_gradient_checkpointing: bool = False
    initializer_range: float = 0.02
    is_reward_model: bool = False
    share_codebook_embeddings: bool = True
    scale_codebook_embeddings: bool = False

    fast_dim: Optional[int] = 1024
    n_fast_layer: int = 4
    fast_n_head: Optional[int] = 16
    fast_n_local_heads: Optional[int] = None
    fast_head_dim: Optional[int] = None
    fast_intermediate_size: Optional[int] = None
    fast_attention_qkv_bias: Optional[bool] = None
    depthwise_wte: Optional[bool] = False
    depthwise_output: Optional[bool] = False
    duplicate_code_0: Optional[bool] = True

    extra: Dict[str, Any] = field(default_factory=dict, repr=False)

    def __post_init__(self) -> None:
        # reference: modeling/model/rq_transformer.py:58-65
        if self.n_local_heads == -1:
            self.n_local_heads = self.n_head
        if self.intermediate_size is None:
            self.intermediate_size = _find_multiple(int(2 * 4 * self.dim / 3), 256)
        self.head_dim = self.dim // self.n_head  # JSON value is overwritten (quirk 8g-5)
        # reference: modeling/model/rq_transformer.py:100-114
        self.fast_dim = self.fast_dim or self.dim
        self.fast_n_head = self.fast_n_head or self.n_head
        self.fast_n_local_heads = self.fast_n_local_heads or self.n_local_heads
        self.fast_head_dim = self.fast_head_dim or self.head_dim
        self.fast_intermediate_size = self.fast_intermediate_size or self.intermediate_size
        if self.fast_attention_qkv_bias is None:
            self.fast_attention_qkv_bias = self.attention_qkv_bias
        if self.duplicate_code_0 is None:
            self.duplicate_code_0 = True

    # ---- derived quantities used by the decode path -----------------------
    @property
    def max_fast_seqlen(self) -> int:
        """Depth-loop length (reference :344-346)."""
        return self.num_codebooks - (0 if self.duplicate_code_0 else 1)

    @property
    def n_rows(self) -> int:
        """Rows of the token grid: 1 text/semantic row + depth rows."""
        return 1 + self.max_fast_seqlen

    @property
    def qkv_rows(self) -> int:
        return (self.n_head + 2 * self.n_local_heads) * self.head_dim

    @property
    def fast_qkv_rows(self) -> int:
        return (self.fast_n_head + 2 * self.fast_n_local_heads) * self.fast_head_dim

    @property
    def fast_embedding_rows(self) -> int:
        return self.codebook_size * ((self.num_codebooks - 1) if self.depthwise_wte else 1)

    def layer_params(self, fast: bool = False) -> int:
        d = self.fast_dim if fast else self.dim
        f = self.fast_intermediate_size if fast else self.intermediate_size
        rows = self.fast_qkv_rows if fast else self.qkv_rows
        return rows * d + d * d + 3 * d * f

    def unique_weight_bytes(self) -> int:
        """W_unique of SURVEY §8(d): every weight byte a frame must touch, once."""
        head = self.vocab_size * self.dim
        depth_heads = (self.max_fast_seqlen if self.depthwise_output else 1) * self.fast_dim * self.codebook_size
        return 2 * (self.n_layer * self.layer_params() + head
                    + self.n_fast_layer * self.layer_params(True) + depth_heads)

    def kv_bytes_per_position(self) -> int:
        return 2 * self.n_local_heads * self.head_dim * 2 * self.n_layer

    def flops_per_frame(self, context: float) -> float:
        macs = (self.n_layer * self.layer_params() + self.vocab_size * self.dim
                + self.max_fast_seqlen * (self.n_fast_layer * self.layer_params(True)
                                          + self.fast_dim * self.codebook_size))
        attn = 4 * self.n_head * self.head_dim * self.n_layer * context
        return 2.0 * macs + attn

    # ---- (de)serialisation ----------------------------------------------------
    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "RQTransformerModelArgs":
        names = {f.name for f in dataclasses.fields(cls)} - {"extra"}
        known = {k: v for k, v in data.items() if k in names}
        extra = {k: v for k, v in data.items() if k not in names}
        return cls(**known, extra=extra)

    @classmethod
    def from_pretrained(cls, pathname: str) -> "RQTransformerModelArgs":
        # reference: modeling/model/rq_transformer.py:67-79
        path = Path(pathname)
        if path.is_dir():
            path = path / "config.json"
        with open(path, "r", encoding="utf-8") as f:
            return cls.from_dict(json.load(f))

    from_json_file = from_pretrained  # MLX twin's name (mlx lm/rq_transformer.py:45-48)

    def to_dict(self) -> Dict[str, Any]:
        d = dataclasses.asdict(self)
        extra = d.pop("extra")
        d.update(extra)
        return d

    def save(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(self.to_dict(), f, indent=4, sort_keys=True, ensure_ascii=False)


# The two configs BASELINE.json names (reference: sample_model_sizes/*.json) plus a
# tiny one for fast tests.  Values restated here because /root/reference does not
# travel to the GPU box.
_COMMON = dict(
    attention_qkv_bias=False, codebook_size=2048, dropout=0.1, fast_attention_qkv_bias=False,
    fast_head_dim=64, head_dim=64, initializer_range=1.0 / 24.0, is_reward_model=False,
    max_seq_len=2048, model_type="dual_ar", n_fast_layer=4, n_layer=10, depthwise_wte=True,
    depthwise_output=True, norm_eps=1e-5, num_codebooks=8, rope_base=100000,
    scale_codebook_embeddings=False, share_codebook_embeddings=True, tie_word_embeddings=True,
    use_gradient_checkpointing=True, vocab_size=2368,
)

MODEL_SIZES: Dict[str, Dict[str, Any]] = {
    "smoltts_byte_150m": dict(_COMMON, dim=768, fast_dim=768, intermediate_size=3072,
                              fast_intermediate_size=3072, n_head=12, fast_n_head=12,
                              n_local_heads=4, fast_n_local_heads=4),
    "smoltts_byte_70m": dict(_COMMON, dim=576, fast_dim=576, intermediate_size=1536,
                             fast_intermediate_size=1536, n_head=9, fast_n_head=9,
                             n_local_heads=3, fast_n_local_heads=3),
    "smoltts_byte_tiny": dict(_COMMON, dim=128, fast_dim=128, intermediate_size=256,
                              fast_intermediate_size=256, n_head=2, fast_n_head=2,
                              n_local_heads=1, fast_n_local_heads=1, n_layer=2, n_fast_layer=2),
}


def named_config(name: str, **overrides: Any) -> RQTransformerModelArgs:
    if name not in MODEL_SIZES:
        raise KeyError(f"unknown model size {name!r}; have {sorted(MODEL_SIZES)}")
    d = dict(MODEL_SIZES[name])
    d.update(overrides)
    return RQTransformerModelArgs.from_dict(d)
