"""CPU oracle of the Mimi streaming decoder (codes -> PCM) -- TEST INFRASTRUCTURE, never imported by the product.

Restates, with torch CPU fp32 ops, the decode half of the reference's codec
(`R` = /root/reference/mlx_inference/src/smoltts_mlx/codec/):

* `embed`            R/rvq.py:118-130 (sum of the codebook rows, 1x1 output projection), :171-186 (semantic + acoustic
                     halves), :30-45 (rows = embed_sum / max(cluster_usage, eps))
* `upsample_*`       R/conv.py:225-282 (grouped transposed convolution, kernel 4, stride 2, right trim), R/mimi.py:46-55
* transformer        R/transformer.py:36-150 (LayerNorm, q/k/v/o without bias, half-split RoPE, exact GELU, layer scale)
* SEANet decoder     R/seanet.py:99-161, R/conv.py:64-160 (causal Conv1d, streaming `step` with the carried input tail),
                     R/conv.py:162-221 (ConvTranspose1d, streaming `step` with the carried output tail), R/seanet.py:9-50
* `decode` / `decode_step`   R/mimi.py:73-104

The reference runs on MLX (Apple only, absent here), and its weights are kyutai/mimi's.  The same model exists in
Hugging Face transformers (`MimiModel`, transformers 5.5.0 in this image), which the MLX code was ported from and whose
state-dict keys it loads: `tests/test_mimi_oracle_golden.py` pins this file against goldens produced from it
(`tools/make_mimi_goldens.py`).

Two upsampling rules exist in the reference and both are kept:
* `decode` (whole sequence, R/mimi.py:88-100): the transposed convolution sees the whole sequence, frame t's last two
  taps land on frame t + 1's outputs.
* `decode_step` (R/mimi.py:102-104, the path SmolTTS.stream uses): `upsample` has no streaming form, every frame is
  upsampled ALONE and the two trailing taps are trimmed away (conv.py:278-281) -- `carry=False` below.  The SEANet and the
  transformer do carry their state, so everything after the upsampler is the whole-sequence computation.
The transformer's cache is the plain grow-by-concat KVCache (R/../__init__.py:86): no 250-position window (`window=0`);
`window=250` gives kyutai's / transformers' sliding window instead.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


@dataclass
class MimiDims:
    n_q: int = 8
    codebook_size: int = 2048
    codebook_dim: int = 256
    dim: int = 512
    n_layers: int = 8
    n_heads: int = 8
    head_dim: int = 64
    ffn: int = 2048
    n_filters: int = 64
    ratios: tuple = (8, 6, 5, 4)
    kernel: int = 7
    res_kernel: int = 3
    last_kernel: int = 3
    up_kernel: int = 4
    up_stride: int = 2
    rope_theta: float = 10000.0
    norm_eps: float = 1e-5
    cb_eps: float = 1e-5

    @property
    def samples_per_frame(self) -> int:
        return self.up_stride * int(math.prod(self.ratios))


def rope_table(dims: MimiDims, n_pos: int) -> torch.Tensor:
    """[n_pos][head_dim]: cos (first half) | sin (second half) of pos * theta^(-2i/head_dim)  (nn.RoPE traditional=False)."""
    half = dims.head_dim // 2
    inv = 1.0 / (dims.rope_theta ** (torch.arange(0, dims.head_dim, 2, dtype=torch.float32) / dims.head_dim))
    ang = torch.arange(n_pos, dtype=torch.float32)[:, None] * inv[None, :]
    return torch.cat([ang.cos(), ang.sin()], dim=1).contiguous()


@dataclass
class StreamState:
    """Everything `decode_step` carries between frames for one batch of streams."""
    k: List[Optional[torch.Tensor]] = field(default_factory=list)       # per layer [B, H, L, hd]
    v: List[Optional[torch.Tensor]] = field(default_factory=list)
    prev_in: Dict[str, Optional[torch.Tensor]] = field(default_factory=dict)    # Conv1d.step: carried input tail
    padded: Dict[str, bool] = field(default_factory=dict)
    prev_out: Dict[str, Optional[torch.Tensor]] = field(default_factory=dict)   # ConvTranspose1d.step: carried output tail
    up_prev: Optional[torch.Tensor] = None                                      # carry=True only: previous frame's embedding
    offset: int = 0


class MimiOracle:
    def __init__(self, sd: Dict[str, torch.Tensor], dims: Optional[MimiDims] = None, window: int = 0):
        self.d = dims or MimiDims()
        self.sd = {k: v.detach().to(torch.float32) for k, v in sd.items()}
        self.window = window

    # ---- RVQ: codes [B, n_q, T] -> [B, dim, T] -------------------------------------------------------------------
    def embed(self, codes: torch.Tensor) -> torch.Tensor:
        d, sd = self.d, self.sd

        def half(prefix: str, cs: torch.Tensor) -> torch.Tensor:
            q = None
            for i in range(cs.shape[1]):
                es = sd[f"{prefix}.layers.{i}.codebook.embed_sum"]
                cu = sd[f"{prefix}.layers.{i}.codebook.cluster_usage"]
                emb = es / torch.clamp(cu, min=d.cb_eps)[:, None]
                e = F.embedding(cs[:, i], emb)
                q = e if q is None else q + e
            return F.conv1d(q.transpose(1, 2), sd[f"{prefix}.output_proj.weight"])

        return (half("quantizer.semantic_residual_vector_quantizer", codes[:, :1])
                + half("quantizer.acoustic_residual_vector_quantizer", codes[:, 1:]))

    # ---- upsample: [B, dim, T] -> [B, dim, 2T] ---------------------------------------------------------------------
    def upsample_full(self, x: torch.Tensor) -> torch.Tensor:
        d = self.d
        y = F.conv_transpose1d(x, self.sd["upsample.conv.weight"], stride=d.up_stride, groups=x.shape[1])
        return y[..., : y.shape[-1] - (d.up_kernel - d.up_stride)]

    # ---- transformer: [B, T, dim] ----------------------------------------------------------------------------------------
    def _rope(self, x: torch.Tensor, offset: int) -> torch.Tensor:   # x [B, H, T, hd]
        T, hd = x.shape[2], self.d.head_dim
        tab = rope_table(self.d, offset + T)[offset:]
        cos, sin = tab[:, : hd // 2], tab[:, hd // 2:]
        x1, x2 = x[..., : hd // 2], x[..., hd // 2:]
        return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)

    def transformer(self, x: torch.Tensor, st: Optional[StreamState] = None) -> torch.Tensor:
        d, sd = self.d, self.sd
        B, T, _ = x.shape
        offset = st.offset if st is not None else 0
        for l in range(d.n_layers):
            p = f"decoder_transformer.layers.{l}."
            h = F.layer_norm(x, (d.dim,), sd[p + "input_layernorm.weight"], sd[p + "input_layernorm.bias"], d.norm_eps)
            q = (h @ sd[p + "self_attn.q_proj.weight"].t()).view(B, T, d.n_heads, d.head_dim).transpose(1, 2)
            k = (h @ sd[p + "self_attn.k_proj.weight"].t()).view(B, T, d.n_heads, d.head_dim).transpose(1, 2)
            v = (h @ sd[p + "self_attn.v_proj.weight"].t()).view(B, T, d.n_heads, d.head_dim).transpose(1, 2)
            q, k = self._rope(q, offset), self._rope(k, offset)
            if st is not None:
                if len(st.k) <= l:
                    st.k.append(None); st.v.append(None)
                k = k if st.k[l] is None else torch.cat([st.k[l], k], dim=2)
                v = v if st.v[l] is None else torch.cat([st.v[l], v], dim=2)
                st.k[l], st.v[l] = k, v
            L = k.shape[2]
            s = (q @ k.transpose(-1, -2)) * (d.head_dim ** -0.5)
            qpos = torch.arange(L - T, L)[:, None]
            kpos = torch.arange(L)[None, :]
            mask = kpos <= qpos
            if self.window > 0:
                mask = mask & (kpos > qpos - self.window)
            s = s.masked_fill(~mask, float("-inf"))
            a = torch.softmax(s, dim=-1) @ v
            a = a.transpose(1, 2).reshape(B, T, d.dim) @ sd[p + "self_attn.o_proj.weight"].t()
            x = x + a * sd[p + "self_attn_layer_scale.scale"]
            h = F.layer_norm(x, (d.dim,), sd[p + "post_attention_layernorm.weight"], sd[p + "post_attention_layernorm.bias"], d.norm_eps)
            h = F.gelu(h @ sd[p + "mlp.fc1.weight"].t()) @ sd[p + "mlp.fc2.weight"].t()
            x = x + h * sd[p + "mlp_layer_scale.scale"]
        if st is not None:
            st.offset = offset + T
        return x

    # ---- SEANet decoder: [B, dim, T] -> [B, 1, T * prod(ratios)] -------------------------------------------------------
    def _seanet_layers(self):
        """(kind, key, kernel, stride) in execution order; kinds: conv, elu, convtr, res."""
        d = self.d
        out = [("conv", "decoder.layers.0", d.kernel, 1)]
        idx = 1
        for r in d.ratios:
            out.append(("elu", "", 0, 0))
            out.append(("convtr", f"decoder.layers.{idx + 1}", 2 * r, r))
            out.append(("res", f"decoder.layers.{idx + 2}", d.res_kernel, 1))
            idx += 3
        out.append(("elu", "", 0, 0))
        out.append(("conv", f"decoder.layers.{idx + 1}", d.last_kernel, 1))
        return out

    def _conv_full(self, key: str, x: torch.Tensor, k: int) -> torch.Tensor:
        return F.conv1d(F.pad(x, (k - 1, 0)), self.sd[key + ".conv.weight"], self.sd[key + ".conv.bias"])

    def _convtr_full(self, key: str, x: torch.Tensor, k: int, s: int) -> torch.Tensor:
        y = F.conv_transpose1d(x, self.sd[key + ".conv.weight"], self.sd[key + ".conv.bias"], stride=s)
        return y[..., : y.shape[-1] - (k - s)]

    def seanet_full(self, x: torch.Tensor) -> torch.Tensor:
        for kind, key, k, s in self._seanet_layers():
            if kind == "elu":
                x = F.elu(x)
            elif kind == "conv":
                x = self._conv_full(key, x, k)
            elif kind == "convtr":
                x = self._convtr_full(key, x, k, s)
            else:
                h = self._conv_full(key + ".block.1", F.elu(x), k)
                h = self._conv_full(key + ".block.3", F.elu(h), 1)
                x = x + h
        return x

    # streaming forms (conv.py:133-160 and :207-221)
    def _conv_step(self, st: StreamState, key: str, x: torch.Tensor, k: int) -> torch.Tensor:
        if not st.padded.get(key, False):
            st.padded[key] = True
            x = F.pad(x, (k - 1, 0))
        prev = st.prev_in.get(key)
        x_long = x if prev is None else torch.cat([prev, x], dim=-1)
        n_frames = max(x_long.shape[-1] + 1 - k, 0)
        assert n_frames > 0
        st.prev_in[key] = x_long[..., n_frames:]
        return F.conv1d(x_long[..., : n_frames - 1 + k], self.sd[key + ".conv.weight"], self.sd[key + ".conv.bias"])

    def _convtr_step(self, st: StreamState, key: str, x: torch.Tensor, k: int, s: int) -> torch.Tensor:
        bias = self.sd[key + ".conv.bias"]
        ys = F.conv_transpose1d(x, self.sd[key + ".conv.weight"], bias, stride=s)
        prev = st.prev_out.get(key)
        if prev is not None:
            n = prev.shape[-1]
            ys = torch.cat([ys[..., :n] + (prev - bias[None, :, None]), ys[..., n:]], dim=-1)
        split = ys.shape[-1] - (k - s)
        st.prev_out[key] = ys[..., split:]
        return ys[..., :split]

    def seanet_step(self, st: StreamState, x: torch.Tensor) -> torch.Tensor:
        for kind, key, k, s in self._seanet_layers():
            if kind == "elu":
                x = F.elu(x)
            elif kind == "conv":
                x = self._conv_step(st, key, x, k)
            elif kind == "convtr":
                x = self._convtr_step(st, key, x, k, s)
            else:
                h = self._conv_step(st, key + ".block.1", F.elu(x), k)
                h = self._conv_step(st, key + ".block.3", F.elu(h), 1)
                x = x + h
        return x

    # ---- the two entry points of the reference -----------------------------------------------------------------------------
    def decode(self, codes: torch.Tensor) -> torch.Tensor:
        """Whole sequence (mimi.py:88-100): codes [B, n_q, T] -> PCM [B, 1, T * samples_per_frame]."""
        x = self.upsample_full(self.embed(codes))
        x = self.transformer(x.transpose(1, 2)).transpose(1, 2)
        return self.seanet_full(x)

    def decode_step(self, codes: torch.Tensor, st: StreamState, carry: bool = False) -> torch.Tensor:
        """One frame (mimi.py:102-104): codes [B, n_q, 1] -> PCM [B, 1, samples_per_frame].  carry=False is the reference's
        rule (the frame is upsampled alone); carry=True adds the previous frame's trailing taps, which makes a run of
        steps equal to `decode` of the whole sequence."""
        d = self.d
        e = self.embed(codes)
        x = self.upsample_full(e)
        if carry:
            if st.up_prev is not None:
                w = self.sd["upsample.conv.weight"]   # [dim, 1, 4]
                x = x + st.up_prev * w[None, :, 0, d.up_stride:]
            st.up_prev = e
        x = self.transformer(x.transpose(1, 2), st).transpose(1, 2)
        return self.seanet_step(st, x)
