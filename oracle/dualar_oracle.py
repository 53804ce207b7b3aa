"""CPU oracle for the DualAR per-frame decode step.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch CPU tensor ops, no CUDA) of the reference's
algorithm for the hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the product
package ``smoltts_b200`` never does.

What it follows (paths relative to the reference checkout):
  * numerics and rounding points: PyTorch ``modeling/model/rq_transformer.py``
    (eager, i.e. TORCHDYNAMO_DISABLE=1): embed :205-221, RMSNorm :601-613,
    Attention :535-570, RoPE :616-640, FeedForward :573-582, block :492-501,
    slow head :250-259, fast half :409-448, DepthwiseLinear :585-598;
  * incremental (KV-cached) structure, loop and token plumbing: the MLX decode
    ``mlx_inference/src/smoltts_mlx/lm/rq_transformer.py:150-220,266-295``,
    ``lm/cache.py:6-22`` and ``lm/generate.py:59-171`` (MLX itself cannot run here).

Pinning: the reference's own tests hold no golden vectors for this path
(SURVEY §4, §8(c)), so the oracle is pinned against outputs of the reference
itself: ``tools/make_goldens.py`` imports the unmodified reference from
/root/reference in the build container, runs ``RQTransformer.forward`` (fp32 and
bf16) and a literal greedy loop on seeded weights, and commits the results under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against them.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """reference :607-613 — fp32 normalise, cast to x.dtype, then multiply by weight."""
    xf = x.float()
    out = (xf * torch.rsqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + eps)).type_as(x)
    return out * w


def rope_table(seq_len: int, n_elem: int, base: float) -> torch.Tensor:
    """reference :616-624 — [S, n_elem/2, 2] (cos, sin), rounded to bf16."""
    freqs = 1.0 / (base ** (torch.arange(0, n_elem, 2)[: (n_elem // 2)].float() / n_elem))
    t = torch.arange(seq_len)
    freqs = torch.outer(t, freqs)
    cis = torch.polar(torch.ones_like(freqs), freqs)
    return torch.stack([cis.real, cis.imag], dim=-1).to(torch.bfloat16)


def apply_rope(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """reference :627-640 — interleaved pairs, fp32 math, cast back.  x: [B,S,H,hd]."""
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    fc = freqs.view(1, xs.size(1), 1, xs.size(3), 2)
    out = torch.stack([xs[..., 0] * fc[..., 0] - xs[..., 1] * fc[..., 1],
                       xs[..., 1] * fc[..., 0] + xs[..., 0] * fc[..., 1]], -1)
    return out.flatten(3).type_as(x)


@dataclass
class LayerCache:
    """Grow-by-concat KV cache, post-RoPE keys (MLX lm/cache.py:6-22)."""
    k: Optional[torch.Tensor] = None  # [B, Hkv, L, hd]
    v: Optional[torch.Tensor] = None
    offset: int = 0

    def update_and_fetch(self, k: torch.Tensor, v: torch.Tensor):
        if self.k is None:
            self.k, self.v = k, v
        else:
            self.k = torch.cat([self.k, k], dim=2)
            self.v = torch.cat([self.v, v], dim=2)
        self.offset += k.shape[2]
        return self.k, self.v


@dataclass
class OracleSettings:
    """GenerationSettings of MLX lm/generate.py:12-16 (+ the north-star's top-k/top-p/seed)."""
    default_temp: float = 0.7
    default_fast_temp: Optional[float] = 0.7
    min_p: Optional[float] = None
    max_new_tokens: int = 1024
    top_k: int = 0
    top_p: float = 1.0
    seed: int = 0


@dataclass
class Frame:
    semantic_code: int
    audio_codes: Optional[List[int]]
    vq: List[int]
    margins: List[float] = field(default_factory=list)


class DualAROracle:
    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], dtype=torch.bfloat16,
                 max_seq_len: Optional[int] = None, mlx_embed_mask: bool = False,
                 semantic_start: int = 320):
        self.cfg = cfg
        self.dtype = dtype
        self.mlx_embed_mask = mlx_embed_mask
        self.semantic_start = semantic_start
        self.semantic_end = semantic_start + cfg.codebook_size - 1
        w = {k: v.to(dtype) for k, v in state_dict.items()}
        fo = w["fast_output.weight"]
        if cfg.depthwise_output and fo.ndim == 3:  # trainer form [i, D, C] -> [(i C), D]
            fo = fo.permute(0, 2, 1).reshape(-1, fo.shape[1]).contiguous()
        w["fast_output.weight"] = fo
        self.w = w
        S = max_seq_len or cfg.max_seq_len
        # the table stays bf16 even when the model runs fp32 (reference :624)
        self.freqs = rope_table(S, cfg.dim // cfg.n_head, cfg.rope_base)
        self.fast_freqs = rope_table(cfg.max_fast_seqlen, cfg.fast_dim // cfg.fast_n_head, cfg.rope_base)
        N, C = cfg.num_codebooks, cfg.codebook_size
        off = torch.arange(0, N * C, C)
        self.semantic_offset = off if cfg.duplicate_code_0 else off[1:]  # reference :211-215

    # ------------------------------------------------------------------ caches
    def new_cache(self) -> List[LayerCache]:
        return [LayerCache() for _ in range(self.cfg.n_layer)]

    def new_fast_cache(self) -> List[LayerCache]:
        return [LayerCache() for _ in range(self.cfg.n_fast_layer)]

    # ------------------------------------------------------------------ pieces
    def embed(self, cols: torch.Tensor) -> torch.Tensor:
        """cols [B, R, S] int64 -> [B, S, D] (reference :205-221)."""
        w = self.w
        text = F.embedding(cols[:, 0, :], w["embeddings.weight"])
        vq = F.embedding(cols[:, 1:, :] + self.semantic_offset.view(1, -1, 1), w["codebook_embeddings.weight"])
        vq_sum = vq.sum(dim=1)
        if self.mlx_embed_mask:  # MLX rule (mlx lm/rq_transformer.py:162-169)
            keep = (cols[:, 0] >= self.semantic_start) & (cols[:, 0] <= self.semantic_end)
            vq_sum = vq_sum * keep.unsqueeze(-1).to(vq_sum.dtype)
        else:  # PyTorch rule (reference :219)
            vq_sum[cols[:, 1] == 0] = 0
        return text + vq_sum

    def _attention(self, x, p: str, n_head: int, n_kv: int, hd: int, freqs, cache: LayerCache, trace, tag):
        w = self.w
        B, S, D = x.shape
        qkv = F.linear(x, w[p + "attention.wqkv.weight"])
        q, k, v = qkv.split([n_head * hd, n_kv * hd, n_kv * hd], dim=-1)
        q = q.view(B, S, n_head, hd)
        k = k.view(B, S, n_kv, hd)
        v = v.view(B, S, n_kv, hd)
        q = apply_rope(q, freqs)
        k = apply_rope(k, freqs)
        if trace is not None:
            trace[tag + "q"] = q[:, -1].reshape(B, -1).clone()
            trace[tag + "k"] = k[:, -1].reshape(B, -1).clone()
            trace[tag + "v"] = v[:, -1].reshape(B, -1).clone()
        q, k, v = (t.transpose(1, 2) for t in (q, k, v))
        fresh = cache.offset == 0
        k, v = cache.update_and_fetch(k, v)
        if not fresh and S != 1:
            raise NotImplementedError("oracle supports (empty cache, any S) or (any cache, S == 1)")
        rep = n_head // n_kv
        ke = k.repeat_interleave(rep, dim=1)
        ve = v.repeat_interleave(rep, dim=1)
        y = F.scaled_dot_product_attention(q, ke, ve, dropout_p=0.0, is_causal=(S > 1))
        y = y.transpose(1, 2).contiguous().view(B, S, n_head * hd)
        if trace is not None:
            trace[tag + "attn"] = y[:, -1].clone()
        return F.linear(y, w[p + "attention.wo.weight"])

    def _block(self, x, p: str, fast: bool, freqs, cache: LayerCache, trace=None, tag=""):
        cfg, w = self.cfg, self.w
        n_head = cfg.fast_n_head if fast else cfg.n_head
        n_kv = cfg.fast_n_local_heads if fast else cfg.n_local_heads
        hd = (cfg.fast_dim // cfg.fast_n_head) if fast else (cfg.dim // cfg.n_head)
        xn = rmsnorm(x, w[p + "attention_norm.weight"], cfg.norm_eps)
        if trace is not None:
            trace[tag + "xn"] = xn[:, -1].clone()
        h = x + self._attention(xn, p, n_head, n_kv, hd, freqs, cache, trace, tag)
        hn = rmsnorm(h, w[p + "ffn_norm.weight"], cfg.norm_eps)
        a = F.silu(F.linear(hn, w[p + "feed_forward.w1.weight"])) * F.linear(hn, w[p + "feed_forward.w3.weight"])
        out = h + F.linear(a, w[p + "feed_forward.w2.weight"])
        if trace is not None:
            trace[tag + "h"] = h[:, -1].clone()
            trace[tag + "act"] = a[:, -1].clone()
            trace[tag + "out"] = out[:, -1].clone()
        return out

    def slow_forward(self, cols: torch.Tensor, cache: List[LayerCache], trace=None,
                     all_positions: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """cols [B, R, S] -> (token_logits, hidden).  Last position only (as the MLX
        ``forward_generate`` :173-192) unless ``all_positions``.  ``hidden`` is the
        PRE-norm stream (reference :259; quirk §8(g)-2)."""
        cfg, w = self.cfg, self.w
        S = cols.shape[2]
        pos0 = cache[0].offset
        x = self.embed(cols)
        if trace is not None:
            trace["embed"] = x[:, -1].clone()
        freqs = self.freqs[pos0:pos0 + S]
        for l in range(cfg.n_layer):
            x = self._block(x, f"layers.{l}.", False, freqs, cache[l], trace, f"l{l}.")
        if not all_positions:
            x = x[:, -1:, :]
        slow_out = rmsnorm(x, w["norm.weight"], cfg.norm_eps)
        head = w["embeddings.weight"] if cfg.tie_word_embeddings else w["output.weight"]
        logits = F.linear(slow_out, head)
        if "fast_project_in.weight" in w:
            x = F.linear(x, w["fast_project_in.weight"], w["fast_project_in.bias"])
        if all_positions:
            return logits, x
        return logits[:, 0], x[:, 0]

    def fast_step(self, x: torch.Tensor, i: int, fcache: List[LayerCache], trace=None) -> torch.Tensor:
        """x [B, Df] at depth position i -> codebook logits [B, C]
        (MLX ``forward_generate_fast`` :194-220; numerics of reference :438-448, :594-598)."""
        cfg, w = self.cfg, self.w
        x = x[:, None, :]
        freqs = self.fast_freqs[i:i + 1]
        for l in range(cfg.n_fast_layer):
            x = self._block(x, f"fast_layers.{l}.", True, freqs, fcache[l], trace, f"f{i}.l{l}.")
        out = rmsnorm(x, w["fast_norm.weight"], cfg.norm_eps)
        C = cfg.codebook_size
        fo = w["fast_output.weight"]
        if cfg.depthwise_output:
            fo = fo[i * C:(i + 1) * C]
        return F.linear(out, fo)[:, 0]

    def fast_embed(self, codes: torch.Tensor, i: int) -> torch.Tensor:
        """Embedding of depth code c_i as the input of depth step i+1
        (MLX lm/generate.py:136-140; reference :355-361,418-421)."""
        cfg = self.cfg
        off = 0
        if cfg.depthwise_wte:
            off = (i if cfg.duplicate_code_0 else i + 1) * cfg.codebook_size
        return F.embedding(codes + off, self.w["fast_embeddings.weight"])

    # ------------------------------------------------------------ teacher forcing
    def teacher_forced(self, grid: torch.Tensor, stepwise_from: Optional[int] = None,
                       fast_positions: Optional[List[int]] = None):
        """grid [B, R, S].  Prefill positions [0, stepwise_from) in one pass, then one
        cached step per position.  Returns token_logits [B, S, V] and
        codebook_logits {t: [B, Nf, C]} for t in fast_positions, where entry i is
        the logits for grid[:, 1+i, t+1] (SURVEY §3.2 verified mapping)."""
        cfg = self.cfg
        B, R, S = grid.shape
        s0 = S if stepwise_from is None else stepwise_from
        cache = self.new_cache()
        toks, hids = [], []
        if s0 > 0:
            lg, hd = self.slow_forward(grid[:, :, :s0], cache, all_positions=True)
            toks.append(lg)
            hids.append(hd)
        for t in range(s0, S):
            lg, hd = self.slow_forward(grid[:, :, t:t + 1], cache)
            toks.append(lg[:, None])
            hids.append(hd[:, None])
        token_logits = torch.cat(toks, dim=1)
        hidden = torch.cat(hids, dim=1)
        cb: Dict[int, torch.Tensor] = {}
        for t in (fast_positions or []):
            fcache = self.new_fast_cache()
            x = hidden[:, t]
            outs = []
            for i in range(cfg.max_fast_seqlen):
                outs.append(self.fast_step(x, i, fcache))
                if i + 1 < cfg.max_fast_seqlen:
                    x = self.fast_embed(grid[:, 1 + i, t + 1], i)
            cb[t] = torch.stack(outs, dim=1)
        return token_logits, cb

    # ------------------------------------------------------------------- decoding
    @staticmethod
    def _greedy(logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        lf = logits.float()
        top2 = lf.topk(2, dim=-1).values
        return torch.argmax(logits, dim=-1), top2[:, 0] - top2[:, 1]

    def decode_frame(self, cols: torch.Tensor, cache: List[LayerCache], settings: OracleSettings,
                     frame_index: int = 0, sampler=None, force: Optional[torch.Tensor] = None, trace=None):
        """One frame for B sequences.  cols [B, R, S] (S == 1 except at prefill).
        Returns next column [B, R] (row 0 = vocab id, rows 1.. = depth codes), the
        slow logits, the stacked depth logits and the greedy margins.
        ``force`` [B, R] overrides the chosen ids after their logits are computed
        (teacher forcing / greedy-with-resync)."""
        cfg = self.cfg
        logits, hidden = self.slow_forward(cols, cache, trace=trace)
        B = logits.shape[0]
        margins = []
        if settings.default_temp == 0.0:
            tok, m = self._greedy(logits)
            margins.append(m)
        else:
            tok = sampler(logits.float(), settings.default_temp, settings, frame_index, 0)
        if force is not None:
            tok = force[:, 0]
        out = [tok]
        x = hidden
        fcache = self.new_fast_cache()
        depth_logits = []
        for i in range(cfg.max_fast_seqlen):
            fl = self.fast_step(x, i, fcache, trace=trace)
            depth_logits.append(fl)
            ft = settings.default_fast_temp
            if ft is not None and ft > 0:
                code = sampler(fl.float(), ft, settings, frame_index, 1 + i)
            else:
                code, m = self._greedy(fl)
                margins.append(m)
            if force is not None:
                code = force[:, 1 + i]
            out.append(code)
            if i + 1 < cfg.max_fast_seqlen:  # the reference's extra OOB embed is not copied (§8(g)-7)
                x = self.fast_embed(code, i)
        nxt = torch.stack(out, dim=1)
        return nxt, logits, torch.stack(depth_logits, dim=1), (torch.stack(margins, dim=1) if margins else None)

    def generate(self, prompt: torch.Tensor, settings: OracleSettings, audio_only: bool = True,
                 fixed_frames: Optional[int] = None, sampler=None, im_end_id: int = 270) -> List[Frame]:
        """bs=1 loop of MLX lm/generate.py:59-171.  prompt [R, S] or [1, R, S].
        ``fixed_frames`` disables the stop rule and the max_new_tokens bound (benchmarks)."""
        cfg = self.cfg
        cols = prompt if prompt.ndim == 3 else prompt[None]
        cache = self.new_cache()
        frames: List[Frame] = []
        input_pos = 0
        while True:
            if fixed_frames is not None:
                if len(frames) >= fixed_frames:
                    break
            elif input_pos > settings.max_new_tokens or cols is None:
                break
            nxt, _, _, margins = self.decode_frame(cols, cache, settings, len(frames), sampler)
            vq = [int(v) for v in nxt[0]]
            slow = vq[0]
            audio = None
            if self.semantic_start <= slow <= self.semantic_end:
                audio = vq[1:] if cfg.duplicate_code_0 else [slow - self.semantic_start, *vq[1:]]
            frames.append(Frame(slow, audio, vq, [float(m) for m in margins[0]] if margins is not None else []))
            input_pos += 1
            stop = audio_only and slow == im_end_id and fixed_frames is None
            cols = None if stop else nxt[:, :, None]
        return frames
