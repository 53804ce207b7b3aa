#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Imports ``modeling.model.rq_transformer`` from /root/reference (read-only, eager:
TORCHDYNAMO_DISABLE=1 because its @torch.compile sites cannot build here), loads the
seeded synthetic weights of ``smoltts_b200.synth`` into it, and records
  * teacher-forced ``RQTransformer.forward`` logits in fp32 and bf16
    (reference: modeling/model/rq_transformer.py:401-479), and
  * a literal greedy decode driven through the same ``forward`` (9 full forwards
    per frame: 1 for the slow token, then one per depth code through a
    provisional column), bf16 and fp32, with the top-2 margin of every decision.
The reference cannot travel to the GPU box, so these vectors are what pins
``oracle/dualar_oracle.py``.

usage: TORCHDYNAMO_DISABLE=1 python tools/make_goldens.py [--sizes tiny 70m 150m]
"""
from __future__ import annotations

import argparse
import hashlib
import os
import sys
import tempfile

os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import torch  # noqa: E402

from smoltts_b200.config import named_config  # noqa: E402
from smoltts_b200.synth import byte_prompt, flat_to_depthwise, make_state_dict, prompt_grid, teacher_grid  # noqa: E402
from tools.make_init import write_init  # noqa: E402

SPEC = {
    # size: (n_text, n_audio, batch, greedy_prompt_bytes, greedy_frames)
    "smoltts_byte_tiny": (20, 10, 2, 12, 16),
    "smoltts_byte_70m": (28, 8, 2, 12, 16),
    "smoltts_byte_150m": (28, 8, 2, 12, 16),
}
SUB = 8  # fp32 logits are stored at every SUB-th vocabulary entry


def weights_digest(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().view(torch.int16).numpy().tobytes())
    return h.hexdigest()


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).contiguous().view(torch.int16).numpy()


def load_reference(size: str, seed: int, dtype):
    from modeling.model.rq_transformer import RQTransformer

    cfg = named_config(size)
    with tempfile.TemporaryDirectory() as d:
        write_init(size, d, weights=False)
        model = RQTransformer.from_pretrained(d, load_weights=False)
    sd = make_state_dict(cfg, seed=seed, norm_jitter=0.05)
    sd_ref = dict(sd)
    sd_ref["fast_output.weight"] = flat_to_depthwise(sd["fast_output.weight"], cfg)
    missing = model.load_state_dict(sd_ref, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.to(dtype).eval()
    return cfg, sd, model


@torch.no_grad()
def literal_greedy(model, cfg, prompt: torch.Tensor, n_frames: int):
    """prompt [1, R, S].  Returns ids [n_frames, R] and margins [n_frames, R]."""
    grid = prompt.clone()
    R = cfg.n_rows
    ids = np.zeros((n_frames, R), dtype=np.int64)
    margins = np.zeros((n_frames, R), dtype=np.float32)
    for f in range(n_frames):
        L = grid.shape[2]
        res = model(inp=grid)
        lg = res.token_logits[0, L - 1].float()
        top2 = lg.topk(2).values
        tok = int(torch.argmax(res.token_logits[0, L - 1]))
        ids[f, 0], margins[f, 0] = tok, float(top2[0] - top2[1])
        col = torch.zeros(1, R, 1, dtype=torch.int64)
        col[0, 0, 0] = tok
        grid = torch.cat([grid, col], dim=2)
        for i in range(cfg.max_fast_seqlen):
            res = model(inp=grid)
            cl = res.codebook_logits[0, L - 1, i]
            top2 = cl.float().topk(2).values
            code = int(torch.argmax(cl))
            ids[f, 1 + i], margins[f, 1 + i] = code, float(top2[0] - top2[1])
            grid[0, 1 + i, L] = code
    return ids, margins


@torch.no_grad()
def make_one(size: str, seed: int, out_dir: str) -> None:
    n_text, n_audio, batch, gp_bytes, gframes = SPEC[size]
    out = {}
    for tag, dtype in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        cfg, sd, model = load_reference(size, seed, dtype)
        grid = teacher_grid(cfg, n_text, n_audio, batch=batch, seed=7, zero_code_at=3)
        res = model(inp=grid)
        tl = res.token_logits            # [B, S, V]
        t0 = n_text - 1                  # first position whose next column is audio
        cl = res.codebook_logits[:, t0:grid.shape[2] - 1]  # [B, n_audio, Nf, C]
        if tag == "bf16":
            out["tok_bf16"] = bf16_bits(tl)
            out["cb_bf16"] = bf16_bits(cl)
        else:
            out["tok_f32_sub"] = tl[..., ::SUB].contiguous().numpy()
            out["cb_f32_sub"] = cl[..., ::SUB].contiguous().numpy()
            t2 = tl.topk(2, dim=-1)
            out["tok_f32_argmax"] = t2.indices[..., 0].numpy().astype(np.int32)
            out["tok_f32_margin"] = (t2.values[..., 0] - t2.values[..., 1]).numpy()
            c2 = cl.topk(2, dim=-1)
            out["cb_f32_argmax"] = c2.indices[..., 0].numpy().astype(np.int32)
            out["cb_f32_margin"] = (c2.values[..., 0] - c2.values[..., 1]).numpy()
        prompt = prompt_grid(byte_prompt(gp_bytes, seed=1), cfg)[None]
        ids, margins = literal_greedy(model, cfg, prompt, gframes)
        out[f"greedy_ids_{tag}"] = ids.astype(np.int32)
        out[f"greedy_margin_{tag}"] = margins
        out["greedy_prompt"] = prompt[0].numpy().astype(np.int32)
        out["grid"] = grid.numpy().astype(np.int32)
        out["cb_t0"] = np.int32(t0)
    out["weights_sha256"] = np.frombuffer(weights_digest(sd).encode(), dtype=np.uint8)
    out["seed"] = np.int32(seed)
    out["sub"] = np.int32(SUB)
    path = os.path.join(out_dir, f"{size}.npz")
    np.savez_compressed(path, **out)
    print(f"{size}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", nargs="*", default=list(SPEC))
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    torch.manual_seed(0)
    for s in a.sizes:
        name = s if s.startswith("smoltts_byte_") else f"smoltts_byte_{s}"
        make_one(name, a.seed, a.out)


if __name__ == "__main__":
    main()
