#!/usr/bin/env python
"""Soak of the data-flow kernel's hand-off protocol: long launches at batch 1 and as teams (batch 2..8), repeated; every
repeat must reproduce the first one's codes bit for bit (a missed or stale word would change an id; a lost word traps).

usage: python tools/soak.py [--model smoltts_byte_150m] [--frames 1024] [--repeats 6]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from smoltts_b200 import GenerationSettings, RQTransformer, generate_batch, named_config  # noqa: E402
from smoltts_b200.synth import byte_prompt, make_state_dict, prompt_grid  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="smoltts_byte_150m")
    ap.add_argument("--frames", type=int, default=1024)
    ap.add_argument("--repeats", type=int, default=6)
    a = ap.parse_args()
    cfg = named_config(a.model)
    model = RQTransformer(cfg, max_batch=8, max_seq_len=64 + 12 + a.frames + 8)
    model.load_state_dict(make_state_dict(cfg, seed=0))
    ok = True
    for B, sampled in ((1, False), (1, True), (3, True), (8, False), (8, True)):
        gs = GenerationSettings(default_temp=0.7 if sampled else 0.0, default_fast_temp=0.7 if sampled else 0.0,
                                top_k=50 if sampled else 0, top_p=0.9 if sampled else 1.0, seed=3)
        prompts = [prompt_grid(byte_prompt(24 + 5 * b, seed=40 + b), cfg) for b in range(B)]
        first = None
        t0 = time.perf_counter()
        for r in range(a.repeats):
            outs = generate_batch(model, prompts, gs, audio_only=False, fixed_frames=a.frames, chunk=a.frames if r % 2 == 0 else 64)
            if first is None:
                first = outs
            elif not all(torch.equal(x, y) for x, y in zip(first, outs)):
                ok = False
                print(f"MISMATCH bs={B} sampled={sampled} repeat {r}")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{a.model} bs={B} {'sampled' if sampled else 'greedy '}: {a.repeats} x {a.frames} frames (launches of {a.frames} / 64 frames alternating) "
              f"identical: {ok}; {a.repeats * a.frames * B / dt:.0f} frames/s wall incl. prefill and host copies")
    print("SOAK", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
