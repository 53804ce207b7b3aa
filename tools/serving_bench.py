#!/usr/bin/env python
"""Continuous batching against static batches on ragged work (smoltts_b200/serving.py): N utterances whose lengths are
drawn from [lo, hi] frames, decoded (a) by ContinuousBatcher with `slots` resident sequences and (b) by generate_batch in
static groups of `slots` utterances that all run as long as their longest member.  Prints useful frames/s of both.

usage: python tools/serving_bench.py [--model smoltts_byte_150m] [--slots 64] [--n 384] [--lo 32] [--hi 256] [--chunk 16]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from smoltts_b200 import ContinuousBatcher, GenerationSettings, RQTransformer, generate_batch, named_config  # noqa: E402
from smoltts_b200.synth import byte_prompt, make_state_dict, prompt_grid  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="smoltts_byte_150m")
    ap.add_argument("--slots", type=int, default=64)
    ap.add_argument("--n", type=int, default=384)
    ap.add_argument("--lo", type=int, default=32)
    ap.add_argument("--hi", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=16)
    ap.add_argument("--prompt-bytes", type=int, default=64)
    a = ap.parse_args()
    cfg = named_config(a.model)
    max_prompt = a.prompt_bytes + 12
    model = RQTransformer(cfg, max_batch=a.slots, max_seq_len=max_prompt + a.hi + 1 + a.chunk + 8)
    model.load_state_dict(make_state_dict(cfg, seed=0))
    g = torch.Generator().manual_seed(5)
    budgets = torch.randint(a.lo, a.hi + 1, (a.n,), generator=g).tolist()
    prompts = [prompt_grid(byte_prompt(a.prompt_bytes, seed=900 + i), cfg) for i in range(a.n)]
    gs = GenerationSettings(default_temp=0.7, default_fast_temp=0.7, top_k=50, top_p=0.9, max_new_tokens=a.hi)
    useful = sum(budgets)

    def continuous():
        cb = ContinuousBatcher(model, gs, slots=a.slots, max_prompt=max_prompt, chunk=a.chunk, ignore_stop=True)
        for i in range(a.n):
            cb.submit(prompts[i], max_new_tokens=budgets[i] - 1, uid=i)
        out = dict(cb.run())
        stats = dict(cb.stats)
        cb.close()
        return out, stats

    def static():
        out = {}
        for i in range(0, a.n, a.slots):
            sel = list(range(i, min(i + a.slots, a.n)))
            res = generate_batch(model, [prompts[j] for j in sel], gs, fixed_frames=max(budgets[j] for j in sel), seq_ids=sel)
            for j, r in zip(sel, res):
                out[j] = r
        return out

    for name, fn in (("warm-up", continuous), ("continuous", continuous), ("static", static)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if name == "warm-up":
            continue
        extra = ""
        if name == "continuous":
            st = res[1]
            extra = f"; slot occupancy {st['frames_decoded'] / max(st['slot_frames'], 1):.2f}, {st['chunks']} chunks of {a.chunk} frames"
        print(f"{a.model} slots={a.slots} n={a.n} frames/utterance {a.lo}..{a.hi} (sum {useful}): {name:10s} {dt * 1e3:8.1f} ms wall "
              f"-> {useful / dt:9.0f} useful frames/s{extra}")


if __name__ == "__main__":
    main()
