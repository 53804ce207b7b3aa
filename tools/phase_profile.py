#!/usr/bin/env python
"""Per-phase time breakdown of the persistent decode kernels: the barrier kernel accumulates %globaltimer per phase
(CTA 0); the data-flow kernel (bs=1, default) records a %clock64 trace of the last frame (CTA 0 and CTA n/2).

usage: python tools/phase_profile.py [--model smoltts_byte_150m] [--batch 1] [--frames 256] [--out profiles/x.json]
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from smoltts_b200 import RQTransformer, named_config  # noqa: E402
from smoltts_b200.generate import GenerationSettings, _sampling, pack_prompts  # noqa: E402
from smoltts_b200.synth import byte_prompt, make_state_dict, prompt_grid  # noqa: E402

KINDS = ["QKV", "ATTN", "WO", "W13", "W2", "HEAD", "SAMPLE"]


def phase_kind(p, n_layer, n_flayer):
    n_slow = 5 * n_layer
    if p < n_slow:
        return "slow." + KINDS[p % 5]
    if p == n_slow:
        return "slow.HEAD"
    if p == n_slow + 1:
        return "slow.SAMPLE"
    r = (p - n_slow - 2) % (4 * n_flayer + 2)
    if r < 4 * n_flayer:
        return "fast." + ["QKV", "WO", "W13", "W2"][r % 4]
    return "fast.HEAD" if r == 4 * n_flayer else "fast.SAMPLE"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="smoltts_byte_150m")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--prompt-bytes", type=int, default=200)
    ap.add_argument("--sampled", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--repeat", type=int, default=0, help="profiling: run every weight-phase body 1+repeat times")
    ap.add_argument("--set", action="append", default=[], help="engine option name=value (smol_set_option), repeatable")
    a = ap.parse_args()
    cfg = named_config(a.model)
    need = a.prompt_bytes + 12 + a.frames + 8
    model = RQTransformer(cfg, max_batch=a.batch, max_seq_len=max(need, 256))
    model.load_state_dict(make_state_dict(cfg, seed=0))
    for kv in a.set:
        k, v = kv.split("=")
        model.set_option(k, int(v))
    prompts = [prompt_grid(byte_prompt(a.prompt_bytes, seed=1 + b), cfg) for b in range(a.batch)]
    padded, lens = pack_prompts(model, prompts)
    batch = model.new_batch(a.batch, max_positions=need, max_frames=a.frames)
    gs = (GenerationSettings(default_temp=0.7, default_fast_temp=0.7, top_k=50, top_p=0.9, seed=1234) if a.sampled
          else GenerationSettings(default_temp=0.0, default_fast_temp=0.0))
    s = _sampling(model, gs, True, ignore_stop=True)
    model.prefill(batch, padded, lens)
    model.decode_frames(batch, s, 8)  # warm
    torch.cuda.synchronize()
    model.set_option("repeat", a.repeat)
    prof = model.set_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    model.decode_frames(batch, s, a.frames)
    e1.record()
    torch.cuda.synchronize()
    model.set_profile(False)
    raw = prof.cpu().numpy().astype(float) / a.frames  # per frame
    ns = raw[: 2 * model.phase_count].reshape(-1, 2)
    seg = raw[2 * 512: 2 * 512 + 64].reshape(16, 4)
    trace = prof.cpu().numpy()[2 * 512 + 64: 2 * 512 + 64 + 2 * 512 * 8].reshape(2, 512, 8)
    skew = prof.cpu().numpy()[2 * 512 + 64 + 2 * 512 * 8:].reshape(256, 512, 2).astype(float)
    agg = {}
    for p in range(ns.shape[0]):
        k = phase_kind(p, cfg.n_layer, cfg.n_fast_layer)
        w, b, n = agg.get(k, (0.0, 0.0, 0))
        agg[k] = (w + ns[p, 0], b + ns[p, 1], n + 1)
    total = ns.sum()
    print(f"{a.model} bs={a.batch} frames={a.frames}: {e0.elapsed_time(e1) * 1e3 / a.frames:.1f} us/frame by events"
          + (f", {total / 1e3:.1f} us/frame by in-kernel timers (CTA 0)" if total > 0 else " (profiling build of the data-flow kernel)"))
    rows = {}
    if total > 0:  # barrier kernel: globaltimer accumulators (work, barrier wait) and sub-phase segments
        print(f"{'phase':14s} {'count':>5s} {'work us':>9s} {'wait us':>9s} {'per-phase work':>15s} {'per-phase wait':>15s}")
        for k, (w, b, n) in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][1])):
            print(f"{k:14s} {n:5d} {w / 1e3:9.1f} {b / 1e3:9.1f} {w / n / 1e3:15.2f} {b / n / 1e3:15.2f}")
            rows[k] = {"count": n, "work_us": w / 1e3, "wait_us": b / 1e3}
        print("sub-phase segments of the weight phases, us per frame (CTA 0 thread 0): prologue | stage wait | gemv+epilogue | final sync")
        for k in range(16):
            if seg[k].sum() > 0:
                name = ("fast." if k >= 8 else "slow.") + KINDS[k % 8]
                print(f"  {name:12s} " + " ".join(f"{v / 1e3:8.1f}" for v in seg[k]))
    if trace.any():
        names = ["poll", "norm/fill", "bar", "setup", "gemv", "publish", "release", "total"]
        print("cycle trace of the last frame (data-flow kernel), mean cycles per phase kind: " + " | ".join(names))
        for slot, label in ((0, "CTA 0"), (1, "CTA n/2")):
            acc = {}
            for p in range(model.phase_count):
                t = trace[slot, p].astype(float)
                if t[0] == 0:
                    continue
                k = phase_kind(p, cfg.n_layer, cfg.n_fast_layer)
                if t[2] == 0:  # attention / sample: start and end only
                    d = [0, 0, 0, 0, 0, 0, 0, t[6] - t[0]]
                else:
                    t3 = t[3] if t[3] else t[2]
                    t4 = t[4] if t[4] else t3
                    t5 = t[5] if t[5] else t4
                    d = [t[1] - t[0], t[7] - t[1], t[2] - t[7], t3 - t[2], t4 - t3, t5 - t4, t[6] - t5, t[6] - t[0]]
                a0, n0 = acc.get(k, ([0.0] * 8, 0))
                acc[k] = ([x + y for x, y in zip(a0, d)], n0 + 1)
            tot = 0.0
            for k, (v, n) in acc.items():
                print(f"  {label:8s} {k:12s} x{n:3d} " + " ".join(f"{x / n:8.0f}" for x in v))
                tot += v[7]
            print(f"  {label}: {tot:.0f} cycles per frame in phases")
    n_ctas = int((skew[:, 0, 1] > 0).sum())
    if n_ctas > 1:
        # per phase kind (norm-type phases stamp 'inputs arrived'): spread of the CTAs' phase ends, and the gap between the
        # LAST CTA finishing phase p and the inputs of phase p+1 arriving at the median CTA
        print(f"skew trace of the last frame over {n_ctas} CTAs (ns): phase-end spread p50/max | last end -> median arrival in the next norm phase")
        acc = {}
        P = model.phase_count
        for p in range(P - 1):
            end = skew[:n_ctas, p, 1]
            if (end <= 0).any():
                continue
            k = phase_kind(p, cfg.n_layer, cfg.n_fast_layer)
            spread = end.max() - np.median(end)
            arr = skew[:n_ctas, p + 1, 0]
            gap = (np.median(arr[arr > 0]) - end.max()) if (arr > 0).any() else float("nan")
            late = int(end.argmax())
            a0 = acc.setdefault(k, [[], [], []])
            a0[0].append(spread); a0[1].append(gap); a0[2].append(late)
        for k, (sp, gp, late) in acc.items():
            gp2 = [g for g in gp if g == g]
            print(f"  {k:12s} end(max-median) mean {np.mean(sp):7.0f} max {np.max(sp):7.0f} | last-end -> next arrival "
                  f"{(np.mean(gp2) if gp2 else float('nan')):7.0f} | most often last: {collections.Counter(late).most_common(5)}")
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"model": a.model, "batch": a.batch, "frames": a.frames,
                       "us_per_frame_events": e0.elapsed_time(e1) * 1e3 / a.frames, "by_phase": rows}, f, indent=1)


if __name__ == "__main__":
    main()
