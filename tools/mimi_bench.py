"""Mimi streaming decoder: time per decode step (CUDA events on the launching stream, graph replay) and the roofline beside it.

Algorithmic bytes of one step: every packed fp32 weight of the decode half once (the 8 codebook rows instead of the
codebooks) + the KV the step reads and writes.   python tools/mimi_bench.py [--batch 1,8,64] [--frames 256]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def weight_bytes(sd, n_q=8):
    tot = 0
    for k, v in sd.items():
        if "codebook" in k:
            continue
        tot += v.numel() * 4
    return tot + n_q * 256 * 4


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", default="1,8,64")
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"])
    args = ap.parse_args()
    from smoltts_b200.mimi import MimiModel
    from smoltts_b200.synth import make_mimi_state_dict

    sd = make_mimi_state_dict(0)
    wb = weight_bytes(sd)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6450.9))
    for B in [int(x) for x in args.batch.split(",")]:
        m = MimiModel(max_streams=B, max_frames=args.frames + 8, window=args.window, mode=args.mode)
        m.load_state_dict(sd)
        caches = [m.make_cache() for _ in range(B)]
        g = torch.Generator().manual_seed(1)
        codes = torch.randint(0, 2048, (B, 8, args.frames + 4), generator=g).cuda()
        for t in range(4):
            m.decode_step(codes[:, :, t], caches)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(4, 4 + args.frames):
            m.decode_step(codes[:, :, t], caches)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.frames
        mean_pos = 2 * (4 + args.frames / 2)
        kv = B * 8 * 2 * 512 * 4 * (2 * (mean_pos if args.window == 0 else min(mean_pos, args.window)) + 2)
        bytes_step = wb + kv
        print(json.dumps({"workload": f"mimi decode_step bs={B}, {args.frames} frames, window {args.window}, {args.mode}", "us_per_step": round(us, 1),
                          "frames_per_s": round(B * 1e6 / us, 1), "realtime_factor": round(B * 1e6 / us / 12.5, 1),
                          "launches_per_step": m.launches_per_step, "algorithmic_bytes_per_step": int(bytes_step),
                          "roofline": {"bound": "hbm", "achieved": round(bytes_step / us / 1e3, 1), "peak": hbm, "unit": "GB/s",
                                       "frac": round(bytes_step / us / 1e3 / hbm, 4)}}))
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
