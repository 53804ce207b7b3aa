#!/usr/bin/env python
"""Runs a few frames with one launch per phase (mode 1, no graph) so that ncu can attribute time and
stall reasons to individual phases.  usage: python tools/ncu_phase.py [--frames 3] [--batch 1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from smoltts_b200 import RQTransformer, named_config  # noqa: E402
from smoltts_b200.generate import GenerationSettings, _sampling, pack_prompts  # noqa: E402
from smoltts_b200.synth import byte_prompt, make_state_dict, prompt_grid  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="smoltts_byte_150m")
ap.add_argument("--frames", type=int, default=3)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--prompt-bytes", type=int, default=200)
a = ap.parse_args()
cfg = named_config(a.model)
need = a.prompt_bytes + 12 + a.frames + 16
model = RQTransformer(cfg, max_batch=a.batch, max_seq_len=max(need, 256))
model.load_state_dict(make_state_dict(cfg, seed=0))
prompts = [prompt_grid(byte_prompt(a.prompt_bytes, seed=1 + b), cfg) for b in range(a.batch)]
padded, lens = pack_prompts(model, prompts)
batch = model.new_batch(a.batch, max_positions=need, max_frames=a.frames + 8)
s = _sampling(model, GenerationSettings(default_temp=0.0, default_fast_temp=0.0), True, ignore_stop=True)
model.prefill(batch, padded, lens)          # 1 persistent launch
torch.cuda.synchronize()
model.set_option("mode", 1)
torch.cuda.nvtx.range_push("frames")
for f in range(a.frames):
    for p in range(model.phase_count):       # explicit per-phase launches (no graph): ncu sees each one
        model.run_phases(batch, s, p, p + 1)
torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("launches", model.launch_count, "phases/frame", model.phase_count, "tokens", batch.tokens.tolist())
