"""Text -> PCM serving throughput: SmolTTS.serve / serve_stream (continuous batching with audio out) on smoltts_byte_150m + the
Mimi decoder, seeded weights.  Wall-clock audio frames per second over the whole run (prompt encoding, prefill, decode, codec,
PCM to the host).   python tools/tts_serving_bench.py [--slots 8,64] [--n 96] [--frames 64]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slots", default="8,64")
    ap.add_argument("--n", type=int, default=96)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--model", default="smoltts_byte_150m")
    args = ap.parse_args()
    from smoltts_b200 import GenerationSettings, MimiModel, PromptEncoder, RQTransformer, SmolTTS, byte_level_tokenizer, named_config
    from smoltts_b200.synth import make_mimi_state_dict, make_state_dict

    cfg = named_config(args.model)
    sd, msd = make_state_dict(cfg, seed=0), make_mimi_state_dict(0)
    texts = [("Utterance %d. " % i) * (1 + i % 5) for i in range(args.n)]
    for slots in [int(x) for x in args.slots.split(",")]:
        lm = RQTransformer(cfg, max_batch=slots, max_seq_len=256 + args.frames)
        lm.load_state_dict(sd)
        codec = MimiModel(max_streams=slots, max_frames=args.frames + 8, upsample_carry=True)
        codec.load_state_dict(msd)
        enc = PromptEncoder.from_model(byte_level_tokenizer(cfg.codebook_size), lm)
        gs = GenerationSettings(default_temp=0.7, default_fast_temp=0.7, top_k=50, top_p=0.9, seed=7, max_new_tokens=args.frames - 1)
        tts = SmolTTS(lm, enc, codec, codec, gs)
        for mode in ("serve", "serve_stream"):
            for rep in range(2):        # first pass warms everything up (graphs per batch size, allocator)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                samples = 0
                first = None
                it = tts.serve(texts, slots=slots, chunk=16, max_prompt=128) if mode == "serve" else tts.serve_stream(texts, slots=slots, chunk=16, max_prompt=128)
                for item in it:
                    pcm = item[1]
                    samples += pcm.shape[0]
                    if first is None and pcm.shape[0]:
                        first = time.perf_counter() - t0
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            print(json.dumps({"workload": f"{mode}: {args.n} utterances x {args.frames} frames (sampled), {slots} slots, {args.model} + Mimi",
                              "audio_frames_per_s": round(samples / 1920 / dt, 1), "realtime_factor": round(samples / 24000 / dt, 1),
                              "wall_s": round(dt, 3), "first_audio_s": round(first or 0.0, 4), "audio_s": round(samples / 24000, 1)}))
        del tts, lm, codec
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
