#!/usr/bin/env python
"""Benchmark of the DualAR decode step: Mimi frames/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [...]                          # the reference's CPU arithmetic (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...           # one rank per GPU, utterance-sharded

A *step* is one pass of the hot path over one batch of synthetic input: ``--frames`` Mimi frames
(default 1024 = BASELINE configs[1]) decoded for ``--batch`` utterances per GPU, starting from a
prefilled 200-byte synthetic prompt.  ``value`` is whole-job frames/s with everything resident in HBM
(CUDA events around the step launches, max over ranks); ``e2e`` is the same metric through the public
``generate_batch`` API from host prompts (H2D of the prompt grid, prefill, decode, D2H of the codes).
The line also carries the roofline of the decode kernel (algorithmic bytes of SURVEY §8(d) over its
event-timed duration, against MEASURED_PEAKS.json), the CPU baseline (the oracle timed on the host
cores on a bounded sample), the SM clocks sampled during the timed region and the launch count.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="smoltts_byte_150m")
    ap.add_argument("--batch", type=int, default=1, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=1024, help="Mimi frames per step")
    ap.add_argument("--prompt-bytes", type=int, default=200)
    ap.add_argument("--sampled", action="store_true", help="configs[2]: temp 0.7 / top-k 50 / top-p 0.9, fast temp 0.7")
    ap.add_argument("--mode", type=int, default=2, help="2 data-flow persistent kernel (default; batch <= 8, barrier kernel above), "
                    "0 grid-barrier persistent kernel, 1 per-phase launches in a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU baseline sample (0 = auto)")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def load_tensor_peak():
    """Dense bf16 TFLOP/s: the sustained cuBLAS figure of MEASURED_PEAKS.json (the kernel is timed inside a long step)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), "measured sustained (MEASURED_PEAKS.json)"
    return 2250.0, "nominal dense bf16 (B200_PROFILING.md fallback)"


def load_traffic():
    """dram bytes per frame of the data-flow decode kernel from the committed ncu --set full capture (bs=1, 150m)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            with open(self.path) as f:
                for line in f:
                    p = [x.strip() for x in line.split(",")]
                    if len(p) < 9:
                        continue
                    try:
                        sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
                    except ValueError:
                        continue
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_prompts(cfg, batch: int, n_bytes: int, first_seq: int):
    from smoltts_b200.synth import byte_prompt, prompt_grid

    return [prompt_grid(byte_prompt(n_bytes, seed=1 + first_seq + b), cfg) for b in range(batch)]


def settings_for(args):
    from smoltts_b200 import GenerationSettings

    if args.sampled:
        return GenerationSettings(default_temp=0.7, default_fast_temp=0.7, top_k=50, top_p=0.9, seed=1234)
    return GenerationSettings(default_temp=0.0, default_fast_temp=0.0)


def workload_name(args) -> str:
    samp = "top-k50/top-p0.9/temp0.7 sampled" if args.sampled else "greedy"
    return (f"{args.model} {samp} decode bs={args.batch}/GPU, {args.frames} frames/step, "
            f"{args.prompt_bytes}-byte synthetic prompt ({args.prompt_bytes + 12} tokens)")


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a port of the reference's arithmetic with a KV cache) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_decode_rate(args, n_frames: int, batch: int = 1):
    import torch

    from oracle.dualar_oracle import DualAROracle, OracleSettings
    from oracle.sampler_oracle import OracleSampler
    from smoltts_b200.config import named_config
    from smoltts_b200.synth import make_state_dict

    torch.set_num_threads(os.cpu_count() or 1)
    cfg = named_config(args.model)
    sd = make_state_dict(cfg, seed=0)
    orc = DualAROracle(cfg, sd, dtype=torch.bfloat16, max_seq_len=max(cfg.max_seq_len, args.prompt_bytes + 12 + n_frames + 8))
    prompts = synth_prompts(cfg, batch, args.prompt_bytes, 0)
    cols = torch.stack(prompts, 0)
    st = (OracleSettings(default_temp=0.7, default_fast_temp=0.7, top_k=50, top_p=0.9, seed=1234) if args.sampled
          else OracleSettings(default_temp=0.0, default_fast_temp=0.0))
    sampler = OracleSampler(seq_ids=list(range(batch))) if args.sampled else None
    with torch.no_grad():
        cache = orc.new_cache()
        t0 = time.perf_counter()
        nxt, *_ = orc.decode_frame(cols, cache, st, 0, sampler)  # prefill + first frame (excluded, as G:187-193)
        t1 = time.perf_counter()
        for f in range(1, n_frames + 1):
            nxt, *_ = orc.decode_frame(nxt[:, :, None], cache, st, f, sampler)
        t2 = time.perf_counter()
    return {"frames_per_s": batch * n_frames / (t2 - t1), "prefill_s": t1 - t0, "decode_s": t2 - t1,
            "threads": torch.get_num_threads(), "frames": n_frames, "batch": batch}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_frames = args.cpu_frames or 128  # ~3 s of CPU work per step
    per_step = []
    for i in range(args.warmup + args.steps):
        r = cpu_decode_rate(args, n_frames, batch=args.batch)
        if i >= args.warmup:
            per_step.append(r)
    total_frames = sum(r["frames"] * r["batch"] for r in per_step)
    total_s = sum(r["decode_s"] for r in per_step)
    value = total_frames / total_s
    sample = (f"{n_frames} frames after a {args.prompt_bytes + 12}-token prefill, bs={args.batch}, bf16 eager on CPU "
              f"(oracle port of modeling/ arithmetic with a KV cache; prefill excluded as in lm/generate.py:187-214)")
    line = {
        "impl": "reference", "metric": "Mimi frames/sec", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / len(per_step),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": workload_name(args), "frames_per_cpu_step": n_frames},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": per_step[0]["threads"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from smoltts_b200 import RQTransformer, generate_batch, named_config
    from smoltts_b200.generate import _sampling, pack_prompts
    from smoltts_b200.synth import make_state_dict

    cfg = named_config(args.model)
    n_prompt = args.prompt_bytes + 12
    need = n_prompt + args.frames + 8
    model = RQTransformer(cfg, max_batch=args.batch, max_seq_len=max(need, 256),
                          kv_pages=2 * args.batch * ((max(need, 256) + 31) // 32))
    model.load_state_dict(make_state_dict(cfg, seed=0))
    model.set_option("mode", args.mode)
    gs = settings_for(args)
    prompts = synth_prompts(cfg, args.batch, args.prompt_bytes, rank * args.batch)
    seq_ids = [rank * args.batch + b for b in range(args.batch)]
    dev = model.device

    # ---- device-resident arm: prefill once, then every step decodes `frames` frames from the same state
    padded, lens = pack_prompts(model, prompts)
    batch = model.new_batch(args.batch, max_positions=need, max_frames=args.frames, seq_ids=seq_ids)
    sampling = _sampling(model, gs, audio_only=True, ignore_stop=True)
    model.prefill(batch, padded, lens)
    torch.cuda.synchronize()
    tokens0, len0 = batch.tokens.clone(), batch.seq_len.clone()
    host_len0 = list(batch.host_len)

    def reset():
        batch.tokens.copy_(tokens0)
        batch.seq_len.copy_(len0)
        batch.step.zero_()
        batch.finished.zero_()
        batch.host_len = list(host_len0)

    def step():
        reset()
        model.decode_frames(batch, sampling, args.frames)

    frame_clock = model.set_frame_clock(args.frames) if args.batch == 1 else None  # device-side per-frame stamps (sequence 0)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = model.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        reset()
        ev[i][0].record()
        model.decode_frames(batch, sampling, args.frames)
        ev[i][1].record()
    stop.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    launches = model.launch_count - launches0
    clocks = sampler.stop()
    total_ms = start.elapsed_time(stop)
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    codes_check = int(batch.out_codes.sum().item())  # the result is really produced
    frame_lat = None
    if frame_clock is not None:  # stamps of the last timed step: p50 / p99 of the per-frame latency (BASELINE.json's second metric)
        ns = frame_clock.cpu().numpy().astype("float64")
        d = (ns[1:] - ns[:-1]) * 1e-3
        d = d[d > 0]
        if d.size:
            d.sort()
            frame_lat = {"p50_us": float(d[int(0.50 * (d.size - 1))]), "p99_us": float(d[int(0.99 * (d.size - 1))]),
                         "max_us": float(d[-1]), "frames": int(d.size) + 1, "clock": "%globaltimer at frame assembly, on the device"}
        model.set_frame_clock(0)

    # ---- end-to-end arm: public API from host prompts, H2D + prefill + decode + D2H every step
    host_prompts = [p.pin_memory() for p in prompts]
    for _ in range(max(1, min(args.warmup, 2))):
        generate_batch(model, host_prompts, gs, audio_only=False, fixed_frames=args.frames, chunk=args.frames, seq_ids=seq_ids)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        outs = generate_batch(model, host_prompts, gs, audio_only=False, fixed_frames=args.frames, chunk=args.frames,
                              seq_ids=seq_ids)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = int(padded.numel() * 4 + lens.numel() * 4)
    d2h = int(args.batch * args.frames * cfg.n_rows * 4 + args.batch * 4)

    # ---- max over ranks
    t = torch.tensor([total_ms, e2e_s * 1e3, statistics.mean(kernel_ms)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kernel_ms_mean = (float(v) for v in t.tolist())

    if rank == 0:
        frames_total = args.steps * args.frames * args.batch * world
        value = frames_total / (total_ms * 1e-3)
        e2e_value = e2e_steps * args.frames * args.batch * world / (e2e_ms * 1e-3)
        hbm_gbs, peak_src = load_peaks()
        l_mean = n_prompt + (args.frames - 1) / 2.0
        w_unique = cfg.unique_weight_bytes()
        bytes_per_frame = w_unique + args.batch * cfg.kv_bytes_per_position() * (l_mean + 1)
        bytes_per_launch = bytes_per_frame * args.frames
        achieved = bytes_per_launch / (kernel_ms_mean * 1e-3) / 1e9
        traffic = load_traffic()
        dataflow = args.mode == 2 and args.batch <= 1
        tensor = not dataflow and args.batch >= model.get_option("tc_min_batch") > 0
        kernel_name = ("smol_ll_kernel" if dataflow else "smol_decode_kernel<0> (tcgen05 tiles, TMA operand ring)" if tensor
                       else "smol_decode_kernel")
        launch_mode = ("data-flow persistent kernel (LL flag words, TMA producer warp), one launch per step" if dataflow
                       else "persistent cooperative kernel with grid barriers" if args.mode != 1
                       else "per-phase launches in a CUDA graph")
        # which roof binds: bytes over HBM bandwidth vs flops over the bf16 tensor peak
        tf_peak, tf_src = load_tensor_peak()
        flops_per_launch = args.batch * cfg.flops_per_frame(l_mean) * args.frames
        t_bytes_us = bytes_per_frame / (hbm_gbs * 1e9) * 1e6
        t_flops_us = args.batch * cfg.flops_per_frame(l_mean) / (tf_peak * 1e12) * 1e6
        tensor_bound = tensor and t_flops_us >= t_bytes_us
        line = {
            "metric": "Mimi frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {
                "workload": workload_name(args), "model": args.model, "batch_per_gpu": args.batch,
                "frames_per_step": args.frames, "prompt_tokens": n_prompt, "mean_context": l_mean,
                "sharding": f"utterance-parallel x{world}, no collective on the decode path",
                "launch_mode": launch_mode,
                "l2": "no flush: each frame streams 271 MB of weights (> 126 MB L2) plus the KV cache",
                "us_per_frame": 1e3 * kernel_ms_mean / args.frames,
                "frame_latency_bs1": frame_lat,
                "e2e_includes": "H2D prompt grid + prefill (tensor-core tiles of up to 128 prompt positions) + decode + D2H codes, via generate_batch()",
            },
            "roofline": ({
                "bound": "tensor", "achieved": flops_per_launch / (kernel_ms_mean * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                "frac": flops_per_launch / (kernel_ms_mean * 1e-3) / 1e12 / tf_peak, "peak_source": tf_src, "kernel": kernel_name,
                "algorithmic_flops_per_launch": flops_per_launch, "launch_ms": kernel_ms_mean,
                "roof_us_per_step": {"flops": t_flops_us, "bytes": t_bytes_us},
                "traffic": (((traffic or {}).get("tensor_core_bs256") or {}).get("dram_bytes_per_frame", 0) * args.frames or None)
                           if args.model == "smoltts_byte_150m" and args.batch == 256 and args.prompt_bytes == 64 else None,
            } if tensor_bound else {
                "bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                "peak_source": peak_src, "kernel": kernel_name,
                "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": kernel_ms_mean,
                "roof_us_per_step": {"flops": t_flops_us, "bytes": t_bytes_us},
                "traffic": ((traffic or {}).get("dram_bytes_per_frame") * args.frames
                            if traffic and dataflow and args.model == "smoltts_byte_150m" and args.batch == 1 else None),
            }),
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "clocks": clocks,
            "gpu_launches": launches,
            "check": codes_check,
        }
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = args.cpu_frames or 480  # ~10 s of CPU work at ~46 frames/s
            r = cpu_decode_rate(args, n_cpu, batch=args.batch)
            line["cpu_baseline"] = {
                "value": r["frames_per_s"], "unit": "frames/s", "cores": r["threads"], "kind": "port",
                "sample": (f"{n_cpu} frames after a {n_prompt}-token prefill ({r['prefill_s']:.1f} s), bs={args.batch}, "
                           f"bf16 eager CPU oracle, {r['decode_s']:.1f} s of decode")}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
