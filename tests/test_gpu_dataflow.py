"""The data-flow (LL flag) kernel against the grid-barrier kernel: same phase program, same arithmetic,
same summation order -> every logit, every id, every cached K/V value must be BIT-identical, for
greedy and sampled decoding, batch 1..8, contexts that cross the split-KV and page boundaries, and
for prefill.  This is the regression net for the hand-off protocol (a missed or stale word shows up as
a bit difference)."""
import pytest
import torch

from gpu_util import model_and_oracle
from smoltts_b200.generate import pack_prompts
from smoltts_b200.synth import byte_prompt, prompt_grid

pytestmark = pytest.mark.gpu


def _run(model, mode, prompts, n_frames, chunk, seq_ids, **skw):
    model.set_option("mode", mode)
    model.set_option("ll_version", 1)   # this file pins the FIRST-generation data-flow kernel (same summation order as the barrier kernel)
    try:
        B = len(prompts)
        padded, lens = pack_prompts(model, prompts)
        batch = model.new_batch(B, max_positions=int(padded.shape[2]) + n_frames + 1, max_frames=n_frames, seq_ids=seq_ids)
        try:
            model.kv_view()[torch.tensor(batch.pages, device=model.device)] = 0  # leftovers of earlier runs
            model.prefill(batch, padded, lens)
            torch.cuda.synchronize()
            kv_prefill = model.kv_view()[torch.tensor(batch.pages, device=model.device)].clone()
            s = model.sampling(ignore_stop=True, **skw)
            done = 0
            logits = []
            while done < n_frames:
                n = min(chunk, n_frames - done)
                model.decode_frames(batch, s, n)
                torch.cuda.synchronize()
                logits.append((model.debug_buffer("token_logits", B).clone(), model.debug_buffer("depth_logits", B).clone()))
                done += n
            out = dict(codes=batch.out_codes.clone(), tokens=batch.tokens.clone(), seq_len=batch.seq_len.clone(),
                       step=batch.step.clone(), kv_prefill=kv_prefill,
                       kv=model.kv_view()[torch.tensor(batch.pages, device=model.device)].clone(), logits=logits)
        finally:
            batch.release()
        return out
    finally:
        model.set_option("mode", 2)
        model.set_option("ll_version", 2)


def _same(a, b, what):
    for k in ("codes", "tokens", "seq_len", "step"):
        assert torch.equal(a[k], b[k]), f"{what}: {k} differ"
    assert torch.equal(a["kv_prefill"].view(torch.int16), b["kv_prefill"].view(torch.int16)), f"{what}: prefill KV differs"
    assert torch.equal(a["kv"].view(torch.int16), b["kv"].view(torch.int16)), f"{what}: KV differs"
    for i, ((ta, da), (tb, db)) in enumerate(zip(a["logits"], b["logits"])):
        assert torch.equal(ta, tb), f"{what}: token logits differ after chunk {i}"
        assert torch.equal(da, db), f"{what}: depth logits differ after chunk {i}"


@pytest.mark.parametrize("size,B,n_prompt,n_frames,chunk", [
    ("smoltts_byte_tiny", 1, 20, 40, 7),
    ("smoltts_byte_tiny", 3, 50, 30, 30),
    ("smoltts_byte_tiny", 8, 30, 12, 5),
    ("smoltts_byte_70m", 1, 100, 24, 24),
    ("smoltts_byte_70m", 2, 60, 10, 3),
    ("smoltts_byte_150m", 1, 180, 16, 16),
    ("smoltts_byte_150m", 4, 40, 6, 6),
])
def test_dataflow_equals_barrier_kernel_greedy(size, B, n_prompt, n_frames, chunk):
    cfg, sd, model, orc = model_and_oracle(size)
    assert model.get_option("mode") == 2
    prompts = [prompt_grid(byte_prompt(n_prompt + 5 * b, seed=40 + b), cfg) for b in range(B)]
    ids = list(range(3, 3 + B))
    ref = _run(model, 0, prompts, n_frames, chunk, ids)
    got = _run(model, 2, prompts, n_frames, chunk, ids)
    assert model.get_option("ll1_ready") == 1 or B > 1
    _same(ref, got, f"{size} B={B}")


@pytest.mark.parametrize("size,B", [("smoltts_byte_tiny", 1), ("smoltts_byte_tiny", 5), ("smoltts_byte_70m", 1)])
def test_dataflow_equals_barrier_kernel_sampled(size, B):
    cfg, sd, model, orc = model_and_oracle(size)
    prompts = [prompt_grid(byte_prompt(25 + 4 * b, seed=60 + b), cfg) for b in range(B)]
    kw = dict(temp=0.8, fast_temp=0.6, top_k=40, top_p=0.9, seed=99)
    ref = _run(model, 0, prompts, 14, 4, None, **kw)
    got = _run(model, 2, prompts, 14, 4, None, **kw)
    _same(ref, got, f"{size} sampled B={B}")


@pytest.mark.parametrize("size,n_prompt", [("smoltts_byte_tiny", 2100), ("smoltts_byte_70m", 2100), ("smoltts_byte_tiny", 700)])
def test_dataflow_long_context_many_splits(size, n_prompt):
    """Context past 64 * splits and across many KV pages: the combiner path with up to 32 splits per head; on 70m
    (9 heads x 32 splits = 288 attention units) some CTAs run two units per phase."""
    cfg, sd, model, orc = model_and_oracle(size, max_batch=1, max_seq_len=2304)
    prompts = [prompt_grid(byte_prompt(n_prompt, seed=70), cfg)]
    ref = _run(model, 0, prompts, 6, 6, None)
    got = _run(model, 2, prompts, 6, 6, None)
    assert model.get_option("ll1_ready") == 1
    _same(ref, got, f"long context {size} {n_prompt}")


def test_dataflow_depth7_without_duplicate_code_0():
    """duplicate_code_0 = False (the kokoro_v1 data config): 7 depth steps, shifted embedding offsets."""
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=1, duplicate_code_0=False)
    assert cfg.max_fast_seqlen == 7
    prompts = [prompt_grid(byte_prompt(30, seed=75), cfg)]
    ref = _run(model, 0, prompts, 10, 4, None)
    got = _run(model, 2, prompts, 10, 4, None)
    assert model.get_option("ll1_ready") == 1
    _same(ref, got, "depth 7")


def test_dataflow_fewer_ctas_and_mixed_sampling():
    """A grid smaller than the GPU (more units per CTA, two GEMV rounds per warp) and the mixed case: the slow id sampled
    by CTA 0 (published per CTA), the depth codes greedy (resolved by every CTA from the candidate words)."""
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=1)
    prompts = [prompt_grid(byte_prompt(40, seed=77), cfg)]
    kw = dict(temp=0.9, fast_temp=0.0, top_k=30, top_p=0.95, seed=5)
    ref = _run(model, 0, prompts, 12, 5, [9], **kw)
    got = _run(model, 2, prompts, 12, 5, [9], **kw)
    _same(ref, got, "mixed sampling")
    model.set_option("n_ctas", 40)
    try:
        few = _run(model, 2, prompts, 12, 5, [9], **kw)
        assert model.get_option("n_ctas") == 40
    finally:
        model.set_option("n_ctas", 0)
    for k in ("codes", "tokens", "seq_len", "step"):
        assert torch.equal(ref[k], few[k]), f"40 CTAs: {k} differ"


def test_dataflow_stop_rule_single_sequence():
    """<|im_end|> forced in frame 2 of a 6-frame schedule: the sequence must freeze (tokens, seq_len, step, codes) in the
    middle of a multi-frame launch of the data-flow kernel exactly as in the barrier kernel."""
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=1)
    R = cfg.n_rows
    prompts = [prompt_grid(byte_prompt(14, seed=81), cfg)]
    res = {}
    for mode in (0, 2):
        model.set_option("mode", mode)
        model.set_option("ll_version", 1)
        padded, lens = pack_prompts(model, prompts)
        batch = model.new_batch(1, max_positions=64, max_frames=8)
        try:
            model.prefill(batch, padded, lens)
            s = model.sampling(audio_only=True)
            model.decode_frames(batch, s, 1)
            force = torch.zeros(1, R, dtype=torch.int32, device=model.device)
            force[0, 0] = model.token_config.im_end_id
            model.set_force(force)
            model.decode_frames(batch, s, 3)   # the stop fires in the first of these three frames
            model.set_force(None)
            model.decode_frames(batch, s, 2)
            torch.cuda.synchronize()
            res[mode] = (batch.tokens.clone(), batch.seq_len.clone(), batch.step.clone(), batch.finished.clone(), batch.out_codes.clone())
        finally:
            model.set_force(None)
            model.set_option("mode", 2)
            model.set_option("ll_version", 2)
            batch.release()
    assert res[0][3].tolist() == [1] and res[0][2].tolist() == [2]
    for a, b in zip(res[0], res[2]):
        assert torch.equal(a, b)


def test_dataflow_stop_rule_and_force():
    """<|im_end|> forced on sequence 1 in frame 2: it must freeze (tokens, seq_len, step) exactly as in
    the barrier kernel while sequence 0 keeps decoding inside the same launch."""
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny")
    B, R = 2, cfg.n_rows
    prompts = [prompt_grid(byte_prompt(12 + b, seed=80 + b), cfg) for b in range(B)]
    res = {}
    for mode in (0, 2):
        model.set_option("mode", mode)
        model.set_option("ll_version", 1)
        padded, lens = pack_prompts(model, prompts)
        batch = model.new_batch(B, max_positions=64, max_frames=8)
        try:
            model.prefill(batch, padded, lens)
            s = model.sampling(audio_only=True)
            model.decode_frames(batch, s, 1)
            force = torch.zeros(B, R, dtype=torch.int32, device=model.device)
            force[:, 0] = 400
            force[1, 0] = model.token_config.im_end_id
            model.set_force(force)
            model.decode_frames(batch, s, 1)
            model.set_force(None)
            model.decode_frames(batch, s, 4)
            torch.cuda.synchronize()
            res[mode] = (batch.tokens.clone(), batch.seq_len.clone(), batch.step.clone(), batch.finished.clone(), batch.out_codes.clone())
        finally:
            model.set_force(None)
            model.set_option("mode", 2)
            model.set_option("ll_version", 2)
            batch.release()
    assert res[0][3].tolist() == [0, 1] and res[0][2].tolist() == [6, 2]
    for a, b in zip(res[0], res[2]):
        assert torch.equal(a, b)


def test_frame_clock_is_monotonic_and_identical_ids():
    """smol_set_frame_clock: one %globaltimer stamp per frame, strictly increasing inside a launch, and switching it on
    does not change what is decoded."""
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=1)
    prompts = [prompt_grid(byte_prompt(20, seed=90), cfg)]
    ref = _run(model, 2, prompts, 16, 16, None)
    clock = model.set_frame_clock(16)
    try:
        got = _run(model, 2, prompts, 16, 16, None)
        torch.cuda.synchronize()
        ns = clock.cpu().tolist()
    finally:
        model.set_frame_clock(0)
    assert all(b > a > 0 for a, b in zip(ns, ns[1:])), ns
    assert (ns[-1] - ns[0]) / 15 < 5e6, "more than 5 ms per frame on the tiny model"
    assert torch.equal(ref["codes"], got["codes"])


@pytest.mark.parametrize("size,lens", [("smoltts_byte_tiny", [45]), ("smoltts_byte_70m", [70]), ("smoltts_byte_tiny", [33, 9, 58])])
def test_prefill_tiles_and_kernels_write_identical_kv(size, lens):
    """Prefill three ways -- 8 prompt positions per iteration on the barrier kernel, one position per iteration
    on the barrier kernel, one position per iteration on the data-flow kernel (bs=1) -- must leave bit-identical K/V,
    seq_len and pending tokens, also for ragged prompt lengths (rows of a tile past the end of a prompt are inert)."""
    cfg, sd, model, orc = model_and_oracle(size)
    B = len(lens)
    prompts = [prompt_grid(byte_prompt(n, seed=110 + b), cfg) for b, n in enumerate(lens)]
    variants = [(0, 8 // B), (0, 1)] + ([(2, 1), (0, 3)] if B == 1 else [])   # all below 9 rows: the CUDA-core variants
    got = []
    for mode, tile in variants:
        model.set_option("mode", mode)
        model.set_option("ll_version", 1)
        model.set_option("prefill_tile", tile)
        padded, lens_t = pack_prompts(model, prompts)
        batch = model.new_batch(B, max_positions=128, max_frames=4)
        try:
            pages = torch.tensor(batch.pages, device=model.device)
            model.kv_view()[pages] = 0
            model.prefill(batch, padded, lens_t)
            model.decode_frames(batch, model.sampling(ignore_stop=True), 2)
            torch.cuda.synchronize()
            got.append((model.kv_view()[pages].clone(), batch.seq_len.clone(), batch.tokens.clone(), batch.out_codes.clone()))
        finally:
            model.set_option("mode", 2)
            model.set_option("ll_version", 2)
            model.set_option("prefill_tile", 0)
            batch.release()
    assert got[0][1].tolist() == [n + 12 - 1 + 2 for n in lens]
    for (mode, tile), g in zip(variants[1:], got[1:]):
        assert torch.equal(g[0].view(torch.int16), got[0][0].view(torch.int16)), f"KV differs (mode {mode}, tile {tile})"
        for a, b in zip(g[1:], got[0][1:]):
            assert torch.equal(a, b), f"state differs (mode {mode}, tile {tile})"
