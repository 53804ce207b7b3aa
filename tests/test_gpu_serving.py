"""Continuous batching (smoltts_b200/serving.py) on the GPU: an utterance decodes to the same codes in the slot scheduler
-- admitted late, next to changing neighbours, on recycled KV pages -- as in a plain ``generate_batch`` call."""
import pytest
import torch

from gpu_util import model_and_oracle
from smoltts_b200 import ContinuousBatcher, GenerationSettings, generate_batch
from smoltts_b200.synth import TOK_IM_END, byte_prompt, prompt_grid

pytestmark = pytest.mark.gpu


def _prompts(cfg, n, lo=14, seed=700):
    return [prompt_grid(byte_prompt(lo - 12 + (7 * i) % 23, seed=seed + i), cfg) for i in range(n)]


def _reference(model, prompts, budgets, gs, group, uids):
    """generate_batch in groups of `group` utterances (same kernel as a batcher with `group` slots), fixed frame counts."""
    want = {}
    for i in range(0, len(prompts), group):
        sel = list(range(i, min(i + group, len(prompts))))
        while len(sel) < group:          # pad the last group to the same batch size (same kernel variant)
            sel.append(sel[-1])
        outs = generate_batch(model, [prompts[j] for j in sel], gs, audio_only=False, fixed_frames=max(budgets[j] for j in sel),
                              seq_ids=[uids[j] for j in sel])
        for j, o in zip(sel, outs):
            want[uids[j]] = o
    return want


@pytest.mark.parametrize("size,slots,sampled", [("smoltts_byte_tiny", 16, False), ("smoltts_byte_70m", 12, True),
                                                ("smoltts_byte_tiny", 4, False), ("smoltts_byte_70m", 3, True)])
def test_continuous_batcher_equals_generate_batch(size, slots, sampled):
    cfg, sd, model, orc = model_and_oracle(size, max_batch=32, max_seq_len=512)
    n = 3 * slots + 2
    prompts = _prompts(cfg, n)
    budgets = [5 + (11 * i) % 19 for i in range(n)]                  # frames per utterance: 5 .. 23, ragged
    gs = GenerationSettings(default_temp=0.7 if sampled else 0.0, default_fast_temp=0.7 if sampled else 0.0, top_k=20 if sampled else 0,
                            top_p=0.9 if sampled else 1.0, seed=11, max_new_tokens=23)
    uids = [100 + i for i in range(n)]
    want = _reference(model, prompts, budgets, gs, slots, uids)
    free0 = len(model._free_pages)
    cb = ContinuousBatcher(model, gs, slots=slots, max_prompt=64, chunk=4, audio_only=False, ignore_stop=True)
    for i in range(n):
        cb.submit(prompts[i], max_new_tokens=budgets[i] - 1, uid=uids[i])
    got, order = {}, []
    for uid, codes in cb.run():
        got[uid] = codes
        order.append(uid)
    assert sorted(got) == sorted(uids) and cb.pending == 0 and cb.running == 0
    assert order != sorted(order) or n <= slots, "utterances with small budgets must retire before earlier, longer ones"
    for i, uid in enumerate(uids):
        w = want[uid][:, : budgets[i]]
        assert got[uid].shape == w.shape, f"utterance {uid}: {tuple(got[uid].shape)} vs {tuple(w.shape)}"
        assert torch.equal(got[uid], w), f"utterance {uid} (budget {budgets[i]}) differs from generate_batch"
    assert cb.stats["admitted"] == cb.stats["retired"] == n
    assert cb.stats["frames_decoded"] < cb.stats["slot_frames"]        # the tail of the run has idle slots
    cb.close()
    assert len(model._free_pages) == free0, "every KV page must be back in the pool"


def test_pages_are_recycled_under_a_small_pool():
    """A pool too small for all slots at their maximum: admission waits for retirements, everything still completes."""
    from smoltts_b200 import RQTransformer, named_config
    from smoltts_b200.synth import make_state_dict

    cfg = named_config("smoltts_byte_tiny")
    ps = 32
    model = RQTransformer(cfg, max_batch=16, max_seq_len=256, page_size=ps, kv_pages=1 + 10 * 3)   # 10 utterances' worth (+ scratch)
    model.load_state_dict(make_state_dict(cfg, seed=0))
    gs = GenerationSettings(default_temp=0.0, default_fast_temp=0.0, max_new_tokens=31)
    prompts = _prompts(cfg, 40, lo=26)            # 26 .. 48 columns + 32 frames + 8: three pages each
    cb = ContinuousBatcher(model, gs, slots=16, max_prompt=48, chunk=8, audio_only=False, ignore_stop=True)
    outs = cb.generate(prompts)
    assert cb.stats["retired"] == 40 and all(o.shape == (cfg.n_rows, 32) for o in outs)
    # at most 10 sequences ever fit the pool at once: the 16 slots were never all busy
    assert cb.stats["frames_decoded"] <= 10 * cb.stats["chunks"] * 8
    cb.close()
    again = ContinuousBatcher(model, gs, slots=9, max_prompt=48, chunk=8, audio_only=False, ignore_stop=True).generate(prompts[:9])
    for a, b in zip(again, outs[:9]):
        assert torch.equal(a, b), "same utterance ids, same codes whatever the scheduler did"


def test_stop_rule_retires_rows_and_frees_their_slots():
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=32, max_seq_len=512)
    gs = GenerationSettings(default_temp=0.0, default_fast_temp=0.0, max_new_tokens=40)
    prompts = _prompts(cfg, 12)
    cb = ContinuousBatcher(model, gs, slots=12, max_prompt=64, chunk=4)
    for p in prompts:
        cb.submit(p)
    assert cb.step_chunk() == [] and cb.running == 12
    force = torch.zeros(12, cfg.n_rows, dtype=torch.int32, device=model.device)
    force[:, 0] = TOK_IM_END                     # every sequence emits <|im_end|> as its next slow token
    model.set_force(force)
    try:
        done = cb.step_chunk()
    finally:
        model.set_force(None)
    assert len(done) == 12 and cb.running == 0
    assert cb.step.tolist() == [5] * 12          # 4 frames, then the <|im_end|> frame; the counters froze there
    for uid, codes in done:
        # audio_only keeps the columns whose slow token is a semantic id: never the <|im_end|> frame (lm/generate.py:143-171)
        assert codes.shape[0] == cfg.num_codebooks and codes.shape[1] <= 4, codes.shape
    cb.close()
