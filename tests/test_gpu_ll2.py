"""The second-generation data-flow kernel (ll2_kernel.cu, the bs=1 bench path) against itself and against the sampler spec.

Its parity with the oracle is checked where every kernel's is (tests/test_gpu_decode.py: teacher-forced logits against the
reference goldens, greedy-with-resync against the oracle; tests/test_gpu_scale_oracle.py: contexts of 2100 / 4300
positions) -- mode 2 is the default, so those tests run this kernel at batch 1.  Here: the hand-off protocol and the
launch geometry cannot change a bit.  A missed or stale word, a ring slot read before its bytes landed, a wrong epoch
after a relaunch all show up as a difference between two runs that must be bit-identical:

  * one launch of n frames == the same frames in launches of 5 / 4 / 3 (epochs continue across launches);
  * 148 CTAs == 40 CTAs (K is split over warps, rows over CTAs: the summation order does not depend on the grid);
  * long contexts (several softmax blocks of 512, score units > CTAs) in one launch == frame by frame;
  * sampled decoding: every id equals what the CPU sampler specification picks from the logits the kernel dumped;
  * forced ids / the stop rule inside a multi-frame launch; the frame clock.
"""
import pytest
import torch

from gpu_util import model_and_oracle
from oracle.sampler_oracle import sample_row
from smoltts_b200.generate import pack_prompts
from smoltts_b200.synth import byte_prompt, prompt_grid

pytestmark = pytest.mark.gpu


def _run(model, prompt, n_frames, chunks, seq_ids=None, **skw):
    padded, lens = pack_prompts(model, [prompt])
    batch = model.new_batch(1, max_positions=int(padded.shape[2]) + n_frames + 1, max_frames=n_frames, seq_ids=seq_ids)
    try:
        pages = torch.tensor(batch.pages, device=model.device)
        model.kv_view()[pages] = 0
        model.prefill(batch, padded, lens)
        s = model.sampling(ignore_stop=True, **skw)
        logits = []
        for n in chunks:
            model.decode_frames(batch, s, n)
            torch.cuda.synchronize()
            logits.append((model.debug_buffer("token_logits", 1).clone(), model.debug_buffer("depth_logits", 1).clone()))
        assert sum(chunks) == n_frames
        assert model.get_option("ll_ready") == 1, "the data-flow kernel did not run"
        return dict(codes=batch.out_codes.clone(), tokens=batch.tokens.clone(), seq_len=batch.seq_len.clone(), step=batch.step.clone(),
                    kv=model.kv_view()[pages].clone(), last_logits=logits[-1])
    finally:
        batch.release()


def _same(a, b, what):
    for k in ("codes", "tokens", "seq_len", "step"):
        assert torch.equal(a[k], b[k]), f"{what}: {k} differ"
    assert torch.equal(a["kv"].view(torch.int16), b["kv"].view(torch.int16)), f"{what}: KV differs"
    assert torch.equal(a["last_logits"][0], b["last_logits"][0]) and torch.equal(a["last_logits"][1], b["last_logits"][1]), f"{what}: logits differ"


@pytest.mark.parametrize("size,n_prompt", [("smoltts_byte_tiny", 20), ("smoltts_byte_70m", 100), ("smoltts_byte_150m", 180)])
def test_ll2_relaunch_and_grid_size_do_not_change_a_bit(size, n_prompt):
    cfg, sd, model, orc = model_and_oracle(size, max_batch=1)
    prompt = prompt_grid(byte_prompt(n_prompt, seed=40), cfg)
    one = _run(model, prompt, 12, [12])
    _same(one, _run(model, prompt, 12, [5, 4, 3]), f"{size}: one launch vs three")
    _same(one, _run(model, prompt, 12, [1] * 12), f"{size}: one launch vs frame by frame")
    model.set_option("n_ctas", 40)
    try:
        few = _run(model, prompt, 12, [12])
        assert model.get_option("n_ctas") == 40
    finally:
        model.set_option("n_ctas", 0)
    _same(one, few, f"{size}: 148 vs 40 CTAs")
    model.set_option("ll_holdoff", 0)
    try:
        _same(one, _run(model, prompt, 12, [12]), f"{size}: hold-off 0")
    finally:
        model.set_option("ll_holdoff", 400)


@pytest.mark.parametrize("size,n_prompt", [("smoltts_byte_tiny", 2100), ("smoltts_byte_70m", 1300), ("smoltts_byte_tiny", 505)])
def test_ll2_long_context_blocks(size, n_prompt):
    """Contexts across the 512-position softmax blocks (505 -> 517 crosses the first boundary), more score units than CTAs
    (70m: 3 kv heads x 21 segments; tiny at 2100: 33 segments on a 40-CTA grid -> several units per CTA)."""
    cfg, sd, model, orc = model_and_oracle(size, max_batch=1, max_seq_len=2304)
    prompt = prompt_grid(byte_prompt(n_prompt, seed=70), cfg)
    one = _run(model, prompt, 14, [14])
    _same(one, _run(model, prompt, 14, [1] * 14), f"{size} {n_prompt}: one launch vs frame by frame")
    model.set_option("n_ctas", 40)
    try:
        _same(one, _run(model, prompt, 14, [14]), f"{size} {n_prompt}: 40 CTAs")
    finally:
        model.set_option("n_ctas", 0)


def test_ll2_depth7_without_duplicate_code_0():
    """duplicate_code_0 = False (the kokoro_v1 data config): 7 depth steps, shifted embedding offsets -- against
    the barrier kernel through forced ids: the two kernels sum in different orders, so logits are compared, not bits."""
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=1, duplicate_code_0=False)
    assert cfg.max_fast_seqlen == 7
    prompt = prompt_grid(byte_prompt(30, seed=75), cfg)
    one = _run(model, prompt, 10, [10])
    _same(one, _run(model, prompt, 10, [4, 3, 3]), "depth 7")
    model.set_option("mode", 0)
    try:
        padded, lens = pack_prompts(model, [prompt])
        batch = model.new_batch(1, max_positions=64, max_frames=10)
        model.prefill(batch, padded, lens)
        # replay the first 9 frames with forced ids, then compare the logits of frame 10 (same history in both kernels)
        s = model.sampling(ignore_stop=True)
        for f in range(10):
            model.set_force(one["codes"][:, f].contiguous())
            model.decode_frames(batch, s, 1)
        torch.cuda.synchronize()
        tl, dl = model.debug_buffer("token_logits", 1).clone(), model.debug_buffer("depth_logits", 1).clone()
        batch.release()
    finally:
        model.set_force(None)
        model.set_option("mode", 2)
    assert (tl - one["last_logits"][0]).abs().max().item() <= 0.25 and (dl - one["last_logits"][1]).abs().max().item() <= 0.35
    assert torch.equal(batch.tokens.cpu(), one["tokens"].cpu())


def test_ll2_sampled_ids_follow_the_sampler_spec():
    """Slow id sampled by the team's CTA 0 (top-k / top-p / temperature, published per CTA), depth codes sampled too; and the
    mixed case (slow sampled, depth greedy from the candidate words)."""
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=1)
    prompt = prompt_grid(byte_prompt(24, seed=8), cfg)
    for fast_temp in (0.6, 0.0):
        padded, lens = pack_prompts(model, [prompt])
        batch = model.new_batch(1, max_positions=128, max_frames=12, seq_ids=[5])
        try:
            model.prefill(batch, padded, lens)
            s = model.sampling(temp=0.7, fast_temp=fast_temp, top_k=50, top_p=0.9, seed=1234, ignore_stop=True)
            for f in range(12):
                model.decode_frames(batch, s, 1)
                torch.cuda.synchronize()
                tl = model.debug_buffer("token_logits", 1).cpu().numpy()
                dl = model.debug_buffer("depth_logits", 1).cpu().numpy()
                got = batch.tokens.cpu().tolist()[0]
                want = [sample_row(tl[0], 0.7, 50, 0.9, 0.0, 1234, f, 5, 0)]
                want += [sample_row(dl[0, i], fast_temp, 0, 1.0, 0.0, 1234, f, 5, 1 + i) if fast_temp > 0 else int(dl[0, i].argmax())
                         for i in range(cfg.max_fast_seqlen)]
                assert got == want, f"fast_temp {fast_temp} frame {f}: {got} != {want}"
            # and the same ids when the 12 frames run inside one launch
            multi = _run(model, prompt, 12, [12], seq_ids=[5], temp=0.7, fast_temp=fast_temp, top_k=50, top_p=0.9, seed=1234)
            assert torch.equal(multi["codes"], batch.out_codes), f"fast_temp {fast_temp}: 12 frames in one launch differ from frame by frame"
        finally:
            batch.release()


def test_ll2_stop_rule_and_forced_ids_inside_a_launch():
    """<|im_end|> forced in the first of three frames of one launch: the sequence must freeze (tokens, seq_len, step, codes)
    for the rest of that launch and for the next one."""
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=1)
    R = cfg.n_rows
    prompt = prompt_grid(byte_prompt(14, seed=81), cfg)
    padded, lens = pack_prompts(model, [prompt])
    batch = model.new_batch(1, max_positions=64, max_frames=8)
    try:
        model.prefill(batch, padded, lens)
        s = model.sampling(audio_only=True)
        model.decode_frames(batch, s, 1)
        torch.cuda.synchronize()
        len1 = int(batch.seq_len.item())
        force = torch.zeros(1, R, dtype=torch.int32, device=model.device)
        force[0, 0] = model.token_config.im_end_id
        model.set_force(force)
        model.decode_frames(batch, s, 3)   # the stop fires in the first of these three frames
        model.set_force(None)
        model.decode_frames(batch, s, 2)
        torch.cuda.synchronize()
        assert batch.finished.tolist() == [1] and batch.step.tolist() == [2]
        assert int(batch.seq_len.item()) == len1 + 1
        assert batch.tokens[0, 0].item() == model.token_config.im_end_id
        assert batch.out_codes[0, 1, 0].item() == model.token_config.im_end_id and int(batch.out_codes[0, 2:].abs().sum()) == 0
    finally:
        model.set_force(None)
        batch.release()


def test_ll2_frame_clock_is_monotonic():
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=1)
    prompt = prompt_grid(byte_prompt(20, seed=90), cfg)
    ref = _run(model, prompt, 16, [16])
    clock = model.set_frame_clock(16)
    try:
        got = _run(model, prompt, 16, [16])
        torch.cuda.synchronize()
        ns = clock.cpu().tolist()
    finally:
        model.set_frame_clock(0)
    assert all(b > a > 0 for a, b in zip(ns, ns[1:])), ns
    assert (ns[-1] - ns[0]) / 15 < 5e6, "more than 5 ms per frame on the tiny model"
    assert torch.equal(ref["codes"], got["codes"])


@pytest.mark.parametrize("size,B", [("smoltts_byte_tiny", 8), ("smoltts_byte_70m", 3), ("smoltts_byte_150m", 5)])
def test_ll2_teams_decode_every_sequence_as_if_alone(size, B):
    """Batches of 2..8: one team of CTAs per sequence, each with its own word regions.  Rows over CTAs, K over warps: the
    arithmetic does not depend on the team size, so every sequence must decode bit for bit as it does alone on 148 CTAs --
    greedy and sampled (counters keyed by the global utterance id)."""
    from smoltts_b200 import GenerationSettings, generate_batch

    cfg, sd, model, orc = model_and_oracle(size, max_batch=8)
    assert model.get_option("ll_max_batch") == 8
    prompts = [prompt_grid(byte_prompt(14 + 9 * b, seed=120 + b), cfg) for b in range(B)]
    ids = list(range(50, 50 + B))
    for gs in (GenerationSettings(default_temp=0.0, default_fast_temp=0.0),
               GenerationSettings(default_temp=0.8, default_fast_temp=0.6, top_k=40, top_p=0.9, seed=77)):
        together = generate_batch(model, prompts, gs, audio_only=False, fixed_frames=10, chunk=4, seq_ids=ids)
        assert model.get_option("ll_ready") == 1
        for b in range(B):
            alone = generate_batch(model, prompts[b:b + 1], gs, audio_only=False, fixed_frames=10, seq_ids=ids[b:b + 1])[0]
            assert torch.equal(together[b], alone), f"{size}: sequence {b} of a batch of {B} differs from its solo decode"


def test_ll2_team_stop_rule_freezes_one_sequence():
    """<|im_end|> forced on sequence 1 in frame 2: its team freezes (tokens, seq_len, step) while sequence 0 keeps decoding
    inside the same launches."""
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=8)
    B, R = 2, cfg.n_rows
    prompts = [prompt_grid(byte_prompt(12 + b, seed=80 + b), cfg) for b in range(B)]
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
        assert batch.finished.tolist() == [0, 1] and batch.step.tolist() == [6, 2]
        assert batch.seq_len.tolist() == [int(lens[0]) - 1 + 6, int(lens[1]) - 1 + 2]
        assert batch.tokens[1, 0].item() == model.token_config.im_end_id
    finally:
        model.set_force(None)
        batch.release()
