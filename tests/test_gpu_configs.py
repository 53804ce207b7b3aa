"""Scaled-down runs of BASELINE.json's configs 3-5 (the batched / sampled / long-form cases) checked through
size-independent properties, since the CPU oracle cannot run them at full size in seconds:

  * config 3 (bs=64, top-k / top-p / temperature): every id the batch kernel emits equals what the CPU sampler
    specification picks from the logits the kernel dumped, with the counters (seed, step, seq_id, stream) it must use;
  * config 4 (bs=256 per GPU, sharded): a sequence decodes to the same ids inside the batch of 256, inside its
    128-sequence shard (what a second rank would hold: no collective on the decode path) and inside a batch of 16
    (another tile position and weight-block size of the tensor-core variant);
  * config 5 (long-form, paged KV): thousands of frames across hundreds of KV pages at bs=32; a probe sequence equals its
    decode inside a batch of 16, and the cached positions / page table arithmetic hold (seq_len, step).

Batches of 9+ run on the tcgen05 variant, whose dot products sum in the tensor core's order: results are bit-identical
across batch compositions, tile positions and launch modes of that variant (checked here), and agree with the bs=1
data-flow kernel / the oracle up to near-tie decisions (tests/test_gpu_tc.py).
"""
import pytest
import torch

from gpu_util import model_and_oracle
from oracle.sampler_oracle import sample_row
from smoltts_b200 import GenerationSettings, generate_batch
from smoltts_b200.generate import pack_prompts
from smoltts_b200.synth import byte_prompt, prompt_grid

pytestmark = pytest.mark.gpu


def test_config3_bs64_sampled_ids_follow_the_sampler_spec():
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_150m", max_batch=64, max_seq_len=320)
    B, n_frames = 64, 3
    seq_ids = list(range(100, 100 + B))
    prompts = [prompt_grid(byte_prompt(200, seed=1 + b), cfg) for b in range(B)]
    padded, lens = pack_prompts(model, prompts)
    batch = model.new_batch(B, max_positions=256, max_frames=n_frames, seq_ids=seq_ids)
    try:
        model.prefill(batch, padded, lens)
        s = model.sampling(temp=0.7, fast_temp=0.7, top_k=50, top_p=0.9, seed=1234, ignore_stop=True)
        for f in range(n_frames):
            model.decode_frames(batch, s, 1)
            torch.cuda.synchronize()
            tl = model.debug_buffer("token_logits", B).cpu().numpy()
            dl = model.debug_buffer("depth_logits", B).cpu().numpy()
            got = batch.tokens.cpu().tolist()
            for b in range(0, B, 7):
                want = [sample_row(tl[b], 0.7, 50, 0.9, 0.0, 1234, f, seq_ids[b], 0)]
                want += [sample_row(dl[b, i], 0.7, 0, 1.0, 0.0, 1234, f, seq_ids[b], 1 + i) for i in range(cfg.max_fast_seqlen)]
                assert got[b] == want, f"frame {f} seq {b}: {got[b]} != {want}"
        assert batch.step.tolist() == [n_frames] * B
        probe = 17
        ids_in_batch = batch.out_codes[probe, :n_frames].cpu()
    finally:
        batch.release()
    # the same utterance inside another batch (other tile row, other weight-block size): sampling counters are keyed by
    # the global seq_id, so the ids must not depend on batch composition
    gs = GenerationSettings(default_temp=0.7, default_fast_temp=0.7, top_k=50, top_p=0.9, seed=1234)
    sub = generate_batch(model, prompts[probe - 5:probe + 11], gs, audio_only=False, fixed_frames=n_frames,
                         seq_ids=seq_ids[probe - 5:probe + 11])
    assert torch.equal(sub[5].t().contiguous(), ids_in_batch)


def test_config4_bs256_shards_are_independent():
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_150m", max_batch=256, max_seq_len=128)
    B, n_frames = 256, 2
    prompts = [prompt_grid(byte_prompt(64, seed=1 + b), cfg) for b in range(B)]   # the 64-byte-prompt point of SURVEY 8(d)
    gs = GenerationSettings(default_temp=0.0, default_fast_temp=0.0)
    whole = generate_batch(model, prompts, gs, audio_only=False, fixed_frames=n_frames, seq_ids=list(range(B)))
    shard1 = generate_batch(model, prompts[128:], gs, audio_only=False, fixed_frames=n_frames, seq_ids=list(range(128, B)))
    for b in range(128, B):
        assert torch.equal(whole[b], shard1[b - 128]), f"sequence {b} differs between the batch of 256 and its shard"
    for lo in (0, 120, 240):
        sub = generate_batch(model, prompts[lo:lo + 16], gs, audio_only=False, fixed_frames=n_frames, seq_ids=list(range(lo, lo + 16)))
        for i in range(16):
            assert torch.equal(sub[i], whole[lo + i]), f"sequence {lo + i}: batch of 256 vs batch of 16"
    assert model.get_option("tc_ready") == 1


def test_config5_long_form_paged_kv_bs32():
    cfg, sd, model, orc = model_and_oracle("smoltts_byte_tiny", max_batch=32, max_seq_len=4400)
    B, n_frames = 32, 4096
    prompts = [prompt_grid(byte_prompt(100 + b, seed=300 + b), cfg) for b in range(B)]
    gs = GenerationSettings(default_temp=0.0, default_fast_temp=0.0)
    outs, batch = generate_batch(model, prompts, gs, audio_only=False, fixed_frames=n_frames, chunk=1024, return_batch=True)
    try:
        assert batch.step.tolist() == [n_frames] * B
        assert batch.seq_len.tolist() == [100 + b + 12 - 1 + n_frames for b in range(B)]   # prompt columns - 1 + frames
        assert batch.max_pages * model.page_size >= 4200 and batch.max_pages > 128        # hundreds of pages per sequence
    finally:
        batch.release()
    probe = 5
    sub = generate_batch(model, prompts[:16], gs, audio_only=False, fixed_frames=n_frames, chunk=512, seq_ids=list(range(16)))
    same = (sub[probe] == outs[probe]).all(dim=0)
    assert bool(same.all()), f"long-form probe diverges from its decode in a batch of 16 at frame {int((~same).nonzero()[0])}"
