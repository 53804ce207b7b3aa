"""The product multi-GPU entry point on real engines: worker processes with a CUDA engine each, utterances sharded,
codes gathered on the caller -- identical to one in-process generate_batch call (sampling is keyed by the global
utterance id, so the sharding cannot change an id)."""
import pytest
import torch

from smoltts_b200.synth import byte_prompt, prompt_grid

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)
def test_generate_sharded_equals_in_process_generate_batch():
    from smoltts_b200 import GenerationSettings, RQTransformer, generate_batch, generate_sharded, named_config
    from smoltts_b200.synth import make_state_dict

    cfg = named_config("smoltts_byte_tiny")
    prompts = [prompt_grid(byte_prompt(10 + 3 * b, seed=200 + b), cfg) for b in range(5)]
    gs = GenerationSettings(default_temp=0.8, default_fast_temp=0.7, top_k=40, top_p=0.9, seed=7)
    model = RQTransformer(cfg, max_batch=8, max_seq_len=128)
    model.load_state_dict(make_state_dict(cfg, seed=0))
    want = generate_batch(model, prompts, gs, audio_only=False, fixed_frames=6, seq_ids=list(range(5)))
    gpus = min(2, torch.cuda.device_count())
    spec = {"model": "smoltts_byte_tiny", "seed": 0, "engine": {"max_batch": 8, "max_seq_len": 128}}
    got = generate_sharded(spec, prompts, gs, gpus=gpus, audio_only=False, fixed_frames=6)
    assert len(got) == 5
    for b, (g, w) in enumerate(zip(got, want)):
        assert torch.equal(g, w.cpu()), f"utterance {b} differs between the sharded and the in-process run"
