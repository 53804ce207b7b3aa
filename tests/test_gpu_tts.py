"""Text -> PCM through both engines (smoltts_b200/tts.py: SmolTTS, the mirror of the reference's front object): the codes of
the DualAR decode step go through the Mimi streaming decoder, checked against the Mimi oracle on the same codes."""
import numpy as np
import pytest
import torch

from gpu_util import model_and_oracle

pytestmark = pytest.mark.gpu


def _tts(max_streams=4, max_frames=32):
    from smoltts_b200 import GenerationSettings, MimiModel, PromptEncoder, SmolTTS, byte_level_tokenizer
    from smoltts_b200.synth import make_mimi_state_dict

    cfg, sd, lm, _ = model_and_oracle("smoltts_byte_tiny", max_batch=8, max_seq_len=256)
    enc = PromptEncoder.from_model(byte_level_tokenizer(cfg.codebook_size), lm)
    msd = make_mimi_state_dict(0)
    stream = MimiModel(max_streams=max_streams, max_frames=max_frames)
    stream.load_state_dict(msd)
    full = MimiModel(max_streams=max_streams, max_frames=max_frames, upsample_carry=True)
    full.load_state_dict(msd)
    gs = GenerationSettings(default_temp=0.0, default_fast_temp=0.0, max_new_tokens=9)
    return SmolTTS(lm, enc, stream, full, gs), msd


def _audio_codes(tts, text, voice):
    from smoltts_b200 import generate_blocking

    return generate_blocking(tts.lm, tts._get_prompt(text, voice), tts.settings).cpu().long()   # [1, 8, T]


def test_call_and_stream_follow_the_two_codec_rules_of_the_reference():
    from oracle.mimi_oracle import MimiOracle, StreamState

    tts, msd = _tts()
    orc = MimiOracle(msd)
    text, voice = "Hello there, B200.", "nova"
    codes = _audio_codes(tts, text, voice)
    T = codes.shape[-1]
    assert T >= 1, "the tiny random-init model emitted no audio frame; pick another text"
    with torch.no_grad():
        want_full = orc.decode(codes).flatten().numpy()
        st = StreamState()
        want_stream = [orc.decode_step(codes[:, :, t:t + 1], st).flatten().numpy() for t in range(T)]
    pcm = tts(text, voice)
    assert pcm.shape == (T * 1920,) and pcm.dtype == np.float32
    assert np.abs(pcm - want_full).max() <= 1e-4 * np.abs(want_full).max()
    chunks = list(tts.stream(text, voice))
    assert len(chunks) == T and all(c.shape == (1920,) for c in chunks)
    for t, (c, w) in enumerate(zip(chunks, want_stream)):
        assert np.abs(c - w).max() <= 1e-4 * np.abs(np.concatenate(want_stream)).max(), f"chunk {t}"
    assert len(tts.codec._free) == tts.codec.max_streams, "stream() must give its codec slot back"
    st = {}
    plain = list(tts.stream(text, voice, overlap=False, stats=st))      # codec step after the decode step instead of beside the next
    assert st["audio_frames"] == T and len(plain) == T and all(np.array_equal(a, b) for a, b in zip(plain, chunks))
    assert tts.lm.get_option("n_ctas_override") == 0, "stream() must restore the decode kernel's grid"


def test_batch_synthesis_equals_one_utterance_at_a_time():
    tts, _ = _tts()
    texts = ["One.", "A second, longer utterance.", "Three!"]
    voices = ["heart", "sky", "liam"]
    outs = tts.synthesize_batch(texts, voices, fixed_frames=6)
    assert len(outs) == 3
    from smoltts_b200 import generate_batch

    prompts = [tts._get_prompt(t, v)[0] for t, v in zip(texts, voices)]
    codes = generate_batch(tts.lm, prompts, tts.settings, fixed_frames=6)
    assert any(c.shape[-1] > 0 for c in codes)
    for b in range(3):
        solo = tts.decode_codes([codes[b]])[0]
        assert outs[b].shape == (codes[b].shape[-1] * 1920,)
        assert np.array_equal(outs[b], solo), f"utterance {b}: PCM differs between the batch of streams and a stream alone"


def test_serving_with_continuous_batching_yields_the_same_audio():
    """SmolTTS.serve: 7 utterances through 3 decode slots (admitted late, retired early) and then the codec, against
    synthesize_batch of the same texts -- the same PCM, bit for bit (greedy decoding, both code paths keep an utterance's codes
    and every codec call here stays below 12 streams, one kernel class)."""
    tts, _ = _tts()
    texts = [f"Utterance number {i}, of some length {'.' * (3 * i)}" for i in range(7)]
    voices = ["heart", "bella", "nova", "sky", "sarah", "michael", "liam"]
    got = dict(tts.serve(texts, voices, slots=3, chunk=4, max_prompt=128))
    assert sorted(got) == list(range(7))
    want = tts.synthesize_batch(texts[:4], voices[:4]) + tts.synthesize_batch(texts[4:], voices[4:])
    for i in range(7):
        assert got[i].shape == want[i].shape, (i, got[i].shape, want[i].shape)
        assert np.array_equal(got[i], want[i]), f"utterance {i}: served audio differs from batch synthesis"


def test_streamed_serving_chunks_concatenate_to_the_served_audio():
    """SmolTTS.serve_stream: a codec stream per decode slot, audio handed out after every chunk of 4 frames; the chunks of an
    utterance concatenate to what serve / synthesize_batch return for it, bit for bit."""
    tts, _ = _tts()
    texts = [f"Streamed utterance {i} {'!' * i}" for i in range(5)]
    voices = ["heart", "nova", "sky", "liam", "emma"]
    parts, finished = {i: [] for i in range(5)}, set()
    n_items = 0
    for i, pcm, done in tts.serve_stream(texts, voices, slots=3, chunk=4, max_prompt=128):
        assert i not in finished, "audio after an utterance was reported done"
        parts[i].append(pcm)
        n_items += 1
        if done:
            finished.add(i)
    assert finished == set(range(5)) and n_items > 5, "every utterance must finish, in more than one chunk"
    want = tts.synthesize_batch(texts[:3], voices[:3]) + tts.synthesize_batch(texts[3:], voices[3:])
    for i in range(5):
        got = np.concatenate(parts[i]) if parts[i] else np.zeros(0, dtype=np.float32)
        assert got.shape == want[i].shape, (i, got.shape, want[i].shape)
        assert np.array_equal(got, want[i]), f"utterance {i}: streamed chunks differ from batch synthesis"
    assert len(tts.codec_full._free) == tts.codec_full.max_streams, "serve_stream must give its codec slots back"
