"""Prompt encoding (CPU): the in-memory byte tokenizer of the reference recipe and the PromptEncoder mirror of
mlx_inference/src/smoltts_mlx/lm/utils/prompt.py:10-63 / smoltts_mlx/__init__.py:97-151."""
import pytest
import torch

from smoltts_b200 import named_config
from smoltts_b200.serving import PromptEncoder, byte_level_tokenizer
from smoltts_b200.synth import (TOK_ASSISTANT, TOK_IM_END, TOK_IM_START, TOK_SEMANTIC0, TOK_SPEAKER0, TOK_SYSTEM, TOK_USER,
                                byte_prompt, prompt_grid)

tokenizers = pytest.importorskip("tokenizers")


@pytest.fixture(scope="module")
def tok():
    return byte_level_tokenizer(2048)


def test_byte_tokenizer_ids_follow_the_reference_recipe(tok):
    # data_pipeline/scripts/create_bytelevel_init.py:15-57: 256 bytes, 15 control + 49 speaker tokens, 2048 semantic tokens
    assert tok.get_vocab_size() == named_config("smoltts_byte_150m").vocab_size == 2368
    assert [tok.token_to_id(chr(i)) for i in (0, 10, 65, 233, 255)] == [0, 10, 65, 233, 255]
    assert tok.token_to_id("system") == TOK_SYSTEM and tok.token_to_id("assistant") == TOK_ASSISTANT
    assert tok.token_to_id("<|im_start|>") == TOK_IM_START and tok.token_to_id("<|im_end|>") == TOK_IM_END
    assert tok.token_to_id("<|speaker:0|>") == TOK_SPEAKER0 and tok.token_to_id("<|speaker:48|>") == TOK_SPEAKER0 + 48
    assert tok.token_to_id("<|semantic:0|>") == TOK_SEMANTIC0 and tok.token_to_id("<|semantic:2047|>") == TOK_SEMANTIC0 + 2047


@pytest.mark.parametrize("dup0", [True, False])
def test_prompt_encoder_grids(tok, dup0):
    cfg = named_config("smoltts_byte_150m", duplicate_code_0=dup0)
    enc = PromptEncoder(tok, semantic_offset=TOK_SEMANTIC0, num_codebooks=cfg.num_codebooks, duplicate_code_0=dup0)
    assert enc.depth == cfg.n_rows - 1
    turn = enc.encode_text_turn("user", "Hi")
    assert turn.shape == (cfg.n_rows, 6) and turn.dtype == torch.int64
    assert turn[0].tolist() == [TOK_IM_START, TOK_USER, 10, ord("H"), ord("i"), TOK_IM_END]
    assert int(turn[1:].abs().sum()) == 0
    open_turn = enc.encode_text_turn("assistant")
    assert open_turn[0].tolist() == [TOK_IM_START, TOK_ASSISTANT, 10]
    # a spoken turn: row 0 = semantic_offset + code 0 of every frame, the last `depth` codebook rows below, then <|im_end|>\n
    codes = torch.arange(8 * 5).view(8, 5) % 2048
    vq = enc.encode_vq(codes)
    assert vq.shape == (cfg.n_rows, 5 + 2)
    assert vq[0, :5].tolist() == (codes[0] + TOK_SEMANTIC0).tolist()
    assert torch.equal(vq[1:, :5], codes[8 - enc.depth:])
    assert vq[0, 5:].tolist() == [TOK_IM_END, 10] and int(vq[1:, 5:].abs().sum()) == 0
    with pytest.raises(ValueError):
        enc.encode_vq(codes[None])


def test_tts_prompt_equals_the_synthetic_bench_prompt(tok):
    """The bench's synthetic prompt (synth.byte_prompt) is this framing with random latin-1 bytes as the text."""
    cfg = named_config("smoltts_byte_150m")
    enc = PromptEncoder(tok, semantic_offset=TOK_SEMANTIC0)
    row0 = byte_prompt(40, seed=3, speaker=5)
    text = "".join(chr(b) for b in row0[8:48])
    got = enc.tts_prompt(text, speaker=5)
    assert torch.equal(got, prompt_grid(row0, cfg))
    # a voice-clone prefix replaces the speaker token (smoltts_mlx/__init__.py:97-115)
    sample = {"text": "hello", "codes": torch.randint(0, 2048, (8, 6))}
    prefix = enc.create_speaker([sample], system_prompt="narrator")
    cloned = enc.tts_prompt("world", sysprompt=prefix)
    assert cloned.shape[1] == prefix.shape[1] + enc.encode_text_turn("user", "world").shape[1] + 3
    assert cloned[0, 0] == TOK_IM_START and cloned[0, 1] == TOK_SYSTEM
    with pytest.raises(ValueError):
        enc.create_speaker([{"text": "no audio"}])
