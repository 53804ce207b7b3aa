#!/usr/bin/env python
"""Fabricate a checkpoint directory for a named model size: ``config.json``,
byte-level tokenizer files and seeded random ``model.safetensors`` in the
reference's exported layout.

The tokenizer follows the recipe of the reference's
data_pipeline/scripts/create_bytelevel_init.py:15-57 (256 byte tokens, 15 control
tokens, 49 speaker tokens, ``codebook_size`` semantic tokens) so that the
reference's ``RQTransformer.from_pretrained`` accepts the directory.

usage: python tools/make_init.py --size smoltts_byte_70m --out /tmp/init70 [--seed 0] [--no-weights]
"""
from __future__ import annotations

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from smoltts_b200.config import named_config  # noqa: E402
from smoltts_b200.synth import CONTROL_TOKENS, make_state_dict  # noqa: E402

CHATML_TEMPLATE = (
    "{% if not add_generation_prompt is defined %}{% set add_generation_prompt = false %}{% endif %}"
    "{% for message in messages %}{{'<|im_start|>' + message['role'] + '\n' + message['content'] + "
    "'<|im_end|>' + '\n'}}{% endfor %}{% if add_generation_prompt %}{{ '<|im_start|>assistant\n' }}{% endif %}"
)


def build_tokenizer(codebook_size: int):
    from tokenizers import Tokenizer, decoders, models
    from tokenizers.trainers import BpeTrainer

    tok = Tokenizer(models.BPE())
    trainer = BpeTrainer(vocab_size=256, special_tokens=[])
    tok.train_from_iterator([bytes([i]).decode("latin-1") for i in range(256)], trainer=trainer)
    tok.pre_tokenizer = None
    tok.normalizer = None
    tok.decoder = decoders.ByteLevel()
    speakers = [f"<|speaker:{i}|>" for i in range(64 - len(CONTROL_TOKENS))]
    semantic = [f"<|semantic:{i}|>" for i in range(codebook_size)]
    tok.add_special_tokens([*CONTROL_TOKENS, *speakers, *semantic])
    return tok


def write_init(size: str, out_dir: str, seed: int = 0, weights: bool = True, **overrides) -> None:
    from transformers import PreTrainedTokenizerFast

    cfg = named_config(size, **overrides)
    os.makedirs(out_dir, exist_ok=True)
    cfg.save(os.path.join(out_dir, "config.json"))
    tok = build_tokenizer(cfg.codebook_size)
    PreTrainedTokenizerFast(
        tokenizer_object=tok, bos_token="<|im_start|>", eos_token="<|endoftext|>",
        unk_token="<|unknown|>", pad_token="<|pad|>", chat_template=CHATML_TEMPLATE,
    ).save_pretrained(out_dir)
    if weights:
        from safetensors.torch import save_file

        save_file(make_state_dict(cfg, seed=seed), os.path.join(out_dir, "model.safetensors"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="smoltts_byte_70m")
    ap.add_argument("--out", "-o", required=True)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-weights", action="store_true")
    a = ap.parse_args()
    write_init(a.size, a.out, a.seed, not a.no_weights)
    print(f"wrote {a.out}")


if __name__ == "__main__":
    main()
