"""Host-side logic that needs no GPU: config derivation, checkpoint normalisation, synthetic inputs,
roofline denominators, RoPE tables, sampler specification and the Philox generator."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as so
from oracle.dualar_oracle import rope_table
from smoltts_b200.config import RQTransformerModelArgs, named_config
from smoltts_b200.shard import partition, shard
from smoltts_b200.synth import (byte_prompt, depthwise_to_flat, flat_to_depthwise, make_state_dict, prompt_grid,
                                state_dict_shapes, teacher_grid)


def test_config_derivation_follows_the_reference_rules():
    cfg = named_config("smoltts_byte_150m")
    assert (cfg.dim, cfg.n_head, cfg.n_local_heads, cfg.head_dim, cfg.intermediate_size) == (768, 12, 4, 64, 3072)
    assert cfg.duplicate_code_0 is True and cfg.max_fast_seqlen == 8 and cfg.n_rows == 9
    assert cfg.qkv_rows == 1280 and cfg.fast_embedding_rows == 14336
    c2 = RQTransformerModelArgs.from_dict({**cfg.to_dict(), "head_dim": 999, "duplicate_code_0": False, "unknown_key": 3})
    assert c2.head_dim == 64                      # JSON value is overwritten (modeling/...:65)
    assert c2.max_fast_seqlen == 7 and c2.n_rows == 8
    assert c2.extra == {"unknown_key": 3} and c2.to_dict()["unknown_key"] == 3


def test_roofline_denominators_match_the_survey():
    c150, c70 = named_config("smoltts_byte_150m"), named_config("smoltts_byte_70m")
    assert c150.layer_params() == 8_650_752
    assert c150.unique_weight_bytes() == 271_024_128
    assert c150.kv_bytes_per_position() == 10_240
    assert c150.flops_per_frame(0) == 755_466_240 and c150.flops_per_frame(100) == 755_466_240 + 30_720 * 100
    assert c70.unique_weight_bytes() == 120_692_736 and c70.kv_bytes_per_position() == 7_680
    assert c70.flops_per_frame(0) == 318_873_600


def test_checkpoint_layout_and_normalisation():
    from smoltts_b200.model import normalise_state_dict

    cfg = named_config("smoltts_byte_tiny")
    sd = make_state_dict(cfg, seed=3)
    shapes = state_dict_shapes(cfg)
    assert set(sd) == set(shapes) and all(tuple(sd[k].shape) == shapes[k] for k in sd)
    assert all(v.dtype == torch.bfloat16 for v in sd.values())
    # trainer form: 3-D fast_output, _orig_mod. prefixes, split wq/wk/wv
    hd = cfg.head_dim
    raw = {}
    for k, v in sd.items():
        if k == "fast_output.weight":
            raw["_orig_mod." + k] = flat_to_depthwise(v, cfg)
        elif k.endswith("attention.wqkv.weight"):
            fast = k.startswith("fast_")
            nh = cfg.fast_n_head if fast else cfg.n_head
            nkv = cfg.fast_n_local_heads if fast else cfg.n_local_heads
            q, kk, vv = v.split([nh * hd, nkv * hd, nkv * hd], dim=0)
            pre = "_orig_mod." + k[: -len("wqkv.weight")]
            raw[pre + "wq.weight"], raw[pre + "wk.weight"], raw[pre + "wv.weight"] = q, kk, vv
        else:
            raw["_orig_mod." + k] = v
    back = normalise_state_dict(raw, cfg)
    assert set(back) == set(sd)
    for k in sd:
        assert torch.equal(back[k], sd[k]), k
    w3 = flat_to_depthwise(sd["fast_output.weight"], cfg)
    assert w3.shape == (cfg.max_fast_seqlen, cfg.fast_dim, cfg.codebook_size)
    assert torch.equal(depthwise_to_flat(w3), sd["fast_output.weight"])
    # row i*C + k of the flat form is column k of W_i (train/convert_safetensors.py:12-15)
    assert torch.equal(sd["fast_output.weight"][1 * cfg.codebook_size + 5], w3[1, :, 5])


def test_convert_cli_round_trip(tmp_path):
    """python -m smoltts_b200.convert: trainer checkpoint (.pt, 3-D fast_output, _orig_mod. prefixes) -> config.json +
    model.safetensors in the exported layout; wrong shapes are refused (train/convert_safetensors.py:6-16 hard-codes 768)."""
    from safetensors.torch import load_file

    from smoltts_b200.convert import main as convert_main

    cfg = named_config("smoltts_byte_tiny")          # hidden size 128: the reference converter would mis-shape this one
    sd = make_state_dict(cfg, seed=5)
    raw = {"_orig_mod." + k: (flat_to_depthwise(v, cfg) if k == "fast_output.weight" else v).float() for k, v in sd.items()}
    ckpt = tmp_path / "checkpoint.pt"
    torch.save({"model_state_dict": raw, "step": 7}, ckpt)
    cfg.save(str(tmp_path / "config.json"))
    out = tmp_path / "export"
    assert convert_main([str(ckpt), "--config", str(tmp_path / "config.json"), "-o", str(out)]) == 0
    got = load_file(str(out / "model.safetensors"))
    assert set(got) == set(sd) and (out / "config.json").exists()
    for k in sd:
        assert got[k].dtype == torch.bfloat16 and torch.equal(got[k], sd[k]), k
    raw["_orig_mod.norm.weight"] = torch.ones(cfg.dim + 1)
    torch.save(raw, ckpt)                              # a bare state dict is accepted too; the bad shape is not
    with pytest.raises(ValueError, match="norm.weight"):
        convert_main([str(ckpt), "--config", str(tmp_path / "config.json"), "-o", str(out)])


def test_synthetic_inputs_are_deterministic_and_well_formed():
    cfg = named_config("smoltts_byte_150m")
    p = byte_prompt(200, seed=1)
    assert len(p) == 212 and p == byte_prompt(200, seed=1) and p != byte_prompt(200, seed=2)
    assert p[0] == 269 and p[4] == 270 and p[-4:] == [270, 269, 258, 10]
    g = prompt_grid(p, cfg)
    assert g.shape == (9, 212) and int(g[1:].abs().sum()) == 0
    t = teacher_grid(cfg, 10, 6, batch=2, zero_code_at=3)
    assert t.shape == (2, 9, 16)
    assert torch.equal(t[:, 0, 10:], 320 + t[:, 1, 10:])      # dup0: row 0 = 320 + first depth code
    assert int(t[0, 1, 13]) == 0 and int(t[:, 1:, :10].abs().sum()) == 0


def test_rope_table_is_bit_identical_to_the_oracle_restatement():
    from smoltts_b200.model import precompute_freqs_cis

    a = precompute_freqs_cis(300, 64, 100000)
    b = rope_table(300, 64, 100000)
    assert a.dtype == torch.bfloat16 and torch.equal(a.view(torch.int16), b.view(torch.int16))


def test_partition_is_contiguous_and_balanced():
    for n in (0, 1, 7, 8, 9, 256, 2048):
        for w in (1, 2, 4, 8):
            parts = partition(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    items, ids = shard(list("abcdefghij"), 1, 4)
    assert items == ["d", "e", "f"] and ids == [3, 4, 5]


# ---- sampler specification ---------------------------------------------------------------------------
def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32, 10 rounds."""
    assert so.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = 0xFFFFFFFF
    assert so.philox4x32_10((f, f, f, f), (f, f)) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert so.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_exp_det_weights_track_exp_and_are_monotone():
    z = np.linspace(-25, 0, 2001).astype(np.float32)
    w = so.exp_det_weights(z).astype(np.float64)
    ref = np.exp(z.astype(np.float64)) * 2.0 ** 30
    live = z >= -21.5
    assert w[-1] == 2 ** 30 and (w[~live] == 0).all()
    assert np.abs(w[live] - ref[live]).max() <= 1.0 + 2e-6 * ref[live].max()
    assert (np.diff(w) >= 0).all()


def test_filters_and_draws():
    rng = np.random.default_rng(0)
    logits = rng.normal(size=500).astype(np.float32) * 2
    assert so.sample_row(logits, 0.0, 0, 1.0, 0.0, 1, 0, 0, 0) == int(np.argmax(logits))
    for step in range(20):  # top_k = 1 is argmax whatever the draw
        assert so.sample_row(logits, 1.0, 1, 1.0, 0.0, 7, step, 3, 0) == int(np.argmax(logits))
    w = so.exp_det_weights(logits)
    keep = so.filter_weights(w, top_k=10)
    assert keep.sum() == 10 and w[keep].min() >= w[~keep].max()
    keep = so.filter_weights(w, top_p=0.5)
    kept, total = w[keep].astype(np.float64).sum(), w.astype(np.float64).sum()
    assert kept >= 0.5 * total - 1 and (kept - w[keep].min()) < 0.5 * total   # smallest prefix reaching top_p
    p = so.kept_distribution(logits, 0.8, 50, 0.9, 0.0)
    assert abs(p.sum() - 1) < 1e-12 and (p > 0).sum() <= 50
    draws = {so.sample_row(logits, 0.8, 50, 0.9, 0.0, 11, s, 0, 0) for s in range(200)}
    assert len(draws) > 5 and all(p[d] > 0 for d in draws)
    # counters: same (seed, step, seq, stream) -> same id; any change -> (almost surely) another stream of draws
    a = [so.sample_row(logits, 1.0, 0, 1.0, 0.0, 5, s, 2, 1) for s in range(50)]
    assert a == [so.sample_row(logits, 1.0, 0, 1.0, 0.0, 5, s, 2, 1) for s in range(50)]
    assert a != [so.sample_row(logits, 1.0, 0, 1.0, 0.0, 5, s, 3, 1) for s in range(50)]
