"""Mimi streaming decoder on the GPU (csrc/mimi_kernels.cu through include/smoltts_b200_mimi.h) against the goldens made from
transformers' MimiModel and against the CPU oracle (oracle/mimi_oracle.py) on the same seeded weights.

Tolerance: everything is fp32 on both sides and differs only in summation order (dot products of up to 3584 terms):
max |error| <= 1e-4 x max |reference| over the compared tensor (measured: a few 1e-6)."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu

REL = 1e-4
_CACHE = {}


def _sd():
    from smoltts_b200.synth import make_mimi_state_dict

    if "sd" not in _CACHE:
        _CACHE["sd"] = make_mimi_state_dict(0)
    return _CACHE["sd"]


def _model(**kw):
    from smoltts_b200.mimi import MimiModel

    key = tuple(sorted(kw.items()))
    if key not in _CACHE:
        m = MimiModel(**kw)
        m.load_state_dict(_sd())
        _CACHE[key] = m
    return _CACHE[key]


def _close(name, got, want, rel=REL):
    got = torch.as_tensor(got).float().cpu()
    want = torch.as_tensor(want).float().cpu()
    assert got.shape == want.shape, (name, tuple(got.shape), tuple(want.shape))
    assert torch.isfinite(got).all(), f"{name}: non-finite samples"
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    print(f"{name}: max abs error {err:.3g} (scale {scale:.3g}, rel {err / scale:.2g})")
    assert err <= rel * scale, f"{name}: max abs error {err:.3g} vs scale {scale:.3g}"


def _run(m, codes, caches=None):
    """codes [B, n_q, T] through decode_step, frame by frame."""
    own = caches is None
    caches = caches or [m.make_cache() for _ in range(codes.shape[0])]
    out = [m.decode_step(codes[:, :, t:t + 1].cuda(), caches).clone() for t in range(codes.shape[-1])]
    if own:
        for c in caches:
            m.release_cache(c)
    return torch.cat(out, dim=-1)


def test_steps_match_transformers_goldens_under_both_upsampling_rules():
    g = np.load(f"{GOLDEN_DIR}/mimi.npz")
    codes = torch.from_numpy(g["codes"]).long()
    ref_rule = _model(max_streams=4, max_frames=64)
    pcm = _run(ref_rule, codes)
    assert pcm.shape == (2, 1, 4 * 1920)
    _close("decode_step rule vs transformers (frame-wise upsample)", pcm, g["stream"])
    assert ref_rule.launches_per_step == 2 + 5 * 8 + 1 + 3 * 4 + 1 + 1
    carry = _model(max_streams=4, max_frames=64, upsample_carry=True)
    _close("carry rule vs MimiModel.decode", _run(carry, codes), g["full"])
    _close("decode() helper", carry.decode(codes), g["full"])


def test_transformer_output_matches_goldens():
    g = np.load(f"{GOLDEN_DIR}/mimi.npz")
    codes = torch.from_numpy(g["codes"]).long()
    m = _model(max_streams=4, max_frames=64)
    caches = [m.make_cache() for _ in range(2)]
    for t in range(codes.shape[-1]):
        m.decode_step(codes[:, :, t:t + 1].cuda(), caches)
        for b, c in enumerate(caches):
            _close(f"xf frame {t} row {b}", m.debug_rows("xf", c.slot, 2, 512), g["xf_stream"][b, 2 * t: 2 * t + 2])
    for c in caches:
        m.release_cache(c)


def test_long_run_matches_the_oracle_with_and_without_window():
    from oracle.mimi_oracle import MimiOracle, StreamState

    gen = torch.Generator().manual_seed(11)
    T = 24
    codes = torch.randint(0, 2048, (1, 8, T), generator=gen)
    for window in (0, 6):
        orc = MimiOracle(_sd(), window=window)
        st = StreamState()
        with torch.no_grad():
            want = torch.cat([orc.decode_step(codes[:, :, t:t + 1], st) for t in range(T)], dim=-1)
        m = _model(max_streams=2, max_frames=32, window=window)
        _close(f"{T} frames, window {window}", _run(m, codes), want)
    # the window really bites
    assert not torch.allclose(_run(_model(max_streams=2, max_frames=32, window=6), codes), _run(_model(max_streams=2, max_frames=32, window=0), codes))


def _shuffled_run(m, codes, perm, split):
    """The same streams stepped in permuted row order on even frames and as two separate calls on odd frames."""
    B, T = codes.shape[0], codes.shape[-1]
    caches = [m.make_cache() for _ in range(B)]
    inv = torch.argsort(torch.tensor(perm)).cuda()
    out = []
    for t in range(T):
        if t % 2 == 0:
            out.append(m.decode_step(codes[perm, :, t:t + 1].cuda(), [caches[i] for i in perm]).clone()[inv])
        else:
            a = m.decode_step(codes[:split, :, t:t + 1].cuda(), caches[:split]).clone()
            b2 = m.decode_step(codes[split:, :, t:t + 1].cuda(), caches[split:]).clone()
            out.append(torch.cat([a, b2], dim=0))
    for c in caches:
        m.release_cache(c)
    return torch.cat(out, dim=-1)


def test_history_longer_than_one_attention_split_matches_the_oracle():
    """270 frames = 540 positions: the history crosses the 512-position split of the attention kernel (two CTAs per head whose
    partial softmaxes are combined), without a window and with one of 20 positions that slides past the first split."""
    from oracle.mimi_oracle import MimiOracle, StreamState

    gen = torch.Generator().manual_seed(13)
    T = 270
    codes = torch.randint(0, 2048, (1, 8, T), generator=gen)
    for window in (0, 20):
        orc = MimiOracle(_sd(), window=window)
        st = StreamState()
        with torch.no_grad():
            want = torch.cat([orc.decode_step(codes[:, :, t:t + 1], st) for t in range(T)], dim=-1)
        m = _model(max_streams=2, max_frames=272, window=window)
        got = _run(m, codes)
        _close(f"{T} frames, window {window}: all", got, want)
        _close(f"{T} frames, window {window}: frames past position 512", got[..., 257 * 1920:], want[..., 257 * 1920:])
        assert not m.overflowed()


def test_a_stream_decodes_to_the_same_bits_alone_in_any_batch_slot_and_launch_mode():
    """Two kernel classes, chosen by the batch size alone: fewer than 12 streams (rows_kernel everywhere) and 12+ (the SEANet's
    many-row stages on tile_kernel, the last convolution on rowdot_kernel).  Inside a class a stream's PCM is bit-identical
    whatever shares the launch; across the classes only the summation order differs."""
    gen = torch.Generator().manual_seed(5)
    T = 5
    m = _model(max_streams=40, max_frames=16)
    eager = _model(max_streams=40, max_frames=16, mode="eager")
    # ---- few streams ----
    codes = torch.randint(0, 2048, (5, 8, T), generator=gen)
    batch = _run(m, codes)
    assert torch.equal(_run(eager, codes), batch), "graph replay and plain launches differ"
    for b in (0, 2, 4):
        assert torch.equal(_run(m, codes[b:b + 1]), batch[b:b + 1]), f"stream {b} decodes differently alone"
    assert torch.equal(_shuffled_run(m, codes, [3, 0, 4, 1, 2], 2), batch)
    # ---- many streams: 27 (ragged 64-row tiles and 16-row chunks), split 13 + 14 on odd frames ----
    codes27 = torch.cat([codes, torch.randint(0, 2048, (22, 8, T), generator=gen)], dim=0)
    big = _run(m, codes27)
    assert torch.equal(_run(eager, codes27), big), "graph replay and plain launches differ (tile kernel)"
    assert torch.equal(_run(m, codes27[:13]), big[:13]) and torch.equal(_run(m, codes27[13:]), big[13:])
    perm = [3, 0, 8, 1, 7, 2, 6, 4, 5, 16, 9, 15, 10, 14, 11, 13, 12, 26, 17, 25, 18, 24, 19, 23, 20, 22, 21]
    assert torch.equal(_shuffled_run(m, codes27, perm, 13), big)
    codes40 = torch.cat([codes27, torch.randint(0, 2048, (13, 8, T), generator=gen)], dim=0)
    huge = _run(m, codes40)
    assert torch.equal(huge[:27], big), "streams decode differently in a batch of 40 than in one of 27"
    # ---- across the classes: same streams, different summation order ----
    _close("5 streams inside a batch of 27 vs their own batch", big[:5], batch, rel=1e-5)


def test_slot_reuse_and_capacity():
    gen = torch.Generator().manual_seed(3)
    codes = torch.randint(0, 2048, (1, 8, 4), generator=gen)
    m = _model(max_streams=1, max_frames=4)
    first = _run(m, codes)
    assert torch.equal(_run(m, codes), first), "a reused slot remembers its previous stream"
    c = m.make_cache()
    for t in range(4):
        m.decode_step(codes[:, :, t:t + 1].cuda(), c)
    with pytest.raises(RuntimeError, match="max_frames"):
        m.decode_step(codes[:, :, :1].cuda(), c)
    with pytest.raises(RuntimeError, match="slots are in use"):
        m.make_cache()
    m.release_cache(c)
    assert not m.overflowed()


def test_a_differently_shaped_codec_matches_the_oracle():
    """Nothing in the kernels is specific to kyutai/mimi's sizes: a small codec (4 codebooks of 64 x 32, d = 128, 2 heads, 3
    transformer layers, SEANet 8 filters with ratios 4, 3, kernels 5 / 3 / 5: 24 samples per frame, channel counts down to 4)
    against the oracle, in both kernel classes (2 streams; 13 streams: tile_kernel / rowdot_kernel where they apply)."""
    from oracle.mimi_oracle import MimiDims, MimiOracle, StreamState
    from smoltts_b200.mimi import MimiConfig, MimiModel, MimiTransformerConfig, RVQConfig, SeanetConfig
    from smoltts_b200.synth import make_mimi_state_dict

    sd = make_mimi_state_dict(5, n_q=4, codebook_size=64, codebook_dim=32, dim=128, n_layers=3, ffn=256, n_filters=8, ratios=(4, 3),
                              kernel=5, res_kernel=3, last_kernel=5)
    dims = MimiDims(n_q=4, codebook_size=64, codebook_dim=32, dim=128, n_layers=3, n_heads=2, head_dim=64, ffn=256, n_filters=8,
                    ratios=(4, 3), kernel=5, res_kernel=3, last_kernel=5)
    cfg = MimiConfig(seanet=SeanetConfig(dimension=128, n_filters=8, kernel_size=5, residual_kernel_size=3, last_kernel_size=5, ratios=[4, 3]),
                     transformer=MimiTransformerConfig(d_model=128, num_heads=2, head_dim=64, num_layers=3, dim_feedforward=256),
                     rvq=RVQConfig(codebook_size=64, codebook_dim=32, num_quantizers=4, hidden_dim=128))
    gen = torch.Generator().manual_seed(17)
    T = 6
    for B, carry in ((2, False), (13, True)):
        codes = torch.randint(0, 64, (B, 4, T), generator=gen)
        orc, st = MimiOracle(sd, dims), StreamState()
        with torch.no_grad():
            want = torch.cat([orc.decode_step(codes[:, :, t:t + 1], st, carry=carry) for t in range(T)], dim=-1)
        m = MimiModel(cfg, num_codebooks=4, max_streams=B, max_frames=8, upsample_carry=carry)
        m.load_state_dict(sd)
        assert m.samples_per_frame == 24
        got = _run(m, codes)
        assert got.shape == (B, 1, T * 24)
        _close(f"small codec, {B} streams, carry {carry}", got, want)


def test_load_mimi_reads_a_kyutai_style_safetensors_file(tmp_path):
    """load_mimi(path) (codec/mimi.py:107-156): a model.safetensors with kyutai/mimi's keys -- here the seeded weights plus keys the
    decode half ignores (encoder, input_proj, `initialized` buffers) -- gives the same PCM as load_state_dict."""
    from safetensors.torch import save_file

    from smoltts_b200.mimi import load_mimi

    sd = dict(_sd())
    sd["encoder.layers.0.conv.weight"] = torch.zeros(64, 1, 7)
    sd["quantizer.semantic_residual_vector_quantizer.input_proj.weight"] = torch.zeros(256, 512, 1)
    sd["quantizer.semantic_residual_vector_quantizer.layers.0.codebook.initialized"] = torch.ones(1)
    path = str(tmp_path / "model.safetensors")
    save_file({k: v.contiguous() for k, v in sd.items()}, path)
    m = load_mimi(path, max_streams=2, max_frames=8)
    codes = torch.randint(0, 2048, (2, 8, 3), generator=torch.Generator().manual_seed(21))
    assert torch.equal(_run(m, codes), _run(_model(max_streams=2, max_frames=8), codes))
    with pytest.raises(ValueError, match="fp32"):
        load_mimi(path, format="bf16")
    bad = {k: v for k, v in sd.items() if k != "upsample.conv.weight"}
    with pytest.raises(KeyError, match="upsample.conv.weight"):
        _model(max_streams=2, max_frames=8).load_state_dict(bad)
