"""Shared helpers of the GPU parity tests: seeded weights -> (CUDA model, CPU oracle)."""
from __future__ import annotations

import numpy as np
import torch

from oracle.dualar_oracle import DualAROracle
from smoltts_b200.config import named_config
from smoltts_b200.synth import make_state_dict

_CACHE = {}


def model_and_oracle(size: str, seed: int = 0, max_batch: int = 8, max_seq_len: int = 512, dtype=torch.bfloat16,
                     **cfg_overrides):
    """The CUDA engine and the CPU oracle on bit-identical seeded weights."""
    from smoltts_b200 import RQTransformer

    key = (size, seed, max_batch, max_seq_len, dtype, tuple(sorted(cfg_overrides.items())))
    if key not in _CACHE:
        cfg = named_config(size, **cfg_overrides)
        sd = make_state_dict(cfg, seed=seed, norm_jitter=0.05)
        model = RQTransformer(cfg, max_batch=max_batch, max_seq_len=max_seq_len)
        model.load_state_dict(sd)
        orc = DualAROracle(cfg, sd, dtype=dtype, max_seq_len=max_seq_len)
        _CACHE[key] = (cfg, sd, model, orc)
    return _CACHE[key]


def bf16_ulp_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """|a - b| in units of bf16 ulps of max(|a|, |b|) (both bf16-representable)."""
    a = a.float().cpu()
    b = b.float().cpu()
    mag = torch.maximum(a.abs(), b.abs()).clamp_min(2.0 ** -126)
    ulp = torch.exp2(torch.floor(torch.log2(mag)) - 7)
    return (a - b).abs() / ulp


def close_report(name: str, got: torch.Tensor, want: torch.Tensor, max_ulp: float = 2.0, min_exact: float = 0.98,
                 outlier_frac: float = 0.0):
    """outlier_frac: share of elements that may miss the ulp bar as long as they stay within one bf16 ulp of the tensor's
    LARGEST value -- a one-ulp flip of an intermediate (pre-RoPE value, residual operand) seen after cancellation."""
    got = got.float().cpu()
    want = want.float().cpu()
    assert got.shape == want.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    assert torch.isfinite(got).all(), f"{name}: non-finite values"
    ulps = bf16_ulp_diff(got, want)
    exact = (got == want).float().mean().item()
    # elements near zero can differ by many "ulps of themselves" after cancellation: also allow
    # an absolute slack of one ulp of the tensor's typical magnitude
    scale = want.abs().mean().item() + 1e-12
    bad = (ulps > max_ulp) & ((got - want).abs() > scale * 2.0 ** -7)
    if outlier_frac > 0 and 0 < int(bad.sum()) <= max(1, int(outlier_frac * got.numel())):
        big = want.abs().max().item()
        bad = bad & ((got - want).abs() > 2.0 ** (np.floor(np.log2(max(big, 1e-30))) - 7) * 1.5)
    if bad.any():
        idx = bad.nonzero()[:6].tolist()
        detail = "; ".join(f"{tuple(i)}: got {got[tuple(i)].item():.6g} want {want[tuple(i)].item():.6g}" for i in idx)
        raise AssertionError(f"{name}: {int(bad.sum())} elements differ by more than {max_ulp} bf16 ulps "
                             f"(max {ulps.max().item():.1f} ulps, max abs {float((got - want).abs().max()):.3g}): {detail}")
    assert exact >= min_exact, f"{name}: only {exact:.4f} of elements bit-exact (need {min_exact})"
    return exact, float(ulps.max())


def np_rms(a, b) -> float:
    d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
    return float(np.sqrt((d ** 2).mean()))
