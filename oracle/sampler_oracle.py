"""CPU oracle for the fused sampling kernel.  TEST INFRASTRUCTURE ONLY.

numpy restatement of the sampling specification in DESIGN.md §"Sampling".  The
reference samples with ``argmax`` (temp == 0) or ``mx.random.categorical(logits /
temp)`` (mlx_inference/src/smoltts_mlx/lm/generate.py:88-99,118-132); its "min-p"
branch is a no-op filter (lm/utils/samplers.py:24-28).  MLX's RNG
cannot run here, so the *distribution* is what is shared with the reference; the
bit-exact contract is between this file and the CUDA kernel, which both implement
the same integer-weight algorithm:

  z_i = logit_i * inv_temp                       (fp32 multiply)
  w_i = trunc(exp_det(z_i - max z) * 2^30)       (fp32, IEEE mul/add only; uint32)
  top-k : keep w_i >= (k-th largest w)           (ties kept)
  top-p : keep w_i >= tau, tau = max{t : sum_{w>=t} w >= max(1, (T*P32)>>32)}
  min-p : keep w_i >= trunc(min_p * 2^30)        ("intended" semantics only)
  draw  : r = Philox4x32-10(key=seed, ctr=(step, seq_id, stream, 0)) -> 64 bits;
          target = (r * T') >> 64; pick first i (index order) whose inclusive
          running sum of kept weights exceeds target.

Every step after exp_det is exact integer arithmetic, hence order-independent and
bit-reproducible between CPU and GPU.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = 0xFFFFFFFF

LOG2E = np.float32(1.4426950408889634)
# 2^f on [-0.5, 0.5]: c_k = ln(2)^k / k!, k = 0..7, Horner with separate mul and add
EXP2_COEF = [np.float32(c) for c in (
    1.0, 0.6931471805599453, 0.2402265069591007, 0.05550410866482158,
    0.009618129107628477, 0.0013333558146428443, 0.00015403530393381608,
    1.5252733804059841e-05)]
W_ONE = np.float32(2.0 ** 30)
X_CUTOFF = np.float32(-21.5)  # exp(x) * 2^30 < 1 below this: weight 0


def philox4x32_10(counter, key):
    c = [int(x) & MASK32 for x in counter]
    k = [int(x) & MASK32 for x in key]
    for r in range(10):
        p0 = PHILOX_M0 * c[0]
        p1 = PHILOX_M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK32, p1 & MASK32,
             ((p0 >> 32) ^ c[3] ^ k[1]) & MASK32, p0 & MASK32]
        k = [(k[0] + PHILOX_W0) & MASK32, (k[1] + PHILOX_W1) & MASK32]
    return c


def draw64(seed: int, step: int, seq_id: int, stream: int) -> int:
    r = philox4x32_10((step, seq_id, stream, 0), (seed & MASK32, (seed >> 32) & MASK32))
    return (r[0] << 32) | r[1]


def exp_det_weights(z: np.ndarray) -> np.ndarray:
    """z float32 [V] -> uint32 weights trunc(exp(z - max) * 2^30), IEEE mul/add only."""
    z = z.astype(np.float32)
    x = z - z.max()                                  # <= 0, fp32 subtract
    live = x >= X_CUTOFF
    xs = np.where(live, x, np.float32(0.0)).astype(np.float32)
    t = xs * LOG2E
    n = np.rint(t).astype(np.float32)
    f = (t - n).astype(np.float32)
    p = np.full_like(f, EXP2_COEF[7])
    for c in EXP2_COEF[6::-1]:
        p = (p * f).astype(np.float32)
        p = (p + c).astype(np.float32)
    scale = ((n.astype(np.int32) + 127).astype(np.uint32) << np.uint32(23)).view(np.float32)
    val = (p * scale).astype(np.float32)
    w = (val * W_ONE).astype(np.float32)
    w = np.where(live, w, np.float32(0.0))
    return w.astype(np.uint32)  # truncation toward zero


def filter_weights(w: np.ndarray, top_k: int = 0, top_p: float = 1.0, min_p: float = 0.0) -> np.ndarray:
    """Returns the kept mask (bool [V]) after top-k -> top-p -> min-p."""
    V = w.shape[0]
    keep = np.ones(V, dtype=bool)
    if 0 < top_k < V:
        kth = np.sort(w)[::-1][top_k - 1]
        keep &= w >= kth
    if top_p < 1.0:
        P32 = min(int(np.floor(np.float64(np.float32(top_p)) * 4294967296.0)), MASK32)
        ws = np.sort(w[keep])[::-1].astype(object)
        T = int(sum(int(v) for v in ws))
        need = max(1, (T * P32) >> 32)
        acc = 0
        tau = 0
        for v in ws:
            acc += int(v)
            if acc >= need:
                tau = int(v)
                break
        keep &= w >= tau
    if min_p > 0.0:
        thr = int(np.float32(np.float32(min_p) * W_ONE))
        keep &= w >= thr
    return keep


def sample_row(logits: np.ndarray, temp: float, top_k: int, top_p: float, min_p: float,
               seed: int, step: int, seq_id: int, stream: int) -> int:
    if temp == 0.0:
        return int(np.argmax(logits))  # first maximal index
    inv_temp = np.float32(1.0) / np.float32(temp)
    z = (logits.astype(np.float32) * inv_temp).astype(np.float32)
    w = exp_det_weights(z)
    keep = filter_weights(w, top_k, top_p, min_p)
    wk = np.where(keep, w, 0).astype(np.uint64)
    total = int(wk.sum(dtype=np.uint64))
    target = (draw64(seed, step, seq_id, stream) * total) >> 64
    csum = np.cumsum(wk, dtype=np.uint64)
    return int(np.searchsorted(csum, np.uint64(target), side="right"))


def kept_distribution(logits: np.ndarray, temp: float, top_k: int, top_p: float, min_p: float) -> np.ndarray:
    """Exact sampling distribution of the spec (for chi-square tests)."""
    inv_temp = np.float32(1.0) / np.float32(temp)
    w = exp_det_weights((logits.astype(np.float32) * inv_temp).astype(np.float32))
    wk = np.where(filter_weights(w, top_k, top_p, min_p), w, 0).astype(np.float64)
    return wk / wk.sum()


class OracleSampler:
    """Callable handed to DualAROracle.decode_frame: (logits[B,V] torch fp32, temp,
    settings, frame_index, stream) -> int64 ids [B]."""

    def __init__(self, seq_ids=None, min_p_intended: bool = False):
        self.seq_ids = seq_ids
        self.min_p_intended = min_p_intended

    def __call__(self, logits, temp, settings, frame_index, stream):
        import torch

        arr = logits.detach().cpu().float().numpy()
        B = arr.shape[0]
        ids = self.seq_ids if self.seq_ids is not None else list(range(B))
        top_k, top_p = (settings.top_k, settings.top_p) if stream == 0 else (0, 1.0)
        min_p = float(settings.min_p) if (self.min_p_intended and settings.min_p) else 0.0
        out = [sample_row(arr[b], float(temp), top_k, top_p, min_p, settings.seed, frame_index,
                          ids[b], stream) for b in range(B)]
        return torch.tensor(out, dtype=torch.int64)
