"""The Mimi oracle (oracle/mimi_oracle.py) against the goldens made from transformers' MimiModel
(tools/make_mimi_goldens.py): whole-sequence decode, and the streaming form under both upsampling rules."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle.mimi_oracle import MimiOracle, StreamState
from smoltts_b200.synth import make_mimi_state_dict


@pytest.fixture(scope="module")
def setup():
    g = np.load(f"{GOLDEN_DIR}/mimi.npz")
    return MimiOracle(make_mimi_state_dict(0)), g


def _close(name, got, want, rel=2e-5):
    got, want = torch.as_tensor(got), torch.as_tensor(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    err = (got - want).abs().max().item()
    assert err <= rel * want.abs().max().item(), f"{name}: max abs error {err:.3g} vs scale {want.abs().max().item():.3g}"


def test_embed_and_full_decode_match_transformers(setup):
    orc, g = setup
    codes = torch.from_numpy(g["codes"]).long()
    with torch.no_grad():
        _close("emb", orc.embed(codes), g["emb"])
        _close("full", orc.decode(codes), g["full"])


def test_streaming_steps_match_both_rules(setup):
    orc, g = setup
    codes = torch.from_numpy(g["codes"]).long()
    T = codes.shape[-1]
    with torch.no_grad():
        for carry, key in ((False, "stream"), (True, "full")):
            st = StreamState()
            pcm = torch.cat([orc.decode_step(codes[:, :, t:t + 1], st, carry=carry) for t in range(T)], dim=-1)
            assert pcm.shape[-1] == T * orc.d.samples_per_frame
            _close(f"steps carry={carry}", pcm, g[key])
    # the two rules really differ (the reference's streaming path drops the upsampler's trailing taps)
    assert np.abs(g["full"] - g["stream"]).max() > 1e-2 * np.abs(g["full"]).max()


def test_streams_are_independent_and_window_is_a_no_op_below_its_length(setup):
    orc, g = setup
    codes = torch.from_numpy(g["codes"]).long()
    with torch.no_grad():
        st = StreamState()
        solo = torch.cat([orc.decode_step(codes[1:2, :, t:t + 1], st) for t in range(codes.shape[-1])], dim=-1)
        _close("row 1 alone", solo, g["stream"][1:2])
        w = MimiOracle(orc.sd, window=250)
        _close("window 250", w.decode(codes), g["full"])
        w3 = MimiOracle(orc.sd, window=3)
        assert (w3.decode(codes) - torch.from_numpy(g["full"])).abs().max() > 1e-3
