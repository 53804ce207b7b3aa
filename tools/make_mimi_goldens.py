"""Golden vectors of the Mimi decoder (tests/golden/mimi.npz), produced HERE (build container) from Hugging Face
transformers' MimiModel (transformers 5.5.0) -- the implementation the reference's MLX codec was ported from and whose
state-dict keys it loads (mlx_inference/src/smoltts_mlx/codec/mimi.py:107-156; MLX itself is Apple-only and absent).

Weights: smoltts_b200.synth.make_mimi_state_dict(seed 0) loaded into MimiModel(num_quantizers=8).
  full    = MimiModel.decode(codes)                                   -- the reference's `decode` (whole sequence)
  stream  = quantizer.decode + upsample per frame (each frame ALONE, as the reference's `decode_step` does,
            mimi.py:73-86), then decoder_transformer and the SEANet decoder over the concatenation -- what a run of
            `decode_step` calls computes (the transformer and the SEANet carry their state; the upsampler has none)

    python tools/make_mimi_goldens.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from smoltts_b200.synth import make_mimi_state_dict  # noqa: E402


def main():
    from transformers import MimiConfig, MimiModel

    sd = make_mimi_state_dict(0)
    hf = MimiModel(MimiConfig(num_quantizers=8)).eval()
    res = hf.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    assert not [k for k in res.missing_keys if k.startswith(("decoder", "upsample"))]
    B, T = 2, 4
    g = torch.Generator().manual_seed(7)
    codes = torch.randint(0, 2048, (B, 8, T), generator=g)
    with torch.no_grad():
        full = hf.decode(codes).audio_values
        x = torch.cat([hf.upsample(hf.quantizer.decode(codes[:, :, t:t + 1])) for t in range(T)], dim=-1)
        y = hf.decoder_transformer(x.transpose(1, 2))[0]
        stream = hf.decoder(y.transpose(1, 2))
        emb = hf.quantizer.decode(codes)
    out = os.path.join(ROOT, "tests", "golden", "mimi.npz")
    np.savez_compressed(out, codes=codes.numpy().astype(np.int32), full=full.numpy().astype(np.float32),
                        stream=stream.numpy().astype(np.float32), emb=emb.numpy().astype(np.float32),
                        xf_stream=y.numpy().astype(np.float32))
    print(out, os.path.getsize(out), "bytes; |full| max", float(full.abs().max()), "full vs stream", float((full - stream).abs().max()))


if __name__ == "__main__":
    main()
