"""Trainer checkpoint (.pt) -> exported checkpoint directory (config.json + model.safetensors).

Replaces train/convert_safetensors.py:6-16 of the reference: same result for the shipped models, without its hard-coded
hidden size (``reshape(768, -1)``), and it also accepts what the trainer can leave behind -- ``_orig_mod.`` prefixes of a
compiled model, legacy split ``wq / wk / wv`` projections, a bare state dict instead of ``{"model_state_dict": ...}`` --
and checks every tensor's name and shape against the config before writing.

usage: python -m smoltts_b200.convert checkpoint.pt --config config.json [-o out_dir] [--dtype bfloat16]
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
from typing import Dict, Optional

import torch

from .config import RQTransformerModelArgs


def convert_checkpoint(checkpoint_path: str, config: RQTransformerModelArgs, out_dir: str, dtype: Optional[torch.dtype] = torch.bfloat16,
                       config_path: Optional[str] = None) -> Dict[str, torch.Tensor]:
    from safetensors.torch import save_file

    from .model import expected_shapes, normalise_state_dict

    data = torch.load(checkpoint_path, map_location="cpu", weights_only=True)
    sd = data["model_state_dict"] if isinstance(data, dict) and "model_state_dict" in data else data
    if not isinstance(sd, dict) or not all(isinstance(v, torch.Tensor) for v in sd.values()):
        raise ValueError("checkpoint holds neither a state dict nor {'model_state_dict': state dict}")
    sd = normalise_state_dict(sd, config)
    want = expected_shapes(config)
    missing = sorted(set(want) - set(sd))
    extra = sorted(set(sd) - set(want))
    if missing or extra:
        raise ValueError(f"checkpoint does not match the config: missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                         f"unexpected {extra[:5]}{'...' if len(extra) > 5 else ''}")
    for k, shape in want.items():
        if tuple(sd[k].shape) != tuple(shape):
            raise ValueError(f"{k}: shape {tuple(sd[k].shape)}, the config implies {tuple(shape)}")
    out = {k: (v.to(dtype) if dtype is not None else v).contiguous() for k, v in sd.items()}
    os.makedirs(out_dir, exist_ok=True)
    save_file(out, os.path.join(out_dir, "model.safetensors"))
    if config_path is not None and os.path.abspath(config_path) != os.path.abspath(os.path.join(out_dir, "config.json")):
        shutil.copyfile(config_path, os.path.join(out_dir, "config.json"))
    elif config_path is None:
        config.save(os.path.join(out_dir, "config.json"))
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("checkpoint")
    ap.add_argument("--config", required=True, help="config.json of the model (RQTransformerModelArgs fields)")
    ap.add_argument("-o", "--out-dir", default=".")
    ap.add_argument("--dtype", default="bfloat16", choices=["bfloat16", "float16", "float32", "keep"])
    a = ap.parse_args(argv)
    with open(a.config) as f:
        cfg = RQTransformerModelArgs.from_dict(json.load(f))
    dtype = None if a.dtype == "keep" else getattr(torch, a.dtype)
    out = convert_checkpoint(a.checkpoint, cfg, a.out_dir, dtype, config_path=a.config)
    n = sum(v.numel() for v in out.values())
    print(f"wrote {os.path.join(a.out_dir, 'model.safetensors')}: {len(out)} tensors, {n / 1e6:.1f} M parameters")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
