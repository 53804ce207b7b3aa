"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol that
include/smoltts_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = "".join(open(os.path.join(ROOT, "include", h)).read() for h in sorted(os.listdir(os.path.join(ROOT, "include"))) if h.endswith(".h"))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(smol_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from smoltts_b200 import _capi, build

    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported by {path}"
    assert set(names) == set(_capi.PROTOTYPES), (set(names) ^ set(_capi.PROTOTYPES))
    assert _capi.load().smol_abi_version() == _capi.SMOL_ABI_VERSION


def test_create_validates_shapes_without_a_gpu():
    """smol_create is pure host code: shape checks and error strings work on the CPU box."""
    from smoltts_b200 import _capi
    from smoltts_b200.config import named_config

    lib = _capi.load()
    cfg = named_config("smoltts_byte_150m")

    def make(**over):
        kw = dict(dim=cfg.dim, n_layer=cfg.n_layer, n_head=cfg.n_head, n_local_heads=cfg.n_local_heads, head_dim=64,
                  intermediate_size=cfg.intermediate_size, vocab_size=cfg.vocab_size, fast_dim=cfg.fast_dim,
                  n_fast_layer=cfg.n_fast_layer, fast_n_head=cfg.fast_n_head, fast_n_local_heads=cfg.fast_n_local_heads,
                  fast_head_dim=64, fast_intermediate_size=cfg.fast_intermediate_size, codebook_size=2048, num_codebooks=8,
                  duplicate_code_0=1, depthwise_wte=1, depthwise_output=1, tie_word_embeddings=1, max_seq_len=2048,
                  max_batch=4, page_size=32, semantic_start_id=320, semantic_end_id=2367, im_end_id=270, mlx_embed_mask=0,
                  norm_eps=1e-5)
        kw.update(over)
        return _capi.SmolConfig(**kw)

    h = ctypes.c_void_p()
    assert lib.smol_create(ctypes.byref(make()), ctypes.byref(h)) == 0
    assert lib.smol_phase_count(h) == 5 * 10 + 2 + 8 * (4 * 4 + 2) == 196
    assert lib.smol_kv_page_bytes(h) == 10 * 2 * 4 * 32 * 64 * 2
    assert lib.smol_workspace_bytes(h) > 0
    # compute calls before binding fail loudly instead of touching a device
    b = _capi.SmolBatch()
    assert lib.smol_slow_step(h, ctypes.byref(b), 1, 1, None) == _capi.SMOL_ERR_UNBOUND
    assert b"bound" in lib.smol_last_error()
    lib.smol_destroy(h)
    h2 = ctypes.c_void_p()
    assert lib.smol_create(ctypes.byref(make(fast_dim=512, fast_n_head=8)), ctypes.byref(h2)) == _capi.SMOL_ERR_UNSUPPORTED
    assert b"fast_project_in" in lib.smol_last_error()
    assert lib.smol_create(ctypes.byref(make(head_dim=128)), ctypes.byref(h2)) == _capi.SMOL_ERR_UNSUPPORTED


def test_mimi_create_validates_shapes_without_a_gpu():
    """smol_mimi_create is pure host code too."""
    from smoltts_b200 import _capi

    lib = _capi.load()

    def make(**over):
        c = _capi.SmolMimiConfig(n_q=8, codebook_size=2048, codebook_dim=256, dim=512, n_layers=8, n_heads=8, head_dim=64, ffn=2048,
                                 n_filters=64, n_ratios=4, kernel=7, res_kernel=3, last_kernel=3, max_streams=2, max_positions=512,
                                 window=0, upsample_carry=0, use_graph=1, norm_eps=1e-5, codebook_eps=1e-5)
        for i, r in enumerate((8, 6, 5, 4)):
            c.ratios[i] = r
        for k, v in over.items():
            setattr(c, k, v)
        return c

    h = ctypes.c_void_p()
    assert lib.smol_mimi_create(ctypes.byref(make()), ctypes.byref(h)) == 0
    assert lib.smol_mimi_samples_per_frame(h) == 1920
    ws = lib.smol_mimi_workspace_bytes(h)
    # packed weights of the decode half (44.4 M parameters) + two streams' KV caches and arenas
    assert 44.3e6 * 4 < ws < 44.3e6 * 4 + 2 * (8 * 2 * 512 * 512 * 4 + 3.0e6) + 4e6
    assert lib.smol_mimi_decode_step(h, ctypes.c_void_p(8), None, 1, ctypes.c_void_p(8), None) == _capi.SMOL_ERR_UNBOUND
    lib.smol_mimi_destroy(h)
    h2 = ctypes.c_void_p()
    assert lib.smol_mimi_create(ctypes.byref(make(head_dim=48)), ctypes.byref(h2)) == _capi.SMOL_ERR_INVALID
    assert lib.smol_mimi_create(ctypes.byref(make(max_positions=7)), ctypes.byref(h2)) == _capi.SMOL_ERR_INVALID
    assert lib.smol_mimi_create(ctypes.byref(make(max_positions=4000000)), ctypes.byref(h2)) == _capi.SMOL_ERR_CAPACITY
