"""The C-ABI boundary without a GPU: libvitae_b200.so builds for sm_100a, loads, and exports every entry point that
include/vitae_b200.h declares; the ctypes table binds exactly that set; calls that need a device fail loudly."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vitae_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vitae_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from vit_ae_plus_plus_b200 import build, _lib
    build.build(verbose=False)
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from vit_ae_plus_plus_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/vitae_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)
    assert lib.vitae_abi_version() == 1


def test_epilogue_struct_layout_matches_header():
    from vit_ae_plus_plus_b200._lib import GemmEpilogue
    text = open(HEADER).read()
    body = re.search(r"typedef struct vitae_gemm_epilogue \{(.*?)\} vitae_gemm_epilogue;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [re.findall(r"(\w+)\s*;", line)[0] for line in body.splitlines() if ";" in line]
    assert fields == [f[0] for f in GemmEpilogue._fields_]


def test_host_side_queries_work_without_a_device(lib):
    assert lib.vitae_gemm_workspace_bytes(516, 768, 4) == 4 * 516 * 768 * 4
    assert lib.vitae_gemm_workspace_bytes(516, 768, 1) == 0
    assert lib.vitae_layernorm_bwd_blocks(516) == 17 and lib.vitae_layernorm_bwd_blocks(100000) == 64
    assert lib.vitae_colsum_workspace_bytes(2052, 16384) >= 16384 * 4
    assert lib.vitae_optim_workspace_bytes() > 0
    assert lib.vitae_launch_count() >= 0


def test_no_device_is_an_error_not_a_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vit_ae_plus_plus_b200 import _lib
    assert lib.vitae_check_device() != 0
    assert lib.vitae_last_error()
    # argument validation happens before any launch
    rc = lib.vitae_layernorm_fwd(None, None, None, None, None, None, None, 4, 128, 1e-6, None)
    assert rc != 0 and b"null" in lib.vitae_last_error()
    with pytest.raises(_lib.VitaeError):
        _lib.check(rc, "vitae_layernorm_fwd")


def test_model_refuses_cpu():
    import argparse
    import torch
    from vit_ae_plus_plus_b200._lib import VitaeError
    from vit_ae_plus_plus_b200.model.vit_autoenc import MaskedAutoencoderViT
    m = MaskedAutoencoderViT(volume_size=32, patch_size=8, in_chans=1, embed_dim=128, depth=1, num_heads=4,
                             decoder_embed_dim=64, decoder_depth=1, decoder_num_heads=4,
                             args=argparse.Namespace(perceptual_weight=0))
    with pytest.raises(VitaeError):
        m(torch.zeros(1, 1, 32, 32, 32))


def test_ring_depths_defaults_and_overrides(monkeypatch):
    """Gradient rings of the backward lanes: one slot per use in the whole backward by default, env overrides with floors."""
    from vit_ae_plus_plus_b200 import engine
    monkeypatch.delenv("VITAE_RING", raising=False)
    monkeypatch.delenv("VITAE_RING_BLOCK", raising=False)
    assert engine.ring_depths(12, 8) == (42, 12)
    assert engine.ring_depths(24, 8) == (66, 24)
    monkeypatch.setenv("VITAE_RING", "2")
    monkeypatch.setenv("VITAE_RING_BLOCK", "1")
    assert engine.ring_depths(12, 8) == (4, 2)


def test_block_colreduce_workspace_covers_every_subset_of_the_jobs():
    """The plan sizes ONE workspace for the six column-reduction jobs of a block; launches with fewer jobs (first / last
    block, encoder-only backward) use more row slices and must still fit.  Host-side check only: without a GPU the call
    gets past the workspace check and fails at the launch (regression: batch 16 failed with 'workspace too small')."""
    import ctypes
    import itertools
    from vit_ae_plus_plus_b200 import _lib, ops
    lib = _lib.load()
    for D, hid in ((768, 3072), (512, 2048), (1024, 4096), (128, 512)):
        cols = [hid, 3 * D, D, D, D, D]
        for rows in (34, 516, 2052, 2064, 8208, 16416, 65664):
            ws_bytes = ops.block_colreduce_workspace_bytes(rows, cols)
            for n in range(1, 7):
                for sub in itertools.combinations(cols, n):
                    jobs = (_lib.ColJob * n)()
                    for j, c in zip(jobs, sub):
                        j.a, j.out0, j.cols, j.ld, j.a_is_bf16 = 0x1000, 0x2000, c, c, 1
                    rc = lib.vitae_block_colreduce(jobs, n, rows, 0, ctypes.c_void_p(0x3000), ws_bytes, None)
                    msg = lib.vitae_last_error().decode() if rc else ""
                    assert "workspace too small" not in msg, (D, rows, sub, msg)
