"""CPU-side checks: the C-ABI library loads, exports every symbol include/mtl_b200.h declares, and the
host-only entry points (layout, workspace planning) agree with the oracle's parameter inventory."""
import ctypes as C
import os
import re

import pytest

import mtl_b200
from mtl_b200 import lib as L
from oracle import ref_asr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.library_path()):
        import __graft_entry__ as g
        g.build()
    return L.get_lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "mtl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mtl_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    assert lib.mtl_abi_version() == 2


@pytest.mark.parametrize("cfg", [ref_asr.SMALL, ref_asr.CFG2])
def test_arena_layout_matches_reference_parameter_order(lib, cfg):
    c = L.ModelCfg(cfg.n_enc, cfg.n_dec, cfg.d_model, cfg.n_heads, cfg.d_k, cfg.d_v, cfg.d_inner, cfg.rank,
                   cfg.vocab, cfg.n_freq)
    h = C.c_void_p()
    L.check(lib.mtl_session_create(C.byref(c), C.byref(h)))
    specs = ref_asr.param_specs(cfg)
    assert lib.mtl_param_count(h) == len(specs)
    from gpu_util import spec_of
    assert mtl_b200.param_specs(spec_of(cfg)) == [(n, tuple(s)) for n, s in specs]
    off, num = C.c_longlong(), C.c_longlong()
    prev_end = 0
    for i, (name, shape) in enumerate(specs):
        L.check(lib.mtl_param_info(h, i, C.byref(off), C.byref(num)))
        n = 1
        for s in shape:
            n *= s
        assert num.value == n, name
        assert off.value % 64 == 0 and off.value >= prev_end, name
        prev_end = off.value + n
    total = lib.mtl_param_arena_floats(h)
    assert total >= prev_end and total % 64 == 0
    if cfg is ref_asr.CFG2:
        assert total == 14022080            # no padding needed at cfg 2
    ws = lib.mtl_workspace_bytes(h, 8, 101, 33)
    assert ws > 0
    assert lib.mtl_workspace_bytes(h, 8, 2, 33) < 0      # shorter than 4 frames -> error, not a crash
    assert b"4 frames" in lib.mtl_last_error()
    lib.mtl_session_destroy(h)


def test_bad_config_is_rejected(lib):
    c = L.ModelCfg(2, 4, 510, 8, 64, 64, 512, 100, 3765, 161)     # d_model % 4 != 0
    h = C.c_void_p()
    assert lib.mtl_session_create(C.byref(c), C.byref(h)) < 0
    assert b"d_model" in lib.mtl_last_error()


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gpu_util import spec_of
    with pytest.raises(L.MtlError):
        mtl_b200.Session(spec_of(ref_asr.SMALL))
