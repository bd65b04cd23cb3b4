"""CPU: the C-ABI library loads and exports every symbol include/fervit_b200.h declares; host-only queries work."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fervit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fervit_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ("fervit_plan_create", "fervit_plan_forward", "fervit_plan_backward", "fervit_cross_entropy",
              "fervit_linear_forward", "fervit_attention_forward", "fervit_layernorm_forward",
              "fervit_premodules_forward"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from fer_vit_b200 import _lib
    handle = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, missing
    # and the ctypes signature table covers the header exactly
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_host_only_queries():
    from fer_vit_b200 import _lib as L
    lib = L.lib()
    assert lib.fervit_abi_version() == 1
    cfg = L.Config()
    cfg.mode, cfg.input_kind, cfg.L, cfg.Din, cfg.E, cfg.depth, cfg.H, cfg.F, cfg.C = L.BF16, 0, 18, 512, 768, 12, 12, 3072, 7
    cfg.norm_first, cfg.act, cfg.eps_block, cfg.eps_head, cfg.adapter_dim = 1, L.ACT_GELU, 1e-6, 1e-5, 64
    cfg.head_dropout = 0.1
    h = ctypes.c_void_p()
    L.check(lib.fervit_plan_create(ctypes.byref(cfg), ctypes.byref(h)))
    assert lib.fervit_plan_num_slots(h) == 16 + 12 * 17
    assert lib.fervit_plan_num_stages(h) == 14
    assert lib.fervit_plan_slot_numel(h, L.G_IN_W) == 768 * 512
    assert lib.fervit_plan_slot_numel(h, L.bslot(3, L.B_QKV_W)) == 3 * 768 * 768
    # bf16 W and W^T of every GEMM weight: 2 * 2 bytes * (85.0M backbone GEMM weights + adapters + input proj)
    n_w = 12 * (3 * 768 * 768 + 768 * 768 + 2 * 768 * 3072 + 2 * 64 * 768) + 768 * 512
    assert lib.fervit_plan_wcache_bytes(h) >= 4 * n_w
    ws_train = lib.fervit_plan_workspace_bytes(h, 256, 1)
    ws_infer = lib.fervit_plan_workspace_bytes(h, 256, 0)
    assert 0 < ws_infer < ws_train < 8 * 2 ** 30
    lib.fervit_plan_destroy(h)


def test_bad_config_is_reported_not_thrown():
    from fer_vit_b200 import _lib as L
    lib = L.lib()
    cfg = L.Config()
    cfg.mode, cfg.L, cfg.Din, cfg.E, cfg.depth, cfg.H, cfg.F, cfg.C, cfg.act = L.F32, 18, 512, 100, 2, 3, 256, 7, 1
    h = ctypes.c_void_p()
    assert lib.fervit_plan_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    with pytest.raises(RuntimeError, match="head"):
        L.check(1)
