"""CPU: the C-ABI library loads, exports every symbol include/dgrm.h declares, and fails loudly (no CPU
fallback) when there is no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from dissect_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "dgrm.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dgrm_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_config_struct_layout_matches_the_binding():
    """dgrm_config in the header and the ctypes mirror: same fields, same order, all int32 (64 bytes)."""
    txt = open(os.path.join(ROOT, "include", "dgrm.h")).read()
    body = re.search(r"typedef struct dgrm_config \{(.*?)\} dgrm_config;", txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"int32_t\s+([a-z_]+)(\[\d+\])?\s*;", body)
    want = [(n, int(a[1:-1]) if a else 1) for n, a in fields]
    got = [(n, C.sizeof(t) // 4) for n, t in _lib.Config._fields_]
    assert got == want
    assert C.sizeof(_lib.Config) == 4 * sum(k for _, k in want) == 64


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.dgrm_abi_version() == 1


def test_no_cpu_fallback():
    """Without a usable sm_100 device dgrm_create must fail with DGRM_E_CUDA and say why."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    cfg = _lib.Config()
    assert lib.dgrm_default_config(C.byref(cfg)) == 0
    assert (cfg.chunk_snps, cfg.cta_group, cfg.world) == (16384, 2, 1)
    h = C.c_void_p()
    rc = lib.dgrm_create(C.byref(cfg), C.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.dgrm_last_error(None) or b"CUDA" in lib.dgrm_last_error(None)


def test_bad_config_rejected_before_touching_the_gpu():
    lib = _lib.load()
    cfg = _lib.Config()
    lib.dgrm_default_config(C.byref(cfg))
    h = C.c_void_p()
    cfg.chunk_snps = 128          # must be a multiple of 256
    assert lib.dgrm_create(C.byref(cfg), C.byref(h)) == -1
    cfg.chunk_snps = 16384
    cfg.rank, cfg.world = 3, 2
    assert lib.dgrm_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.dgrm_set_individuals(None, 10, None) == -1


def test_engine_import_is_the_cuda_library():
    from dissect_b200 import lib_path
    assert lib_path.endswith("dissect_b200/libdgrm.so") and os.path.exists(lib_path)
    # the product package never imports the oracle
    pkg = os.path.join(ROOT, "dissect_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f
