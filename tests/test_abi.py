"""The C-ABI library loads and exports every symbol include/heaac_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

import ffmpeg_heaac_b200 as hb
from common import ROOT, abi


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "heaac_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(heaac_(?:b200|fe|dec|adts|esds)_\w+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ("heaac_b200_create", "heaac_b200_destroy", "heaac_b200_synth", "heaac_b200_mdct_init",
                 "heaac_b200_imdct_half", "heaac_b200_mdct_end", "heaac_b200_sbr_tables"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(hb.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared in include/heaac_b200.h but not exported: %s" % missing


def test_library_exports_the_experimental_header_too():
    src = open(os.path.join(ROOT, "include", "heaac_b200_experimental.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    lib = ctypes.CDLL(hb.LIB_PATH)
    syms = sorted(set(re.findall(r"\b(heaac_b200_exp_\w+)\s*\(", src)))
    assert syms and all(hasattr(lib, s) for s in syms)


def test_python_mirror_lists_every_declared_symbol():
    assert sorted(hb.EXPORTS) == declared_symbols()


def test_struct_sizes_match_the_library():
    L = hb.lib()   # hb.lib() itself asserts abi.SIZES against heaac_b200_sizeof
    assert L.heaac_b200_sizeof(b"HeaacConfig") == ctypes.sizeof(abi.Config)
    assert L.heaac_b200_sizeof(b"HeaacBatch") == ctypes.sizeof(abi.Batch)
    assert L.heaac_b200_sizeof(b"nonsense") == -1


def test_no_cpu_path(monkeypatch):
    """Without a CUDA device the product must fail loudly (HEAAC_E_NODEV), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = abi.make_config(1, [abi.TYPE_CPE], 4)
    with pytest.raises(hb.HeaacError) as e:
        hb.Context(cfg)
    assert e.value.code == -19


def test_create_rejects_bad_arguments():
    L = hb.lib()
    h = ctypes.c_void_p()
    cfg = abi.make_config(1, [abi.TYPE_CPE], 4)
    cfg.abi_version = 99
    assert L.heaac_b200_create(ctypes.byref(cfg), ctypes.byref(h)) == -22
    cfg = abi.make_config(0, [abi.TYPE_CPE], 4)
    assert L.heaac_b200_create(ctypes.byref(cfg), ctypes.byref(h)) == -22
    cfg = abi.make_config(1, [abi.TYPE_CPE], 13)
    assert L.heaac_b200_create(ctypes.byref(cfg), ctypes.byref(h)) == -22
    cfg = abi.make_config(1, [abi.TYPE_CPE], 6, sbr=True, ps=True)      # PS needs a mono SCE
    assert L.heaac_b200_create(ctypes.byref(cfg), ctypes.byref(h)) == -22
    assert L.heaac_b200_strerror(-19).decode().startswith("no usable")
