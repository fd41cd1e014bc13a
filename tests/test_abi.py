"""CPU: the C-ABI library builds, loads, and exports every symbol include/t4s.h declares (no compute calls)."""
import ctypes
import os

import pytest


def test_library_builds_and_exports_header_symbols():
    from transformer4sed_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _lib.exported_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/t4s.h but not exported"
    assert lib.t4s_version() >= 100


def test_every_bound_symbol_is_declared_in_header():
    from transformer4sed_b200 import _lib
    lib = _lib.load()
    declared = set(_lib.exported_symbols())
    bound = set(_lib._declare(lib).keys())
    assert bound <= declared, bound - declared
    assert declared <= bound, declared - bound


def test_ops_fail_loudly_without_cuda():
    import torch
    from transformer4sed_b200 import _lib
    from transformer4sed_b200.src_models.passt.passt_feature_extraction import PasstFeatureExtractor
    ext = PasstFeatureExtractor()
    with pytest.raises(_lib.T4sError):
        ext(torch.zeros(1, 32000))
