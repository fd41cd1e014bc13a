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


def test_packed_gemm_descriptor_matches_ctypes_layout():
    """ops.gemm fills T4sGemm with one struct.pack_into call: the bytes must equal the ctypes structure built field by field."""
    import torch
    from transformer4sed_b200 import ops
    from transformer4sed_b200._lib import Gemm, Matrix, Operand
    a, b = torch.zeros(40, 24, dtype=torch.bfloat16), torch.zeros(56, 24, dtype=torch.bfloat16)
    c, x, r = torch.zeros(40, 56), torch.zeros(40, 56, dtype=torch.bfloat16), torch.zeros(40, 56, dtype=torch.bfloat16)
    bias, cs = torch.zeros(56), torch.zeros(56)
    A, B = ops.Op(a, 40, 24, 3, nb1=2, stride1=5, nb2=3, stride2=7), ops.Op(b, 56, 24, 1, mn_major=True)
    C, X, R = ops.Out(c, 56, 2, 11, 13), ops.Out(x, 56), ops.Out(r, 56, 4)
    vals = ops._gemm_args(A, B, C, 40, 56, 24, 2, 3, bias, X, R, 0.5, ops.ACT_GELU, 4, 99, cs, (7, 30))
    buf = (ctypes.c_char * ops._GEMM_PACK.size)()
    ops._GEMM_PACK.pack_into(buf, 0, *vals)
    g = Gemm()
    g.M, g.N, g.K, g.in_dtype, g.nb1, g.nb2, g.split_k, g.c_split_stride = 40, 56, 24, ops.BF16, 2, 3, 4, 99
    g.A, g.B, g.C, g.aux, g.residual = A.c(), B.c(), C.c(), X.c(), R.c()
    g.bias, g.alpha, g.act, g.colsum = ctypes.c_void_p(bias.data_ptr()), 0.5, ops.ACT_GELU, ctypes.c_void_p(cs.data_ptr())
    g.band_lo, g.band_hi = 7, 30
    assert ops._GEMM_PACK.size == ctypes.sizeof(Gemm)
    assert bytes(buf) == bytes(g)
    # absent optional parts are all-zero, as the zero-initialised ctypes structure has them
    vals = ops._gemm_args(A, B, C, 40, 56, 24, 1, 1, None, None, None, 1.0, ops.ACT_NONE, 1, 0, None, None)
    ops._GEMM_PACK.pack_into(buf, 0, *vals)
    g2 = Gemm()
    g2.M, g2.N, g2.K, g2.in_dtype, g2.nb1, g2.nb2, g2.split_k = 40, 56, 24, ops.BF16, 1, 1, 1
    g2.A, g2.B, g2.C, g2.alpha = A.c(), B.c(), C.c(), 1.0
    assert bytes(buf) == bytes(g2)
