"""Python-side launchers for the libt4s kernels (raw device pointers through ctypes; torch is plumbing only).

Everything here runs on the current CUDA stream of the tensor's device and raises `T4sError` on any
failure; there is no CPU implementation.
"""
import ctypes
import struct
import threading

import torch

from . import _lib
from ._lib import Gemm, Matrix, Operand

F32, BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_GELU_GRAD = 0, 1, 2


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise _lib.T4sError(f"unsupported dtype {dt}: libt4s kernels take float32 or bfloat16")


def _esize(t):
    return t.element_size()


class Op:
    """Strided view of a GEMM operand inside tensor `t` (K-major unless mn_major)."""

    def __init__(self, t, rows, ld, offset=0, nb1=1, stride1=0, nb2=1, stride2=0, mn_major=False):
        self.t, self.rows, self.ld, self.offset = t, rows, ld, offset
        self.nb1, self.stride1, self.nb2, self.stride2, self.mn_major = nb1, stride1, nb2, stride2, mn_major

    def c(self):
        return Operand(ctypes.c_void_p(self.t.data_ptr() + self.offset * _esize(self.t)), self.rows, self.ld, self.nb1,
                       self.stride1, self.nb2, self.stride2, int(self.mn_major))


class Out:
    """Strided view of an output / residual / aux matrix."""

    def __init__(self, t, ld, offset=0, stride1=0, stride2=0):
        self.t, self.ld, self.offset, self.stride1, self.stride2 = t, ld, offset, stride1, stride2

    def c(self):
        return Matrix(ctypes.c_void_p(self.t.data_ptr() + self.offset * _esize(self.t)), dtype_code(self.t.dtype), self.ld,
                      self.stride1, self.stride2)


_NULL_MAT = Matrix(ctypes.c_void_p(0), 0, 0, 0, 0)


# The T4sGemm descriptor is filled with ONE struct.pack_into call instead of ~45 ctypes attribute stores (11 us -> 2 us of host time per
# GEMM; the 8-clip DASM step issues ~500 GEMMs and is host-bound).  The format string is derived from the ctypes layout, so the two
# cannot drift apart (tests/test_abi.py compares the bytes).
def _flat_fields(cls, base=0, prefix=""):
    out = []
    for name, typ in cls._fields_:
        off = base + getattr(cls, name).offset
        if isinstance(typ, type) and issubclass(typ, ctypes.Structure):
            out += _flat_fields(typ, off, prefix + name + ".")
        else:
            out.append((prefix + name, off, typ))
    return out


def _packer(cls):
    code = {ctypes.c_int: "i", ctypes.c_int64: "q", ctypes.c_void_p: "Q", ctypes.c_float: "f"}
    fmt, pos = "=", 0
    for _, off, typ in _flat_fields(cls):
        if off > pos:
            fmt += f"{off - pos}x"
            pos = off
        fmt += code[typ]
        pos += ctypes.sizeof(typ)
    if ctypes.sizeof(cls) > pos:
        fmt += f"{ctypes.sizeof(cls) - pos}x"
    return struct.Struct(fmt)


_GEMM_PACK = _packer(Gemm)
_tls = threading.local()


def _gemm_args(A, B, C, M, N, K, nb1, nb2, bias, aux, residual, alpha, act, split_k, c_split_stride, colsum, band):
    """Field values of T4sGemm in declaration order (include/t4s.h)."""
    at, bt, ct = A.t, B.t, C.t
    vals = [M, N, K, dtype_code(at.dtype), nb1, nb2, split_k, c_split_stride,
            at.data_ptr() + A.offset * at.element_size(), A.rows, A.ld, A.nb1, A.stride1, A.nb2, A.stride2, int(A.mn_major),
            bt.data_ptr() + B.offset * bt.element_size(), B.rows, B.ld, B.nb1, B.stride1, B.nb2, B.stride2, int(B.mn_major),
            ct.data_ptr() + C.offset * ct.element_size(), dtype_code(ct.dtype), C.ld, C.stride1, C.stride2]
    for m in (aux, residual):
        if m is None:
            vals += [0, 0, 0, 0, 0]
        else:
            vals += [m.t.data_ptr() + m.offset * m.t.element_size(), dtype_code(m.t.dtype), m.ld, m.stride1, m.stride2]
    vals += [bias.data_ptr() if bias is not None else 0, float(alpha), act, colsum.data_ptr() if colsum is not None else 0,
             int(band[0]) if band is not None else 0, int(band[1]) if band is not None else 0]
    return vals


def gemm(A: Op, B: Op, C: Out, M, N, K, nb1=1, nb2=1, bias=None, aux: Out = None, residual: Out = None, alpha=1.0,
         act=ACT_NONE, split_k=1, c_split_stride=0, colsum=None, band=None):
    """C[z] = act(alpha * A[z] @ B[z]^T + bias) + residual[z] on tensor cores (tcgen05).  `colsum` (fp32 [N], zeroed by the caller)
    receives the column sums of C from the epilogue (bf16 un-split, un-batched outputs only).  `band` = (lo, hi): the caller
    guarantees A[z][m, k] == 0 unless lo <= m + k < hi; k-blocks outside an M tile's band are skipped."""
    _lib.ensure_device(A.t)
    if A.t.dtype != B.t.dtype:
        raise _lib.T4sError("gemm operands must share a dtype")
    if bias is not None and bias.dtype != torch.float32:
        raise _lib.T4sError("gemm bias must be float32")
    buf = getattr(_tls, "gemm_buf", None)
    if buf is None:
        buf = _tls.gemm_buf = (ctypes.c_char * _GEMM_PACK.size)()
    _GEMM_PACK.pack_into(buf, 0, *_gemm_args(A, B, C, M, N, K, nb1, nb2, bias, aux, residual, alpha, act, split_k, c_split_stride, colsum, band))
    with _lib.device_guard(A.t.device):
        if _lib.profiler is not None:
            key = (M, N, K, nb1 * nb2, "T" if A.mn_major else "N", "T" if B.mn_major else "N", split_k)
            rc = _lib.profiler.timed("t4s_gemm", key, lambda: _lib.load().t4s_gemm(buf, _lib.stream_ptr()))
        else:
            rc = _lib.load().t4s_gemm(buf, _lib.stream_ptr())
        _lib.check(rc, "t4s_gemm")


def reduce_splits(ws, splits, n, out, accumulate=False):
    _lib.ensure_device(ws)
    with _lib.device_guard(ws.device):
        _lib.check(_lib.load().t4s_reduce_splits(_lib.ptr(ws), splits, n, _lib.ptr(out), int(accumulate), _lib.stream_ptr()),
                   "t4s_reduce_splits")


def linear_nt(x, w, bias=None, out_dtype=None, act=ACT_NONE, residual=None, aux_dtype=None, alpha=1.0):
    """y[M,N] = act(alpha * x[M,K] @ w[N,K]^T + bias) + residual; returns (y, aux or None)."""
    M, K = x.shape
    N = w.shape[0]
    out_dtype = out_dtype or x.dtype
    y = torch.empty(M, N, dtype=out_dtype, device=x.device)
    aux = torch.empty(M, N, dtype=aux_dtype, device=x.device) if aux_dtype is not None else None
    gemm(Op(x, M, x.stride(0)), Op(w, N, w.stride(0)), Out(y, N), M, N, K, bias=bias,
         aux=Out(aux, N) if aux is not None else None,
         residual=Out(residual, residual.stride(0)) if residual is not None else None, alpha=alpha, act=act)
    return y, aux
