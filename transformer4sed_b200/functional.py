"""Autograd-visible ops of the SED hot path.  Forward AND backward of every op run in libt4s kernels
(tcgen05 GEMMs + fused row kernels); torch supplies only tensor allocation, views and the autograd tape.

Precision modes (`set_precision`):
  "bf16"   activations and GEMM operands bf16, fp32 accumulate / statistics / master weights  (performance mode)
  "tf32"   activations fp32, GEMMs on tcgen05 kind::tf32
  "tf32x3" as tf32 but every GEMM is error-compensated (hi/lo split, 3 tensor-core products): fp32-class accuracy;
           this is the strict-parity mode the <=1e-3 / argmax-exact contract is checked in.
"""
import ctypes
import math
import os

import torch

from . import _lib, ops
from .ops import Op, Out

_MODE = "bf16"
_FUSED_ATTN = True
_FUSED_ATTN_BWD = True   # one-kernel attention backward (dQ via TMA reduce-add); False: deterministic two-kernel backward


def set_precision(mode: str):
    global _MODE
    if mode not in ("bf16", "tf32", "tf32x3"):
        raise ValueError(f"unknown precision mode {mode!r}")
    _MODE = mode


def get_precision():
    return _MODE


def act_dtype():
    return torch.bfloat16 if _MODE == "bf16" else torch.float32


def _lib_call(name, *args, _key=None):
    fn = getattr(_lib.load(), name)
    if _lib.profiler is not None:
        _lib.check(_lib.profiler.timed(name, _key, lambda: fn(*args)), name)
    else:
        _lib.check(fn(*args), name)


def _st():
    return _lib.stream_ptr()


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _pad8(n):
    return (n + 7) // 8 * 8


# ------------------------------------------------------------------------------------------------------------------
# operand preparation
# ------------------------------------------------------------------------------------------------------------------
_wcache = {}


def _arena_shadow(w):
    """bf16 shadow of a parameter that lives in a `training.ParamArena` (None otherwise).  The fused AdamW / EMA kernels write the
    fp32 master and the shadow together; anything that changes the parameter through torch instead (`load_state_dict`, `copy_`:
    resuming from a checkpoint after the arena was built, reference recipes/desed/finetune/passt/main.py:63-69) bumps its version
    counter, and the shadow is re-converted here before it is used."""
    shadow = getattr(w, "_t4s_shadow", None)
    if shadow is not None and getattr(w, "_t4s_shadow_version", w._version) != w._version:
        convert(w.detach(), shadow)
        w._t4s_shadow_version = w._version
    return shadow


def cast_weight(w: torch.Tensor) -> torch.Tensor:
    """fp32 master weight -> GEMM operand dtype (bf16 copy cached until the parameter is modified)."""
    if _MODE != "bf16":
        return w.detach()
    shadow = _arena_shadow(w)
    if shadow is not None:
        return shadow
    if w._base is not None and w._base.dtype == torch.float32:  # a slice of a parameter: cast the parameter once, re-slice the copy
        c = cast_weight(w._base)    # (the copy may start elsewhere in its own storage than the base does: offsets are relative)
        return c.as_strided(w.shape, w.stride(), c.storage_offset() + w.storage_offset() - w._base.storage_offset())
    key = id(w)
    ent = _wcache.get(key)
    ver = (w._version, w.data_ptr(), w.device)
    if ent is not None and ent[0] == ver and ent[2]() is w:
        return ent[1]
    import weakref
    out = torch.empty(w.shape, dtype=torch.bfloat16, device=w.device)
    convert(w.detach(), out)
    _wcache[key] = (ver, out, weakref.ref(w, lambda _r, k=key: _wcache.pop(k, None)))
    return out


def convert(src, dst):
    _lib.ensure_device(src)
    with _lib.device_guard(src.device):
        _lib_call("t4s_convert", _p(src), ops.dtype_code(src.dtype), _p(dst), ops.dtype_code(dst.dtype), src.numel(), _st())
    return dst


class _Cast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src_dtype = x.dtype
        x = x.contiguous()
        return convert(x, torch.empty(x.shape, dtype=dtype, device=x.device))

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        return convert(dy, torch.empty(dy.shape, dtype=ctx.src_dtype, device=dy.device)), None


def cast(x, dtype):
    return x if x.dtype == dtype else _Cast.apply(x, dtype)


def to_act(x):
    """Bring a tensor to the activation dtype (autograd-visible; no-op when it already is)."""
    return cast(x, act_dtype())


def _split3(op: Op, K, pattern, nb1, nb2):
    """tf32x3: gather a strided operand into a contiguous K-major [nb2, nb1, rows, 3K] hi/lo buffer."""
    b1 = nb1 if op.nb1 > 1 else 1
    b2 = nb2 if op.nb2 > 1 else 1
    ld = (3 * K + 3) // 4 * 4  # 16-byte row pitch for TMA
    dst = torch.empty(b2, b1, op.rows, ld, dtype=torch.float32, device=op.t.device)
    c = op.c()
    _lib_call("t4s_split_tf32", ctypes.byref(c), K, _p(dst), ld, pattern, _st())
    return Op(dst, op.rows, ld, 0, nb1=b1, stride1=op.rows * ld, nb2=b2, stride2=b1 * op.rows * ld)


def mm(A: Op, B: Op, C: Out, M, N, K, nb1=1, nb2=1, **kw):
    """Tensor-core GEMM honouring the precision mode."""
    if _MODE == "tf32x3" and A.t.dtype == torch.float32:
        with _lib.device_guard(A.t.device):
            A3, B3 = _split3(A, K, 0, nb1, nb2), _split3(B, K, 1, nb1, nb2)
        kw.pop("band", None)   # the [hi | lo | hi] operand has three copies of every k: the band hint does not carry over
        ops.gemm(A3, B3, C, M, N, 3 * K, nb1=nb1, nb2=nb2, **kw)
    else:
        ops.gemm(A, B, C, M, N, K, nb1=nb1, nb2=nb2, **kw)


def _split_k_for(M, N, K, batches=1):
    tiles = ((M + 127) // 128) * ((N + 255) // 256 if N > 128 else 1) * batches
    bk = 64 if _MODE == "bf16" else 32
    kblocks = max(1, (K + bk - 1) // bk)
    want = max(1, (2 * 148 + tiles - 1) // tiles)
    s = max(1, min(want, kblocks // 4 if kblocks >= 8 else 1, 64))
    per = (kblocks + s - 1) // s
    return (kblocks + per - 1) // per  # no empty trailing split


def weight_grad(dy2d, x2d, n_out, k_in):
    """dW[n_out, k_in] = dy^T x over the token dimension: both operands consumed MN-major (no transposes), split-K."""
    Mtok = dy2d.shape[0]
    S = _split_k_for(n_out, k_in, Mtok)
    dev = dy2d.device
    if S == 1:
        dw = torch.empty(n_out, k_in, dtype=torch.float32, device=dev)
        mm(Op(dy2d, n_out, dy2d.stride(0), mn_major=True), Op(x2d, k_in, x2d.stride(0), mn_major=True), Out(dw, k_in), n_out, k_in, Mtok)
        return dw
    ws = torch.empty(S, n_out, k_in, dtype=torch.float32, device=dev)
    if _MODE == "tf32x3":
        # split along the contraction by hand (the 3K interleave does not commute with the kernel's K ranges)
        chunk = (Mtok + S - 1) // S
        chunk = (chunk + 3) // 4 * 4
        S = (Mtok + chunk - 1) // chunk
        ws = ws[:S]
        for s in range(S):
            r0, r1 = s * chunk, min(Mtok, (s + 1) * chunk)
            mm(Op(dy2d, n_out, dy2d.stride(0), offset=r0 * dy2d.stride(0), mn_major=True),
               Op(x2d, k_in, x2d.stride(0), offset=r0 * x2d.stride(0), mn_major=True), Out(ws, k_in, offset=s * n_out * k_in),
               n_out, k_in, r1 - r0)
    else:
        mm(Op(dy2d, n_out, dy2d.stride(0), mn_major=True), Op(x2d, k_in, x2d.stride(0), mn_major=True), Out(ws, k_in), n_out, k_in, Mtok,
           split_k=S, c_split_stride=n_out * k_in)
    dw = torch.empty(n_out, k_in, dtype=torch.float32, device=dev)
    ops.reduce_splits(ws, S, n_out * k_in, dw)
    return dw


_FUSE_BIAS_GRAD = True


def set_fused_bias_grad(flag: bool):
    """A/B switch: False recomputes every bias gradient with a separate column-sum pass over the output gradient."""
    global _FUSE_BIAS_GRAD
    _FUSE_BIAS_GRAD = bool(flag)


def _attached_colsum(dy, n):
    """Column sums that the kernel which produced `dy` accumulated on the way out (LayerNorm backward: residual-stream gradients,
    fused attention backward: dqkv).  Found on the tensor object autograd hands over; anything else (a sum of several gradients,
    a dtype conversion, a different producer) has no attribute and the caller falls back to a column-sum pass."""
    cs = getattr(dy, "_t4s_colsum", None)
    if cs is not None and _FUSE_BIAS_GRAD and cs.numel() == n and cs.device == dy.device:
        return cs
    return None


def colsum(x2d, cols=None, ld=None, rows=None):
    rows = x2d.shape[0] if rows is None else rows
    cols = x2d.shape[1] if cols is None else cols
    ld = x2d.stride(0) if ld is None else ld
    lib = _lib.load()
    with _lib.device_guard(x2d.device):
        nbytes = lib.t4s_colsum_workspace(rows, cols)
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=x2d.device)
        out = torch.empty(cols, dtype=torch.float32, device=x2d.device)
        _lib_call("t4s_colsum", _p(x2d), ops.dtype_code(x2d.dtype), rows, cols, ld, _p(ws), nbytes, _p(out), 0, _st())
    return out


# ------------------------------------------------------------------------------------------------------------------
# Linear (+bias, +GELU, +residual)                  reference: nn.Linear / timm Mlp / attention projections
# ------------------------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, residual, act, out_dtype):
        _lib.ensure_device(x)
        shp = x.shape
        K = shp[-1]
        x2 = x.reshape(-1, K)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        M, N = x2.shape[0], w.shape[0]
        wq = cast_weight(w)
        out_dtype = out_dtype or x.dtype
        y = torch.empty(M, N, dtype=out_dtype, device=x.device)
        aux = torch.empty(M, N, dtype=x.dtype, device=x.device) if act == ops.ACT_GELU else None
        res2 = residual.reshape(M, N) if residual is not None else None
        with _lib.device_guard(x.device):
            mm(Op(x2, M, x2.stride(0)), Op(wq, N, wq.stride(0)), Out(y, N), M, N, K, bias=b.detach() if b is not None else None,
               aux=Out(aux, N) if aux is not None else None, residual=Out(res2, res2.stride(0)) if res2 is not None else None, act=act)
        ctx.save_for_backward(x2, w, aux)
        ctx.has_bias = b is not None
        ctx.has_res = residual is not None
        ctx.act = act
        ctx.in_shape = shp
        ctx.res_dtype = residual.dtype if residual is not None else None
        return y.reshape(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, aux = ctx.saved_tensors
        M, K = x2.shape
        N = w.shape[0]
        dy2 = dy.reshape(M, N)
        if dy2.stride(-1) != 1:
            dy2 = dy2.contiguous()
        dev = x2.device
        d_res = None
        with _lib.device_guard(dev):
            if ctx.has_res:
                d_res = dy if dy.dtype == ctx.res_dtype else convert(dy2, torch.empty(dy.shape, dtype=ctx.res_dtype, device=dev))
            if dy2.dtype != x2.dtype:
                dy2 = convert(dy2, torch.empty(M, N, dtype=x2.dtype, device=dev))
            if ctx.act == ops.ACT_GELU:
                dh = torch.empty(M, N, dtype=x2.dtype, device=dev)
                _lib_call("t4s_gelu_bwd", _p(dy2), _p(aux), _p(dh), M * N, ops.dtype_code(x2.dtype), _st())
            else:
                dh = dy2
            if (dh.stride(0) * dh.element_size()) % 16 or dh.data_ptr() % 16:
                # narrow heads (e.g. 10 classes): TMA needs a 16-byte row pitch -> padded copy, pad columns are never read
                Np = _pad8(N)
                dhp = torch.empty(M, Np, dtype=dh.dtype, device=dev)
                _lib_call("t4s_add2", _p(dh), dh.stride(0), _p(dh), dh.stride(0), _p(dhp), Np, M, N, 1.0, 0.0, ops.dtype_code(dh.dtype), _st())
                dh = dhp[:, :N]
            dx = dw = db = None
            if ctx.needs_input_grad[0]:
                wq = cast_weight(w)
                dx = torch.empty(M, K, dtype=x2.dtype, device=dev)
                mm(Op(dh, M, dh.stride(0)), Op(wq, K, wq.stride(0), mn_major=True), Out(dx, K), M, K, N)
                dx = dx.reshape(ctx.in_shape)
            if ctx.needs_input_grad[1]:
                dw = weight_grad(dh, x2, N, K)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = _attached_colsum(dy, N) if ctx.act != ops.ACT_GELU else None
                if db is None:
                    db = colsum(dh)
        return dx, dw, db, d_res, None, None


def linear(x, w, b=None, residual=None, act=ops.ACT_NONE, out_dtype=None):
    return _Linear.apply(x, w, b, residual, act, out_dtype)


class _Mlp(torch.autograd.Function):
    """y = fc2(gelu(fc1(x))) + residual (timm Mlp, passt.py:270-276) as one autograd node: the backward applies gelu' inside the
    epilogue of the fc2 dgrad GEMM (T4S_ACT_GELU_GRAD), so the hidden-size gradient is written once and never re-read by an
    element-wise kernel."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual):
        _lib.ensure_device(x)
        shp = x.shape
        K = shp[-1]
        x2 = x.reshape(-1, K)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        M, Hd, N = x2.shape[0], w1.shape[0], w2.shape[0]
        w1q, w2q = cast_weight(w1), cast_weight(w2)
        dt, dev = x.dtype, x.device
        with _lib.device_guard(dev):
            h = torch.empty(M, Hd, dtype=dt, device=dev)
            pre = torch.empty(M, Hd, dtype=dt, device=dev)
            mm(Op(x2, M, x2.stride(0)), Op(w1q, Hd, w1q.stride(0)), Out(h, Hd), M, Hd, K, bias=b1.detach(), aux=Out(pre, Hd), act=ops.ACT_GELU)
            y = torch.empty(M, N, dtype=dt, device=dev)
            res2 = residual.reshape(M, N) if residual is not None else None
            mm(Op(h, M, Hd), Op(w2q, N, w2q.stride(0)), Out(y, N), M, N, Hd, bias=b2.detach(),
               residual=Out(res2, res2.stride(0)) if res2 is not None else None)
        ctx.save_for_backward(x2, w1, w2, pre, h)
        ctx.in_shape = shp
        ctx.has_res = residual is not None
        return y.reshape(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w1, w2, pre, h = ctx.saved_tensors
        M, K = x2.shape
        Hd, N = w1.shape[0], w2.shape[0]
        dev = x2.device
        dy2 = dy.reshape(M, N)
        if dy2.stride(-1) != 1:
            dy2 = dy2.contiguous()
        ng = ctx.needs_input_grad
        with _lib.device_guard(dev):
            if dy2.dtype != x2.dtype:
                dy2 = convert(dy2, torch.empty(M, N, dtype=x2.dtype, device=dev))
            w1q, w2q = cast_weight(w1), cast_weight(w2)
            dpre = torch.empty(M, Hd, dtype=x2.dtype, device=dev)
            # bf16: the fc1 bias gradient (column sums of dpre) is accumulated by the epilogue of the GEMM that produces dpre
            fused_db1 = ng[2] and x2.dtype == torch.bfloat16 and M >= 32 and Hd % 8 == 0
            db1 = torch.zeros(Hd, dtype=torch.float32, device=dev) if fused_db1 else None
            mm(Op(dy2, M, dy2.stride(0)), Op(w2q, Hd, w2q.stride(0), mn_major=True), Out(dpre, Hd), M, Hd, N, residual=Out(pre, Hd),
               act=ops.ACT_GELU_GRAD, colsum=db1)
            dw2 = weight_grad(dy2, h, N, Hd) if ng[3] else None
            db2 = None
            if ng[4]:
                db2 = _attached_colsum(dy, N)
                if db2 is None:
                    db2 = colsum(dy2)
            dx = None
            if ng[0]:
                dx = torch.empty(M, K, dtype=x2.dtype, device=dev)
                mm(Op(dpre, M, Hd), Op(w1q, K, w1q.stride(0), mn_major=True), Out(dx, K), M, K, Hd)
                dx = dx.reshape(ctx.in_shape)
            dw1 = weight_grad(dpre, x2, Hd, K) if ng[1] else None
            if ng[2] and not fused_db1:
                db1 = colsum(dpre)
        d_res = dy if ctx.has_res else None
        return dx, dw1, db1, dw2, db2, d_res


def mlp(x, w1, b1, w2, b2, residual=None):
    return _Mlp.apply(x, w1, b1, w2, b2, residual)


# ------------------------------------------------------------------------------------------------------------------
# LayerNorm over the last dim of x [B, n, C], optionally skipping the first `skip` tokens of every clip
# ------------------------------------------------------------------------------------------------------------------
class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, in_scale, skip):
        _lib.ensure_device(x)
        x = x.contiguous()
        C = x.shape[-1]
        if skip:
            B, n_all = x.shape[0], x.shape[1]
            n_inner, bstride, off = n_all - skip, n_all * C, skip * C
            out_shape = (B, n_inner, C)
        else:
            n_inner, bstride, off = 0, 0, 0
            out_shape = x.shape
        rows = x.numel() // C if not skip else x.shape[0] * n_inner
        y = torch.empty(out_shape, dtype=x.dtype, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        xp = ctypes.c_void_p(x.data_ptr() + off * x.element_size())
        with _lib.device_guard(x.device):
            _lib_call("t4s_layernorm_fwd", xp, _p(gamma.detach()), _p(beta.detach()), _p(y), _p(mean), _p(rstd), rows, C, eps, in_scale,
                      ops.dtype_code(x.dtype), n_inner, bstride, _st())
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.cfg = (in_scale, skip, rows, C, n_inner, bstride, off)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        in_scale, skip, rows, C, n_inner, bstride, off = ctx.cfg
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = convert(dy, torch.empty(dy.shape, dtype=x.dtype, device=x.device))
        lib = _lib.load()
        dev = x.device
        with _lib.device_guard(dev):
            dx = torch.zeros_like(x) if skip else torch.empty_like(x)
            want = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
            dg = torch.empty(C, dtype=torch.float32, device=dev) if want else None
            db = torch.empty(C, dtype=torch.float32, device=dev) if want else None
            nbytes = lib.t4s_layernorm_bwd_workspace(rows, C) if want else 0
            ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
            es = x.element_size()
            _lib_call("t4s_layernorm_bwd", _p(dy), ctypes.c_void_p(x.data_ptr() + off * es), _p(gamma.detach()), _p(mean), _p(rstd),
                      ctypes.c_void_p(0), ctypes.c_void_p(dx.data_ptr() + off * es), _p(dg), _p(db), ctypes.c_void_p(0), _p(ws), nbytes, rows, C,
                      in_scale, ops.dtype_code(x.dtype), n_inner, bstride, _st())
        return dx, dg, db, None, None, None


def layer_norm(x, gamma, beta, eps=1e-5, in_scale=1.0, skip=0):
    return _LayerNorm.apply(x, gamma, beta, float(eps), float(in_scale), int(skip))


class _LayerNormRes(torch.autograd.Function):
    """Pre-norm residual fan-out as ONE node: (y, x_res) = (LN(x), x).  The sub-layer consumes y and adds x_res back in its GEMM
    epilogue; in the backward the gradient of the residual branch is added inside the LayerNorm-backward kernel (`dx_add`), so
    autograd never launches an element-wise add for the fan-out of x (24 adds of [B, N, D] per MAT-SED step)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        _lib.ensure_device(x)
        x = x.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        with _lib.device_guard(x.device):
            _lib_call("t4s_layernorm_fwd", _p(x), _p(gamma.detach()), _p(beta.detach()), _p(y), _p(mean), _p(rstd), rows, C, eps, 1.0,
                      ops.dtype_code(x.dtype), 0, 0, _st())
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.cfg = (rows, C)
        return y, x

    @staticmethod
    def backward(ctx, dy, dres):
        x, gamma, mean, rstd = ctx.saved_tensors
        rows, C = ctx.cfg
        dev = x.device
        if dy is None:
            return dres, None, None, None
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = convert(dy, torch.empty(dy.shape, dtype=x.dtype, device=dev))
        if dres is not None:
            dres = dres.contiguous()
            if dres.dtype != x.dtype:
                dres = convert(dres, torch.empty(dres.shape, dtype=x.dtype, device=dev))
        lib = _lib.load()
        with _lib.device_guard(dev):
            dx = torch.empty_like(x)
            want = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
            dg = torch.empty(C, dtype=torch.float32, device=dev) if want else None
            db = torch.empty(C, dtype=torch.float32, device=dev) if want else None
            # dx is the output gradient of the layer that wrote the residual stream (proj / fc2): its column sums are that layer's
            # bias gradient; the kernel accumulates them on the way out and the consumer finds them on the tensor (`_t4s_colsum`)
            fuse = _FUSE_BIAS_GRAD and x.dtype == torch.bfloat16 and C % 8 == 0
            dxs = torch.empty(C, dtype=torch.float32, device=dev) if fuse else None
            nbytes = lib.t4s_layernorm_bwd_workspace(rows, C) if (want or fuse) else 0
            ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
            _lib_call("t4s_layernorm_bwd", _p(dy), _p(x), _p(gamma.detach()), _p(mean), _p(rstd), _p(dres), _p(dx), _p(dg), _p(db), _p(dxs),
                      _p(ws), nbytes, rows, C, 1.0, ops.dtype_code(x.dtype), 0, 0, _st())
            if fuse:
                dx._t4s_colsum = dxs
        return dx, dg, db, None


def layer_norm_res(x, gamma, beta, eps=1e-5):
    """(LN(x), x): use the second output as the residual operand of the sub-layer that consumes the first."""
    return _LayerNormRes.apply(x, gamma, beta, float(eps))


# ------------------------------------------------------------------------------------------------------------------
# Multi-head self-attention on a fused qkv buffer [B, N, 3*D]      reference: passt.py:330-341
# ------------------------------------------------------------------------------------------------------------------
def _heads(qkv, B, N, D, H, which, mn=False):
    """Per-(head, clip) operand view of q (0) / k (1) / v (2) inside the fused buffer.  K-major: rows = tokens, K = hd;
    MN-major (`mn`): the token index is the contraction, rows = hd."""
    hd = D // H
    return Op(qkv, hd if mn else N, 3 * D, which * D, nb1=H, stride1=hd, nb2=B, stride2=N * 3 * D, mn_major=mn)


def _tok_heads(t, B, N, D, H, mn=False):
    """Same for a [B, N, D] tensor (attention output / its gradient, q+u, q+v)."""
    hd = D // H
    return Op(t, hd if mn else N, D, 0, nb1=H, stride1=hd, nb2=B, stride2=N * D, mn_major=mn)


class _Attention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, H):
        _lib.ensure_device(qkv)
        qkv = qkv.contiguous()
        B, N, D3 = qkv.shape
        D = D3 // 3
        hd = D // H
        Np = _pad8(N)
        scale = hd ** -0.5
        dt, dev = qkv.dtype, qkv.device
        with _lib.device_guard(dev):
            P = torch.empty(B, H, N, Np, dtype=dt, device=dev)
            mm(_heads(qkv, B, N, D, H, 0), _heads(qkv, B, N, D, H, 1), Out(P, Np, 0, N * Np, H * N * Np), N, N, hd, nb1=H, nb2=B, alpha=scale)
            _lib_call("t4s_softmax_fwd", _p(P), _p(P), B * H * N, N, Np, Np, ops.dtype_code(dt), _st())
            o = torch.empty(B, N, D, dtype=dt, device=dev)
            mm(Op(P, N, Np, 0, nb1=H, stride1=N * Np, nb2=B, stride2=H * N * Np), _heads(qkv, B, N, D, H, 2, mn=True),
               Out(o, D, 0, hd, N * D), N, hd, N, nb1=H, nb2=B)
        ctx.save_for_backward(qkv, P)
        ctx.H = H
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, P = ctx.saved_tensors
        H = ctx.H
        B, N, D3 = qkv.shape
        D = D3 // 3
        hd = D // H
        Np = P.shape[-1]
        scale = hd ** -0.5
        dt, dev = qkv.dtype, qkv.device
        do = do.contiguous()
        if do.dtype != dt:
            do = convert(do, torch.empty(do.shape, dtype=dt, device=dev))
        with _lib.device_guard(dev):
            dqkv = torch.empty_like(qkv)
            dP = torch.empty(B, H, N, Np, dtype=dt, device=dev)
            pmat = dict(nb1=H, stride1=N * Np, nb2=B, stride2=H * N * Np)
            # dP = dO V^T
            mm(_tok_heads(do, B, N, D, H), _heads(qkv, B, N, D, H, 2), Out(dP, Np, 0, N * Np, H * N * Np), N, N, hd, nb1=H, nb2=B)
            # dV = P^T dO
            mm(Op(P, N, Np, 0, mn_major=True, **pmat), _tok_heads(do, B, N, D, H, mn=True), Out(dqkv, 3 * D, 2 * D, hd, N * 3 * D), N, hd, N,
               nb1=H, nb2=B)
            _lib_call("t4s_softmax_bwd", _p(P), _p(dP), B * H * N, N, Np, Np, ops.dtype_code(dt), _st())
            # dQ = scale dS K ; dK = scale dS^T Q
            mm(Op(dP, N, Np, 0, **pmat), _heads(qkv, B, N, D, H, 1, mn=True), Out(dqkv, 3 * D, 0, hd, N * 3 * D), N, hd, N, nb1=H, nb2=B,
               alpha=scale)
            mm(Op(dP, N, Np, 0, mn_major=True, **pmat), _heads(qkv, B, N, D, H, 0, mn=True), Out(dqkv, 3 * D, D, hd, N * 3 * D), N, hd, N, nb1=H,
               nb2=B, alpha=scale)
        return dqkv, None


# fp32 copy of the attention output for delta = rowsum(dO * O) in the backward (T4S_ATTN_O32=0: delta from the bf16 output)
_ATTN_O32 = os.environ.get("T4S_ATTN_O32", "1") != "0"


def _attn_desc(qkv, o, lse, B, N, D, H, scale, o32=None):
    a = _lib.Attn()
    a.batch, a.heads, a.tokens, a.head_dim, a.scale = B, H, N, D // H, scale
    es = qkv.element_size()
    base = qkv.data_ptr()
    a.q, a.k, a.v = base, base + D * es, base + 2 * D * es
    a.q_ld = a.k_ld = a.v_ld = 3 * D
    a.q_bs = a.k_bs = a.v_bs = N * 3 * D
    a.o, a.o_ld, a.o_bs = o.data_ptr(), D, N * D
    a.lse = lse.data_ptr()
    a.o32 = o32.data_ptr() if o32 is not None else None
    return a


class _FlashAttention(torch.autograd.Function):
    """Fused tcgen05 attention (csrc/attn.cu): no [B, H, N, N] tensor is materialised; bf16 mode, head_dim 64."""

    @staticmethod
    def forward(ctx, qkv, H):
        _lib.ensure_device(qkv)
        qkv = qkv.contiguous()
        B, N, D3 = qkv.shape
        D = D3 // 3
        dev = qkv.device
        lib = _lib.load()
        Nl = lib.t4s_attn_padded_len(N)
        with _lib.device_guard(dev):
            o = torch.empty(B, N, D, dtype=qkv.dtype, device=dev)
            lse = torch.empty(B, H, Nl, dtype=torch.float32, device=dev)
            # un-rounded copy of o for the backward's delta (kept only when a backward can follow)
            o32 = torch.empty(B, N, D, dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] and _ATTN_O32 else None
            a = _attn_desc(qkv, o, lse, B, N, D, H, (D // H) ** -0.5, o32)
            _lib_call("t4s_attn_fwd", ctypes.byref(a), _st(), _key=(B, H, N))
        ctx.save_for_backward(qkv, o, lse, o32)
        ctx.H = H
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse, o32 = ctx.saved_tensors
        H = ctx.H
        B, N, D3 = qkv.shape
        D = D3 // 3
        dev = qkv.device
        do = do.contiguous()
        if do.dtype != qkv.dtype:
            do = convert(do, torch.empty(do.shape, dtype=qkv.dtype, device=dev))
        with _lib.device_guard(dev):
            dqkv = torch.empty_like(qkv)
            delta = torch.empty_like(lse)
            g = _lib.AttnBwd()
            g.fwd = _attn_desc(qkv, o, lse, B, N, D, H, (D // H) ** -0.5, o32)
            g.d_o, g.do_ld, g.do_bs = do.data_ptr(), D, N * D
            g.delta = delta.data_ptr()
            es = dqkv.element_size()
            g.dq, g.dk, g.dv = dqkv.data_ptr(), dqkv.data_ptr() + D * es, dqkv.data_ptr() + 2 * D * es
            g.dq_ld = g.dk_ld = g.dv_ld = 3 * D
            g.dq_bs = g.dk_bs = g.dv_bs = N * 3 * D
            dq32 = torch.empty(B, N, D, dtype=torch.float32, device=dev) if _FUSED_ATTN_BWD else None   # zeroed by the call
            g.dq32 = dq32.data_ptr() if dq32 is not None else None
            # the qkv projection's bias gradient = column sums of dqkv: accumulated by the kernels that write dq / dk / dv
            cs = torch.zeros(3 * D, dtype=torch.float32, device=dev) if (_FUSED_ATTN_BWD and _FUSE_BIAS_GRAD and D // 8 <= 256) else None
            g.dqkv_colsum = cs.data_ptr() if cs is not None else None
            _lib_call("t4s_attn_bwd", ctypes.byref(g), _st(), _key=(B, H, N))
            if cs is not None:
                dqkv._t4s_colsum = cs
        return dqkv, None


def set_fused_attention(flag: bool):
    """Debug / A-B switch: False routes bf16 attention through the unfused GEMM + softmax kernels."""
    global _FUSED_ATTN
    _FUSED_ATTN = bool(flag)


def set_fused_attention_backward(flag: bool):
    """False selects the deterministic two-kernel attention backward (dQ summed in a fixed order)."""
    global _FUSED_ATTN_BWD
    _FUSED_ATTN_BWD = bool(flag)


def attention(qkv, num_heads):
    if _FUSED_ATTN and qkv.dtype == torch.bfloat16 and qkv.shape[-1] // 3 // num_heads == 64:
        return _FlashAttention.apply(qkv, num_heads)
    return _Attention.apply(qkv, num_heads)


# ------------------------------------------------------------------------------------------------------------------
# Transformer-XL relative-position attention               reference: transformerXL.py:299-593, rel_shift :254-297
# score[i, j] = ((q_i + u) k_j + (q_i + v) p_{T-1-i+j}) / sqrt(hd)
# ------------------------------------------------------------------------------------------------------------------
class _RelPosAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, p_lin, u, v, H):
        _lib.ensure_device(qkv)
        qkv = qkv.contiguous()
        p_lin = p_lin.contiguous()
        B, T, D3 = qkv.shape
        D = D3 // 3
        hd = D // H
        L = 2 * T - 1
        Tp, Lp = _pad8(T), _pad8(L)
        scale = hd ** -0.5
        dt, dev = qkv.dtype, qkv.device
        code = ops.dtype_code(dt)
        with _lib.device_guard(dev):
            qu = torch.empty(B, T, D, dtype=dt, device=dev)
            qv = torch.empty(B, T, D, dtype=dt, device=dev)
            _lib_call("t4s_add_rowvec", _p(qkv), 3 * D, _p(u.detach().reshape(-1)), _p(qu), B * T, D, 1.0, code, _st())
            _lib_call("t4s_add_rowvec", _p(qkv), 3 * D, _p(v.detach().reshape(-1)), _p(qv), B * T, D, 1.0, code, _st())
            P = torch.empty(B, H, T, Tp, dtype=dt, device=dev)
            BD = torch.empty(B, H, T, Lp, dtype=dt, device=dev)
            mm(_tok_heads(qu, B, T, D, H), _heads(qkv, B, T, D, H, 1), Out(P, Tp, 0, T * Tp, H * T * Tp), T, T, hd, nb1=H, nb2=B, alpha=scale)
            mm(_tok_heads(qv, B, T, D, H), Op(p_lin, L, D, 0, nb1=H, stride1=hd), Out(BD, Lp, 0, T * Lp, H * T * Lp), T, L, hd, nb1=H, nb2=B,
               alpha=scale)
            _lib_call("t4s_relpos_softmax_fwd", _p(P), _p(BD), _p(P), B * H * T, T, Tp, Lp, Tp, code, _st())
            del BD
            o = torch.empty(B, T, D, dtype=dt, device=dev)
            mm(Op(P, T, Tp, 0, nb1=H, stride1=T * Tp, nb2=B, stride2=H * T * Tp), _heads(qkv, B, T, D, H, 2, mn=True),
               Out(o, D, 0, hd, T * D), T, hd, T, nb1=H, nb2=B)
        ctx.save_for_backward(qkv, p_lin, u, v, P)
        ctx.H = H
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, p_lin, u, v, P = ctx.saved_tensors
        H = ctx.H
        B, T, D3 = qkv.shape
        D = D3 // 3
        hd = D // H
        L = 2 * T - 1
        Tp, Lp = P.shape[-1], _pad8(L)
        scale = hd ** -0.5
        dt, dev = qkv.dtype, qkv.device
        code = ops.dtype_code(dt)
        do = do.contiguous()
        if do.dtype != dt:
            do = convert(do, torch.empty(do.shape, dtype=dt, device=dev))
        with _lib.device_guard(dev):
            qu = torch.empty(B, T, D, dtype=dt, device=dev)
            qv = torch.empty(B, T, D, dtype=dt, device=dev)
            _lib_call("t4s_add_rowvec", _p(qkv), 3 * D, _p(u.detach().reshape(-1)), _p(qu), B * T, D, 1.0, code, _st())
            _lib_call("t4s_add_rowvec", _p(qkv), 3 * D, _p(v.detach().reshape(-1)), _p(qv), B * T, D, 1.0, code, _st())
            pm = dict(nb1=H, stride1=T * Tp, nb2=B, stride2=H * T * Tp)
            bm = dict(nb1=H, stride1=T * Lp, nb2=B, stride2=H * T * Lp)
            dqkv = torch.empty_like(qkv)
            dP = torch.empty(B, H, T, Tp, dtype=dt, device=dev)
            dBD = torch.empty(B, H, T, Lp, dtype=dt, device=dev)
            mm(_tok_heads(do, B, T, D, H), _heads(qkv, B, T, D, H, 2), Out(dP, Tp, 0, T * Tp, H * T * Tp), T, T, hd, nb1=H, nb2=B)
            mm(Op(P, T, Tp, 0, mn_major=True, **pm), _tok_heads(do, B, T, D, H, mn=True), Out(dqkv, 3 * D, 2 * D, hd, T * 3 * D), T, hd, T,
               nb1=H, nb2=B)
            _lib_call("t4s_relpos_softmax_bwd", _p(P), _p(dP), _p(dBD), B * H * T, T, Tp, Tp, Lp, code, _st())
            # d(q+u) = scale dAC K ; dK = scale dAC^T (q+u)
            dqu = torch.empty(B, T, D, dtype=dt, device=dev)
            mm(Op(dP, T, Tp, 0, **pm), _heads(qkv, B, T, D, H, 1, mn=True), Out(dqu, D, 0, hd, T * D), T, hd, T, nb1=H, nb2=B, alpha=scale)
            mm(Op(dP, T, Tp, 0, mn_major=True, **pm), _tok_heads(qu, B, T, D, H, mn=True), Out(dqkv, 3 * D, D, hd, T * 3 * D), T, hd, T, nb1=H,
               nb2=B, alpha=scale)
            # d(q+v) = scale dBD p ; dp = scale sum_b dBD^T (q+v)
            dqv = torch.empty(B, T, D, dtype=dt, device=dev)
            # dBD[i, k] is zero unless T-1 <= i + k <= 2T-2 (the kernel only ever writes that band): the GEMMs skip the k-blocks outside it
            band = (T - 1, 2 * T - 1)
            mm(Op(dBD, T, Lp, 0, **bm), Op(p_lin, hd, D, 0, nb1=H, stride1=hd, mn_major=True), Out(dqv, D, 0, hd, T * D), T, hd, L, nb1=H, nb2=B,
               alpha=scale, band=band)
            dp_lin = None
            if ctx.needs_input_grad[1]:
                ws = torch.empty(B, L, D, dtype=torch.float32, device=dev)
                mm(Op(dBD, L, Lp, 0, mn_major=True, **bm), _tok_heads(qv, B, T, D, H, mn=True), Out(ws, D, 0, hd, L * D), L, hd, T, nb1=H, nb2=B,
                   alpha=scale, band=band)
                dp32 = torch.empty(L, D, dtype=torch.float32, device=dev)
                ops.reduce_splits(ws, B, L * D, dp32)
                dp_lin = dp32 if dt == torch.float32 else convert(dp32, torch.empty(L, D, dtype=dt, device=dev))
            _lib_call("t4s_add2", _p(dqu), D, _p(dqv), D, _p(dqkv), 3 * D, B * T, D, 1.0, 1.0, code, _st())
            du = colsum(dqu.reshape(B * T, D)).reshape(u.shape) if ctx.needs_input_grad[2] else None
            dv = colsum(dqv.reshape(B * T, D)).reshape(v.shape) if ctx.needs_input_grad[3] else None
        return dqkv, dp_lin, du, dv, None


_dbd_cache = {}


def _dbd_buffer(B, H, T, Lp, dev):
    """Position-coordinate score gradient [B, H, T, Lp] (bf16).  The fused backward only ever writes the band
    T-1-i <= x <= 2T-2-i of row i, so the buffer is zeroed once and then reused by every layer and step."""
    key = (B, H, T, Lp, str(dev))
    t = _dbd_cache.get(key)
    if t is None:
        _dbd_cache.clear()
        t = torch.zeros(B, H, T, Lp, dtype=torch.bfloat16, device=dev)
        _dbd_cache[key] = t
    return t


def _relattn_desc(qkv, qu, qv, p_lin, o, lse, B, T, D, H, scale, o32=None):
    a = _lib.RelAttn()
    a.batch, a.heads, a.tokens, a.head_dim, a.scale = B, H, T, D // H, scale
    es = qkv.element_size()
    a.qu, a.qu_ld, a.qu_bs = qu.data_ptr(), D, T * D
    a.qv, a.qv_ld, a.qv_bs = qv.data_ptr(), D, T * D
    a.k, a.v = qkv.data_ptr() + D * es, qkv.data_ptr() + 2 * D * es
    a.k_ld = a.v_ld = 3 * D
    a.k_bs = a.v_bs = T * 3 * D
    a.pos, a.pos_ld = p_lin.data_ptr(), p_lin.stride(0)
    a.o, a.o_ld, a.o_bs = o.data_ptr(), D, T * D
    a.lse = lse.data_ptr()
    a.o32 = o32.data_ptr() if o32 is not None else None
    return a


class _FlashRelPosAttention(torch.autograd.Function):
    """Fused tcgen05 Transformer-XL attention (csrc/attn_rel.cu): AC, the position scores and their rel_shift live in
    TMEM / shared memory only.  bf16 mode, head_dim 64."""

    @staticmethod
    def forward(ctx, qkv, p_lin, u, v, H):
        _lib.ensure_device(qkv)
        qkv = qkv.contiguous()
        p_lin = p_lin.contiguous()
        B, T, D3 = qkv.shape
        D = D3 // 3
        dt, dev = qkv.dtype, qkv.device
        code = ops.dtype_code(dt)
        lib = _lib.load()
        Nl = lib.t4s_attn_padded_len(T)
        with _lib.device_guard(dev):
            qu = torch.empty(B, T, D, dtype=dt, device=dev)
            qv = torch.empty(B, T, D, dtype=dt, device=dev)
            _lib_call("t4s_add_rowvec", _p(qkv), 3 * D, _p(u.detach().reshape(-1)), _p(qu), B * T, D, 1.0, code, _st())
            _lib_call("t4s_add_rowvec", _p(qkv), 3 * D, _p(v.detach().reshape(-1)), _p(qv), B * T, D, 1.0, code, _st())
            o = torch.empty(B, T, D, dtype=dt, device=dev)
            lse = torch.empty(B, H, Nl, dtype=torch.float32, device=dev)
            o32 = torch.empty(B, T, D, dtype=torch.float32, device=dev) if any(ctx.needs_input_grad[:4]) and _ATTN_O32 else None
            a = _relattn_desc(qkv, qu, qv, p_lin, o, lse, B, T, D, H, (D // H) ** -0.5, o32)
            _lib_call("t4s_relattn_fwd", ctypes.byref(a), _st(), _key=(B, H, T))
        ctx.save_for_backward(qkv, p_lin, u, v, qu, qv, o, lse, o32)
        ctx.H = H
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, p_lin, u, v, qu, qv, o, lse, o32 = ctx.saved_tensors
        H = ctx.H
        B, T, D3 = qkv.shape
        D = D3 // 3
        hd = D // H
        L = 2 * T - 1
        Lp = _pad8(L)
        scale = hd ** -0.5
        dt, dev = qkv.dtype, qkv.device
        code = ops.dtype_code(dt)
        do = do.contiguous()
        if do.dtype != dt:
            do = convert(do, torch.empty(do.shape, dtype=dt, device=dev))
        with _lib.device_guard(dev):
            dqkv = torch.empty_like(qkv)
            dqu = torch.empty(B, T, D, dtype=dt, device=dev)
            delta = torch.empty_like(lse)
            dBD = _dbd_buffer(B, H, T, Lp, dev)
            g = _lib.RelAttnBwd()
            g.fwd = _relattn_desc(qkv, qu, qv, p_lin, o, lse, B, T, D, H, scale, o32)
            g.d_o, g.do_ld, g.do_bs = do.data_ptr(), D, T * D
            g.delta = delta.data_ptr()
            es = dqkv.element_size()
            g.dqu, g.dqu_ld, g.dqu_bs = dqu.data_ptr(), D, T * D
            g.dk, g.dv = dqkv.data_ptr() + D * es, dqkv.data_ptr() + 2 * D * es
            g.dk_ld = g.dv_ld = 3 * D
            g.dk_bs = g.dv_bs = T * 3 * D
            g.dbd, g.dbd_ld = dBD.data_ptr(), Lp
            _lib_call("t4s_relattn_bwd", ctypes.byref(g), _st(), _key=(B, H, T))
            bm = dict(nb1=H, stride1=T * Lp, nb2=B, stride2=H * T * Lp)
            # d(q+v) = scale dBD p ; dp = scale sum_b dBD^T (q+v)
            dqv = torch.empty(B, T, D, dtype=dt, device=dev)
            # dBD[i, k] is zero unless T-1 <= i + k <= 2T-2 (the kernel only ever writes that band): the GEMMs skip the k-blocks outside it
            band = (T - 1, 2 * T - 1)
            mm(Op(dBD, T, Lp, 0, **bm), Op(p_lin, hd, D, 0, nb1=H, stride1=hd, mn_major=True), Out(dqv, D, 0, hd, T * D), T, hd, L, nb1=H, nb2=B,
               alpha=scale, band=band)
            dp_lin = None
            if ctx.needs_input_grad[1]:
                ws = torch.empty(B, L, D, dtype=torch.float32, device=dev)
                mm(Op(dBD, L, Lp, 0, mn_major=True, **bm), _tok_heads(qv, B, T, D, H, mn=True), Out(ws, D, 0, hd, L * D), L, hd, T, nb1=H, nb2=B,
                   alpha=scale, band=band)
                dp32 = torch.empty(L, D, dtype=torch.float32, device=dev)
                ops.reduce_splits(ws, B, L * D, dp32)
                dp_lin = convert(dp32, torch.empty(L, D, dtype=dt, device=dev))
            _lib_call("t4s_add2", _p(dqu), D, _p(dqv), D, _p(dqkv), 3 * D, B * T, D, 1.0, 1.0, code, _st())
            du = colsum(dqu.reshape(B * T, D)).reshape(u.shape) if ctx.needs_input_grad[2] else None
            dv = colsum(dqv.reshape(B * T, D)).reshape(v.shape) if ctx.needs_input_grad[3] else None
        return dqkv, dp_lin, du, dv, None


def relpos_attention(qkv, p_lin, pos_bias_u, pos_bias_v, num_heads):
    if _FUSED_ATTN and qkv.dtype == torch.bfloat16 and qkv.shape[-1] // 3 // num_heads == 64:
        return _FlashRelPosAttention.apply(qkv, p_lin, pos_bias_u, pos_bias_v, num_heads)
    return _RelPosAttention.apply(qkv, p_lin, pos_bias_u, pos_bias_v, num_heads)


_pos_tables = {}


def rel_pos_table(T, d_model, device, dtype):
    """Sin/cos relative-position table [2T-1, d]: row k <-> relative position T-1-k (transformerXL.py:68-101,120-126).
    A constant of (T, d): built once on the host exactly as the reference builds `pe`, then cached on the device."""
    key = (T, d_model, str(device), dtype)
    t = _pos_tables.get(key)
    if t is None:
        rel = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
        div = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
        pe = torch.zeros(2 * T - 1, d_model)
        pe[:, 0::2] = torch.sin(rel * div)
        pe[:, 1::2] = torch.cos(rel * div)
        t = pe.to(device=device, dtype=dtype)
        _pos_tables[key] = t
    return t


# ------------------------------------------------------------------------------------------------------------------
# Patch embedding + positional tables + cls/dist tokens          reference: passt.py:302-315, 496-569
# ------------------------------------------------------------------------------------------------------------------
class _PatchEmbed(torch.autograd.Function):
    """`windows` = None (whole image, one positional offset) or (starts, t_dim, t_offsets): every time window
    [start, start + (t_dim-1)*stride + patch) of the image becomes one sequence; output rows are window-major
    [(w, b), n_tok, D] (sliding-window fusion, reference encoder_slide_window.py:29-33)."""

    @staticmethod
    def forward(ctx, mel, conv_w, conv_b, time_pos, freq_pos, cls, dist, new_pos, stride, t_offset, windows):
        _lib.ensure_device(mel)
        mel = mel.contiguous()
        B, Hh, W = mel.shape
        D, _, P, _ = conv_w.shape
        F = (Hh - P) // stride + 1
        Tt = time_pos.shape[-1]
        if windows is None:
            Tp_full = (W - P) // stride + 1
            Tp = min(Tp_full, Tt)          # a longer grid is cropped to the table (passt.py:515)
            starts, offsets = None, [int(t_offset)]
        else:
            starts, Tp, offsets = [int(v) for v in windows[0]], int(windows[1]), [int(v) for v in windows[2]]
            if Tp > Tt or len(offsets) != len(starts):
                raise _lib.T4sError("patch_embed: bad window specification")
        nW = len(offsets)
        dt, dev = act_dtype(), mel.device
        n_tok = 2 + F * Tp
        PP = P * P
        with _lib.device_guard(dev):
            A = torch.empty(nW * B * F * Tp, PP, dtype=dt, device=dev)
            if starts is None:
                _lib_call("t4s_patch_im2col", _p(mel), ops.dtype_code(mel.dtype), _p(A), ops.dtype_code(dt), B, Hh, W, P, stride, F, Tp, _st())
            else:
                arr = (ctypes.c_int * nW)(*starts)
                _lib_call("t4s_patch_im2col_windows", _p(mel), ops.dtype_code(mel.dtype), _p(A), ops.dtype_code(dt), B, Hh, W, arr, nW, P, stride,
                          F, Tp, _st())
            same = all(o == offsets[0] for o in offsets)
            n_pos = 1 if same else nW
            pos = torch.empty(n_pos, F * Tp, D, dtype=torch.float32, device=dev)
            for i in range(n_pos):
                _lib_call("t4s_patch_posbias", _p(time_pos.detach()), _p(freq_pos.detach()), ctypes.c_void_p(pos.data_ptr() + i * F * Tp * D * 4),
                          D, F, Tp, Tt, offsets[i], _st())
            x = torch.empty(nW * B, n_tok, D, dtype=dt, device=dev)
            w2 = cast_weight(conv_w).reshape(D, PP)
            mm(Op(A, F * Tp, PP, 0, nb1=B, stride1=F * Tp * PP, nb2=nW, stride2=B * F * Tp * PP), Op(w2, D, PP),
               Out(x, D, 2 * D, n_tok * D, B * n_tok * D), F * Tp, D, PP, nb1=B, nb2=nW, bias=conv_b.detach(),
               residual=Out(pos, D, 0, 0, 0 if same else F * Tp * D))
            _lib_call("t4s_cls_dist_tokens", _p(x), ops.dtype_code(dt), _p(cls.detach()), _p(dist.detach()), _p(new_pos.detach()), nW * B,
                      n_tok * D, D, _st())
        ctx.save_for_backward(A, conv_w)
        ctx.cfg = (B, F, Tp, Tt, offsets, D, P, n_tok, time_pos.shape, freq_pos.shape, cls.shape, new_pos.shape)
        return x

    @staticmethod
    def backward(ctx, dx):
        A, conv_w = ctx.saved_tensors
        B, F, Tp, Tt, offsets, D, P, n_tok, tshape, fshape, cshape, nshape = ctx.cfg
        nW = len(offsets)
        dev = dx.device
        dx = dx.contiguous()
        if dx.dtype != A.dtype:
            dx = convert(dx, torch.empty(dx.shape, dtype=A.dtype, device=dev))
        ng = ctx.needs_input_grad
        PP = P * P
        with _lib.device_guard(dev):
            f32 = dict(dtype=torch.float32, device=dev)
            tmp = torch.empty(n_tok * D, **f32)
            shapes = (tshape if ng[3] else None, fshape if ng[4] else None, (D,) if ng[2] else None, cshape if ng[5] else None,
                      cshape if ng[6] else None, nshape if ng[7] else None)
            same = all(o == offsets[0] for o in offsets)
            groups = 1 if same else nW            # windows sharing a positional offset reduce together (batch = all their clips)
            parts = [torch.empty((groups,) + tuple(sh), **f32) if sh is not None else None for sh in shapes]
            if any(g is not None for g in parts):
                per = nW * B // groups
                es = dx.element_size()
                for i in range(groups):
                    ptrs = [ctypes.c_void_p(g.data_ptr() + i * g[0].numel() * 4) if g is not None else ctypes.c_void_p(0) for g in parts]
                    _lib_call("t4s_patch_small_grads", ctypes.c_void_p(dx.data_ptr() + i * per * n_tok * D * es), ops.dtype_code(dx.dtype), _p(tmp),
                              ptrs[0], ptrs[1], ptrs[2], ptrs[3], ptrs[4], ptrs[5], per, n_tok * D, D, F, Tp, Tt, offsets[i], _st())
            outs = []
            for g in parts:
                if g is None or groups == 1:
                    outs.append(g[0] if g is not None else None)
                else:
                    o = torch.empty(g.shape[1:], **f32)
                    ops.reduce_splits(g, groups, o.numel(), o)
                    outs.append(o)
            d_time, d_freq, d_bias, d_cls, d_dist, d_new = outs
            dw = None
            if ng[1]:
                nb = nW * B
                ws = torch.empty(nb, D, PP, **f32)
                mm(Op(dx, D, D, 2 * D, nb1=nb, stride1=n_tok * D, mn_major=True), Op(A, PP, PP, 0, nb1=nb, stride1=F * Tp * PP, mn_major=True),
                   Out(ws, PP, 0, D * PP), D, PP, F * Tp, nb1=nb)
                dw = torch.empty(D, PP, **f32)
                ops.reduce_splits(ws, nb, D * PP, dw)
                dw = dw.reshape(conv_w.shape)
        return None, dw, d_bias, d_time, d_freq, d_cls, d_dist, d_new, None, None, None


def patch_embed(mel, conv_w, conv_b, time_pos, freq_pos, cls, dist, new_pos, stride=10, t_offset=0, windows=None):
    return _PatchEmbed.apply(mel, conv_w, conv_b, time_pos, freq_pos, cls, dist, new_pos, stride, t_offset, windows)


# ------------------------------------------------------------------------------------------------------------------
# frequency mean-pool, pad + interpolate                          reference: passt_sed.py:199-218, 23-34, 258-259
# ------------------------------------------------------------------------------------------------------------------
class _FpoolMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, F, Tp):
        _lib.ensure_device(y)
        y = y.contiguous()
        B, _, C = y.shape
        out = torch.empty(B, Tp, C, dtype=y.dtype, device=y.device)
        with _lib.device_guard(y.device):
            _lib_call("t4s_fpool_mean_fwd", _p(y), _p(out), ops.dtype_code(y.dtype), B, F, Tp, C, _st())
        ctx.cfg = (B, F, Tp, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, F, Tp, C = ctx.cfg
        dout = dout.contiguous()
        dy = torch.empty(B, F * Tp, C, dtype=dout.dtype, device=dout.device)
        with _lib.device_guard(dout.device):
            _lib_call("t4s_fpool_mean_bwd", _p(dout), _p(dy), ops.dtype_code(dout.dtype), B, F, Tp, C, _st())
        return dy, None, None


def fpool_mean(y, f_dim, t_dim):
    return _FpoolMean.apply(y, f_dim, t_dim)


class _PadInterp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ratio, pad):
        _lib.ensure_device(x)
        x = x.contiguous()
        B, Tin, C = x.shape
        out = torch.empty(B, (Tin + pad) * ratio, C, dtype=x.dtype, device=x.device)
        with _lib.device_guard(x.device):
            _lib_call("t4s_pad_interp_fwd", _p(x), _p(out), ops.dtype_code(x.dtype), B, Tin, ratio, C, pad, _st())
        ctx.cfg = (B, Tin, ratio, C, pad)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, Tin, ratio, C, pad = ctx.cfg
        dout = dout.contiguous()
        dx = torch.empty(B, Tin, C, dtype=dout.dtype, device=dout.device)
        with _lib.device_guard(dout.device):
            _lib_call("t4s_pad_interp_bwd", _p(dout), _p(dx), ops.dtype_code(dout.dtype), B, Tin, ratio, C, pad, _st())
        return dx, None, None


def pad_interpolate(x, ratio, pad=True):
    """(optionally repeat the last frame, then) linear interpolation x ratio along time, align_corners=False."""
    return _PadInterp.apply(x, int(ratio), int(bool(pad)))


# ------------------------------------------------------------------------------------------------------------------
# sliding-window overlap-add mean, global/local blend            reference: encoder_slide_window.py:24-36, passt_sed.py:266-271
# ------------------------------------------------------------------------------------------------------------------
def _segments(groups, B, C):
    """groups: [(tensor [n_w*B, L, C] window-major, out_starts [n_w])] -> ctypes array of T4sWindowSegment."""
    segs = []
    for t, starts in groups:
        L = t.shape[1]
        es = t.element_size()
        for i, st in enumerate(starts):
            segs.append(_lib.WindowSegment(ctypes.c_void_p(t.data_ptr() + i * B * L * C * es), L * C, int(st), L))
    return (_lib.WindowSegment * len(segs))(*segs), len(segs)


class _OverlapAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, frames, B, spec, *locals_):
        _lib.ensure_device(locals_[0])
        locals_ = [t.contiguous() for t in locals_]
        C = locals_[0].shape[-1]
        dt, dev = locals_[0].dtype, locals_[0].device
        segs, n = _segments(list(zip(locals_, spec)), B, C)
        out = torch.empty(B, frames, C, dtype=dt, device=dev)
        with _lib.device_guard(dev):
            _lib_call("t4s_window_overlap_add_fwd", segs, n, _p(out), ops.dtype_code(dt), B, frames, C, _st())
        ctx.cfg = (frames, B, spec, [t.shape for t in locals_], dt)
        return out

    @staticmethod
    def backward(ctx, dout):
        frames, B, spec, shapes, dt = ctx.cfg
        dev = dout.device
        dout = dout.contiguous()
        if dout.dtype != dt:
            dout = convert(dout, torch.empty(dout.shape, dtype=dt, device=dev))
        C = dout.shape[-1]
        grads = [torch.empty(sh, dtype=dt, device=dev) for sh in shapes]
        segs, n = _segments(list(zip(grads, spec)), B, C)
        with _lib.device_guard(dev):
            _lib_call("t4s_window_overlap_add_bwd", _p(dout), segs, n, ops.dtype_code(dt), B, frames, C, _st())
        return (None, None, None) + tuple(grads)


def window_overlap_add(groups, batch, frames):
    """groups: list of (local [n_w*batch, L, C] window-major, out_starts list) -> [batch, frames, C]: mean over the windows
    covering each frame, 0 where none does."""
    spec = tuple(tuple(int(v) for v in st) for _, st in groups)
    return _OverlapAdd.apply(int(frames), int(batch), spec, *[t for t, _ in groups])


class _Lerp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, w):
        _lib.ensure_device(a)
        a, b = a.contiguous(), b.contiguous()
        if b.dtype != a.dtype:
            b = convert(b, torch.empty(b.shape, dtype=a.dtype, device=a.device))
        C = a.shape[-1]
        out = torch.empty_like(a)
        with _lib.device_guard(a.device):
            _lib_call("t4s_add2", _p(a), C, _p(b), C, _p(out), C, a.numel() // C, C, 1.0 - w, w, ops.dtype_code(a.dtype), _st())
        ctx.w = w
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        C = dout.shape[-1]
        code = ops.dtype_code(dout.dtype)
        da = db = None
        with _lib.device_guard(dout.device):
            if ctx.needs_input_grad[0]:
                da = torch.empty_like(dout)
                _lib_call("t4s_add_rowvec", _p(dout), C, ctypes.c_void_p(0), _p(da), dout.numel() // C, C, 1.0 - ctx.w, code, _st())
            if ctx.needs_input_grad[1]:
                db = torch.empty_like(dout)
                _lib_call("t4s_add_rowvec", _p(dout), C, ctypes.c_void_p(0), _p(db), dout.numel() // C, C, ctx.w, code, _st())
        return da, db, None


def lerp(a, b, w):
    """(1 - w) * a + w * b   (mix_rate blend of the global and the sliding-window embeddings, passt_sed.py:271)."""
    return _Lerp.apply(a, b, float(w))


# ------------------------------------------------------------------------------------------------------------------
# MLM frame replacement                                           reference: transformer/mask.py:62-82
# ------------------------------------------------------------------------------------------------------------------
class _MaskRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, token, kind, src):
        _lib.ensure_device(x)
        x = x.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        tok = token.detach().reshape(-1).float().contiguous()
        out = torch.empty_like(x)
        with _lib.device_guard(x.device):
            _lib_call("t4s_mask_rows_fwd", _p(x), _p(tok), _p(kind), _p(src), _p(out), rows, C, ops.dtype_code(x.dtype), _st())
        ctx.save_for_backward(kind, src)
        ctx.cfg = (rows, C, token.shape, x.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        kind, src = ctx.saved_tensors
        rows, C, tshape, dt = ctx.cfg
        dev = dout.device
        dout = dout.contiguous()
        if dout.dtype != dt:
            dout = convert(dout, torch.empty(dout.shape, dtype=dt, device=dev))
        lib = _lib.load()
        copy_rows = torch.nonzero(kind == 2).reshape(-1)  # index bookkeeping (ascending); the arithmetic is in the kernel
        with _lib.device_guard(dev):
            dx = torch.empty_like(dout) if ctx.needs_input_grad[0] else None
            dtok = torch.empty(C, dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
            nbytes = lib.t4s_mask_rows_bwd_workspace(rows, C)
            ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
            _lib_call("t4s_mask_rows_bwd", _p(dout), _p(kind), _p(src), _p(copy_rows), copy_rows.numel(), _p(dx), _p(dtok), _p(ws), nbytes, rows,
                      C, ops.dtype_code(dt), _st())
        return dx, (dtok.reshape(tshape) if dtok is not None else None), None, None


def mask_rows(token_seq, mask_token, mask_mask, random_mask, random_indices):
    """token_seq [B, T, C]; rows (flattened b*T + t) in `mask_mask` become the mask token, rows in `random_mask` become a copy of
    row random_indices[k] (k-th selected row) of the ORIGINAL sequence, all other rows pass through."""
    kind = mask_mask.to(torch.uint8) + 2 * random_mask.to(torch.uint8)
    src = torch.zeros(kind.numel(), dtype=torch.int64, device=kind.device)
    src[random_mask] = random_indices.to(torch.int64)
    return _MaskRows.apply(token_seq, mask_token, kind.contiguous(), src)


# ------------------------------------------------------------------------------------------------------------------
# CNN branch (PMAM / DASM), channels-last                        reference: cnn/base.py:33-113, passt_cnn.py:52-62
# ------------------------------------------------------------------------------------------------------------------
class _Im2col3x3(torch.autograd.Function):
    """x [B, H, W, C] (channels-last) -> col [B*H*W, pad8(9C)].  `mel_layout`: x is the mel image [B, F, T], read in place as the
    [B, T, F, 1] tensor the reference builds with transpose + unsqueeze (passt_cnn.py:53); it gets no gradient."""

    @staticmethod
    def forward(ctx, x, mel_layout):
        _lib.ensure_device(x)
        x = x.contiguous()
        if mel_layout:
            B, Fq, T = x.shape
            H, W, C = T, Fq, 1
            sb, sh, sw = Fq * T, 1, T
        else:
            B, H, W, C = x.shape
            sb, sh, sw = H * W * C, W * C, C
        Kp = _pad8(9 * C)
        dt = act_dtype()
        col = torch.empty(B * H * W, Kp, dtype=dt, device=x.device)
        with _lib.device_guard(x.device):
            _lib_call("t4s_im2col3x3", _p(x), ops.dtype_code(x.dtype), sb, sh, sw, _p(col), ops.dtype_code(dt), B, H, W, C, Kp, _st())
        ctx.cfg = (B, H, W, C, Kp, mel_layout, x.dtype)
        return col

    @staticmethod
    def backward(ctx, dcol):
        B, H, W, C, Kp, mel_layout, xdt = ctx.cfg
        if mel_layout or not ctx.needs_input_grad[0]:
            return None, None
        dcol = dcol.contiguous()
        din = torch.empty(B, H, W, C, dtype=dcol.dtype, device=dcol.device)
        with _lib.device_guard(dcol.device):
            _lib_call("t4s_col2im3x3", _p(dcol), _p(din), ops.dtype_code(dcol.dtype), B, H, W, C, Kp, _st())
        return (din if din.dtype == xdt else cast(din, xdt)), None


def conv3x3(x, weight, bias, mel_layout=False):
    """Conv2d(C_in, C_out, 3, stride 1, padding 1) on channels-last x [B, H, W, C_in] -> [B, H, W, C_out] (im2col + tcgen05 GEMM)."""
    col = _Im2col3x3.apply(x, mel_layout)
    Cout, Cin = weight.shape[0], weight.shape[1]
    w2 = weight.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin)
    if col.shape[1] != 9 * Cin:
        w2 = torch.nn.functional.pad(w2, (0, col.shape[1] - 9 * Cin))   # zero columns matching the zero-padded im2col rows
    y = linear(col, w2.contiguous(), bias)
    if mel_layout:
        B, Fq, T = x.shape
        return y.reshape(B, T, Fq, Cout)
    return y.reshape(x.shape[0], x.shape[1], x.shape[2], Cout)


class _BatchNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, training):
        _lib.ensure_device(x)
        x = x.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        dev = x.device
        lib = _lib.load()
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=dev)
        rstd = torch.empty(C, dtype=torch.float32, device=dev)
        with _lib.device_guard(dev):
            nbytes = lib.t4s_chan_stats_workspace(rows, C)
            ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
            _lib_call("t4s_batchnorm_fwd", _p(x), _p(y), ops.dtype_code(x.dtype), rows, C, _p(gamma.detach()), _p(beta.detach()), _p(running_mean),
                      _p(running_var), eps, momentum, int(training), _p(mean), _p(rstd), _p(ws), nbytes, _st())
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.training = bool(training)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        C = x.shape[-1]
        rows = x.numel() // C
        dev = x.device
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = convert(dy, torch.empty(dy.shape, dtype=x.dtype, device=dev))
        lib = _lib.load()
        dx = torch.empty_like(x)
        dg = torch.empty(C, dtype=torch.float32, device=dev)
        db = torch.empty(C, dtype=torch.float32, device=dev)
        with _lib.device_guard(dev):
            nbytes = lib.t4s_chan_stats_workspace(rows, C)
            ws = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=dev)
            _lib_call("t4s_batchnorm_bwd", _p(dy), _p(x), ops.dtype_code(x.dtype), rows, C, _p(gamma.detach()), _p(mean), _p(rstd),
                      int(ctx.training), _p(dx), _p(dg), _p(db), _p(ws), nbytes, _st())
        return dx, dg, db, None, None, None, None, None


def batch_norm(x, gamma, beta, running_mean, running_var, eps, momentum, training):
    """BatchNorm2d over the last (channel) dim of a channels-last tensor; updates the running buffers in training mode."""
    return _BatchNorm.apply(x, gamma, beta, running_mean, running_var, float(eps), float(momentum), bool(training))


class _Gate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, lin, p, seed):
        _lib.ensure_device(y)
        y, lin = y.contiguous(), lin.contiguous()
        out = torch.empty_like(y)
        with _lib.device_guard(y.device):
            _lib_call("t4s_gate_fwd", _p(y), _p(lin), _p(out), y.numel(), p, seed, ops.dtype_code(y.dtype), _st())
        ctx.save_for_backward(y, lin)
        ctx.cfg = (p, seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        y, lin = ctx.saved_tensors
        p, seed = ctx.cfg
        dout = dout.contiguous()
        if dout.dtype != y.dtype:
            dout = convert(dout, torch.empty(dout.shape, dtype=y.dtype, device=y.device))
        dy, dlin = torch.empty_like(y), torch.empty_like(y)
        with _lib.device_guard(y.device):
            _lib_call("t4s_gate_bwd", _p(dout), _p(y), _p(lin), _p(dy), _p(dlin), y.numel(), p, seed, ops.dtype_code(y.dtype), _st())
        return dy, dlin, None, None


_dropout_calls = 0


def context_gate(y, lin, dropout_p=0.0):
    """y * sigmoid(lin) (ContextGating, cnn/base.py:19-30) followed by inverted dropout with rate `dropout_p` (0 = none).  The
    keep mask is a counter-based hash of (torch.initial_seed(), call index, element index): reproducible, not torch's stream."""
    global _dropout_calls
    seed = 0
    if dropout_p > 0:
        _dropout_calls += 1
        seed = (torch.initial_seed() * 1000003 + _dropout_calls) & 0xFFFFFFFFFFFFFFFF
    return _Gate.apply(y, lin, float(dropout_p), seed)


class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ph, pw):
        _lib.ensure_device(x)
        x = x.contiguous()
        B, H, W, C = x.shape
        out = torch.empty(B, H // ph, W // pw, C, dtype=x.dtype, device=x.device)
        with _lib.device_guard(x.device):
            _lib_call("t4s_avgpool_fwd", _p(x), _p(out), ops.dtype_code(x.dtype), B, H, W, C, ph, pw, _st())
        ctx.cfg = (B, H, W, C, ph, pw)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, H, W, C, ph, pw = ctx.cfg
        dout = dout.contiguous()
        dx = torch.empty(B, H, W, C, dtype=dout.dtype, device=dout.device)
        with _lib.device_guard(dout.device):
            _lib_call("t4s_avgpool_bwd", _p(dout), _p(dx), ops.dtype_code(dout.dtype), B, H, W, C, ph, pw, _st())
        return dx, None, None


def avg_pool(x, ph, pw):
    """AvgPool2d((ph, pw)) on channels-last x [B, H, W, C]."""
    if ph == 1 and pw == 1:
        return x
    return _AvgPool.apply(x, int(ph), int(pw))


class _ScaleAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, w):
        _lib.ensure_device(a)
        a, b = a.contiguous(), b.contiguous()
        if b.dtype != a.dtype:
            b = convert(b, torch.empty(b.shape, dtype=a.dtype, device=a.device))
        w32 = w.detach().reshape(-1).float().contiguous()
        out = torch.empty_like(a)
        with _lib.device_guard(a.device):
            _lib_call("t4s_scale_add_fwd", _p(a), _p(b), _p(w32), _p(out), a.numel(), ops.dtype_code(a.dtype), _st())
        ctx.save_for_backward(b, w32)
        ctx.wshape = w.shape
        return out

    @staticmethod
    def backward(ctx, dout):
        b, w32 = ctx.saved_tensors
        dev = b.device
        dout = dout.contiguous()
        if dout.dtype != b.dtype:
            dout = convert(dout, torch.empty(dout.shape, dtype=b.dtype, device=dev))
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        dw = torch.empty(1, dtype=torch.float32, device=dev)
        ws = torch.empty(256, dtype=torch.float32, device=dev)
        with _lib.device_guard(dev):
            _lib_call("t4s_scale_add_bwd", _p(dout), _p(b), _p(w32), _p(db), _p(dw), _p(ws), b.numel(), ops.dtype_code(b.dtype), _st())
        return (dout if ctx.needs_input_grad[0] else None), db, (dw.reshape(ctx.wshape) if ctx.needs_input_grad[2] else None)


def scale_add(a, b, w):
    """a + w * b with w a one-element tensor (learnable merge weight, passt_cnn.py:60-61); no host sync."""
    return _ScaleAdd.apply(a, b, w)


class _L2Norm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _lib.ensure_device(x)
        x = x.contiguous()
        C = x.shape[-1]
        rows = x.numel() // C
        y = torch.empty_like(x)
        inv = torch.empty(rows, dtype=torch.float32, device=x.device)
        with _lib.device_guard(x.device):
            _lib_call("t4s_l2norm_fwd", _p(x), _p(y), _p(inv), rows, C, ops.dtype_code(x.dtype), _st())
        ctx.save_for_backward(y, inv)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != y.dtype:
            dy = convert(dy, torch.empty(dy.shape, dtype=y.dtype, device=y.device))
        dx = torch.empty_like(y)
        C = y.shape[-1]
        with _lib.device_guard(y.device):
            _lib_call("t4s_l2norm_bwd", _p(dy), _p(y), _p(inv), _p(dx), y.numel() // C, C, ops.dtype_code(y.dtype), _st())
        return dx


def l2_normalize(x):
    """F.normalize(x, dim=-1)."""
    return _L2Norm.apply(x)


class _ProtoAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s, slope, temperature):
        _lib.ensure_device(s)
        s = s.contiguous().float()
        p = torch.empty_like(s)
        with _lib.device_guard(s.device):
            _lib_call("t4s_proto_act_fwd", _p(s), _p(p), s.numel(), slope, temperature, _st())
        ctx.save_for_backward(s, p)
        ctx.cfg = (slope, temperature)
        return p

    @staticmethod
    def backward(ctx, dp):
        s, p = ctx.saved_tensors
        slope, temperature = ctx.cfg
        dp = dp.contiguous().float()
        ds = torch.empty_like(s)
        with _lib.device_guard(s.device):
            _lib_call("t4s_proto_act_bwd", _p(s), _p(p), _p(dp), _p(ds), s.numel(), slope, temperature, _st())
        return ds, None, None


def prototype_predict(logit, prototypes, temperature=0.1, slope=0.2):
    """PMAM prototype head (recipes/desed/pmam/train.py:82-87): sigmoid((leaky_relu(cos-sim to the prototypes, 0.2) * 2 - 1) / T).
    logit [B, T, C], prototypes [K, C] (already L2-normalised GMM means) -> [B, T, K] fp32."""
    z = l2_normalize(logit)
    sim = linear(z, prototypes, None, out_dtype=torch.float32)
    return _ProtoAct.apply(sim, float(slope), float(temperature))


# ------------------------------------------------------------------------------------------------------------------
# LoRA linear                                                    reference: src/models/lora/layers.py:87-153
# ------------------------------------------------------------------------------------------------------------------
def invalidate_weight_cache(w):
    """Forget the cached bf16 copy of `w` (needed after in-place `.data` edits, which do not bump the version counter)."""
    _wcache.pop(id(w), None)
    shadow = getattr(w, "_t4s_shadow", None)
    if shadow is not None:      # parameter-arena bf16 shadow: refresh it in place
        convert(w.detach(), shadow)
        w._t4s_shadow_version = w._version


class _LoraLinear(torch.autograd.Function):
    """y = act(x (W + s B A)^T + b) + residual.  Forward: the rank-r update is folded into an effective weight (one small GEMM), so
    the token GEMM is the ordinary one with its fused epilogue.  Backward: dx through the effective weight; dA, dB through the
    low-rank factors only (u = dh B, t = x A^T, both [tokens, r]) -- no [N, K] weight-gradient GEMM unless W itself trains."""

    @staticmethod
    def forward(ctx, x, w, b, A, B, scaling, residual, act):
        _lib.ensure_device(x)
        shp = x.shape
        K = shp[-1]
        x2 = x.reshape(-1, K)
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        M, N, r = x2.shape[0], w.shape[0], A.shape[0]
        dt, dev = x.dtype, x.device
        with _lib.device_guard(dev):
            Aq, Bq = to_plain(A, dt), to_plain(B, dt)
            weff = torch.empty(N, K, dtype=dt, device=dev)
            wq = cast_weight(w)
            # W_eff = W + s * B A:  [N, r] x [K, r]^T with A consumed MN-major (A is [r, K])
            mm(Op(Bq, N, r), Op(Aq, K, K, mn_major=True), Out(weff, K), N, K, r, alpha=scaling, residual=Out(wq, wq.stride(0)))
            y = torch.empty(M, N, dtype=dt, device=dev)
            aux = torch.empty(M, N, dtype=dt, device=dev) if act == ops.ACT_GELU else None
            res2 = residual.reshape(M, N) if residual is not None else None
            mm(Op(x2, M, x2.stride(0)), Op(weff, N, K), Out(y, N), M, N, K, bias=b.detach() if b is not None else None,
               aux=Out(aux, N) if aux is not None else None, residual=Out(res2, res2.stride(0)) if res2 is not None else None, act=act)
        ctx.save_for_backward(x2, w, A, B, aux, weff)
        ctx.cfg = (scaling, act, shp, b is not None, residual is not None)
        return y.reshape(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w, A, B, aux, weff = ctx.saved_tensors
        scaling, act, shp, has_bias, has_res = ctx.cfg
        M, K = x2.shape
        N, r = w.shape[0], A.shape[0]
        dt, dev = x2.dtype, x2.device
        dy2 = dy.reshape(M, N)
        if dy2.stride(-1) != 1:
            dy2 = dy2.contiguous()
        ng = ctx.needs_input_grad
        with _lib.device_guard(dev):
            if dy2.dtype != dt:
                dy2 = convert(dy2, torch.empty(M, N, dtype=dt, device=dev))
            if act == ops.ACT_GELU:
                dh = torch.empty(M, N, dtype=dt, device=dev)
                _lib_call("t4s_gelu_bwd", _p(dy2), _p(aux), _p(dh), M * N, ops.dtype_code(dt), _st())
            else:
                dh = dy2
            dx = dw = db = dA = dB = None
            if ng[0]:
                dx = torch.empty(M, K, dtype=dt, device=dev)
                mm(Op(dh, M, dh.stride(0)), Op(weff, K, K, mn_major=True), Out(dx, K), M, K, N)
                dx = dx.reshape(shp)
            if ng[1]:
                dw = weight_grad(dh, x2, N, K)
            if has_bias and ng[2]:
                db = colsum(dh)
            if ng[3] or ng[4]:
                Aq, Bq = to_plain(A, dt), to_plain(B, dt)
                rp = _pad8(r)
                if ng[3]:
                    u = torch.zeros(M, rp, dtype=dt, device=dev) if rp != r else torch.empty(M, rp, dtype=dt, device=dev)
                    mm(Op(dh, M, dh.stride(0)), Op(Bq, r, r, mn_major=True), Out(u, rp), M, r, N, alpha=scaling)   # u = s dh B
                    dA = weight_grad(u[:, :r], x2, r, K)                                                          # dA = u^T x
                if ng[4]:
                    t = torch.zeros(M, rp, dtype=dt, device=dev) if rp != r else torch.empty(M, rp, dtype=dt, device=dev)
                    mm(Op(x2, M, x2.stride(0)), Op(Aq, r, K), Out(t, rp), M, r, K, alpha=scaling)                  # t = s x A^T
                    dB = weight_grad(dh, t[:, :r], N, r)                                                          # dB = dh^T t
        d_res = dy if has_res else None
        return dx, dw, db, dA, dB, None, d_res, None


def to_plain(p, dtype):
    """Small parameter -> contiguous tensor of the activation dtype (no caching: LoRA factors change every step)."""
    shadow = _arena_shadow(p)
    if shadow is not None and shadow.dtype == dtype:
        return shadow
    p = p.detach().contiguous()
    return p if p.dtype == dtype else convert(p, torch.empty(p.shape, dtype=dtype, device=p.device))


def lora_linear(x, w, b, lora_A, lora_B, scaling, residual=None, act=ops.ACT_NONE):
    return _LoraLinear.apply(x, w, b, lora_A, lora_B, float(scaling), residual, act)


# ------------------------------------------------------------------------------------------------------------------
# DASM: general multi-head attention (separate query / key-value sequences, boolean mask, dropout), query pooling
# reference: src/models/detect_any_sound/at_adapter.py:7-50 (nn.TransformerDecoderLayer), detect_any_sound.py:376-388
# ------------------------------------------------------------------------------------------------------------------
_drop_calls = 0


def _next_seed():
    global _drop_calls
    _drop_calls += 1
    return (torch.initial_seed() * 1000003 + 7919 * _drop_calls) & 0xFFFFFFFFFFFFFFFF


class _Dropout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        _lib.ensure_device(x)
        x = x.contiguous()
        out = torch.empty_like(x)
        with _lib.device_guard(x.device):
            _lib_call("t4s_dropout", _p(x), _p(out), x.numel(), p, seed, ops.dtype_code(x.dtype), _st())
        ctx.cfg = (p, seed)
        return out

    @staticmethod
    def backward(ctx, dout):
        p, seed = ctx.cfg
        dout = dout.contiguous()
        dx = torch.empty_like(dout)
        with _lib.device_guard(dout.device):
            _lib_call("t4s_dropout", _p(dout), _p(dx), dout.numel(), p, seed, ops.dtype_code(dout.dtype), _st())
        return dx, None, None


def dropout(x, p, training=True):
    """Inverted dropout with a counter-based mask (reproducible from torch.initial_seed(); not torch's Philox stream)."""
    if not training or p <= 0:
        return x
    return _Dropout.apply(x, float(p), _next_seed())


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        _lib.ensure_device(a)
        a, b = a.contiguous(), b.contiguous()
        C = a.shape[-1]
        out = torch.empty_like(a)
        with _lib.device_guard(a.device):
            _lib_call("t4s_add2", _p(a), C, _p(b), C, _p(out), C, a.numel() // C, C, 1.0, 1.0, ops.dtype_code(a.dtype), _st())
        return out

    @staticmethod
    def backward(ctx, dout):
        return dout, dout


def add(a, b):
    return _Add.apply(a, b)


class _CrossAttention(torch.autograd.Function):
    """softmax(q k^T / sqrt(hd) [masked]) v with q [B, Nq, D] and kv [B, Nk, 2D] (k | v per key): the nn.MultiheadAttention core for
    Nq != Nk.  Scores live in a [B, H, Nq, pad8(Nk)] buffer (Nq = 407 queries: small); GEMMs read heads in place."""

    @staticmethod
    def forward(ctx, q, kv, H, mask, drop_p, seed):
        _lib.ensure_device(q)
        q, kv = q.contiguous(), kv.contiguous()
        B, Nq, D = q.shape
        Nk = kv.shape[1]
        hd = D // H
        Np = _pad8(Nk)
        scale = hd ** -0.5
        dt, dev = q.dtype, q.device
        code = ops.dtype_code(dt)
        with _lib.device_guard(dev):
            P = torch.empty(B, H, Nq, Np, dtype=dt, device=dev)
            mm(Op(q, Nq, D, 0, nb1=H, stride1=hd, nb2=B, stride2=Nq * D), Op(kv, Nk, 2 * D, 0, nb1=H, stride1=hd, nb2=B, stride2=Nk * 2 * D),
               Out(P, Np, 0, Nq * Np, H * Nq * Np), Nq, Nk, hd, nb1=H, nb2=B, alpha=scale)
            if mask is not None:
                _lib_call("t4s_mask_scores", _p(P), _p(mask), B * H * Nq, Nk, Np, Nq, code, _st())
            _lib_call("t4s_softmax_fwd", _p(P), _p(P), B * H * Nq, Nk, Np, Np, code, _st())
            Pd = P
            if drop_p > 0:
                Pd = torch.empty_like(P)
                _lib_call("t4s_dropout", _p(P), _p(Pd), P.numel(), drop_p, seed, code, _st())
            o = torch.empty(B, Nq, D, dtype=dt, device=dev)
            mm(Op(Pd, Nq, Np, 0, nb1=H, stride1=Nq * Np, nb2=B, stride2=H * Nq * Np),
               Op(kv, hd, 2 * D, D, nb1=H, stride1=hd, nb2=B, stride2=Nk * 2 * D, mn_major=True), Out(o, D, 0, hd, Nq * D), Nq, hd, Nk, nb1=H, nb2=B)
        ctx.save_for_backward(q, kv, P, Pd if drop_p > 0 else None)
        ctx.cfg = (H, drop_p, seed)
        return o

    @staticmethod
    def backward(ctx, do):
        q, kv, P, Pd = ctx.saved_tensors
        H, drop_p, seed = ctx.cfg
        B, Nq, D = q.shape
        Nk = kv.shape[1]
        hd = D // H
        Np = P.shape[-1]
        scale = hd ** -0.5
        dt, dev = q.dtype, q.device
        code = ops.dtype_code(dt)
        do = do.contiguous()
        if do.dtype != dt:
            do = convert(do, torch.empty(do.shape, dtype=dt, device=dev))
        if Pd is None:
            Pd = P
        with _lib.device_guard(dev):
            dq = torch.empty_like(q)
            dkv = torch.empty_like(kv)
            dP = torch.empty(B, H, Nq, Np, dtype=dt, device=dev)
            pm = dict(nb1=H, stride1=Nq * Np, nb2=B, stride2=H * Nq * Np)
            kh = dict(nb1=H, stride1=hd, nb2=B, stride2=Nk * 2 * D)
            qh = dict(nb1=H, stride1=hd, nb2=B, stride2=Nq * D)
            # dPd = dO V^T ; dV = Pd^T dO
            mm(Op(do, Nq, D, 0, **qh), Op(kv, Nk, 2 * D, D, **kh), Out(dP, Np, 0, Nq * Np, H * Nq * Np), Nq, Nk, hd, nb1=H, nb2=B)
            mm(Op(Pd, Nk, Np, 0, mn_major=True, **pm), Op(do, hd, D, 0, mn_major=True, **qh), Out(dkv, 2 * D, D, hd, Nk * 2 * D), Nk, hd, Nq,
               nb1=H, nb2=B)
            if drop_p > 0:
                _lib_call("t4s_dropout", _p(dP), _p(dP), dP.numel(), drop_p, seed, code, _st())
            _lib_call("t4s_softmax_bwd", _p(P), _p(dP), B * H * Nq, Nk, Np, Np, code, _st())
            # dQ = scale dS K ; dK = scale dS^T Q
            mm(Op(dP, Nq, Np, 0, **pm), Op(kv, hd, 2 * D, 0, mn_major=True, **kh), Out(dq, D, 0, hd, Nq * D), Nq, hd, Nk, nb1=H, nb2=B, alpha=scale)
            mm(Op(dP, Nk, Np, 0, mn_major=True, **pm), Op(q, hd, D, 0, mn_major=True, **qh), Out(dkv, 2 * D, 0, hd, Nk * 2 * D), Nk, hd, Nq,
               nb1=H, nb2=B, alpha=scale)
        return dq, dkv, None, None, None, None


def multi_head_attention(q_in, kv_in, in_proj_weight, in_proj_bias, out_w, out_b, num_heads, attn_mask=None, dropout_p=0.0, training=False,
                         residual=None):
    """nn.MultiheadAttention(batch_first=True)(q_in, kv_in, kv_in, attn_mask=...)[0] (+ residual fused into the out_proj epilogue).
    attn_mask: boolean [Nq, Nk], True = masked out."""
    D = q_in.shape[-1]
    q = linear(q_in, in_proj_weight[:D], in_proj_bias[:D])
    kv = linear(kv_in, in_proj_weight[D:], in_proj_bias[D:])
    m = attn_mask.to(torch.uint8).contiguous() if attn_mask is not None else None
    p = float(dropout_p) if training else 0.0
    o = _CrossAttention.apply(q, kv, num_heads, m, p, _next_seed() if p > 0 else 0)
    return linear(o, out_w, out_b, residual=residual)


class _QueryPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, at_out, temp, pad_mask):
        _lib.ensure_device(score)
        score = score.contiguous().float()
        at_out = at_out.contiguous().float()
        B, T, K = score.shape
        strong = torch.empty(B, K, T, dtype=torch.float32, device=score.device)
        weak = torch.empty(B, K, dtype=torch.float32, device=score.device)
        pm = pad_mask.to(torch.uint8).contiguous() if pad_mask is not None else None
        with _lib.device_guard(score.device):
            _lib_call("t4s_query_pool_fwd", _p(score), _p(at_out), _p(pm), float(temp), _p(strong), _p(weak), B, T, K, _st())
        ctx.save_for_backward(score, at_out, strong, pm)
        ctx.temp = float(temp)
        return strong, weak

    @staticmethod
    def backward(ctx, dstrong, dweak):
        score, at_out, strong, pm = ctx.saved_tensors
        B, T, K = score.shape
        dscore = torch.empty_like(score)
        dat = torch.empty_like(at_out)
        ds = dstrong.contiguous().float() if dstrong is not None else None
        dw = dweak.contiguous().float() if dweak is not None else None
        with _lib.device_guard(score.device):
            _lib_call("t4s_query_pool_bwd", _p(score), _p(at_out), _p(strong), _p(ds), _p(dw), _p(pm), ctx.temp, _p(dscore), _p(dat), B, T, K, _st())
        return dscore, dat, None, None


def query_pool(score, at_out, temp=0.1, pad_mask=None):
    """score [B, T, K] (frame x query logits), at_out [B, K] -> (strong [B, K, T], weak [B, K]) as DASM's head (:379-388)."""
    return _QueryPool.apply(score, at_out, temp, pad_mask)


class _QueryFrameScore(torch.autograd.Function):
    """score[b, t, q] = <x[b, t, :], emb[b, q, :]>  (einsum 'bqc,bct->bqt' transposed, detect_any_sound.py:378): one batched GEMM."""

    @staticmethod
    def forward(ctx, x, emb):
        _lib.ensure_device(x)
        x, emb = x.contiguous(), emb.contiguous()
        if emb.dtype != x.dtype:
            emb = convert(emb, torch.empty(emb.shape, dtype=x.dtype, device=x.device))
        B, T, C = x.shape
        K = emb.shape[1]
        out = torch.empty(B, T, K, dtype=torch.float32, device=x.device)
        with _lib.device_guard(x.device):
            mm(Op(x, T, C, 0, nb1=B, stride1=T * C), Op(emb, K, C, 0, nb1=B, stride1=K * C), Out(out, K, 0, T * K), T, K, C, nb1=B)
        ctx.save_for_backward(x, emb)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, emb = ctx.saved_tensors
        B, T, C = x.shape
        K = emb.shape[1]
        dt, dev = x.dtype, x.device
        Kp = _pad8(K)
        with _lib.device_guard(dev):
            # operand copy of the gradient in the activation dtype with a 16-byte row pitch (K = 407 is not a multiple of 8)
            d2 = torch.zeros(B, T, Kp, dtype=dt, device=dev)
            d32 = dout.contiguous().float()
            tmp = torch.empty(B * T, K, dtype=dt, device=dev)
            convert(d32, tmp)
            _lib_call("t4s_add2", _p(tmp), K, _p(tmp), K, _p(d2), Kp, B * T, K, 1.0, 0.0, ops.dtype_code(dt), _st())
            dx = torch.empty_like(x)
            demb = torch.empty_like(emb)
            # dx[b] = d[b] emb[b]  (contract over q);  demb[b] = d[b]^T x[b]  (contract over t)
            mm(Op(d2, T, Kp, 0, nb1=B, stride1=T * Kp), Op(emb, C, C, 0, nb1=B, stride1=K * C, mn_major=True), Out(dx, C, 0, T * C), T, C, K, nb1=B)
            mm(Op(d2, K, Kp, 0, nb1=B, stride1=T * Kp, mn_major=True), Op(x, C, C, 0, nb1=B, stride1=T * C, mn_major=True), Out(demb, C, 0, K * C),
               K, C, T, nb1=B)
        return dx, demb


def query_frame_score(x, emb):
    return _QueryFrameScore.apply(x, emb)


# ------------------------------------------------------------------------------------------------------------------
# heads and losses
# ------------------------------------------------------------------------------------------------------------------
class _SedPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, temp, pad_mask):
        _lib.ensure_device(logits)
        logits = logits.contiguous().float()
        B, T, K = logits.shape
        strong = torch.empty(B, K, T, dtype=torch.float32, device=logits.device)
        weak = torch.empty(B, K, dtype=torch.float32, device=logits.device)
        pm = pad_mask.to(torch.uint8).contiguous() if pad_mask is not None else None
        with _lib.device_guard(logits.device):
            _lib_call("t4s_sed_pool_fwd", _p(logits), _p(pm), float(temp), _p(strong), _p(weak), B, T, K, _st())
        ctx.save_for_backward(strong, pm)
        ctx.temp = float(temp)
        return strong, weak

    @staticmethod
    def backward(ctx, dstrong, dweak):
        strong, pm = ctx.saved_tensors
        B, K, T = strong.shape
        dl = torch.empty(B, T, K, dtype=torch.float32, device=strong.device)
        ds = dstrong.contiguous().float() if dstrong is not None else None
        dw = dweak.contiguous().float() if dweak is not None else None
        with _lib.device_guard(strong.device):
            _lib_call("t4s_sed_pool_bwd", _p(strong), _p(ds), _p(dw), _p(pm), ctx.temp, _p(dl), B, T, K, _st())
        return dl, None, None


def sed_pool(logits, temp=1.0, pad_mask=None):
    """sigmoid(logits/temp) -> (strong [B,K,T], weak [B,K]) with padded frames zeroed and linear-softmax pooling."""
    return _SedPool.apply(logits, temp, pad_mask)


class _Sigmoid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _lib.ensure_device(x)
        x = x.contiguous().float()
        y = torch.empty_like(x)
        with _lib.device_guard(x.device):
            _lib_call("t4s_sigmoid_fwd", _p(x), _p(y), x.numel(), _st())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = dy.contiguous().float()
        dx = torch.empty_like(y)
        with _lib.device_guard(y.device):
            _lib_call("t4s_sigmoid_bwd", _p(y), _p(dy), _p(dx), y.numel(), _st())
        return dx


def sigmoid(x):
    return _Sigmoid.apply(x)


class _Bce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, y):
        _lib.ensure_device(p)
        p = p.contiguous().float()
        y = y.contiguous().float()
        ws = torch.empty(512, dtype=torch.float32, device=p.device)
        out = torch.empty(2, dtype=torch.float32, device=p.device)
        with _lib.device_guard(p.device):
            _lib_call("t4s_bce_fwd", _p(p), _p(y), p.numel(), _p(ws), _p(out), _st())
        ctx.save_for_backward(p, y)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        p, y = ctx.saved_tensors
        dp = torch.empty_like(p)
        g = g.reshape(1).contiguous().float()
        with _lib.device_guard(p.device):
            _lib_call("t4s_bce_bwd", _p(p), _p(y), _p(g), p.numel(), _p(dp), _st())
        return dp, None


def bce_loss(p, y):
    """torch.nn.BCELoss() (mean reduction, log clamped at -100)."""
    return _Bce.apply(p, y)


class _Mse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, row_mask):
        _lib.ensure_device(a)
        a = a.contiguous()
        b = b.contiguous()
        if b.dtype != a.dtype:
            b = convert(b, torch.empty(b.shape, dtype=a.dtype, device=a.device))
        C = a.shape[-1]
        rows = a.numel() // C
        m = row_mask.reshape(-1).to(torch.uint8).contiguous() if row_mask is not None else None
        ws = torch.empty(512, dtype=torch.float32, device=a.device)
        out = torch.empty(2, dtype=torch.float32, device=a.device)
        with _lib.device_guard(a.device):
            _lib_call("t4s_mse_fwd", _p(a), _p(b), _p(m), rows, C, ops.dtype_code(a.dtype), _p(ws), _p(out), _st())
        ctx.save_for_backward(a, b, m, out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        a, b, m, out = ctx.saved_tensors
        C = a.shape[-1]
        rows = a.numel() // C
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        g = g.reshape(1).contiguous().float()
        with _lib.device_guard(a.device):
            _lib_call("t4s_mse_bwd", _p(a), _p(b), _p(m), rows, C, ops.dtype_code(a.dtype), _p(g), _p(out), _p(da), _p(db), _st())
        return da, db, None


def mse_loss(a, b, row_mask=None):
    """MSELoss(mean) over rows (last dim = features); with `row_mask` [rows] only the selected rows count
    (== MSELoss()(a[mask], b[mask]), reference recipes/desed/mlm/mlm_passt/train.py:38)."""
    return _Mse.apply(a, b, row_mask)


class _AttnPool(torch.autograd.Function):
    """One learned query attending over keys/values packed as kv [items, keys_all, 2C], skipping the first `skip` keys of
    every item (cls/dist tokens); q [C] fp32, pre-scaled."""

    @staticmethod
    def forward(ctx, kv, q, H, skip):
        _lib.ensure_device(kv)
        kv = kv.contiguous()
        q = q.contiguous().float()
        items, K_all, C2 = kv.shape
        C, K = C2 // 2, K_all - skip
        ctxv = torch.empty(items, C, dtype=kv.dtype, device=kv.device)
        probs = torch.empty(items, H, K, dtype=torch.float32, device=kv.device)
        kp = ctypes.c_void_p(kv.data_ptr() + skip * C2 * kv.element_size())
        with _lib.device_guard(kv.device):
            _lib_call("t4s_attnpool_fwd", kp, _p(q), _p(ctxv), _p(probs), items, K, C, H, K_all * C2, ops.dtype_code(kv.dtype), _st())
        ctx.save_for_backward(kv, q, probs)
        ctx.cfg = (H, skip)
        return ctxv

    @staticmethod
    def backward(ctx, dctx):
        kv, q, probs = ctx.saved_tensors
        H, skip = ctx.cfg
        items, K_all, C2 = kv.shape
        C, K = C2 // 2, K_all - skip
        dctx = dctx.contiguous()
        if dctx.dtype != kv.dtype:
            dctx = convert(dctx, torch.empty(dctx.shape, dtype=kv.dtype, device=kv.device))
        dkv = torch.zeros_like(kv) if skip else torch.empty_like(kv)
        dq_part = torch.empty(items, C, dtype=torch.float32, device=kv.device)
        off = skip * C2 * kv.element_size()
        with _lib.device_guard(kv.device):
            _lib_call("t4s_attnpool_bwd", ctypes.c_void_p(kv.data_ptr() + off), _p(q), _p(probs), _p(dctx),
                      ctypes.c_void_p(dkv.data_ptr() + off), _p(dq_part), items, K, C, H, K_all * C2, ops.dtype_code(kv.dtype), _st())
            dq = colsum(dq_part) if ctx.needs_input_grad[1] else None
        return dkv, dq, None, None


def attn_pool(kv, q, num_heads, skip=0):
    return _AttnPool.apply(kv, q, num_heads, skip)


def mha_pool(x, token, in_proj_weight, in_proj_bias, out_w, out_b, num_heads, skip=0):
    """nn.MultiheadAttention(batch_first) with a single learned query (pooling.py:37-51): x [items, keys, C] -> [items, C].
    `skip` leading keys of every item are ignored (keys/values are still projected for them: 2 of 1190 tokens)."""
    items, K, C = x.shape
    hd = C // num_heads
    # q = (token W_q^T + b_q) / sqrt(hd): batch independent, one 1-row GEMM with the scale folded into alpha/bias
    tok = to_act(token.reshape(1, C))
    qv = _ScaledLinear.apply(tok, in_proj_weight[:C], in_proj_bias[:C], hd ** -0.5)
    kv = linear(x.reshape(items * K, C), in_proj_weight[C:], in_proj_bias[C:]).reshape(items, K, 2 * C)
    ctxv = attn_pool(kv, qv.reshape(C), num_heads, skip)
    return linear(ctxv, out_w, out_b)


class _ScaledLinear(torch.autograd.Function):
    """y = s * (x W^T + b), fp32 output (used for the pooled-attention query)."""

    @staticmethod
    def forward(ctx, x, w, b, s):
        _lib.ensure_device(x)
        M, K = x.shape
        N = w.shape[0]
        wq = cast_weight(w.contiguous())
        bs = torch.empty(N, dtype=torch.float32, device=x.device)
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
        with _lib.device_guard(x.device):
            _lib_call("t4s_add_rowvec", _p(b.detach().contiguous()), N, ctypes.c_void_p(0), _p(bs), 1, N, float(s), ops.F32, _st())
            mm(Op(x, M, K), Op(wq, N, K), Out(y, N), M, N, K, bias=bs, alpha=float(s))
        ctx.save_for_backward(x, w)
        ctx.s = float(s)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        M, K = x.shape
        N = w.shape[0]
        dev = x.device
        with _lib.device_guard(dev):
            dys = torch.empty(M, N, dtype=x.dtype, device=dev)
            dy32 = dy.contiguous().float()
            tmp = torch.empty(M, N, dtype=torch.float32, device=dev)
            _lib_call("t4s_add_rowvec", _p(dy32), N, ctypes.c_void_p(0), _p(tmp), M, N, ctx.s, ops.F32, _st())
            convert(tmp, dys)
            dx = None
            if ctx.needs_input_grad[0]:
                wq = cast_weight(w.contiguous())
                dx = torch.empty(M, K, dtype=x.dtype, device=dev)
                mm(Op(dys, M, N), Op(wq, K, K, mn_major=True), Out(dx, K), M, K, N)
            dw = weight_grad(dys, x, N, K) if ctx.needs_input_grad[1] else None
            db = colsum(tmp) if ctx.needs_input_grad[2] else None
        return dx, dw, db, None
