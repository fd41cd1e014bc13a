"""GPU: tcgen05 GEMM (through the C ABI) vs torch fp32/fp64 matmul.  Covers every operand-major combination, both
input dtypes, ragged M/N/K, batching with broadcast, split-K and every epilogue option."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from transformer4sed_b200 import ops
    return ops


def _rand(shape, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda", dtype=torch.float32).to(dtype)


def _tol(dtype, K):
    # bf16 inputs are exact in the reference too (we upcast the same bf16 values); error is fp32 accumulation order only.
    # tf32 rounds fp32 inputs to 10 mantissa bits: rel 2^-11 per operand -> ~ sqrt(K) * 2^-11 * |a||b| absolute.
    return (2e-3 if dtype == torch.bfloat16 else 4e-3 * (K ** 0.5) / 8)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (1190, 768, 768), (300, 72, 200), (77, 1190, 64), (128, 64, 1190)])
def test_gemm_majors(dtype, a_mn, b_mn, M, N, K):
    ops = _ops()
    pad = 8 if dtype == torch.bfloat16 else 4

    def make(rows, mn, seed):
        if mn:  # stored [K][rows_padded]
            ld = (rows + pad - 1) // pad * pad
            t = _rand((K, ld), dtype, seed)
            return t, t[:, :rows].t(), ld
        ld = (K + pad - 1) // pad * pad
        t = _rand((rows, ld), dtype, seed)
        return t, t[:, :K], ld

    ta, a_log, lda = make(M, a_mn, 1)
    tb, b_log, ldb = make(N, b_mn, 2)
    c = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    ops.gemm(ops.Op(ta, M, lda, mn_major=a_mn), ops.Op(tb, N, ldb, mn_major=b_mn), ops.Out(c, N), M, N, K)
    ref = a_log.double() @ b_log.double().t()
    err = (c.double() - ref).abs().max().item()
    assert torch.isfinite(c).all()
    assert err < _tol(dtype, K) * max(1.0, ref.abs().max().item() / 8), (err, ref.abs().max().item())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_epilogue_bias_gelu_residual_aux(dtype):
    ops = _ops()
    M, N, K = 515, 392, 320
    x, w = _rand((M, K), dtype, 3), _rand((N, K), dtype, 4) * 0.1
    bias, res = _rand((N,), torch.float32, 5), _rand((M, N), dtype, 6)
    for out_dtype in (torch.float32, torch.bfloat16):
        y, aux = ops.linear_nt(x, w, bias=bias, out_dtype=out_dtype, act=ops.ACT_GELU, residual=res, aux_dtype=out_dtype, alpha=0.5)
        pre = 0.5 * (x.double() @ w.double().t()) + bias.double()
        ref = torch.nn.functional.gelu(pre) + res.double()
        tol = 3e-3 if out_dtype == torch.float32 else 4e-2
        assert (aux.double() - pre).abs().max().item() < tol
        assert (y.double() - ref).abs().max().item() < tol


def test_gemm_batched_attention_layout():
    """q k^T and p v straight out of a fused qkv buffer [B, N, 3, H, hd] (passt.py:333-341 layout)."""
    ops = _ops()
    B, Nt, H, hd = 2, 1190, 3, 64
    D = H * hd
    for dtype in (torch.bfloat16, torch.float32):
        qkv = _rand((B, Nt, 3 * D), dtype, 7)
        Np = 1192
        S = torch.zeros(B, H, Nt, Np, device="cuda", dtype=torch.float32)
        ops.gemm(ops.Op(qkv, Nt, 3 * D, 0, nb1=H, stride1=hd, nb2=B, stride2=Nt * 3 * D),
                 ops.Op(qkv, Nt, 3 * D, D, nb1=H, stride1=hd, nb2=B, stride2=Nt * 3 * D),
                 ops.Out(S, Np, 0, stride1=Nt * Np, stride2=H * Nt * Np), Nt, Nt, hd, nb1=H, nb2=B, alpha=0.125)
        q, k, v = qkv.double().view(B, Nt, 3, H, hd).permute(2, 0, 3, 1, 4)
        ref = (q @ k.transpose(-1, -2)) * 0.125
        assert (S[..., :Nt].double() - ref).abs().max().item() < (2e-3 if dtype == torch.bfloat16 else 2e-2)
        assert (S[..., Nt:] == 0).all()
        # o = p v with v consumed in place as an MN-major B operand: B[n=hd, k=key] stored [key][hd]
        P = torch.softmax(S[..., :Nt], dim=-1).to(dtype)
        Pp = torch.zeros(B, H, Nt, Np, device="cuda", dtype=dtype)
        Pp[..., :Nt] = P
        O = torch.zeros(B, Nt, D, device="cuda", dtype=torch.float32)
        ops.gemm(ops.Op(Pp, Nt, Np, 0, nb1=H, stride1=Nt * Np, nb2=B, stride2=H * Nt * Np),
                 ops.Op(qkv, hd, 3 * D, 2 * D, nb1=H, stride1=hd, nb2=B, stride2=Nt * 3 * D, mn_major=True),
                 ops.Out(O, D, 0, stride1=hd, stride2=Nt * D), Nt, hd, Nt, nb1=H, nb2=B)
        ref_o = (P.double() @ v).permute(0, 2, 1, 3).reshape(B, Nt, D)
        assert (O.double() - ref_o).abs().max().item() < 3e-3


def test_gemm_split_k_weight_gradient():
    """dW[N_out, K_in] = dy^T x with both operands MN-major (no transposes) and the long contraction split 8 ways."""
    ops = _ops()
    Mtok, Nout, Kin = 4760, 384, 256
    for dtype in (torch.bfloat16, torch.float32):
        dy, x = _rand((Mtok, Nout), dtype, 8), _rand((Mtok, Kin), dtype, 9)
        S = 8
        ws = torch.empty(S, Nout, Kin, device="cuda", dtype=torch.float32)
        ops.gemm(ops.Op(dy, Nout, Nout, mn_major=True), ops.Op(x, Kin, Kin, mn_major=True), ops.Out(ws, Kin), Nout, Kin, Mtok,
                 split_k=S, c_split_stride=Nout * Kin)
        dw = torch.ones(Nout, Kin, device="cuda")
        ops.reduce_splits(ws, S, Nout * Kin, dw, accumulate=True)
        ref = dy.double().t() @ x.double() + 1.0
        assert (dw.double() - ref).abs().max().item() < (5e-3 if dtype == torch.bfloat16 else 0.3)


def test_gemm_anti_diagonal_band():
    """`band` = (lo, hi): A[m, k] == 0 unless lo <= m + k < hi (the un-shifted position-score gradient of the rel-pos attention
    backward, functional._FlashRelPosAttention.backward): the clipped GEMM equals the full one, K-major and MN-major A, batched,
    bf16 (single-CTA and pair tiles) and with a split contraction."""
    ops = _ops()
    T, hd, H, B = 333, 64, 2, 3
    L, Lp = 2 * T - 1, 672
    g = torch.Generator(device="cuda").manual_seed(5)
    dbd = torch.zeros(B, H, T, Lp, device="cuda", dtype=torch.bfloat16)
    i, k = torch.arange(T, device="cuda")[:, None], torch.arange(Lp, device="cuda")[None, :]
    inband = ((i + k >= T - 1) & (i + k < 2 * T - 1)).expand(B, H, T, Lp)
    dbd[inband] = torch.randn(int(inband.sum()), generator=g, device="cuda").to(torch.bfloat16)
    p = _rand((L, H * hd), torch.bfloat16, 6)
    qv = _rand((B, T, H * hd), torch.bfloat16, 7)
    bm = dict(nb1=H, stride1=T * Lp, nb2=B, stride2=H * T * Lp)
    outs = []
    for band in (None, (T - 1, 2 * T - 1)):
        dqv = torch.zeros(B, T, H * hd, device="cuda", dtype=torch.bfloat16)
        ops.gemm(ops.Op(dbd, T, Lp, 0, **bm), ops.Op(p, hd, H * hd, 0, nb1=H, stride1=hd, mn_major=True), ops.Out(dqv, H * hd, 0, hd, T * H * hd),
                 T, hd, L, nb1=H, nb2=B, band=band)
        ws = torch.zeros(B, L, H * hd, device="cuda", dtype=torch.float32)
        ops.gemm(ops.Op(dbd, L, Lp, 0, mn_major=True, **bm), ops.Op(qv, hd, H * hd, 0, nb1=H, stride1=hd, nb2=B, stride2=T * H * hd, mn_major=True),
                 ops.Out(ws, H * hd, 0, hd, L * H * hd), L, hd, T, nb1=H, nb2=B, band=band)
        outs.append((dqv, ws))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref = torch.einsum("bhik,khd->bihd", dbd[..., :L].double(), p.double().view(L, H, hd)).reshape(B, T, H * hd)
    assert (outs[1][0].double() - ref).abs().max().item() < 0.02 * ref.abs().max().item()
    # wide output (pair tiles) and a split contraction over a band that leaves some splits empty
    M, N, K = 512, 256, 1024
    A = torch.zeros(M, K, device="cuda", dtype=torch.bfloat16)
    mi, ki = torch.arange(M, device="cuda")[:, None], torch.arange(K, device="cuda")[None, :]
    msk = (mi + ki >= 700) & (mi + ki < 900)
    A[msk] = torch.randn(int(msk.sum()), generator=g, device="cuda").to(torch.bfloat16)
    Bm = _rand((N, K), torch.bfloat16, 8)
    ref = A.double() @ Bm.double().t()
    C = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(ops.Op(A, M, K), ops.Op(Bm, N, K), ops.Out(C, N), M, N, K, band=(700, 900))
    assert (C.double() - ref).abs().max().item() < 2e-3
    ws = torch.empty(4, M, N, device="cuda", dtype=torch.float32)
    ops.gemm(ops.Op(A, M, K), ops.Op(Bm, N, K), ops.Out(ws, N), M, N, K, split_k=4, c_split_stride=M * N, band=(700, 900))
    assert (ws.sum(0).double() - ref).abs().max().item() < 2e-3


def test_gemm_rejects_bad_arguments():
    from transformer4sed_b200 import _lib
    ops = _ops()
    x = torch.zeros(128, 65, device="cuda", dtype=torch.bfloat16)  # pitch not a multiple of 16 bytes
    w = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
    c = torch.zeros(128, 64, device="cuda")
    with pytest.raises(_lib.T4sError):
        ops.gemm(ops.Op(x, 128, 65), ops.Op(w, 64, 64), ops.Out(c, 64), 128, 64, 64)
    with pytest.raises(_lib.T4sError):
        ops.gemm(ops.Op(w.float(), 64, 64), ops.Op(w, 64, 64), ops.Out(c, 64), 64, 64, 64)
