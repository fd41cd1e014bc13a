"""GPU: every autograd op of the hot path (forward AND backward kernels, through the C ABI) against a plain PyTorch
float64 reference of the same op on the same inputs.  Run in the strict tf32x3 mode with tight tolerances and in the
bf16 performance mode with bf16-sized tolerances."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["tf32x3", "tf32", "bf16"])
def mode(request):
    from transformer4sed_b200 import functional as F
    F.set_precision(request.param)
    yield request.param
    F.set_precision("bf16")


def _F():
    from transformer4sed_b200 import functional as F
    return F


def tol(mode, strict=2e-5, tf32=4e-3, bf16=3e-2):
    return {"tf32x3": strict, "tf32": tf32, "bf16": bf16}[mode]


def rnd(*shape, seed=0, scale=1.0, grad=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randn(*shape, generator=g, device="cuda") * scale
    return t.requires_grad_(grad)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def check(name, ours, ref, t):
    e = rel_err(ours, ref)
    assert e < t, f"{name}: rel err {e:.3e} >= {t:.1e}"


def run_pair(fn_ours, fn_ref, inputs, mode, t, grad_scale=1.0):
    """inputs: list of fp32 leaf tensors (requires_grad as wanted). Compares outputs and all input grads."""
    F = _F()
    ins_o = [i.detach().clone().requires_grad_(i.requires_grad) for i in inputs]
    ins_r = [i.detach().double().requires_grad_(i.requires_grad) for i in inputs]
    out_o = fn_ours(*ins_o)
    out_r = fn_ref(*ins_r)
    outs_o = out_o if isinstance(out_o, tuple) else (out_o,)
    outs_r = out_r if isinstance(out_r, tuple) else (out_r,)
    loss_o, loss_r = 0, 0
    for k, (a, b) in enumerate(zip(outs_o, outs_r)):
        check(f"out{k}", a, b, t)
        g = torch.Generator(device="cuda").manual_seed(100 + k)
        w = torch.randn(b.shape, generator=g, device="cuda") * grad_scale
        loss_o = loss_o + (F.cast(a, torch.float32) * w).sum() if a.dtype != torch.float32 else loss_o + (a * w).sum()
        loss_r = loss_r + (b * w.double()).sum()
    loss_o.backward()
    loss_r.backward()
    for k, (a, b) in enumerate(zip(ins_o, ins_r)):
        if b.requires_grad:
            assert a.grad is not None, f"missing grad for input {k}"
            check(f"grad{k}", a.grad, b.grad, t * 2)


def test_linear_gelu_residual(mode):
    F = _F()
    M, K, N = 300, 192, 264
    x, w, b, r = rnd(3, 100, K, seed=1, grad=True), rnd(N, K, seed=2, scale=K ** -0.5, grad=True), rnd(N, seed=3, grad=True), rnd(3, 100, N, seed=4, grad=True)
    run_pair(lambda x, w, b, r: F.linear(F.to_act(x), w, b, residual=F.to_act(r), act=F.ops.ACT_GELU),
             lambda x, w, b, r: torch.nn.functional.gelu(torch.nn.functional.linear(x, w, b)) + r, [x, w, b, r], mode, tol(mode))


@pytest.mark.parametrize("M1,M2,K,Hd", [(3, 100, 192, 264), (2, 333, 768, 3072)])
def test_mlp_fused_gelu_backward(mode, M1, M2, K, Hd):
    """timm Mlp + residual as one node: gelu' is applied inside the fc2 dgrad GEMM epilogue (T4S_ACT_GELU_GRAD)."""
    F = _F()
    x, r = rnd(M1, M2, K, seed=21, grad=True), rnd(M1, M2, K, seed=22, grad=True)
    w1, b1 = rnd(Hd, K, seed=23, scale=K ** -0.5, grad=True), rnd(Hd, seed=24, scale=0.2, grad=True)
    w2, b2 = rnd(K, Hd, seed=25, scale=Hd ** -0.5, grad=True), rnd(K, seed=26, scale=0.2, grad=True)
    tf = torch.nn.functional
    run_pair(lambda x, w1, b1, w2, b2, r: F.mlp(F.to_act(x), w1, b1, w2, b2, residual=F.to_act(r)),
             lambda x, w1, b1, w2, b2, r: tf.linear(tf.gelu(tf.linear(x, w1, b1)), w2, b2) + r, [x, w1, b1, w2, b2, r], mode,
             tol(mode, strict=5e-5))  # two chained K <= 3072 contractions; the contract itself is 1e-3


@pytest.mark.parametrize("rows,cols", [(5000, 768), (4097, 3072), (777, 10), (64, 2304)])
def test_colsum_strips(rows, cols):
    """bias-gradient column sums (two-stage, deterministic) for bf16 and fp32 inputs, vector and scalar paths."""
    F = _F()
    for dt, t in ((torch.float32, 1e-5), (torch.bfloat16, 1e-5)):
        x = rnd(rows, cols, seed=31).to(dt)
        out = F.colsum(x)
        ref = x.double().sum(0)
        check(f"colsum {dt}", out, ref, t)
        assert torch.equal(out, F.colsum(x))


def test_layer_norm_wide_rows_bf16_fast_path():
    """C = 768, many rows: exercises the 16-byte bf16 LayerNorm backward and the partial-sum reduction."""
    F = _F()
    F.set_precision("bf16")
    x, g, b = rnd(4, 1190, 768, seed=41, grad=True), (1 + 0.1 * rnd(768, seed=42)).requires_grad_(True), rnd(768, seed=43, scale=0.1, grad=True)
    run_pair(lambda x, g, b: F.layer_norm(F.to_act(x), g, b, 1e-6, skip=2),
             lambda x, g, b: torch.nn.functional.layer_norm(x[:, 2:], (768,), g, b, 1e-6), [x, g, b], "bf16", 2e-2)


def test_linear_narrow_head_fp32_out(mode):
    F = _F()
    x, w, b = rnd(2, 500, 192, seed=5, grad=True), rnd(10, 192, seed=6, scale=0.1, grad=True), rnd(10, seed=7, grad=True)
    run_pair(lambda x, w, b: F.linear(F.to_act(x), w, b, out_dtype=torch.float32), lambda x, w, b: torch.nn.functional.linear(x, w, b), [x, w, b],
             mode, tol(mode))


@pytest.mark.parametrize("skip,scale", [(0, 1.0), (2, 1.0), (0, 13.856)])
def test_layer_norm(mode, skip, scale):
    F = _F()
    x, g, b = rnd(3, 50, 192, seed=8, grad=True), (1 + 0.1 * rnd(192, seed=9)).requires_grad_(True), rnd(192, seed=10, scale=0.1, grad=True)
    run_pair(lambda x, g, b: F.layer_norm(F.to_act(x), g, b, 1e-5, in_scale=scale, skip=skip),
             lambda x, g, b: torch.nn.functional.layer_norm(x[:, skip:] * scale, (192,), g, b, 1e-5), [x, g, b], mode,
             tol(mode, strict=1e-5, tf32=1e-5, bf16=2e-2))


@pytest.mark.parametrize("B,N,H,hd", [(2, 1190, 3, 64), (3, 77, 4, 16)])
def test_attention(mode, B, N, H, hd):
    F = _F()
    D = H * hd
    qkv = rnd(B, N, 3 * D, seed=11, scale=0.7, grad=True)

    def ref(qkv):
        q, k, v = qkv.view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        a = ((q @ k.transpose(-1, -2)) * hd ** -0.5).softmax(-1)
        return (a @ v).transpose(1, 2).reshape(B, N, D)

    run_pair(lambda qkv: F.attention(F.to_act(qkv), H), ref, [qkv], mode, tol(mode))


@pytest.mark.parametrize("B,T,H,hd", [(2, 1000, 2, 64), (3, 50, 4, 16)])
def test_relpos_attention(mode, B, T, H, hd):
    F = _F()
    D = H * hd
    qkv = rnd(B, T, 3 * D, seed=12, scale=0.6, grad=True)
    p = rnd(2 * T - 1, D, seed=13, scale=0.5, grad=True)
    u, v = rnd(H, hd, seed=14, scale=0.3, grad=True), rnd(H, hd, seed=15, scale=0.3, grad=True)

    def ref(qkv, p, u, v):
        q, k, val = qkv.view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)  # [B,H,T,hd]
        pp = p.view(2 * T - 1, H, hd).permute(1, 2, 0)               # [H,hd,2T-1]
        ac = (q + u[None, :, None, :]) @ k.transpose(-1, -2)
        bd = (q + v[None, :, None, :]) @ pp
        idx = (T - 1 - torch.arange(T, device="cuda").unsqueeze(1)) + torch.arange(T, device="cuda").unsqueeze(0)
        bd = bd.gather(-1, idx.expand(B, H, T, T))
        a = ((ac + bd) * hd ** -0.5).softmax(-1)
        return (a @ val).transpose(1, 2).reshape(B, T, D)

    run_pair(lambda qkv, p, u, v: F.relpos_attention(F.to_act(qkv), F.to_act(p), u, v, H), ref, [qkv, p, u, v], mode, tol(mode))


def test_patch_embed(mode):
    F = _F()
    B, D, Fd, Tt = 2, 64, 12, 99
    mel = rnd(B, 128, 1000, seed=16)
    w, b = rnd(D, 1, 16, 16, seed=17, scale=1 / 16, grad=True), rnd(D, seed=18, scale=0.1, grad=True)
    tp, fp = rnd(D, Tt, seed=19, scale=0.3, grad=True), rnd(D, Fd, seed=20, scale=0.3, grad=True)
    cls, dist, npos = rnd(D, seed=21, grad=True), rnd(D, seed=22, grad=True), rnd(2, D, seed=23, grad=True)

    def ref(w, b, tp, fp, cls, dist, npos):
        x = torch.nn.functional.conv2d(mel.double().unsqueeze(1), w, b, stride=10)[..., :Tt]
        x = x + tp.view(1, D, 1, Tt) + fp.view(1, D, Fd, 1)
        x = x.flatten(2).transpose(1, 2)
        c = (cls + npos[0]).expand(B, 1, D)
        d = (dist + npos[1]).expand(B, 1, D)
        return torch.cat((c, d, x), dim=1)

    run_pair(lambda w, b, tp, fp, cls, dist, npos: F.patch_embed(mel, w, b, tp, fp, cls, dist, npos, stride=10), ref,
             [w, b, tp, fp, cls, dist, npos], mode, tol(mode))


def test_patch_embed_short_window_offset(mode):
    """512-frame sliding-window crop: 50 time patches, positional table read from an offset (passt.py:506-513)."""
    F = _F()
    B, D, Fd, Tt = 2, 32, 12, 99
    mel = rnd(B, 128, 512, seed=24)
    w, b = rnd(D, 1, 16, 16, seed=25, scale=1 / 16), rnd(D, seed=26, scale=0.1)
    tp, fp = rnd(D, Tt, seed=27), rnd(D, Fd, seed=28)
    cls, dist, npos = rnd(D, seed=29), rnd(D, seed=30), rnd(2, D, seed=31)
    out = F.patch_embed(mel, w, b, tp, fp, cls, dist, npos, stride=10, t_offset=7)
    x = torch.nn.functional.conv2d(mel.double().unsqueeze(1), w.double(), b.double(), stride=10)
    assert x.shape[-1] == 50
    x = x + tp.double()[:, 7:57].view(1, D, 1, 50) + fp.double().view(1, D, Fd, 1)
    check("tokens", out[:, 2:], x.flatten(2).transpose(1, 2), tol(mode))


def test_fpool_and_interp(mode):
    F = _F()
    y = rnd(2, 12 * 99, 48, seed=32, grad=True)
    run_pair(lambda y: F.pad_interpolate(F.fpool_mean(F.to_act(y), 12, 99), 10),
             lambda y: torch.nn.functional.interpolate(
                 torch.cat((y.view(2, 12, 99, 48).mean(1), y.view(2, 12, 99, 48).mean(1)[:, -1:]), 1).transpose(1, 2), scale_factor=10,
                 mode="linear").transpose(1, 2), [y], mode, tol(mode, strict=1e-5, tf32=1e-5, bf16=2e-2))
    x = rnd(2, 100, 48, seed=33, grad=True)
    run_pair(lambda x: F.pad_interpolate(F.to_act(x), 10, pad=False),
             lambda x: torch.nn.functional.interpolate(x.transpose(1, 2), scale_factor=10, mode="linear").transpose(1, 2), [x], mode,
             tol(mode, strict=1e-5, tf32=1e-5, bf16=2e-2))


def test_sed_pool_and_losses():
    F = _F()
    B, T, K = 3, 1000, 10
    logits = rnd(B, T, K, seed=34, scale=2.0, grad=True)
    pad = torch.zeros(B, T, dtype=torch.bool, device="cuda")
    pad[1, 800:] = True
    y = (rnd(B, K, T, seed=35) > 0.8).float()
    yw = (y.sum(-1) > 0).float()

    def ours(logits):
        s, w = F.sed_pool(logits, 0.5, pad)
        return s, w, F.bce_loss(s, y) + 0.5 * F.bce_loss(w, yw)

    def ref(logits):
        p = torch.sigmoid(logits / 0.5).masked_fill(pad.unsqueeze(-1), 0.0)
        w = torch.clamp((p * p).sum(1) / p.sum(1), 1e-7, 1.0)
        s = p.transpose(1, 2)
        bce = torch.nn.BCELoss()
        return s, w, bce(s, y.double()) + 0.5 * bce(w, yw.double())

    run_pair(ours, ref, [logits], "tf32x3", 2e-5)


def test_mse_masked_and_plain(mode):
    F = _F()
    a, b = rnd(2, 1000, 96, seed=36, grad=True), rnd(2, 1000, 96, seed=37, grad=True)
    m = rnd(2, 1000, seed=38) > 0.3
    run_pair(lambda a, b: F.mse_loss(F.to_act(a), F.to_act(b), m), lambda a, b: torch.nn.functional.mse_loss(a[m], b[m]), [a, b], mode,
             tol(mode, strict=1e-5, tf32=1e-5, bf16=2e-2))
    run_pair(lambda a, b: F.mse_loss(F.to_act(a), F.to_act(b)), lambda a, b: torch.nn.functional.mse_loss(a, b), [a, b], mode,
             tol(mode, strict=1e-5, tf32=1e-5, bf16=2e-2))


@pytest.mark.parametrize("items,K,C,H,skip", [(2, 1190, 192, 12, 2), (99, 12, 96, 6, 0)])
def test_mha_pool_matches_nn_multiheadattention(mode, items, K, C, H, skip):
    F = _F()
    x = rnd(items, K, C, seed=39, grad=True)
    mha = torch.nn.MultiheadAttention(C, H, batch_first=True).cuda().double()
    tok = rnd(1, 1, C, seed=40, scale=0.5, grad=True)
    params = [p.detach().float().requires_grad_(True) for p in (mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias)]

    def ours(x, tok, w, b, ow, ob):
        return F.mha_pool(F.to_act(x), tok, w, b, ow, ob, H, skip=skip)

    def ref(x, tok, w, b, ow, ob):
        out, _ = torch.nn.functional.multi_head_attention_forward(
            tok.repeat(items, 1, 1).transpose(0, 1), x[:, skip:].transpose(0, 1), x[:, skip:].transpose(0, 1), C, H, w, b, None, None, False, 0.0,
            ow, ob, training=False, need_weights=False)
        return out.transpose(0, 1).squeeze(1)

    run_pair(ours, ref, [x, tok] + params, mode, tol(mode))
