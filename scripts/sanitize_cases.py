"""Small-shape launches of the mbarrier / TMA / TMEM kernels for compute-sanitizer (scripts/sanitize.sh): tcgen05 GEMM (single-CTA and
CTA-pair tiles, split-K, staged epilogues), fused attention forward / backward (plain and Transformer-XL), the bulk-copy staged LayerNorm kernels and
the post-processing kernels.  Every case is also checked against torch so a sanitizer-clean run is a correct run."""
import sys

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import functional as F, ops  # noqa: E402

torch.manual_seed(0)
F.set_precision("bf16")
dev = "cuda"


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


# ---- GEMMs
for (M, N, K) in ((300, 256, 192), (1000, 768, 768), (2500, 3072, 768), (520, 768, 3072)):
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = torch.randn(N, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    y = F.linear(x, w.float(), b)
    assert rel(y.float(), x.float() @ w.float().t() + b) < 2e-2, (M, N, K)
    xw = x.clone().requires_grad_(True)
    wp = torch.nn.Parameter(w.float())
    F.linear(xw, wp, b).float().sum().backward()          # dgrad + split-K wgrad
    assert rel(wp.grad, torch.ones(M, N, device=dev).t() @ x.float()) < 2e-2
print("gemm cases ok", flush=True)

# ---- fused attention, plain and rel-pos
for (B, N, H) in ((2, 200, 2), (1, 37, 3), (1, 300, 1)):
    D = 64 * H
    qkv = torch.randn(B, N, 3 * D, device=dev).to(torch.bfloat16).requires_grad_(True)
    wgt = torch.randn(B, N, D, device=dev).to(torch.bfloat16)
    o = F.attention(qkv, H)
    o.backward(wgt)
    q, k, v = qkv.detach().float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-1, -2)) / 8.0).softmax(-1) @ v
    assert rel(o.float(), ref.permute(0, 2, 1, 3).reshape(B, N, D)) < 2e-2
    assert torch.isfinite(qkv.grad.float()).all()
    T = N
    qkv2 = (torch.randn(B, T, 3 * D, device=dev) * 0.6).to(torch.bfloat16).requires_grad_(True)
    p = (torch.randn(2 * T - 1, D, device=dev) * 0.5).to(torch.bfloat16).requires_grad_(True)
    u = (torch.randn(H, 64, device=dev) * 0.3).requires_grad_(True)
    vb = (torch.randn(H, 64, device=dev) * 0.3).requires_grad_(True)
    o2 = F.relpos_attention(qkv2, p, u, vb, H)
    o2.backward(wgt)
    assert torch.isfinite(o2.float()).all() and torch.isfinite(qkv2.grad.float()).all() and torch.isfinite(p.grad.float()).all()
print("attention cases ok", flush=True)

# ---- LayerNorm through the bulk-copy staged kernels (ragged row counts, a sliced [B, skip:, C] view, with and without the residual gradient)
for (B, N, C, skip) in ((3, 37, 768, 0), (2, 101, 384, 2), (1, 5, 256, 0)):
    xs = torch.randn(B, N + skip, C, device=dev).to(torch.bfloat16).requires_grad_(True)
    gam, bet = torch.nn.Parameter(torch.rand(C, device=dev) + 0.5), torch.nn.Parameter(torch.randn(C, device=dev) * 0.1)
    y = F.layer_norm(xs, gam, bet, 1e-6, skip=skip)
    y2, res = F.layer_norm_res(y, gam, bet, 1e-6)
    ref = torch.nn.functional.layer_norm(xs.detach().float()[:, skip:], (C,), gam.detach(), bet.detach(), 1e-6)
    assert rel(y.float(), ref) < 2e-2, (B, N, C, skip)
    (y2.float().sum() + res.float().sum()).backward()
    assert torch.isfinite(xs.grad.float()).all() and torch.isfinite(gam.grad).all()
print("layernorm cases ok", flush=True)

# ---- post-processing / scaler / losses
from transformer4sed_b200.src_codec import decoder as D_  # noqa: E402
from transformer4sed_b200.src_preprocess.scaler import TorchScaler  # noqa: E402
from transformer4sed_b200 import training as TR  # noqa: E402
x = torch.rand(3, 10, 156, device=dev)
ev = D_.decode_events(D_.filter_scores(x, [3, 8, 5, 4, 7, 9, 11, 2, 1, 6])[1], torch.rand(3, 10, device=dev), [0.3, 0.6])
assert ev.shape[1] == 5
TorchScaler("instance", "standard")(torch.randn(4, 64, 100, device=dev))
s = [torch.rand(6, 10, 50, device=dev).requires_grad_(), torch.rand(6, 10, device=dev).requires_grad_(), torch.rand(6, 10, device=dev).requires_grad_()]
tot, _ = TR.sed_losses(*s, torch.rand(6, 10, 50, device=dev), torch.rand(6, 10, device=dev), (torch.rand(6, 10, 50, device=dev) > 0.5).float(),
                       (torch.rand(6, 10, device=dev) > 0.5).float(), (0, 2), (2, 4))
tot.backward()
torch.cuda.synchronize()
print("post cases ok", flush=True)
