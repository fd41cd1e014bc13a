"""Diagnostic for csrc/attn.cu on the GPU box: fused attention vs an fp32 torch reference, per output, plus timings."""
import sys
import time

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import functional as F  # noqa: E402


def ref_attn(qkv, H):
    B, N, D3 = qkv.shape
    D = D3 // 3
    hd = D // H
    q, k, v = qkv.float().reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    p = s.softmax(-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B, N, D), s


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def run(B, N, H, scale_in=1.0, seed=0, timing=False):
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv32 = torch.randn(B, N, 3 * D, generator=g, device="cuda") * scale_in
    qkv = qkv32.to(torch.bfloat16).requires_grad_(True)
    w = torch.randn(B, N, D, generator=g, device="cuda").to(torch.bfloat16)
    F.set_precision("bf16")
    o = F.attention(qkv, H)
    torch.cuda.synchronize()
    (o.float() * w.float()).sum().backward()
    torch.cuda.synchronize()
    dqkv = qkv.grad.clone()
    qr = qkv.detach().float().requires_grad_(True)
    o_ref, s = ref_attn(qr, H)
    (o_ref * w.float()).sum().backward()
    dref = qr.grad
    res = {"o": rel(o, o_ref), "dq": rel(dqkv[..., :D], dref[..., :D]), "dk": rel(dqkv[..., D:2 * D], dref[..., D:2 * D]),
           "dv": rel(dqkv[..., 2 * D:], dref[..., 2 * D:])}
    # unfused path of this repo for comparison
    F.set_fused_attention(False)
    q2 = qkv.detach().clone().requires_grad_(True)
    o2 = F.attention(q2, H)
    (o2.float() * w.float()).sum().backward()
    F.set_fused_attention(True)
    res["o_unfused"] = rel(o2, o_ref)
    res["dq_unfused"] = rel(q2.grad[..., :D], dref[..., :D])
    print(f"B={B} N={N} H={H} scale={scale_in}: " + " ".join(f"{k}={v:.2e}" for k, v in res.items()), flush=True)
    if timing:
        for fused in (True, False):
            F.set_fused_attention(fused)
            q3 = qkv.detach().clone().requires_grad_(True)
            for it in range(3):
                o3 = F.attention(q3, H)
                o3.backward(w)
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            for it in range(5):
                o3 = F.attention(q3, H)
            e[1].record()
            for it in range(5):
                o3.backward(w, retain_graph=True)
            e[2].record()
            torch.cuda.synchronize()
            fl = 4.0 * B * H * N * N * 64
            tf, tb = e[0].elapsed_time(e[1]) / 5, e[1].elapsed_time(e[2]) / 5
            print(f"   fused={fused}: fwd {tf:.3f} ms ({fl / tf / 1e9:.0f} TFLOP/s)  bwd {tb:.3f} ms ({2.5 * fl / tb / 1e9:.0f} TFLOP/s algorithmic)",
                  flush=True)
        F.set_fused_attention(True)
    return res


if __name__ == "__main__":
    torch.manual_seed(0)
    run(1, 128, 1)
    run(1, 256, 1)
    run(2, 200, 2)
    run(2, 1190, 12)
    run(2, 1000, 12, scale_in=3.0)
    run(1, 37, 3)
    run(64, 1190, 12, timing=True)
