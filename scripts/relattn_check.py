"""Diagnostic for csrc/attn_rel.cu on the GPU box: fused rel-pos attention vs an fp32 torch reference, per output, + timings."""
import sys

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import functional as F  # noqa: E402


def ref(qkv, p, u, v, H):
    B, T, D3 = qkv.shape
    D = D3 // 3
    hd = D // H
    q, k, val = qkv.view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    pp = p.view(2 * T - 1, H, hd).permute(1, 2, 0)
    ac = (q + u[None, :, None, :]) @ k.transpose(-1, -2)
    bd = (q + v[None, :, None, :]) @ pp
    idx = (T - 1 - torch.arange(T, device="cuda").unsqueeze(1)) + torch.arange(T, device="cuda").unsqueeze(0)
    bd = bd.gather(-1, idx.expand(B, H, T, T))
    a = ((ac + bd) * hd ** -0.5).softmax(-1)
    return (a @ val).transpose(1, 2).reshape(B, T, D)


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def run(B, T, H, scale_in=0.6, timing=False):
    D = H * 64
    g = torch.Generator(device="cuda").manual_seed(T + H)
    mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device="cuda") * sc)  # noqa: E731
    qkv = mk(B, T, 3 * D, sc=scale_in).to(torch.bfloat16)
    p = mk(2 * T - 1, D, sc=0.5).to(torch.bfloat16)
    u, v = mk(H, 64, sc=0.3), mk(H, 64, sc=0.3)
    w = mk(B, T, D).to(torch.bfloat16)
    F.set_precision("bf16")
    res = {}
    outs = {}
    for fused in (True, False):
        F.set_fused_attention(fused)
        ins = [qkv.clone().requires_grad_(True), p.clone().requires_grad_(True), u.clone().requires_grad_(True), v.clone().requires_grad_(True)]
        o = F.relpos_attention(ins[0], ins[1], ins[2], ins[3], H)
        o.backward(w)
        torch.cuda.synchronize()
        outs[fused] = (o, [i.grad for i in ins])
    F.set_fused_attention(True)
    insr = [qkv.float().requires_grad_(True), p.float().requires_grad_(True), u.clone().requires_grad_(True), v.clone().requires_grad_(True)]
    o_ref = ref(*insr, H)
    o_ref.backward(w.float())
    for fused in (True, False):
        tag = "f" if fused else "u"
        o, gr = outs[fused]
        res[f"o_{tag}"] = rel(o, o_ref)
        D_ = D
        res[f"dq_{tag}"] = rel(gr[0][..., :D_], insr[0].grad[..., :D_])
        res[f"dk_{tag}"] = rel(gr[0][..., D_:2 * D_], insr[0].grad[..., D_:2 * D_])
        res[f"dv_{tag}"] = rel(gr[0][..., 2 * D_:], insr[0].grad[..., 2 * D_:])
        res[f"dp_{tag}"] = rel(gr[1], insr[1].grad)
        res[f"du_{tag}"] = rel(gr[2], insr[2].grad)
        res[f"dvb_{tag}"] = rel(gr[3], insr[3].grad)
    print(f"B={B} T={T} H={H}: " + " ".join(f"{k}={v:.1e}" for k, v in res.items()), flush=True)
    if timing:
        for fused in (True, False):
            F.set_fused_attention(fused)
            ins = [qkv.clone().requires_grad_(True), p.clone().requires_grad_(True), u.clone().requires_grad_(True), v.clone().requires_grad_(True)]
            for _ in range(2):
                o = F.relpos_attention(*ins, H)
                o.backward(w)
            torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            for _ in range(3):
                o = F.relpos_attention(*ins, H)
            e[1].record()
            for _ in range(3):
                o.backward(w, retain_graph=True)
            e[2].record()
            torch.cuda.synchronize()
            print(f"   fused={fused}: fwd {e[0].elapsed_time(e[1]) / 3:.3f} ms  bwd {e[1].elapsed_time(e[2]) / 3:.3f} ms", flush=True)
        F.set_fused_attention(True)


if __name__ == "__main__":
    run(1, 128, 1)
    run(1, 256, 1)
    run(2, 200, 2)
    run(1, 37, 3)
    run(2, 1000, 12)
    run(64, 1000, 12, timing=True)
