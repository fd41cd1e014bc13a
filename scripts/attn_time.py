"""Device time of the fused attention forward / backward at the bench shape (B=64, N=1190, H=12), CUDA events over 10 launches each."""
import sys

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import functional as F  # noqa: E402

F.set_precision("bf16")
B, N, H, D = 64, 1190, 12, 768
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, N, 3 * D, generator=g, device="cuda").to(torch.bfloat16).requires_grad_(True)
w = torch.randn(B, N, D, generator=g, device="cuda").to(torch.bfloat16)
for _ in range(3):
    o = F.attention(qkv, H)
    o.backward(w, retain_graph=True)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(10):
    o = F.attention(qkv, H)
e[1].record()
for _ in range(10):
    o.backward(w, retain_graph=True)
e[2].record()
torch.cuda.synchronize()
ref = (qkv.detach().float().reshape(B, N, 3, H, 64).permute(2, 0, 3, 1, 4))
q_, k_, v_ = ref[0][:2], ref[1][:2], ref[2][:2]
o_ref = (((q_ @ k_.transpose(-1, -2)) / 8.0).softmax(-1) @ v_).permute(0, 2, 1, 3).reshape(2, N, D)
err = ((o[:2].float() - o_ref).abs().max() / o_ref.abs().max()).item()
print(f"fwd {e[0].elapsed_time(e[1]) / 10:.3f} ms  bwd {e[1].elapsed_time(e[2]) / 10:.3f} ms  o rel err {err:.2e}")
