"""One rel-pos attention forward + backward at the bench shape (B=64, T=1000, H=12) for ncu / timing runs."""
import sys

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import functional as F  # noqa: E402

B, T, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (64, 1000, 12)))
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s, sc=1.0: torch.randn(*s, generator=g, device="cuda") * sc  # noqa: E731
F.set_precision("bf16")
ins = [mk(B, T, 3 * D, sc=0.6).to(torch.bfloat16).requires_grad_(True), mk(2 * T - 1, D, sc=0.5).to(torch.bfloat16).requires_grad_(True),
       mk(H, 64, sc=0.3).requires_grad_(True), mk(H, 64, sc=0.3).requires_grad_(True)]
w = mk(B, T, D).to(torch.bfloat16)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
for _ in range(reps):
    o = F.relpos_attention(*ins, H)
    o.backward(w)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record()
for _ in range(reps):
    o = F.relpos_attention(*ins, H)
e[1].record()
for _ in range(reps):
    o.backward(w, retain_graph=True)
e[2].record()
torch.cuda.synchronize()
print(f"rel-pos attention B={B} T={T} H={H}: fwd {e[0].elapsed_time(e[1]) / reps:.3f} ms  bwd {e[1].elapsed_time(e[2]) / reps:.3f} ms (incl. position GEMMs)")
