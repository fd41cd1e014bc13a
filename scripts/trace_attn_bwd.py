"""Pipeline timeline of one persistent CTA of the fused attention backward (T4S_TRACE build): global tiles 20..35 of CTA 70."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe", "libt4s_trace.so")
from transformer4sed_b200 import functional as F  # noqa: E402

F.set_precision("bf16")
B, N, H, D = 64, 1190, 12, 768
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(B, N, 3 * D, generator=g, device="cuda").to(torch.bfloat16).requires_grad_(True)
w = torch.randn(B, N, D, generator=g, device="cuda").to(torch.bfloat16)
o = F.attention(qkv, H)
for _ in range(2):
    o.backward(w, retain_graph=True)
torch.cuda.synchronize()
lib = _lib.load()
n = 4096
buf = (ctypes.c_longlong * n)()
lib.t4s_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.t4s_debug_trace(buf, n)
t = list(buf)


def at(w_, j, e):
    return t[(w_ * 16 + j) * 8 + e]


t0 = at(0, 0, 0)
for w_ in (0, 5):
    print(f"softmax warp {w_}: tile: wait-start, S/dP ready, loaded, computed (dS stored), dq drained, handed over")
    for j in range(16):
        print(f"   T={20 + j}: " + " ".join(f"{at(w_, j, e) - t0:7d}" for e in range(6)))
print("item epilogue of warp 0 (after T=29): tail done, AccFull seen, column sums done, dV/dK store issued:",
      at(0, 9, 6) - t0, at(0, 9, 7) - t0, at(0, 10, 6) - t0, at(0, 10, 7) - t0)
print("MMA warp: tile: loop top, S/dP(T+1) issued, P/dS(T) ready, dK/dQ/dV(T) issued")
for j in range(16):
    print(f"   T={20 + j}: " + " ".join(f"{at(9, j, e) - t0:7d}" for e in range(4)))
