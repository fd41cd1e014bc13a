"""Pipeline timeline of one CTA of the fused attention forward (T4S_TRACE build, scripts/trace_attn.sh): clock64 at the hand-over
points of the softmax warps and the MMA warp, printed relative to the CTA start."""
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
qkv = torch.randn(B, N, 3 * D, generator=g, device="cuda").to(torch.bfloat16)
for _ in range(3):
    o = F.attention(qkv, H)
torch.cuda.synchronize()
lib = _lib.load()
n = 4096
buf = (ctypes.c_longlong * n)()
lib.t4s_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = lib.t4s_debug_trace(buf, n)
t = list(buf)


def at(w, j, e):
    return t[(w * 16 + j) * 8 + e]


t0 = at(16, 15, 0)
print("CTA start (after setup) = 0; all times in SM clocks")
for w in (0, 4, 8, 12):
    print(f"softmax warp {w}: tile: wait-start, S ready, S loaded, max done, exp done, P handed over")
    for j in range(10):
        print(f"   j={j}: " + " ".join(f"{at(w, j, e) - t0:7d}" for e in range(6)))
    print(f"   epilogue start {at(w, 15, 0) - t0}, stores done {at(w, 15, 1) - t0}, past final barrier {at(w, 15, 2) - t0}")
print("MMA warp: per tile: [group 0: wait-start, P ready, issued] [group 1: ...]")
for j in range(10):
    print(f"   j={j}: " + " ".join(f"{at(17, j, e) - t0:7d}" for e in (0, 1, 2)) + " | " + " ".join(f"{at(18, j, e) - t0:7d}" for e in (0, 1, 2)))
