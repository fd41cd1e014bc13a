"""Phase timeline of two softmax warps of one CTA of the rel-pos attention forward (T4S_TRACE build, scripts/trace_attn.sh)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe", os.environ.get("T4S_TRACE_LIB", "libt4s_trace.so"))
from transformer4sed_b200 import functional as F  # noqa: E402

F.set_precision("bf16")
B, T, H = 64, 1000, 12
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s, sc=1.0: torch.randn(*s, generator=g, device="cuda") * sc  # noqa: E731
ins = [mk(B, T, 3 * D, sc=0.6).to(torch.bfloat16), mk(2 * T - 1, D, sc=0.5).to(torch.bfloat16), mk(H, 64, sc=0.3), mk(H, 64, sc=0.3)]
with torch.no_grad():
    for _ in range(2):
        o = F.relpos_attention(*ins, H)
torch.cuda.synchronize()
lib = _lib.load()
n = 4096
buf = (ctypes.c_longlong * n)()
lib.t4s_debug_trace_rel.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.t4s_debug_trace_rel(buf, n)
t = list(buf)


def at(w_, j, e):
    return t[(w_ * 16 + j) * 8 + e]


for w_ in (0, 5):
    t0 = at(w_, 0, 0)
    print(f"softmax warp {w_}: top, S ready, S loaded, skewed + max (+ wait P free), exponentials done, P handed over")
    for j in range(8):
        row = [at(w_, j, e) - t0 for e in range(6)]
        print(f"   j={j}: " + " ".join(f"{v:7d}" for v in row) + "   | deltas " + " ".join(f"{b - a:5d}" for a, b in zip(row, row[1:])))
