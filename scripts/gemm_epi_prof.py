"""The two epilogue-heavy GEMMs of the MLP at the bench shape, for ncu / timing: fc1 forward (bias + GELU, two outputs) and the fc2
dgrad with GELU' and the fc1 bias-gradient column sums.  `python scripts/gemm_epi_prof.py [reps]`."""
import sys

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import ops  # noqa: E402

M, D, Hd = 76160, 768, 3072
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device="cuda") * sc).to(torch.bfloat16)  # noqa: E731
x, w1, w2, dy = mk(M, D), mk(Hd, D, sc=0.03), mk(D, Hd, sc=0.03), mk(M, D)
b1 = torch.randn(Hd, generator=g, device="cuda") * 0.1
pre, h, dpre = (torch.empty(M, Hd, dtype=torch.bfloat16, device="cuda") for _ in range(3))
db1 = torch.zeros(Hd, device="cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1


def fc1():
    ops.gemm(ops.Op(x, M, D), ops.Op(w1, Hd, D), ops.Out(h, Hd), M, Hd, D, bias=b1, aux=ops.Out(pre, Hd), act=ops.ACT_GELU)


def dgrad(colsum=True):
    ops.gemm(ops.Op(dy, M, D), ops.Op(w2, Hd, Hd, mn_major=True), ops.Out(dpre, Hd), M, Hd, D, residual=ops.Out(pre, Hd), act=ops.ACT_GELU_GRAD,
             colsum=db1 if colsum else None)


def plain():
    ops.gemm(ops.Op(x, M, D), ops.Op(w1, Hd, D), ops.Out(h, Hd), M, Hd, D)


def bias_gelu_one():
    ops.gemm(ops.Op(x, M, D), ops.Op(w1, Hd, D), ops.Out(h, Hd), M, Hd, D, bias=b1, act=ops.ACT_GELU)


def bias_two():
    ops.gemm(ops.Op(x, M, D), ops.Op(w1, Hd, D), ops.Out(h, Hd), M, Hd, D, bias=b1, aux=ops.Out(pre, Hd))


def resid_add():
    ops.gemm(ops.Op(x, M, D), ops.Op(w1, Hd, D), ops.Out(h, Hd), M, Hd, D, residual=ops.Out(pre, Hd))


for name, fn in (("same shape, no epilogue work, 1 output", plain), ("bias + GELU, 1 output", bias_gelu_one), ("bias, 2 outputs, no GELU", bias_two),
                 ("residual add (TMA-fed), 1 output", resid_add), ("fc1 forward (bias, GELU, 2 outputs)", fc1), ("fc2 dgrad (GELU', colsum)", dgrad), ("fc2 dgrad (GELU', no colsum)", lambda: dgrad(False))):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms * 1e3:.1f} us  {2 * M * D * Hd / ms / 1e9:.0f} TFLOP/s")
