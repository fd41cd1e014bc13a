import sys, os, ctypes
import torch
sys.path.insert(0, ".")
from transformer4sed_b200 import functional as F, _lib, ops
F.set_precision("bf16")
B, T, H = 1, 128, 1
D = 64
g = torch.Generator(device="cuda").manual_seed(1)
qkv = (torch.randn(B, T, 3 * D, generator=g, device="cuda") * 0.6).to(torch.bfloat16)
p = (torch.randn(2 * T - 1, D, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
qu = qkv[..., :D].contiguous(); qv = qkv[..., :D].contiguous()
o = torch.empty(B, T, D, dtype=torch.bfloat16, device="cuda")
lse = torch.empty(B, H, 128, dtype=torch.float32, device="cuda")
a = F._relattn_desc(qkv, qu, qv, p, o, lse, B, T, D, H, 0.125)
lib = _lib.load()
print("fwd rc", lib.t4s_relattn_fwd(ctypes.byref(a), None)); 
try:
    torch.cuda.synchronize(); print("fwd ok", o.float().abs().mean().item(), lse[0,0,:4])
except Exception as e:
    print("fwd FAILED", e); sys.exit(1)
do = torch.randn(B, T, D, device="cuda").to(torch.bfloat16)
dqkv = torch.empty_like(qkv); dqu = torch.empty_like(qu); delta = torch.empty_like(lse)
dbd = torch.zeros(B, H, T, 256, dtype=torch.bfloat16, device="cuda")
gg = _lib.RelAttnBwd(); gg.fwd = a
gg.d_o, gg.do_ld, gg.do_bs = do.data_ptr(), D, T * D
gg.delta = delta.data_ptr()
gg.dqu, gg.dqu_ld, gg.dqu_bs = dqu.data_ptr(), D, T * D
gg.dk, gg.dv = dqkv.data_ptr() + D * 2, dqkv.data_ptr() + 2 * D * 2
gg.dk_ld = gg.dv_ld = 3 * D; gg.dk_bs = gg.dv_bs = T * 3 * D
gg.dbd, gg.dbd_ld = dbd.data_ptr(), 256
print("bwd rc", lib.t4s_relattn_bwd(ctypes.byref(gg), None), lib.t4s_last_error())
try:
    torch.cuda.synchronize(); print("bwd ok", dqu.float().abs().mean().item(), dbd.float().abs().sum().item())
except Exception as e:
    print("bwd FAILED", e)
