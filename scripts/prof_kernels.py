"""One launch of each hot kernel at the bench shapes (B=64), for `ncu --set full` captures (no timing here)."""
import sys

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import functional as F, ops  # noqa: E402

F.set_precision("bf16")
which = sys.argv[1:] or ["attn", "rel", "gemm"]
g = torch.Generator(device="cuda").manual_seed(0)
B, H, D = 64, 12, 768
if "attn" in which:
    qkv = (torch.randn(B, 1190, 3 * D, generator=g, device="cuda")).to(torch.bfloat16).requires_grad_(True)
    w = torch.randn(B, 1190, D, generator=g, device="cuda").to(torch.bfloat16)
    F.attention(qkv, H).backward(w)
if "rel" in which:
    T = 1000
    qkv = (torch.randn(B, T, 3 * D, generator=g, device="cuda") * 0.6).to(torch.bfloat16).requires_grad_(True)
    p = (torch.randn(2 * T - 1, D, generator=g, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    u = (torch.randn(H, 64, generator=g, device="cuda") * 0.3).requires_grad_(True)
    v = (torch.randn(H, 64, generator=g, device="cuda") * 0.3).requires_grad_(True)
    w = torch.randn(B, T, D, generator=g, device="cuda").to(torch.bfloat16)
    F.relpos_attention(qkv, p, u, v, H).backward(w)
if "gemm" in which:
    M = B * 1190
    x = torch.randn(M, 768, device="cuda").to(torch.bfloat16)
    w1 = torch.randn(3072, 768, device="cuda").to(torch.bfloat16)
    bias = torch.randn(3072, device="cuda")
    y = torch.empty(M, 3072, device="cuda", dtype=torch.bfloat16)
    aux = torch.empty_like(y)
    ops.gemm(ops.Op(x, M, 768), ops.Op(w1, 3072, 768), ops.Out(y, 3072), M, 3072, 768, bias=bias)                       # plain + bias
    ops.gemm(ops.Op(x, M, 768), ops.Op(w1, 3072, 768), ops.Out(y, 3072), M, 3072, 768, bias=bias, aux=ops.Out(aux, 3072),
             act=ops.ACT_GELU)                                                                                         # fc1 forward
if "mel" in which:
    from transformer4sed_b200.src_models.passt.passt_feature_extraction import PasstFeatureExtractor
    ext = PasstFeatureExtractor(n_mels=128, sr=32000, win_length=800, hopsize=320, n_fft=1024, htk=False, fmin=0.0, fmax=None,
                                fmin_aug_range=10, fmax_aug_range=2000).cuda().eval()
    wav = torch.randn(B, 320000, generator=g, device="cuda") * 0.1
    ext.logmel(wav)
if "melg" in which:
    from transformer4sed_b200.src_preprocess.feats_extraction import setmelspectrogram
    ms = setmelspectrogram(dict(sample_rate=16000, n_window=2048, hop_length=256, f_min=0, f_max=8000, n_mels=128)).cuda()
    wav = torch.randn(B, 160000, generator=g, device="cuda") * 0.1
    ms.logmel(wav)
    ms.logmel(wav)
torch.cuda.synchronize()
print("done")
