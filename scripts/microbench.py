"""Kernel micro-benchmarks on one B200 (CUDA events on the launching stream, L2 flushed between iterations)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from transformer4sed_b200 import ops  # noqa: E402
from transformer4sed_b200.utils import synth  # noqa: E402


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_gemm(res):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = [("qkv", 76160, 2304, 768), ("proj", 76160, 768, 768), ("fc1", 76160, 3072, 768), ("fc2", 76160, 768, 3072),
              ("square8k", 8192, 8192, 8192)]
    for dtype in (torch.bfloat16, torch.float32):
        for name, M, N, K in shapes:
            x = torch.randn(M, K, device="cuda").to(dtype)
            w = torch.randn(N, K, device="cuda").to(dtype)
            y = torch.empty(M, N, device="cuda", dtype=dtype)
            f = lambda: ops.gemm(ops.Op(x, M, K), ops.Op(w, N, K), ops.Out(y, N), M, N, K)  # noqa: E731
            med, best = timeit(f, flush=flush)
            if name != "square8k":
                bias = torch.randn(N, device="cuda")
                aux = torch.empty_like(y)
                resid = torch.randn(M, N, device="cuda").to(dtype)
                fe = {"bias": lambda: ops.gemm(ops.Op(x, M, K), ops.Op(w, N, K), ops.Out(y, N), M, N, K, bias=bias),
                      "bias_gelu_aux": lambda: ops.gemm(ops.Op(x, M, K), ops.Op(w, N, K), ops.Out(y, N), M, N, K, bias=bias,
                                                        aux=ops.Out(aux, N), act=ops.ACT_GELU),
                      "bias_res": lambda: ops.gemm(ops.Op(x, M, K), ops.Op(w, N, K), ops.Out(y, N), M, N, K, bias=bias,
                                                   residual=ops.Out(resid, N))}
                extra = {k: timeit(v, flush=flush)[0] for k, v in fe.items()}
                del aux, resid
            else:
                extra = {}
            ref = lambda: torch.matmul(x, w.t(), out=y)  # noqa: E731
            torch.backends.cuda.matmul.allow_tf32 = True
            rmed, rbest = timeit(ref, flush=flush)
            fl = 2.0 * M * N * K
            res.append(dict(kernel="gemm", name=name, dtype=str(dtype), M=M, N=N, K=K, ms=med, tflops=fl / med / 1e9,
                            best_tflops=fl / best / 1e9, cublas_ms=rmed, cublas_tflops=fl / rmed / 1e9, epilogue_ms=extra))
            print(res[-1], flush=True)
            del x, w, y


def bench_mel(res):
    from transformer4sed_b200.src_models.passt.passt_feature_extraction import PasstFeatureExtractor
    ext = PasstFeatureExtractor(fmin_aug_range=10, fmax_aug_range=2000).cuda().eval()
    for B in (64, 256):
        wav = synth.synth_wav(8, 320000, seed=1).cuda().repeat(B // 8, 1).contiguous()
        f = lambda: ext.logmel(wav)  # noqa: E731
        med, best = timeit(f, iters=20)   # inputs: B*1.28 MB (82 MB at B=64, 328 MB > L2 at B=256)
        bytes_ = B * (4 * 320000 + 4 * 128 * 1000)
        res.append(dict(kernel="mel", B=B, ms=med, gbs=bytes_ / med / 1e6, best_gbs=bytes_ / best / 1e6, clips_per_s=B / med * 1e3))
        print(res[-1], flush=True)


if __name__ == "__main__":
    res = []
    which = sys.argv[1:] or ["mel", "gemm"]
    if "mel" in which:
        bench_mel(res)
    if "gemm" in which:
        bench_gemm(res)
    json.dump(res, open("gpurun_out/microbench.json", "w"), indent=1)
