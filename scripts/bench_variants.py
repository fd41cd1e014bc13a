"""Informational timings of the other BASELINE.json configurations on one B200 (not bench.py lines): PMAM post-pre-training
(`PaSST_CNN`, config/pmam/post_pretrain.yaml, 32 clips/GPU) and DASM (K = 407 queries, 8 clips/GPU): forward + loss + backward."""
import copy
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from transformer4sed_b200 import functional as F  # noqa: E402
from transformer4sed_b200.utils import synth  # noqa: E402


def timed(step, n=3, warm=2):
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        step()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def pmam(B=32):
    from test_pmam_gpu import CNN_PARAM, PASST_SED_PARAM
    from transformer4sed_b200.src_models.cnn_transformer.passt_cnn import PaSST_CNN
    net = PaSST_CNN(dict(PASST_SED_PARAM, load_pretrained_model=False), dict(CNN_PARAM, conv_dropout=0.5))
    net.load_state_dict(synth.synth_state_dict_like(net, 1))
    net = net.cuda().train()
    ext = net.get_feature_extractor().eval()
    wav = synth.synth_wav(4, 320000, seed=2).repeat(B // 4, 1).cuda()
    protos = torch.nn.functional.normalize(torch.randn(30, 768), dim=-1).cuda()
    labels = (torch.rand(B, 1000, 30, device="cuda") < 0.1).float()

    def step():
        for p in net.parameters():
            p.grad = None
        mel = ext.logmel(wav)
        pred, other = net(mel)
        strong = F.prototype_predict(pred, protos)
        rows = other["mask_id_seq"].reshape(-1)
        loss = F.bce_loss(strong.reshape(-1, 30)[rows], labels.reshape(-1, 30)[rows]) + 0.5 * F.bce_loss(other["at_out"], labels.amax(1))
        loss.backward()

    ms = timed(step)
    return dict(config="PMAM post-pretrain (PaSST_CNN, LoRA r=8, CNN branch, TXL d=384, MLM + prototype BCE)", clips_per_gpu=B, ms_per_step=ms,
                clips_per_s=B / ms * 1e3, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)


def dasm(B=8, K=407):
    from test_dasm_gpu import DASM_KW
    from transformer4sed_b200.src_models.detect_any_sound.detect_any_sound import DASM
    net = DASM(**copy.deepcopy(DASM_KW))
    net.load_state_dict(synth.synth_state_dict_like(net, 1))
    net = net.cuda().train()
    ext = net.get_feature_extractor().eval()
    wav = synth.synth_wav(4, 320000, seed=2).repeat(B // 4, 1).cuda()
    query = (torch.nn.functional.normalize(torch.randn(K, 768), dim=-1) * 3).cuda()
    labels = (torch.rand(B, K, 1000, device="cuda") < 0.05).float()

    def step():
        for p in net.parameters():
            p.grad = None
        mel = ext.logmel(wav)
        s, w, o = net(mel, temp_w=4.0, query=query)
        loss = F.bce_loss(s, labels) + 0.5 * F.bce_loss(w, labels.amax(-1)) + 0.5 * F.bce_loss(o["at_out"], labels.amax(-1))
        loss.backward()

    ms = timed(step)
    return dict(config=f"DASM (K={K} queries, LoRA backbone, CNN branch, 2-layer tagging decoder, TXL d=384)", clips_per_gpu=B, ms_per_step=ms,
                clips_per_s=B / ms * 1e3, peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)


if __name__ == "__main__" and len(sys.argv) == 1:
    F.set_precision("bf16")
    out = []
    for fn in (pmam, dasm):
        torch.cuda.reset_peak_memory_stats()
        t0 = time.time()
        out.append(fn())
        print(json.dumps(out[-1]), f"({time.time() - t0:.0f} s)", flush=True)
        torch.cuda.empty_cache()


def profile_pmam(B=32):
    """Per-op device time of one PMAM step (C-ABI launches bracketed by CUDA events)."""
    from transformer4sed_b200 import _lib
    from test_pmam_gpu import CNN_PARAM, PASST_SED_PARAM
    from transformer4sed_b200.src_models.cnn_transformer.passt_cnn import PaSST_CNN
    net = PaSST_CNN(dict(PASST_SED_PARAM, load_pretrained_model=False), dict(CNN_PARAM, conv_dropout=0.5))
    net.load_state_dict(synth.synth_state_dict_like(net, 1))
    net = net.cuda().train()
    ext = net.get_feature_extractor().eval()
    wav = synth.synth_wav(4, 320000, seed=2).repeat(B // 4, 1).cuda()
    protos = torch.nn.functional.normalize(torch.randn(30, 768), dim=-1).cuda()

    def step():
        for p in net.parameters():
            p.grad = None
        pred, other = net(ext.logmel(wav))
        F.prototype_predict(pred, protos).sum().backward()

    step(); step()
    torch.cuda.synchronize()
    _lib.profiler = _lib.LaunchProfiler()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); step(); t1.record()
    torch.cuda.synchronize()
    summ = _lib.profiler.summary()
    _lib.profiler = None
    print("step ms", t0.elapsed_time(t1), "sum of ops", sum(v[1] for v in summ.values()))
    for k, v in list(summ.items())[:28]:
        print(f"{v[1]:8.3f} {v[0]:4d} {k}")


if len(sys.argv) > 1 and sys.argv[1] == "profile":
    profile_pmam()
