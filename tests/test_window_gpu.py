"""GPU: sliding-window global-local fusion and MLM frame replacement (a5), through the C ABI.

Model level: `PaSST_SED(encoder_win=True)` against golden vectors of the UNMODIFIED reference (validation kwargs in eval mode,
teacher kwargs in train mode with the reference's CPU RNG replayed).  Op level: windowed patch-embed vs per-crop patch-embed,
overlap-add mean and frame replacement (forward + backward) vs plain PyTorch float64 on the same inputs.
"""
import numpy as np
import pytest
import torch

from conftest import checksum
from transformer4sed_b200.utils import synth

pytestmark = pytest.mark.gpu

BASE = dict(passt_feature_layer=10, f_pool="mean_pool", decode_ratio=10, at_adapter=True, decoder="transformerXL", decoder_layer_num=3,
            decoder_pos_emd_len=1000, mlm=False)
PRE = dict(BASE, mlm=True, mlm_dict=dict(strategy="block", block_width=10, mask_rate=0.75, out_dim=768))


def relmax(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def build(kw, seed):
    from transformer4sed_b200.src_models.passt.passt_sed import PaSST_SED
    net = PaSST_SED(load_pretrained_model=False, **kw)
    net.load_state_dict(synth.synth_state_dict_like(net, seed), strict=True)
    return net.cuda()


@pytest.mark.parametrize("mode,tol", [("tf32x3", 1e-3), ("bf16", 6e-2)])
def test_sliding_window_matches_reference(golden, mode, tol):
    from transformer4sed_b200 import functional as F
    g = golden("matsed_window_base.npz")
    F.set_precision(mode)
    try:
        net = build(BASE, 8)
        ext = net.get_feature_extractor().eval()
        wav = synth.synth_wav(1, 320000, seed=9)
        np.testing.assert_allclose(checksum(wav), g["wav_ck"], rtol=1e-12)
        mel = ext.normalize(ext(wav.cuda()))
        cap = {}
        net.slide_window_layer.register_forward_hook(lambda m, i, o: cap.__setitem__("x_local", o))
        net.eval()
        with torch.no_grad():
            s, w, _ = net(mel, encoder_win=True, mix_rate=0.5, win_param=[512, 31], temp_w=0.5)
        r = dict(local=relmax(cap["x_local"].float()[:, ::4, ::4], g["local_val"]), strong=relmax(s, g["strong_val"]), weak=relmax(w, g["weak_val"]))
        print(mode, "val", r)
        assert max(r.values()) < tol, r
        if mode == "tf32x3":
            assert torch.equal(s.argmax(1).cpu(), torch.from_numpy(g["strong_val"]).argmax(1))
        net.train()
        torch.manual_seed(int(g["train_seed"]))    # the per-window time-table offsets come from the global CPU RNG (passt.py:508)
        with torch.no_grad():
            s, w, _ = net(mel, encoder_win=True, mix_rate=0.5, win_param=[512, 49], temp_w=1)
        r = dict(local=relmax(cap["x_local"].float()[:, ::4, ::4], g["local_train"]), strong=relmax(s, g["strong_train"]),
                 weak=relmax(w, g["weak_train"]))
        print(mode, "train", r)
        assert max(r.values()) < tol, r
    finally:
        F.set_precision("bf16")


def test_sliding_window_chunking_and_grad():
    """Chunked window passes give the same embedding, and gradients flow to the backbone through the windows."""
    from transformer4sed_b200 import functional as F
    from transformer4sed_b200.src_models.passt.passt_win import PasstWithSlide
    F.set_precision("tf32")
    try:
        net = build(dict(BASE, decoder_layer_num=1), 3).eval()
        mel = torch.randn(2, 128, 1000, generator=torch.Generator().manual_seed(1)).cuda() * 0.5
        with torch.no_grad():
            a = PasstWithSlide(net, [512, 49])(mel, emb_len=1000)
            sw = PasstWithSlide(net, [512, 49])
            sw.max_sequences = 6
            b = sw(mel, emb_len=1000)
        assert a.shape == (2, 1000, 768) and torch.equal(a, b)
        # per-crop `encode` (reference contract) agrees with the batched path
        with torch.no_grad():
            c = PasstWithSlide(net, [512, 49]).encode(mel[:, :, 49:49 + 512])
            ref = torch.zeros(2, 1000, 768, device="cuda")
            cnt = torch.zeros(1000, device="cuda")
            for wl, width in PasstWithSlide(net, [512, 49]).window_starts(1000):
                o = PasstWithSlide(net, [512, 49]).encode(mel[:, :, wl:wl + width]).float()
                ref[:, wl:wl + o.shape[1]] += o
                cnt[wl:wl + o.shape[1]] += 1
            ref = torch.nan_to_num(ref / cnt[None, :, None], nan=0.0)   # frames no window covers are 0 (encoder_slide_window.py:36)
        assert c.shape == (2, 500, 768)
        assert relmax(a.float(), ref) < 1e-5
        s, w, _ = net(mel, encoder_win=True, win_param=[512, 98])
        (s.sum() + w.sum()).backward()
        gw = net.backbone.patch_embed.proj.weight.grad
        assert gw is not None and torch.isfinite(gw).all() and gw.abs().sum() > 0
        assert net.backbone.time_new_pos_embed.grad is not None
    finally:
        F.set_precision("bf16")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_overlap_add_vs_torch(dtype):
    from transformer4sed_b200 import functional as F
    B, C, frames = 3, 64, 200
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(4 * B, 50, C, generator=g, device="cuda").to(dtype).requires_grad_(True)   # 4 windows of 50 frames
    b = torch.randn(1 * B, 40, C, generator=g, device="cuda").to(dtype).requires_grad_(True)   # a shorter last window, cut at the end
    sa, sb = [0, 30, 60, 90], [170]                                                           # frames 140..169 uncovered -> 0
    out = F.window_overlap_add([(a, sa), (b, sb)], B, frames)
    ref = torch.zeros(B, frames, C, dtype=torch.float64, device="cuda")
    cnt = torch.zeros(frames, dtype=torch.float64, device="cuda")
    ad, bd = a.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    for i, st in enumerate(sa):
        ref[:, st:st + 50] = ref[:, st:st + 50] + ad[i * B:(i + 1) * B]
        cnt[st:st + 50] += 1
    ref[:, 170:200] = ref[:, 170:200] + bd[:, :30]
    cnt[170:200] += 1
    ref = torch.where(cnt[None, :, None] > 0, ref / cnt.clamp_min(1)[None, :, None], torch.zeros_like(ref))
    t = 1e-6 if dtype == torch.float32 else 1e-2
    assert relmax(out.float(), ref) < t
    assert out[:, 140:170].abs().max().item() == 0.0
    wgt = torch.randn(B, frames, C, generator=g, device="cuda")
    (out.float() * wgt).sum().backward()
    (ref * wgt.double()).sum().backward()
    assert relmax(a.grad.float(), ad.grad) < t and relmax(b.grad.float(), bd.grad) < t
    assert b.grad[:, 30:].abs().max().item() == 0.0


def test_windowed_patch_embed_equals_per_crop():
    from transformer4sed_b200 import functional as F
    F.set_precision("tf32")
    try:
        g = torch.Generator(device="cuda").manual_seed(2)
        D, B = 64, 2
        mel = torch.randn(B, 128, 300, generator=g, device="cuda")
        conv_w = (torch.randn(D, 1, 16, 16, generator=g, device="cuda") * 0.05).requires_grad_(True)
        conv_b = torch.randn(D, generator=g, device="cuda").requires_grad_(True)
        tpos = torch.randn(D, 99, generator=g, device="cuda").requires_grad_(True)
        fpos = torch.randn(D, 12, generator=g, device="cuda").requires_grad_(True)
        cls, dist = torch.randn(D, generator=g, device="cuda").requires_grad_(True), torch.randn(D, generator=g, device="cuda").requires_grad_(True)
        npos = torch.randn(2, D, generator=g, device="cuda").requires_grad_(True)
        params = [conv_w, conv_b, tpos, fpos, cls, dist, npos]
        starts, offs, t_dim = [0, 31, 188], [5, 0, 40], 10     # crops of width 106..112 -> 10 patches
        x = F.patch_embed(mel, *params, stride=10, windows=(starts, t_dim, offs))
        assert x.shape == (3 * B, 2 + 12 * t_dim, D)
        wgt = torch.randn(x.shape, generator=g, device="cuda")
        (x * wgt).sum().backward()
        grads = [p.grad.clone() for p in params]
        for p in params:
            p.grad = None
        tot = 0
        for i, (s, o) in enumerate(zip(starts, offs)):
            xi = F.patch_embed(mel[:, :, s:s + 106].contiguous(), *params, stride=10, t_offset=o)
            assert torch.equal(xi, x[i * B:(i + 1) * B])
            tot = tot + (xi * wgt[i * B:(i + 1) * B]).sum()
        tot.backward()
        for p, gw in zip(params, grads):
            assert relmax(gw, p.grad) < 1e-5
    finally:
        F.set_precision("bf16")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mask_rows_vs_torch(dtype):
    from oracle import model as OM
    from transformer4sed_b200 import functional as F
    B, T, C = 4, 1000, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, T, C, generator=g, device="cuda").to(dtype).requires_grad_(True)
    tok = (torch.randn(1, 1, C, generator=g, device="cuda") * 0.02).requires_grad_(True)
    cg = torch.Generator().manual_seed(6)
    mask = OM.block_mask_from_noise(torch.rand(B, 100, generator=cg), 0.75, 10, T)
    probs = torch.rand(B * T, generator=cg)
    m = mask.view(-1)
    mm, rm = m & (probs < 0.8), m & (probs >= 0.8) & (probs < 0.9)
    ridx = torch.randint(0, B * T, (int(rm.sum()),), generator=cg)
    ridx[:5] = ridx[5]                  # several frames copy the same source row (gradient accumulates there)
    out = F.mask_rows(x, tok, mm.cuda(), rm.cuda(), ridx.cuda())
    xd = x.detach().double().cpu().requires_grad_(True)
    td = tok.detach().double().cpu().requires_grad_(True)
    ref = OM.apply_mask(xd, mask, probs, ridx, td, style=(0.8, 0.1, 0.1))
    t = 1e-6 if dtype == torch.float32 else 1e-2
    if dtype == torch.float32:
        assert torch.equal(out.cpu().double(), ref.detach())
    assert relmax(out.float(), ref) < t
    wgt = torch.randn(B, T, C, generator=g, device="cuda")
    (out.float() * wgt).sum().backward()
    (ref * wgt.double().cpu()).sum().backward()
    assert relmax(x.grad.float(), xd.grad) < t
    assert relmax(tok.grad, td.grad) < (1e-5 if dtype == torch.float32 else 2e-2)


def test_mlm_b1_masking_applies_like_reference(golden):
    """B == 1: upstream masking is NOT a no-op (the reshape is a view).  Replay the reference's CPU RNG and match its decoder
    input, prediction and loss."""
    from transformer4sed_b200 import functional as F
    g = golden("matsed_mlm_base_b1.npz")
    assert not bool(g["decoder_in_equals_input"])
    F.set_precision("tf32x3")
    try:
        net = build(PRE, 6).train()
        net.mlm_tool.device = "cpu"     # draw on the CPU generator like the recorded reference run
        ext = net.get_feature_extractor().eval()
        wav = synth.synth_wav(1, 320000, seed=7)
        np.testing.assert_allclose(checksum(wav), g["wav_ck"], rtol=1e-12)
        mel = ext.normalize(ext(wav.cuda()))
        cap = {}
        net.decoder.register_forward_pre_hook(lambda m, i: cap.__setitem__("x", i[0]))
        torch.manual_seed(9)
        pred, other = net(mel)
        mask_ref = torch.from_numpy(np.unpackbits(g["mask"])[:1000].astype(bool)).view(1, 1000)
        assert torch.equal(other["mask_id_seq"].cpu(), mask_ref)
        assert relmax(cap["x"].float()[:, ::8, ::4], g["decoder_in"]) < 1e-3
        assert relmax(pred.float()[:, ::8, ::4], g["pred"]) < 1e-3
        loss = F.mse_loss(other["frame_before_mask"].detach(), pred, other["mask_id_seq"])
        assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < 1e-3
        loss.backward()
        assert net.mask_token.grad is not None and net.mask_token.grad.abs().sum() > 0
    finally:
        F.set_precision("bf16")
