"""GPU: DASM -- query projector, cross-attention-first tagging decoder (boolean tgt_mask), query x frame score
GEMM and the sigmoid * at_out head, through the C ABI against golden vectors of the UNMODIFIED reference; new ops against
plain PyTorch float64."""
import copy

import numpy as np
import pytest
import torch

from conftest import checksum
from transformer4sed_b200 import schema
from transformer4sed_b200.utils import synth

pytestmark = pytest.mark.gpu

DASM_KW = dict(
    cnn_param=dict(n_in_channel=1, activation="cg", conv_dropout=0.0, kernel_size=[3] * 10, padding=[1] * 10, stride=[1] * 10,
                   nb_filters=list(schema.PMAM_FILTERS), pooling=[list(p) for p in schema.PMAM_POOLING]),
    backbone_param=dict(embed_dim=768, passt_feature_layer=10, pretrain_model_path=None, lora_config=dict(r=8, lora_alpha=1, requires_grad_pretrain=False)),
    at_param=dict(at_decoder_layer=2, query_projector=True, query_dim=768, out_type="sigmoid", query=None),
    mlm_dict=None, backbone_upsample_ratio=10, decoder_dim=384, num_heads=12, decoder="transformerXL", decoder_layer_num=3,
    decoder_pos_emd_len=1000, decoder_expand_rate=1, class_num=407)


def relmax(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


# bf16: the query x frame scores are O(10) (synthetic weights) and go through a sigmoid, so bf16 operand rounding shows up as a few %
@pytest.mark.parametrize("mode,tol,gtol", [("tf32x3", 1e-3, 5e-3), ("bf16", 0.12, 0.4)])
def test_dasm_matches_reference(golden, mode, tol, gtol):
    from transformer4sed_b200 import functional as F
    from transformer4sed_b200.src_models.detect_any_sound.detect_any_sound import DASM
    g = golden("dasm_base.npz")
    seed, batch, K = 12, 2, 407
    F.set_precision(mode)
    try:
        net = DASM(**copy.deepcopy(DASM_KW))
        sd = synth.synth_state_dict_like(net, seed)
        assert sorted(sd.keys()) == [str(k) for k in g["sd_keys"]]
        net.load_state_dict(sd, strict=True)
        net = net.cuda()
        assert sorted(n for n, p in net.named_parameters() if p.requires_grad) == [str(k) for k in g["trainable"]]
        np.testing.assert_allclose(checksum(torch.cat([v.flatten().float() for k, v in sorted(sd.items()) if torch.is_floating_point(v)])),
                                   g["sd_ck"], rtol=1e-12)
        ext = net.get_feature_extractor().eval()
        wav = synth.synth_wav(batch, 320000, seed=seed + 1)
        mel = ext.normalize(ext(wav.cuda()))
        query = (torch.nn.functional.normalize(synth.synth_tensor(seed, "queries", (K, 768)), dim=-1) * 3.0).cuda()
        tgt_mask = torch.from_numpy(np.unpackbits(g["tgt_mask"])[:K * K].astype(bool)).view(K, K).cuda()
        pad = torch.zeros(batch, 1000, dtype=torch.bool, device="cuda")
        pad[-1, 900:] = True
        labels = synth.synth_strong_labels(batch, K, 1000, seed + 2).cuda()
        weak_labels = (labels.sum(-1) >= 1).float()
        net.eval()
        with torch.no_grad():
            s, w, o = net(mel, temp_w=4.0, query=query)
            r = dict(strong=relmax(s[:, ::3, ::4], g["eval_strong"]), weak=relmax(w, g["eval_weak"]), at=relmax(o["at_out"], g["eval_at"]))
            mism = float((s.argmax(dim=1).cpu().numpy() != g["eval_argmax"]).mean())
            print(mode, "eval", r, "argmax mismatch", mism)
            assert max(r.values()) < tol, r
            # argmax over 407 near-tied query scores (synthetic weights): exact in the strict mode.  In bf16 the raw flip rate is a
            # tie rate (6 % .. 40 % for the same 6e-2 value error, depending on rounding details), so the check is tie-aware: wherever
            # our argmax differs, the reference's choice must score within the value tolerance of our maximum.
            if mode == "tf32x3":
                assert mism < 1e-3
            else:
                ref_idx = torch.from_numpy(g["eval_argmax"]).long().cuda().unsqueeze(1)
                gap = (s.max(dim=1, keepdim=True).values - s.gather(1, ref_idx)).float()
                real = float((gap > 2 * tol * s.max().item()).float().mean())
                print(mode, "argmax flips beyond the value tolerance", real)
                assert real == 0.0, real
            s, w, o = net(mel, temp_w=4.0, pad_mask=pad, query=query.unsqueeze(0), tgt_mask=tgt_mask.unsqueeze(0))   # DataParallel-style 3-D inputs
            r = dict(strong=relmax(s[:, ::3, ::4], g["evalm_strong"]), weak=relmax(w, g["evalm_weak"]), at=relmax(o["at_out"], g["evalm_at"]))
            print(mode, "eval masked", r)
            assert max(r.values()) < tol, r
            assert s[-1, :, 900:].max().item() <= 1e-7 + 1e-12
        net.train()
        for m in net.modules():          # the golden run zeroed every dropout (torch's stream cannot be replayed)
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                m.dropout = 0.0
        s, w, o = net(mel, temp_w=4.0, query=query)
        loss = F.bce_loss(s, labels) + 0.5 * F.bce_loss(w, weak_labels) + 0.5 * F.bce_loss(o["at_out"], weak_labels)
        loss.backward()
        r = dict(strong=relmax(s[:, ::3, ::4], g["train_strong"]), weak=relmax(w, g["train_weak"]), at=relmax(o["at_out"], g["train_at"]),
                 loss=abs(loss.item() - float(g["train_loss"])) / float(g["train_loss"]))
        print(mode, "train", r)
        assert max(r.values()) < tol, r
        params = dict(net.named_parameters())
        gerr = {}
        for name, norm, head in zip(g["grad_names"], g["grad_norms"], g["grad_heads"]):
            p = params[str(name)]
            assert p.grad is not None, name
            if str(name).startswith("cnn.cnn.conv") and str(name).endswith(".bias"):
                continue
            n = min(8, p.grad.numel())
            den = max(norm, 1e-9)
            gerr[str(name)] = max(abs(p.grad.double().norm().item() - norm) / den,
                                  float(np.abs(p.grad.flatten()[:n].double().cpu().numpy() - head[:n]).max()) / den)
        worst = max(gerr, key=gerr.get)
        print(mode, "grad worst", worst, gerr[worst])
        assert gerr[worst] < gtol, (worst, gerr[worst])
    finally:
        F.set_precision("bf16")


@pytest.mark.parametrize("mode,tol", [("tf32x3", 1e-4), ("bf16", 3e-2)])
def test_multi_head_attention_vs_torch(mode, tol):
    from transformer4sed_b200 import functional as F
    F.set_precision(mode)
    try:
        g = torch.Generator(device="cuda").manual_seed(2)
        B, Nq, Nk, D, H = 2, 37, 101, 96, 6
        mha = torch.nn.MultiheadAttention(D, H, batch_first=True).cuda().double()
        with torch.no_grad():
            mha.in_proj_bias.normal_(0, 0.1)
            mha.out_proj.bias.normal_(0, 0.1)
        q = torch.randn(B, Nq, D, generator=g, device="cuda").requires_grad_(True)
        mem = torch.randn(B, Nk, D, generator=g, device="cuda").requires_grad_(True)
        mask = torch.rand(Nq, Nk, generator=g, device="cuda") < 0.3
        mask[:, 0] = False
        w = [p.detach().float().requires_grad_(True) for p in (mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias)]
        out = F.multi_head_attention(F.to_act(q), F.to_act(mem), w[0], w[1], w[2], w[3], H, attn_mask=mask)
        wgt = torch.randn(out.shape, generator=g, device="cuda")
        (out.float() * wgt).sum().backward()
        qd, md = q.detach().double().requires_grad_(True), mem.detach().double().requires_grad_(True)
        ref, _ = mha(qd, md, md, attn_mask=mask)
        (ref * wgt.double()).sum().backward()
        assert relmax(out.float(), ref) < tol
        assert relmax(q.grad, qd.grad) < 5 * tol and relmax(mem.grad, md.grad) < 5 * tol
        for a, b in zip(w, (mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias)):
            assert relmax(a.grad, b.grad) < 5 * tol
    finally:
        F.set_precision("bf16")


def test_query_head_ops_vs_torch():
    from transformer4sed_b200 import functional as F
    F.set_precision("tf32x3")
    try:
        g = torch.Generator(device="cuda").manual_seed(6)
        B, T, K, C = 2, 120, 23, 64
        x = torch.randn(B, T, C, generator=g, device="cuda").requires_grad_(True)
        emb = torch.randn(B, K, C, generator=g, device="cuda").requires_grad_(True)
        at = torch.rand(B, K, generator=g, device="cuda").requires_grad_(True)
        pad = torch.zeros(B, T, dtype=torch.bool, device="cuda")
        pad[1, 100:] = True
        score = F.query_frame_score(x, emb)
        strong, weak = F.query_pool(score * 0.3, at, 0.5, pad)
        ws, ww = torch.randn(strong.shape, generator=g, device="cuda"), torch.randn(weak.shape, generator=g, device="cuda")
        ((strong * ws).sum() + (weak * ww).sum()).backward()
        xd, ed, ad = [t.detach().double().requires_grad_(True) for t in (x, emb, at)]
        sc = torch.einsum("bqc,bct->bqt", ed, xd.transpose(1, 2)).transpose(1, 2) * 0.3
        sed = torch.sigmoid(sc / 0.5) * ad.unsqueeze(1)
        sed = sed.masked_fill(pad.unsqueeze(-1), 0.0)
        sed = torch.clamp(sed, 1e-7, 1.0)
        wk = torch.clamp((sed * sed).sum(1) / sed.sum(1), 1e-7, 1.0)
        ((sed.transpose(1, 2) * ws.double()).sum() + (wk * ww.double()).sum()).backward()
        assert relmax(strong, sed.transpose(1, 2)) < 1e-4 and relmax(weak, wk) < 1e-4
        assert relmax(x.grad, xd.grad) < 1e-3 and relmax(emb.grad, ed.grad) < 1e-3 and relmax(at.grad, ad.grad) < 1e-3
        # dropout: keep rate and mask reuse between forward and backward
        y = torch.ones(1 << 18, device="cuda", requires_grad=True).reshape(1 << 10, 1 << 8)
        y.retain_grad()
        o = F.dropout(y, 0.1, True)
        assert abs((o != 0).float().mean().item() - 0.9) < 0.01 and abs(o.max().item() - 1 / 0.9) < 1e-6
        o.sum().backward()
        assert torch.equal(y.grad != 0, o != 0)
        assert F.dropout(y, 0.1, False) is y
    finally:
        F.set_precision("bf16")
