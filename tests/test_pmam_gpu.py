"""GPU: PMAM -- `PaSST_CNN` (PaSST + LoRA backbone, CNN branch, attention frequency pooling, TransformerXL d=384,
MLM + prototype head) through the C ABI against golden vectors of the UNMODIFIED reference, plus every new op (3x3 convolution,
BatchNorm, ContextGating, pooling, merge, prototype head, LoRA linear) against plain PyTorch float64 on the same inputs."""
import numpy as np
import pytest
import torch

from conftest import checksum
from transformer4sed_b200 import schema
from transformer4sed_b200.utils import synth

pytestmark = pytest.mark.gpu

# config/pmam/post_pretrain.yaml:48-79 (conv_dropout set to 0 where noted: torch's dropout stream cannot be replayed)
PASST_SED_PARAM = dict(passt_feature_layer=10, class_num=30, f_pool="attention", decode_ratio=10, at_adapter=True, decoder="transformerXL",
                       decoder_layer_num=3, decoder_pos_emd_len=1000, decoder_dim=384, mlm=True,
                       lora_config=dict(r=8, lora_alpha=1, requires_grad_pretrain=False),
                       mlm_dict=dict(strategy="block", block_width=10, mask_rate=0.8, out_dim=768, mask_style=[0.9, 0.05, 0.05]))
CNN_PARAM = dict(n_in_channel=1, activation="cg", conv_dropout=0.0, kernel_size=[3] * 10, padding=[1] * 10, stride=[1] * 10,
                 nb_filters=list(schema.PMAM_FILTERS), pooling=[list(p) for p in schema.PMAM_POOLING])


def relmax(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def build(seed, conv_dropout=0.0):
    from transformer4sed_b200.src_models.cnn_transformer.passt_cnn import PaSST_CNN
    net = PaSST_CNN(dict(PASST_SED_PARAM, load_pretrained_model=False), dict(CNN_PARAM, conv_dropout=conv_dropout))
    sd = synth.synth_state_dict_like(net, seed)
    net.load_state_dict(sd, strict=True)
    return net.cuda(), sd


@pytest.mark.parametrize("mode,tol,gtol", [("tf32x3", 1e-3, 5e-3), ("bf16", 6e-2, 0.3)])
def test_passt_cnn_matches_reference(golden, mode, tol, gtol):
    from transformer4sed_b200 import functional as F
    g = golden("pmam_base.npz")
    seed, batch = 10, 2
    F.set_precision(mode)
    try:
        net, sd = build(seed)
        assert sorted(sd.keys()) == [str(k) for k in g["sd_keys"]]                       # same checkpoint schema as the reference
        assert sorted(n for n, p in net.named_parameters() if p.requires_grad) == [str(k) for k in g["trainable"]]
        np.testing.assert_allclose(checksum(torch.cat([v.flatten().float() for k, v in sorted(sd.items()) if torch.is_floating_point(v)])),
                                   g["sd_ck"], rtol=1e-12)
        net.mlm_tool.device = "cpu"     # the recorded reference run drew its masks from the CPU generator
        ext = net.get_feature_extractor().eval()
        wav = synth.synth_wav(batch, 320000, seed=seed + 1)
        mel = ext.normalize(ext(wav.cuda()))
        protos = torch.nn.functional.normalize(synth.synth_tensor(seed, "prototypes", (30, 768)), dim=-1).cuda()
        labels = synth.synth_strong_labels(batch, 30, 1000, seed + 2).cuda()
        weak_labels = (labels.sum(-1) >= 1).float()
        cap = {}
        net.decoder.register_forward_pre_hook(lambda m, i: cap.__setitem__("x", i[0]))
        # (a) eval mode: BatchNorm running statistics, LoRA merged into the weights
        net.eval()
        torch.manual_seed(seed + 3)
        with torch.no_grad():
            pred, other = net(mel)
            strong = F.prototype_predict(pred, protos)
        assert (np.packbits(other["mask_id_seq"].cpu().numpy()) == g["eval_mask"]).all()
        r = dict(fbm=relmax(other["frame_before_mask"].float()[:, ::8, ::4], g["eval_fbm"]), dec_in=relmax(cap["x"].float()[:, ::8, ::4], g["eval_dec_in"]),
                 pred=relmax(pred.float()[:, ::8, ::4], g["eval_pred"]), at=relmax(other["at_out"], g["eval_at"]),
                 strong=relmax(strong[:, ::8], g["eval_strong"]))
        print(mode, "eval", r)
        assert max(r.values()) < tol, r
        # (b) train mode: batch statistics, un-merged LoRA, masked prototype BCE + AT BCE, gradients of every trainable tensor
        net.train()
        torch.manual_seed(seed + 4)
        pred, other = net(mel)
        m = other["mask_id_seq"]
        assert (np.packbits(m.cpu().numpy()) == g["train_mask"]).all()
        strong = F.prototype_predict(pred, protos)
        # BCELoss over the masked frames == mean over selected rows: weight rows by the mask
        rows = m.reshape(-1)
        p_sel = strong.reshape(-1, 30)[rows]
        y_sel = labels.transpose(1, 2).reshape(-1, 30)[rows]
        loss = F.bce_loss(p_sel, y_sel) + 0.5 * F.bce_loss(other["at_out"], weak_labels)
        loss.backward()
        r = dict(dec_in=relmax(cap["x"].float()[:, ::8, ::4], g["train_dec_in"]), pred=relmax(pred.float()[:, ::8, ::4], g["train_pred"]),
                 at=relmax(other["at_out"], g["train_at"]), loss=abs(loss.item() - float(g["train_loss"])) / float(g["train_loss"]))
        print(mode, "train", r)
        assert max(r.values()) < tol, r
        bn = net.cnn.cnn
        assert relmax(bn.batchnorm0.running_mean, g["bn_cnn.cnn.batchnorm0.running_mean"]) < tol
        assert relmax(bn.batchnorm9.running_var, g["bn_cnn.cnn.batchnorm9.running_var"]) < tol
        assert int(bn.batchnorm0.num_batches_tracked) == 1
        params = dict(net.named_parameters())
        gerr = {}
        for name, norm, head in zip(g["grad_names"], g["grad_norms"], g["grad_heads"]):
            p = params[str(name)]
            assert p.grad is not None, name
            n = min(8, p.grad.numel())
            if str(name).startswith("cnn.cnn.conv") and str(name).endswith(".bias"):
                # a bias in front of a training-mode BatchNorm has a mathematically zero gradient (1e-9 rounding noise upstream)
                assert p.grad.double().norm().item() < 1e-3 * params[str(name)[:-4] + "weight"].grad.double().norm().item(), name
                continue
            den = max(norm, 1e-9)
            gerr[str(name)] = max(abs(p.grad.double().norm().item() - norm) / den,
                                  float(np.abs(p.grad.flatten()[:n].double().cpu().numpy() - head[:n]).max()) / den)
        worst = max(gerr, key=gerr.get)
        print(mode, "grad worst", worst, gerr[worst])
        assert gerr[worst] < gtol, (worst, gerr[worst])
        assert all(p.grad is None for n, p in params.items() if not p.requires_grad)
    finally:
        F.set_precision("bf16")


@pytest.mark.parametrize("mode,tol", [("tf32x3", 1e-3), ("bf16", 0.1)])
def test_passt_cnn_finetune_form_matches_reference(golden, mode, tol):
    """`PaSST_CNN` as the PMAM fine-tuning stages build it (config/pmam/finetune1.yaml:61-82: no LoRA, no MLM, 10 classes, frozen
    merge weight): strong / weak / AT outputs, frame-label argmax, and the pad-mask + temperature path vs the unmodified reference."""
    from transformer4sed_b200 import functional as F
    from transformer4sed_b200.src_models.cnn_transformer.passt_cnn import PaSST_CNN
    g = golden("pmam_finetune.npz")
    seed, batch = 14, 2
    F.set_precision(mode)
    try:
        sed_param = dict(passt_feature_layer=10, f_pool="attention", decode_ratio=10, at_adapter=True, decoder="transformerXL",
                         decoder_layer_num=3, decoder_pos_emd_len=1000, decoder_dim=384, mlm=False, load_pretrained_model=False)
        net = PaSST_CNN(sed_param, dict(CNN_PARAM, conv_dropout=0.5))
        sd = synth.synth_state_dict_like(net, seed)
        net.load_state_dict(sd, strict=True)
        assert sorted(sd.keys()) == [str(k) for k in g["sd_keys"]]
        np.testing.assert_allclose(checksum(torch.cat([v.flatten().float() for k, v in sorted(sd.items()) if torch.is_floating_point(v)])),
                                   g["sd_ck"], rtol=1e-12)
        assert not net.merge_weight.requires_grad
        net = net.cuda().eval()
        ext = net.get_feature_extractor().eval()
        mel = ext.normalize(ext(synth.synth_wav(batch, 320000, seed=seed + 1).cuda()))
        pad = torch.zeros(batch, 1000, dtype=torch.bool, device="cuda")
        pad[-1, 850:] = True
        with torch.no_grad():
            s1, w1, o1 = net(mel, temp_w=1)
            s2, w2, _ = net(mel, temp_w=0.5, pad_mask=pad)
        assert tuple(s1.shape) == (batch, 10, 1000) and tuple(w1.shape) == (batch, 10)
        r = dict(strong=relmax(s1, g["strong"]), weak=relmax(w1, g["weak"]), at=relmax(o1["at_out"], g["at_out"]),
                 strong_pad=relmax(s2, g["strong_pad"]), weak_pad=relmax(w2, g["weak_pad"]))
        print(mode, r)
        assert max(r.values()) < tol, r
        assert float(s2[-1, :, 850:].abs().max()) == 0.0
        if mode == "tf32x3":      # smallest top-2 margin of the recorded output is 3.6 %: the frame decisions must be identical
            assert (s1.argmax(dim=1).cpu().numpy() == g["argmax"]).all()
    finally:
        F.set_precision("bf16")


def test_lora_merge_on_eval_roundtrip():
    from transformer4sed_b200 import functional as F
    from transformer4sed_b200.src_models import lora
    F.set_precision("tf32x3")
    try:
        lin = lora.Linear(64, 96, r=8, lora_alpha=1).cuda()
        with torch.no_grad():
            lin.lora_B.normal_(0, 0.3)
        x = torch.randn(5, 64, device="cuda")
        w0 = lin.weight.detach().clone()
        y_train = lin(x)
        lin.eval()
        assert lin.merged and not torch.equal(lin.weight, w0)
        y_eval = lin(x)
        lin.train()
        assert not lin.merged and relmax(lin.weight, w0) < 1e-6
        ref = torch.nn.functional.linear(x.double(), w0.double(), lin.bias.double()) + (x.double() @ lin.lora_A.double().T @ lin.lora_B.double().T) / 8
        assert relmax(y_train, ref) < 1e-4 and relmax(y_eval, ref) < 1e-4
    finally:
        F.set_precision("bf16")


@pytest.mark.parametrize("mode,tol", [("tf32x3", 2e-4), ("bf16", 4e-2)])
def test_lora_linear_vs_torch(mode, tol):
    from transformer4sed_b200 import functional as F, ops
    F.set_precision(mode)
    try:
        g = torch.Generator(device="cuda").manual_seed(1)
        M_, K, N, r = 300, 128, 192, 8
        x = torch.randn(M_, K, generator=g, device="cuda").requires_grad_(True)
        w = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5)
        b = torch.randn(N, generator=g, device="cuda").requires_grad_(True)
        A = (torch.randn(r, K, generator=g, device="cuda") * 0.3).requires_grad_(True)
        B = (torch.randn(N, r, generator=g, device="cuda") * 0.3).requires_grad_(True)
        res = torch.randn(M_, N, generator=g, device="cuda").requires_grad_(True)
        wgt = torch.randn(M_, N, generator=g, device="cuda")
        for act in (ops.ACT_NONE, ops.ACT_GELU):
            for t in (x, b, A, B, res):
                t.grad = None
            y = F.lora_linear(F.to_act(x), w, b, A, B, 0.125, residual=F.to_act(res), act=act)
            (y.float() * wgt).sum().backward()
            xd, bd, Ad, Bd, rd = [t.detach().double().requires_grad_(True) for t in (x, b, A, B, res)]
            pre = xd @ (w.double() + 0.125 * Bd @ Ad).T + bd
            ref = (torch.nn.functional.gelu(pre) if act == ops.ACT_GELU else pre) + rd
            (ref * wgt.double()).sum().backward()
            assert relmax(y.float(), ref) < tol
            for name, ours, theirs in (("x", x, xd), ("b", b, bd), ("A", A, Ad), ("B", B, Bd), ("res", res, rd)):
                assert relmax(ours.grad, theirs.grad) < 3 * tol, (act, name, relmax(ours.grad, theirs.grad))
    finally:
        F.set_precision("bf16")


@pytest.mark.parametrize("mode,tol", [("tf32x3", 1e-4), ("bf16", 3e-2)])
def test_cnn_ops_vs_torch(mode, tol):
    """One CNN stage (conv3x3 -> BatchNorm(train) -> ContextGating -> AvgPool) in both layouts of the first layer."""
    import torch.nn.functional as TF
    from transformer4sed_b200 import functional as F
    F.set_precision(mode)
    try:
        g = torch.Generator(device="cuda").manual_seed(3)
        B, T, Fq, C1, C2 = 2, 40, 16, 16, 32
        mel = torch.randn(B, Fq, T, generator=g, device="cuda")
        w1 = (torch.randn(C1, 1, 3, 3, generator=g, device="cuda") * 0.3).requires_grad_(True)
        b1 = torch.randn(C1, generator=g, device="cuda").requires_grad_(True)
        w2 = (torch.randn(C2, C1, 3, 3, generator=g, device="cuda") * 0.1).requires_grad_(True)
        b2 = torch.randn(C2, generator=g, device="cuda").requires_grad_(True)
        gam = (1 + 0.1 * torch.randn(C2, generator=g, device="cuda")).requires_grad_(True)
        bet = (0.1 * torch.randn(C2, generator=g, device="cuda")).requires_grad_(True)
        wl = (torch.randn(C2, C2, generator=g, device="cuda") * C2 ** -0.5).requires_grad_(True)
        bl = torch.randn(C2, generator=g, device="cuda").requires_grad_(True)
        params = [w1, b1, w2, b2, gam, bet, wl, bl]
        rm, rv = torch.zeros(C2, device="cuda"), torch.ones(C2, device="cuda")
        x = F.conv3x3(mel, w1, b1, mel_layout=True)                     # [B, T, F, C1]
        x = F.avg_pool(x, 2, 2)
        x = F.conv3x3(x, w2, b2)
        y = F.batch_norm(x, gam, bet, rm, rv, 1e-3, 0.99, True)
        out = F.avg_pool(F.context_gate(y, F.linear(y, wl, bl)), 1, 2)
        wgt = torch.randn(out.shape, generator=g, device="cuda")
        (out.float() * wgt).sum().backward()
        grads = [p.grad.clone() for p in params]
        pd = [p.detach().double().requires_grad_(True) for p in params]
        rmd, rvd = torch.zeros(C2, device="cuda", dtype=torch.float64), torch.ones(C2, device="cuda", dtype=torch.float64)
        xr = TF.conv2d(mel.double().transpose(1, 2).unsqueeze(1), pd[0], pd[1], padding=1)   # [B, C1, T, F]
        xr = TF.avg_pool2d(xr, (2, 2))
        xr = TF.conv2d(xr, pd[2], pd[3], padding=1)
        yr = TF.batch_norm(xr, rmd, rvd, pd[4], pd[5], True, 0.99, 1e-3)
        lin = TF.linear(yr.permute(0, 2, 3, 1), pd[6], pd[7]).permute(0, 3, 1, 2)
        outr = TF.avg_pool2d(yr * torch.sigmoid(lin), (1, 2))
        (outr.permute(0, 2, 3, 1) * wgt.double()).sum().backward()
        assert relmax(out.float(), outr.permute(0, 2, 3, 1)) < tol
        assert relmax(rm, rmd) < tol and relmax(rv, rvd) < tol
        for i, (a, b) in enumerate(zip(grads, pd)):
            if i == 3:   # b2 sits in front of a training-mode BatchNorm: its true gradient is 0, what is left is rounding noise
                assert a.double().norm().item() < 10 * tol * grads[2].double().norm().item()
                continue
            assert relmax(a, b.grad) < 5 * tol, (i, relmax(a, b.grad))
        # eval mode: running statistics
        with torch.no_grad():
            ye = F.batch_norm(x.detach(), gam, bet, rm, rv, 1e-3, 0.99, False)
            yre = TF.batch_norm(xr.detach(), rmd, rvd, pd[4], pd[5], False, 0.99, 1e-3)
        assert relmax(ye.float(), yre.permute(0, 2, 3, 1)) < tol
    finally:
        F.set_precision("bf16")


def test_merge_prototype_and_dropout_ops():
    import torch.nn.functional as TF
    from transformer4sed_b200 import functional as F
    F.set_precision("tf32x3")
    try:
        g = torch.Generator(device="cuda").manual_seed(4)
        a = torch.randn(3, 50, 64, generator=g, device="cuda").requires_grad_(True)
        b = torch.randn(3, 50, 64, generator=g, device="cuda").requires_grad_(True)
        w = torch.tensor([0.37], device="cuda", requires_grad=True)
        out = F.scale_add(a, b, w)
        wgt = torch.randn(out.shape, generator=g, device="cuda")
        (out * wgt).sum().backward()
        assert relmax(out, a + w * b) < 1e-6 and relmax(a.grad, wgt) < 1e-6 and relmax(b.grad, 0.37 * wgt) < 1e-6
        assert relmax(w.grad, (wgt * b).sum().reshape(1)) < 1e-5
        # prototype head
        x = torch.randn(2, 40, 96, generator=g, device="cuda").requires_grad_(True)
        protos = TF.normalize(torch.randn(30, 96, generator=g, device="cuda"), dim=-1)
        p = F.prototype_predict(x, protos)
        wp = torch.randn(p.shape, generator=g, device="cuda")
        (p * wp).sum().backward()
        xd = x.detach().double().requires_grad_(True)
        s = TF.normalize(xd, dim=-1) @ protos.double().T
        pr = torch.sigmoid((TF.leaky_relu(s, 0.2) * 2 - 1) / 0.1)
        (pr * wp.double()).sum().backward()
        assert relmax(p, pr) < 1e-4 and relmax(x.grad, xd.grad) < 1e-3
        # dropout fused in the gate: keep rate, scaling, and a backward that reuses the same mask
        y = torch.ones(1 << 16, 16, device="cuda", requires_grad=True)
        lin = torch.zeros(1 << 16, 16, device="cuda", requires_grad=True)
        o = F.context_gate(y, lin, dropout_p=0.5)
        kept = (o != 0)
        assert abs(kept.float().mean().item() - 0.5) < 0.01
        assert torch.allclose(o[kept], torch.full_like(o[kept], 0.5 / 0.5))      # y * sigmoid(0) / (1 - p)
        o.sum().backward()
        assert torch.equal(y.grad != 0, kept)
        o2 = F.context_gate(y, lin, dropout_p=0.5)
        assert not torch.equal(o2 != 0, kept)                                   # a fresh mask per call
    finally:
        F.set_precision("bf16")
