"""GPU parity proper: the drop-in `PaSST_SED` (CUDA kernels through the C ABI) against the golden vectors recorded from
the UNMODIFIED reference (tests/golden/, oracle/make_golden.py) and against the CPU oracle on fresh inputs.

Contract (BASELINE.json north_star): bit-exact frame-label argmax, <= 1e-3 relative on activations / loss.  That contract
is checked in the strict `tf32x3` mode (error-compensated tensor-core GEMMs, fp32 activations).  The bf16 performance mode
is checked against the same vectors with bf16-sized tolerances and its argmax tie-rate is reported, not asserted exact.
"""
import numpy as np
import pytest
import torch

from conftest import checksum
from transformer4sed_b200 import schema
from transformer4sed_b200.utils import synth

pytestmark = pytest.mark.gpu

SMALL = dict(embed_dim=192, decoder_dim=192, decoder="transformerXL", decoder_layer_num=1, at_adapter=True, f_pool="mean_pool", mlm=False)
BASE = dict(passt_feature_layer=10, f_pool="mean_pool", decode_ratio=10, at_adapter=True, decoder="transformerXL", decoder_layer_num=3,
            decoder_pos_emd_len=1000, mlm=False)
PRE = dict(BASE, mlm=True, mlm_dict=dict(strategy="block", block_width=10, mask_rate=0.75, out_dim=768))


def build(kw, seed):
    from transformer4sed_b200.src_models.passt.passt_sed import PaSST_SED
    net = PaSST_SED(load_pretrained_model=False, **kw)
    sd = synth.synth_state_dict_like(net, seed)
    net.load_state_dict(sd, strict=True)
    return net.cuda(), sd


def relmax(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def run_golden(tag, kw, seed, batch, golden, mode):
    from transformer4sed_b200 import functional as F
    g = golden(f"matsed_{tag}.npz")
    F.set_precision(mode)
    try:
        net, sd = build(kw, seed)
        np.testing.assert_allclose(checksum(torch.cat([v.flatten() for _, v in sorted(sd.items())])), g["sd_ck"], rtol=1e-12)
        net.eval()
        ext = net.get_feature_extractor().eval()
        wav = synth.synth_wav(batch, 320000, seed=seed + 1)
        np.testing.assert_allclose(checksum(wav), g["wav_ck"], rtol=1e-12)
        mel = ext.normalize(ext(wav.cuda()))
        labels = synth.synth_strong_labels(batch, 10, 1000, seed + 2).cuda()
        weak_labels = (labels.sum(-1) > 0).float()
        cap = {}
        net.decoder.register_forward_hook(lambda m, i, o: cap.__setitem__("decoder_out", o))
        net.interpolate_module.register_forward_hook(lambda m, i, o: cap.__setitem__("interp", o))
        strong, weak, other = net(mel, temp_w=1)
        loss = F.bce_loss(strong, labels) + 0.5 * F.bce_loss(weak, weak_labels) + 2.0 * F.bce_loss(other["at_out"], weak_labels)
        loss.backward()
        res = dict(
            strong=relmax(strong, g["strong"]), weak=relmax(weak, g["weak"]), at_out=relmax(other["at_out"], g["at_out"]),
            fbm=relmax(other["frame_before_mask"].float()[:, ::8, ::4], g["frame_before_mask"]),
            dec=relmax(cap["decoder_out"].float()[:, ::8, ::4], g["decoder_out"]),
            loss=abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])),
            argmax_mismatch=float((strong.argmax(dim=1).cpu().numpy() != g["argmax"]).mean()),
        )
        assert torch.equal(cap["interp"], other["frame_before_mask"])
        gerr = {}
        params = dict(net.named_parameters())
        for name, norm, head in zip(g["grad_names"], g["grad_norms"], g["grad_heads"]):
            p = params[str(name)]
            assert p.grad is not None, name
            gn = p.grad.double().norm().item()
            n = min(8, p.grad.numel())
            gerr[str(name)] = max(abs(gn - norm) / max(norm, 1e-9),
                                  float(np.abs(p.grad.flatten()[:n].double().cpu().numpy() - head[:n]).max()) / max(norm, 1e-9))
        res["grad_worst"] = max(gerr.values())
        res["grad_worst_name"] = max(gerr, key=gerr.get)
        with torch.no_grad():
            pad = torch.zeros(batch, 1000, dtype=torch.bool, device="cuda")
            pad[-1, 900:] = True
            sp, wp, _ = net(mel, temp_w=1, pad_mask=pad)
            s5, w5, _ = net(mel, temp_w=0.5)
        res["pad"] = max(relmax(sp, g["strong_pad"]), relmax(wp, g["weak_pad"]))
        res["t05"] = max(relmax(s5, g["strong_t05"]), relmax(w5, g["weak_t05"]))
        return res
    finally:
        F.set_precision("bf16")


@pytest.mark.parametrize("tag,kw,seed,batch", [("small", SMALL, 3, 2), ("base", BASE, 4, 1)])
def test_strict_mode_meets_contract(golden, tag, kw, seed, batch):
    r = run_golden(tag, kw, seed, batch, golden, "tf32x3")
    print(tag, "tf32x3", r)
    for k in ("strong", "weak", "at_out", "fbm", "dec", "loss", "pad", "t05"):
        assert r[k] < 1e-3, (k, r)
    assert r["argmax_mismatch"] == 0.0, r
    assert r["grad_worst"] < 5e-3, r


@pytest.mark.parametrize("tag,kw,seed,batch", [("small", SMALL, 3, 2), ("base", BASE, 4, 1)])
def test_tf32_mode(golden, tag, kw, seed, batch):
    r = run_golden(tag, kw, seed, batch, golden, "tf32")
    print(tag, "tf32", r)
    for k in ("strong", "weak", "at_out", "loss"):
        assert r[k] < 1e-2, (k, r)
    assert r["argmax_mismatch"] < 0.01, r


@pytest.mark.parametrize("tag,kw,seed,batch", [("small", SMALL, 3, 2), ("base", BASE, 4, 1)])
def test_bf16_mode(golden, tag, kw, seed, batch):
    """The benchmarked mode.  Thresholds = 3 x the values measured on B200 (probabilities 8e-3, gradients 1e-2, no argmax
    flips on the base fixture), so a regression of the bf16 path fails here and not only in the strict mode."""
    r = run_golden(tag, kw, seed, batch, golden, "bf16")
    print(tag, "bf16", r)
    for k in ("strong", "weak", "at_out"):
        assert r[k] < 2.5e-2, (k, r)
    assert r["loss"] < 1e-2, r
    # the 192-wide synthetic model has near-tied class probabilities on ~2 % of its frames (flips there are ties: its value errors
    # are the same 7e-3 as the base model's); the base fixture must not flip at all
    assert r["argmax_mismatch"] <= (0.05 if tag == "small" else 0.0), r
    assert r["grad_worst"] < 4e-2, r


def test_batch_invariance_at_bench_shape():
    """Nothing above checks the B=64 shape the benchmark runs (GEMM tiles, split-K, CTA-pair selection and the attention grids all depend
    on M = B x 1190).  64 clips = 8 distinct clips repeated 8 times: every row must equal the B=1 result of the same clip, and the
    gradient of the mean loss must equal the gradient of the 8-clip batch (bf16 mode, the benchmarked configuration)."""
    from transformer4sed_b200 import functional as F
    F.set_precision("bf16")
    net, _ = build(BASE, 4)
    net.train()
    ext = net.get_feature_extractor().eval()
    wav8 = synth.synth_wav(8, 320000, seed=21).cuda()
    y8 = synth.synth_strong_labels(8, 10, 1000, 22).cuda()
    yw8 = (y8.sum(-1) > 0).float()

    def run(wav, y, yw):
        for p in net.parameters():
            p.grad = None
        strong, weak, other = net(ext.logmel(wav))
        loss = F.bce_loss(strong, y) + 0.5 * F.bce_loss(weak, yw) + 2.0 * F.bce_loss(other["at_out"], yw)
        loss.backward()
        return strong.detach().float(), weak.detach().float(), loss.item(), {n: p.grad.double().norm().item() for n, p in net.named_parameters()
                                                                                if p.grad is not None}

    s64, w64, l64, g64 = run(wav8.repeat(8, 1), y8.repeat(8, 1, 1), yw8.repeat(8, 1))
    s8, w8, l8, g8 = run(wav8, y8, yw8)
    s1, w1, _, _ = run(wav8[3:4], y8[3:4], yw8[3:4])
    assert relmax(s64[:8], s8) < 1e-2 and relmax(s64[56:], s8) < 1e-2, (relmax(s64[:8], s8), relmax(s64[56:], s8))
    assert relmax(s64[3:4], s1) < 1e-2 and relmax(w64[3:4], w1) < 1e-2
    assert float((s64[:8].argmax(1) != s8.argmax(1)).float().mean()) <= 2e-3
    assert abs(l64 - l8) / abs(l8) < 2e-3, (l64, l8)
    assert set(g64) == set(g8)
    worst = max(abs(g64[n] - g8[n]) / max(g8[n], 1e-9) for n in g8 if g8[n] > 1e-7)
    assert worst < 3e-2, worst


def test_fused_bias_gradients_match_column_sum_pass():
    """Bias gradients accumulated by the producers of the output gradients (LayerNorm backward -> proj / fc2, fused attention backward ->
    qkv, the fc2-dgrad GEMM epilogue -> fc1) against the separate column-sum pass they replace (bf16 mode, base width)."""
    from transformer4sed_b200 import functional as F
    F.set_precision("bf16")
    net, _ = build(BASE, 4)
    net.train()
    ext = net.get_feature_extractor().eval()
    mel = ext.logmel(synth.synth_wav(2, 320000, seed=31).cuda())
    y = synth.synth_strong_labels(2, 10, 1000, 32).cuda()

    def grads(fused):
        F.set_fused_bias_grad(fused)
        try:
            for p in net.parameters():
                p.grad = None
            strong, weak, other = net(mel)
            (F.bce_loss(strong, y) + F.bce_loss(weak, (y.sum(-1) > 0).float())).backward()
            return {n: p.grad.double().clone() for n, p in net.named_parameters() if p.grad is not None and n.endswith("bias")}
        finally:
            F.set_fused_bias_grad(True)

    a, b = grads(True), grads(False)
    assert set(a) == set(b) and len(a) > 60
    for n in a:
        scale = b[n].abs().max().clamp_min(1e-12)
        assert ((a[n] - b[n]).abs().max() / scale).item() < 2e-3, n      # same bf16 values summed in fp32, different order only


def test_mlm_pretrain_forward_matches_reference(golden):
    """MAT-SED pre-training (mlm=True): same mask as the reference for the same torch seed; B>1 keeps the upstream no-op."""
    from transformer4sed_b200 import functional as F
    g = golden("matsed_mlm_base_b2.npz")
    F.set_precision("tf32x3")
    try:
        net, _ = build(PRE, 6)
        net.train()
        ext = net.get_feature_extractor().eval()
        wav = synth.synth_wav(2, 320000, seed=7)
        mel = ext.normalize(ext(wav.cuda()))
        # the reference draws on the CPU generator; replay its first draw to fix the mask, then check ours reproduces the loss
        torch.manual_seed(9)
        pred, other = net(mel)
        mask_ref = torch.from_numpy(np.unpackbits(g["mask"])[:2000].astype(bool)).view(2, 1000)
        from oracle import model as OM
        mask_from_noise = OM.block_mask_from_noise(torch.from_numpy(g["noise"]), 0.75, 10, 1000)
        assert torch.equal(mask_ref, mask_from_noise)
        assert abs(other["mask_id_seq"].float().mean().item() - float(g["masked_frac"])) < 1e-6   # 76 of 100 blocks
        assert relmax(pred.float()[:, ::8, ::4], g["pred"]) < 1e-3
        assert relmax(other["at_out"], g["at_out"]) < 1e-3
        fbm = other["frame_before_mask"]
        loss = F.mse_loss(fbm.detach(), pred, mask_ref.cuda())
        assert abs(loss.item() - float(g["loss"])) / float(g["loss"]) < 1e-3
        loss.backward()
        assert net.mask_token.grad is None   # upstream no-op masking: the mask token never reaches the decoder
        assert net.mlm_mlp[2].weight.grad is not None
        # with the sliding-window fusion the mixed frame sequence is contiguous upstream (`mix_rate * x_local + ...` takes x_local's layout),
        # `clone().reshape(-1, C)` is a view and the mask IS written: the mask token takes part and receives a gradient
        net.zero_grad(set_to_none=True)
        torch.manual_seed(9)
        pred_w, other_w = net(mel, encoder_win=True)
        assert torch.equal(other_w["mask_id_seq"].cpu(), other["mask_id_seq"].cpu())     # same draws, same mask
        F.mse_loss(other_w["frame_before_mask"].detach(), pred_w, mask_ref.cuda()).backward()
        assert net.mask_token.grad is not None and float(net.mask_token.grad.abs().sum()) > 0
    finally:
        F.set_precision("bf16")
